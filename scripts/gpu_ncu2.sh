#!/bin/bash
# ncu --set full of one kernel for two builds of the library.  Usage: bash scripts/gpu_ncu2.sh <tag> <kernel regex> <workload> <libA> <libB>
TAG=$1; KRE=$2; W=$3; OUT=gpurun_out/$TAG; mkdir -p $OUT
for lib in $4 $5; do
  n=$(basename $lib .so)
  WEBRADIO_B200_LIB=$PWD/$lib timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 1 -o $OUT/prof_$n -f \
    python bench.py --no-cpu-baseline --no-e2e --subs none --steps 4 --warmup 3 --workload $W > $OUT/prof_$n.log 2>&1
  tail -2 $OUT/prof_$n.log
done
ls -la $OUT
