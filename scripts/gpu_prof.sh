#!/bin/bash
# ncu --set full of the fused channel kernel for one variant/workload.  Usage: gpu_prof.sh tag variant workload [extra bench args]
TAG=$1; V=$2; W=$3; shift 3
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 6 -c 2 -o $OUT/prof_${W}_v$V -f \
  python bench.py --workload $W --variant $V --steps 6 --warmup 3 --no-cpu-baseline "$@" > $OUT/ncu_${W}_v$V.log 2>&1
tail -3 $OUT/ncu_${W}_v$V.log
