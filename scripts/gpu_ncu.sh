#!/bin/bash
# One ncu --set full capture of a kernel in a bench run.  Usage: bash scripts/gpu_ncu.sh <tag> <kernel regex> <name> [bench flags...]
TAG=$1; KRE=$2; NAME=$3; shift 3
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 1 -o $OUT/$NAME -f \
  python bench.py --no-cpu-baseline --no-e2e --subs none --steps 4 --warmup 3 "$@" > $OUT/$NAME.log 2>&1
tail -3 $OUT/$NAME.log; ls -la $OUT
