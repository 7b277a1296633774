#!/bin/bash
TAG=${1:-fft}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spectrum_kernel -s 3 -c 1 -o $OUT/prof_cfg4 -f \
  python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_cfg4.log 2>&1
tail -2 $OUT/ncu_cfg4.log
