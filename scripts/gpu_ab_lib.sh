#!/bin/bash
# A/B of two builds of the library on one box: build/libwebradio_b200_prev.so against the in-tree one.
# Usage: bash scripts/gpu_ab_lib.sh "<workloads>" [extra bench flags]
for w in ${1:-cfg2 cfg3}; do
  for lib in prev new prev new; do
    if [ $lib = prev ]; then export WEBRADIO_B200_LIB=$PWD/build/libwebradio_b200_prev.so; else unset WEBRADIO_B200_LIB; fi
    timeout 600 python bench.py --workload $w --no-cpu-baseline $2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w $lib: value %.0f MS/s  step %.4f ms  chan %.4f ms  audio %.4f ms  frac %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['audio_kernel_ms'], r['frac'], d['e2e']['value']))"
  done
done
