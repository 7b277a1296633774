#!/bin/bash
# A/B of two builds of the library on ONE box: libwebradio_b200_exp.so (A, the previous build) against
# libwebradio_b200.so (B), alternating.  Usage: bash scripts/gpu_ablib.sh <tag> <workload> [rounds]
TAG=${1:-ablib}; W=${2:-cfg3}; N=${3:-3}; OUT=gpurun_out/$TAG; mkdir -p $OUT
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$1: value %.0f step %.2f us  kernel %.2f us  second %.2f us  frac %.3f' % (d['value'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r.get('audio_kernel_ms',0)*1e3, r['frac']))
except Exception as e: print('$1: FAILED', e)"; }
for i in $(seq 1 $N); do
  WEBRADIO_B200_LIB=$PWD/webradio_b200/libwebradio_b200_exp.so timeout 300 python bench.py --workload $W --subs none --no-cpu-baseline --no-e2e 2>>$OUT/err.log | show "A(prev) $W" | tee -a $OUT/results.txt
  timeout 300 python bench.py --workload $W --subs none --no-cpu-baseline --no-e2e 2>>$OUT/err.log | show "B(new)  $W" | tee -a $OUT/results.txt
done
