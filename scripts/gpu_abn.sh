#!/bin/bash
# A/B/n of several builds of the library on ONE box.  Usage: bash scripts/gpu_abn.sh <tag> "<workloads>" <rounds> <lib.so> [<lib.so> ...]
TAG=$1; WS=$2; N=$3; shift 3; OUT=gpurun_out/$TAG; mkdir -p $OUT
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('%-34s value %.0f step %.2f us  kernel %.2f us  second %.2f us  frac %.3f parity %s' % ('$1', d['value'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r.get('audio_kernel_ms',0)*1e3, r['frac'], (d.get('parity') or {}).get('bit_exact')))
except Exception as e: print('$1: FAILED', e)"; }
for i in $(seq 1 $N); do for w in $WS; do for lib in "$@"; do
  WEBRADIO_B200_LIB=$PWD/$lib timeout 300 python bench.py --workload $w --subs none --no-cpu-baseline --no-e2e 2>>$OUT/err.log | show "$w $(basename $lib)" | tee -a $OUT/results.txt
done; done; done
