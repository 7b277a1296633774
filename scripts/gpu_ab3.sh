#!/bin/bash
# A/B of several builds of the library on ONE box, alternating: libwebradio_b200.so (base) against the
# experimental twins named on the command line.  Usage: bash scripts/gpu_ab3.sh <tag> <workload> <rounds> <lib suffix>...
TAG=$1; W=$2; N=$3; shift 3; OUT=gpurun_out/$TAG; mkdir -p $OUT
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$1: value %.0f step %.2f us  kernel %.2f us  frac %.3f  parity %s' % (d['value'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['frac'], (d.get('parity') or {}).get('max_error_over_frame_peak', (d.get('parity') or {}).get('bit_exact'))))
except Exception as e: print('$1: FAILED', e)"; }
for i in $(seq 1 $N); do
  timeout 300 python bench.py --workload $W --subs none --no-cpu-baseline --no-e2e 2>>$OUT/err.log | show "base $W" | tee -a $OUT/results.txt
  for s in "$@"; do
    WEBRADIO_B200_LIB=$PWD/webradio_b200/libwebradio_b200_$s.so timeout 300 python bench.py --workload $W --subs none --no-cpu-baseline --no-e2e 2>>$OUT/err.log | show "$s $W" | tee -a $OUT/results.txt
  done
done
