#!/bin/bash
# v4 on cfg3: warps per receiver (WR_V4_G) and anything else passed as "ENV=val ENV=val" strings.
# Usage (under gpurun): bash scripts/gpu_v4g.sh <tag> ["ENV=.. ENV=.." ...]
TAG=${1:-v4g}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$1: step %.2f us  chan %.2f us  demod %.2f us  frac %.3f  variant %s parity %s' % (d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['audio_kernel_ms']*1e3, r['frac'], d['kernel_variant'], (d.get('parity') or {}).get('bit_exact')))
except Exception as e: print('$1: FAILED', e)"; }
B="python bench.py --workload cfg3 --subs none --no-cpu-baseline --no-e2e"
for e in "WR_V4_G=0" "$@"; do
  env $e timeout 300 $B 2>>$OUT/err_env.log | show "cfg3 [$e]" | tee -a $OUT/results.txt
done
tail -5 $OUT/err_env.log
