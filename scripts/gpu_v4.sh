#!/bin/bash
# v4 channel kernel: parity, then cfg3 device-timed A/B against v3, optional env sweeps.
# Usage: bash scripts/gpu_v4.sh <tag> ["ENV=.. ENV=.." ...]
TAG=${1:-v4}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x --timeout=600 -k "v4" 2>&1 | tail -8 | tee $OUT/pytest_v4.log
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$1: step %.2f us  chan %.2f us  demod %.2f us  frac %.3f  variant %s parity %s' % (d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['audio_kernel_ms']*1e3, r['frac'], d['kernel_variant'], (d.get('parity') or {}).get('bit_exact')))
except Exception as e: print('$1: FAILED', e)"; }
B="python bench.py --workload cfg3 --subs none --no-cpu-baseline --no-e2e"
for v in 3 4; do
  timeout 300 $B --variant $v 2>$OUT/err_$v.log | tee $OUT/bench_cfg3_v$v.json | show "cfg3 v$v"
done
for e in "$@"; do
  env $e timeout 300 $B --variant 4 2>>$OUT/err_env.log | show "cfg3 v4 [$e]"
done
tail -3 $OUT/err_4.log
