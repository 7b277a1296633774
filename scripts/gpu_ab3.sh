#!/bin/bash
# Parity tests, then cfg3 device-resident with env variants.  Usage: bash scripts/gpu_ab3.sh <tag> ["ENV=.." ...]
TAG=${1:-ab3}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -8 | tee $OUT/pytest.log
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$1: value %.0f step %.2f us  chan %.2f us  demod %.2f us  frac %.3f  variant %s parity %s' % (d['value'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r.get('audio_kernel_ms',0)*1e3, r['frac'], d.get('kernel_variant'), (d.get('parity') or {}).get('bit_exact')))
except Exception as e: print('$1: FAILED', e)"; }
for e in "WR_NOP=1" "$@"; do
  env $e timeout 300 python bench.py --workload cfg3 --subs none --no-cpu-baseline --no-e2e 2>>$OUT/err.log | show "cfg3 [$e]" | tee -a $OUT/results.txt
done
tail -3 $OUT/err.log
