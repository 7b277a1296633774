#!/bin/bash
TAG=${1:-v4cfg5}; OUT=gpurun_out/$TAG; mkdir -p $OUT
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$1: value %.0f step %.2f us  kernel %.2f us  second %.2f us  variant %s parity %s' % (d['value'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r.get('audio_kernel_ms',0)*1e3, d.get('kernel_variant'), (d.get('parity') or {}).get('bit_exact')))
except Exception as e: print('$1: FAILED', e)"; }
for w in cfg5 cfg2; do for v in 0 4; do
  timeout 300 python bench.py --workload $w --variant $v --subs none --no-cpu-baseline --no-e2e 2>>$OUT/err.log | show "$w variant $v" | tee -a $OUT/results.txt
done; done
tail -3 $OUT/err.log
