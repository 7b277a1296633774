#!/bin/bash
# v4: knob sweep on cfg3 + one ncu --set full capture
TAG=${1:-v4b}; OUT=gpurun_out/$TAG; mkdir -p $OUT
show() { python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$1: step %.2f us  chan %.2f us  demod %.2f us  frac %.3f  variant %s parity %s' % (d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['audio_kernel_ms']*1e3, r['frac'], d['kernel_variant'], (d.get('parity') or {}).get('bit_exact')))
except Exception as e: print('$1: FAILED', e)"; }
B="python bench.py --workload cfg3 --subs none --no-cpu-baseline --no-e2e --variant 4"
for e in "WR_V4_NOW8=1" "WR_V4_PF=0" "WR_V4_PF=8" "WR_V4_PF=16" "WR_V4_PF=32" "WR_V4_PF=16 WR_V4_PFD=64" "WR_V4_PF=16 WR_V4_G=2" "WR_V4_PF=8 WR_V4_NS=2"; do
  env $e timeout 300 $B 2>>$OUT/err_env.log | show "cfg3 v4 [$e]"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chan_kernel_v4 -s 4 -c 1 -o $OUT/prof_v4_cfg3 -f $B --steps 4 --warmup 3 > $OUT/ncu_v4.log 2>&1
ls -la $OUT | head
