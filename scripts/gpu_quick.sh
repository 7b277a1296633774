#!/bin/bash
# Quick A/B of the channel-kernel families on the BASELINE workloads (no CPU baseline).
# Usage: bash scripts/gpu_quick.sh "<variants>" "<workloads>" [extra bench.py flags]
for w in ${2:-cfg3 cfg2 cfg5 cfg1}; do
  for v in ${1:-2 3}; do
    timeout 600 python bench.py --workload $w --variant $v --no-cpu-baseline $3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w v$v $3: value %.0f MS/s  step %.4f ms  chan %.4f ms  audio %.4f ms  frac %.3f  rxframes/s %.1f G  e2e %.0f  variant %d' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['audio_kernel_ms'], r['frac'], r['receiver_frames_per_s']/1e9, d['e2e']['value'], d['kernel_variant']))"
  done
done
