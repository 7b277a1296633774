#!/bin/bash
# The default bench line under a list of environment settings, two repetitions each.
# Usage: bash scripts/gpu_env.sh <workload> "ENV=VAL ..." "ENV=VAL ..." ...
w=$1; shift
for rep in 1 2; do
for e in "$@"; do
  env $e timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 1000 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w [$e]: value %.0f  step %.2f us  chan %.2f  audio %.2f  e2e %.0f' % (d['value'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['audio_kernel_ms']*1e3, d['e2e']['value']))"
done; done
