#!/bin/bash
# Sweep of v2 launch parameters (env WR_V2_NC / WR_V2_RB) on cfg2 / cfg3, plus one ncu capture.
TAG=${1:-sweep}; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() { # workload, env...
  w=$1; shift
  env "$@" timeout 300 python bench.py --workload $w --variant 2 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$w $*: value %.0f MS/s  kernel_ms %.4f  frac %.3f' % (d['value'], r['kernel_ms'], r['frac']))" | tee -a $OUT/sweep.log
}
for rb in 1 2 4 8; do run cfg2 WR_V2_RB=$rb; done
for nc in 12 16 20; do run cfg2 WR_V2_NC=$nc; done
for nc in 12 16; do run cfg3 WR_V2_NC=$nc; done
bash scripts/gpu_prof.sh $TAG 2 cfg2
bash scripts/gpu_prof.sh $TAG 2 cfg3
