#!/bin/bash
# A/B of launch-level choices on one box, two repetitions each (cfg2 by default).
# Usage: bash scripts/gpu_ab.sh [workload] [steps]
w=${1:-cfg2}; n=${2:-1000}
for rep in 1 2; do
for late in 0 1; do for pers in 0 1; do for pair in 0:0 1:0; do
  i=${pair%%:*}; o=${pair##*:}
  WR_WAIT_LATE=$late WR_DEMOD_PERSIST=$pers WR_HAND_IN=$i WR_HAND_OUT=$o timeout 300 python bench.py --workload $w --no-cpu-baseline --steps $n 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('$w late=$late persist=$pers in=$i out=$o: value %.0f  step %.2f us  e2e %.0f  sync %.0f' % (d['value'], d['ms_per_step']*1e3, e['value'], e['sync_value']))"
done; done; done; done
