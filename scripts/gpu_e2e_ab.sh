#!/bin/bash
# e2e of the default line (float IQ in) under the copy-in hand-over schemes, several repetitions on one box.
for rep in 1 2 3; do
for e in "WR_HAND_IN=1 WR_POLL_NS=1000" "WR_HAND_IN=1 WR_POLL_NS=4000" "WR_HAND_IN=0" "WR_HAND_IN=1 WR_HAND_OUT=1"; do
  env $e timeout 300 python bench.py --no-cpu-baseline --steps 2000 $1 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('[$e]: value %.0f  e2e %.0f  sync %.0f' % (d['value'], e['value'], e['sync_value']))"
done; done
