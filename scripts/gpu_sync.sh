#!/bin/bash
TAG=${1:-sync}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
for e in "A=0" "$@"; do
env $e timeout 600 python bench.py --workload cfg2 --subs none --no-cpu-baseline --steps 500 2>$OUT/err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('[$e] cfg2 value %.0f step %.2f us | e2e sync %.0f pipelined %.0f u8 %.0f u8pipe %.0f plugin %s' % (d['value'], d['ms_per_step']*1e3, e['value'], e['pipelined_value'], e['u8_value'], e['u8_pipelined_value'], e.get('plugin_value')))"
done
