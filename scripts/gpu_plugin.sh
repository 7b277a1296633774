#!/bin/bash
# Plug-in path: block tests, then the cfg2 record (sync / pipelined / plug-in figures).
TAG=${1:-plugin}; shift; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests/test_blocks_gpu.py tests/test_zz_lookback_gpu.py tests/test_parity_gpu.py -q -x --timeout=600 \
   -k "blocks or lookback or upload or state_moves or reserve or sharded or pipelined" 2>&1 | tail -8 | tee $OUT/pytest.log
for e in "WR_SYNC_SPLIT=0" "$@"; do
env $e timeout 600 python bench.py --workload cfg2 --subs none --no-cpu-baseline --steps 500 2>$OUT/err.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('[$e] cfg2 value %.0f step %.2f us | e2e sync %.0f pipelined %.0f u8 %.0f u8pipe %.0f plugin %s %s' % (d['value'], d['ms_per_step']*1e3, e['value'], e['pipelined_value'], e['u8_value'], e['u8_pipelined_value'], e.get('plugin_value'), e.get('plugin_parity')))"
done
tail -3 $OUT/err.log
