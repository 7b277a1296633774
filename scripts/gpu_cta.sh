#!/bin/bash
# value / e2e of the default bench line, then a per-CTA trace of six blocks (scripts/cta_summary.py reads it).
# Usage: bash scripts/gpu_cta.sh [tag] [extra env assignments...]
tag=${1:-x}; shift
mkdir -p gpurun_out
for rep in 1 2; do
env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 1000 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('$tag: value %.0f step %.2f us e2e %.0f sync %.0f' % (d['value'], d['ms_per_step']*1e3, e['value'], e['sync_value']))"
done
rm -f gpurun_out/cta_$tag.csv gpurun_out/tr_$tag.csv
env "$@" WR_TRACE=gpurun_out/tr_$tag.csv WR_TRACE_CTA=gpurun_out/cta_$tag.csv timeout 300 python bench.py --no-cpu-baseline --steps 500 >/dev/null 2>&1
