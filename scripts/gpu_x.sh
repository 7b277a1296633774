#!/bin/bash
# e2e of the default line under experiment knobs, with the block trace summarised.
# Usage: bash scripts/gpu_x.sh tag [ENV=VAL ...]
tag=$1; shift
rm -f gpurun_out/tr_$tag.csv
env "$@" WR_TRACE=gpurun_out/tr_$tag.csv timeout 300 python bench.py --no-cpu-baseline --steps 1000 $BENCH_FLAGS 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('$tag: value %.0f step %.2f us e2e %.0f sync %.0f' % (d['value'], d['ms_per_step']*1e3, e['value'], e['sync_value']))"
python - <<PY
import numpy as np
rows=[[int(x) for x in l.split(",")] for l in open("gpurun_out/tr_$tag.csv") if l[0].isdigit()]
a=np.array(rows,dtype=np.int64)
hp=np.where(a[:,1]>0)[0]
seg=a[hp[0]+20:hp[0]+900]
print("   pipelined: period %.2f  input wait %.2f  mixing after input %.2f  ce->ds %.2f  demod %.2f  overlap %.2f  host submit %.2f"%(np.diff(seg[:,4]).mean()/1e3,(seg[:,5]-seg[:,4]).mean()/1e3,(seg[:,6]-seg[:,5]).mean()/1e3,(seg[:,7]-seg[:,6]).mean()/1e3,(seg[:,8]-seg[:,7]).mean()/1e3,(seg[:-1,8]-seg[1:,4]).mean()/1e3,(seg[:,2]-seg[:,1]).mean()/1e3))
PY
