#!/bin/bash
# cfg1 (one receiver) under the launch hand-over knobs, with the block timeline of the default.
TAG=${1:-cfg1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
show() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']; r=d['roofline']
print('$1: value %.0f  step %.2f us  chan %.2f us  demod %.2f us  e2e %.0f' % (d['value'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['audio_kernel_ms']*1e3, e['value']))"; }
run() { name=$1; shift; env "$@" timeout 200 python bench.py --workload cfg1 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_$name.json | show "cfg1 $*"; }
run default WR_NONE=0
run waitearly WR_WAIT_LATE=0
run demod0 WR_DEMOD_PER_SM=0
run demod0_waitearly WR_DEMOD_PER_SM=0 WR_WAIT_LATE=0
run demod1 WR_DEMOD_PER_SM=1
run nopdl WR_V3_PDL=0
WR_TRACE=$OUT/trace_cfg1.csv WR_TRACE_CTA=$OUT/cta_cfg1.csv timeout 200 python bench.py --workload cfg1 --no-cpu-baseline --steps 500 > /dev/null 2>&1
{ python scripts/trace_summary.py $OUT/trace_cfg1.csv 20 480; echo; python scripts/cta_summary.py $OUT/cta_cfg1.csv; } > $OUT/timeline_cfg1.txt 2>&1
cat $OUT/timeline_cfg1.txt | head -40
