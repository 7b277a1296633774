#!/bin/bash
# Host-path follow-up visit: e2e of the multi-stream workloads with every pipeline slot warmed up,
# and the default line under the copy-out hand-over schemes (WR_HAND_OUT: 0 event, 1 stream
# wait-value, 2 the demodulator kernel stores straight into the caller's pinned buffer).
# Usage (under gpurun): bash scripts/gpu_e2e_knobs.sh [tag]
TAG=${1:-e2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
show() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('$1: value %.0f  step %.2f us  e2e %.0f (%s)  sync %.0f' % (d['value'], d['ms_per_step']*1e3, e['value'], ' '.join('%.0f' % v for v in e.get('values', [])), e['sync_value']))"; }
for w in cfg3 cfg5; do
  timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | tee $OUT/bench_$w.json | show $w
done
timeout 300 python bench.py --workload cfg3 --input u8 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_cfg3_u8.json | show cfg3_u8
timeout 300 python bench.py --workload cfg5 --input u8 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_cfg5_u8.json | show cfg5_u8
for ho in 0 1 2; do
  WR_HAND_OUT=$ho timeout 300 python bench.py --no-cpu-baseline 2>/dev/null | tee $OUT/bench_cfg2_ho$ho.json | show "cfg2 f32 WR_HAND_OUT=$ho"
  WR_HAND_OUT=$ho timeout 300 python bench.py --input u8 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_cfg2_u8_ho$ho.json | show "cfg2 u8 WR_HAND_OUT=$ho"
done
