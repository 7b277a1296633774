#!/bin/bash
# The pipelined host path's hand-over schemes side by side (cfg2 by default):
#   WR_HAND_IN  0 event | 1 stream write-value | 2 4-byte copy      (copy-in stream -> channel kernel)
#   WR_HAND_OUT 0 event + copy | 1 stream wait-value + copy | 2 the kernel stores into the pinned buffer
# Usage: bash scripts/gpu_e2e.sh [workload] ["in:out pairs"] [extra bench flags]
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "pipelined" 2>&1 | tail -3
w=${1:-cfg2}
for pair in ${2:-1:0 2:0 0:0 1:2}; do
  i=${pair%%:*}; o=${pair##*:}
  rm -f gpurun_out/trace_$i$o.csv
  WR_TRACE=gpurun_out/trace_$i$o.csv WR_HAND_IN=$i WR_HAND_OUT=$o timeout 300 python bench.py --workload $w --no-cpu-baseline $3 2>gpurun_out/e2e_$i$o.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); e=d['e2e']
print('$w in=$i out=$o: value %.0f MS/s  step %.4f ms  e2e %.0f (%s)  sync %.0f' % (d['value'], d['ms_per_step'], e['value'], e['mode'], e['sync_value']))"
  python scripts/trace_summary.py gpurun_out/trace_$i$o.csv 2>&1 | tail -12
done
