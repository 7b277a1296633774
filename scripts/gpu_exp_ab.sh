#!/bin/bash
# A/B of the experimental twin of the library (make lib-exp; EXPFLAGS picks the experiment) against
# the default build on one box: parity subset on the twin first, then alternating bench lines.
# Usage (under gpurun, after `make lib-exp` in the container):  bash scripts/gpu_exp_ab.sh [tag] [workloads]
TAG=${1:-exp}; OUT=gpurun_out/$TAG; mkdir -p $OUT
EXP=$PWD/webradio_b200/libwebradio_b200_exp.so
[ -f "$EXP" ] || { echo "no $EXP: run make lib-exp first"; exit 1; }
echo "== parity on the experimental build"
WEBRADIO_B200_LIB=$EXP timeout 600 python -m pytest tests/test_parity_gpu.py -q -x --timeout=300 \
  -k "golden_chain or ragged or cfg2_full or cfg3_reduced or cfg5_mixed or pipelined" 2>&1 | tail -4 | tee $OUT/pytest_exp.log
show() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1: value %.0f  step %.2f us  chan %.2f us  demod %.2f us  frac %.3f' % (d['value'], d['ms_per_step']*1e3, r['kernel_ms']*1e3, r['audio_kernel_ms']*1e3, r['frac']))"; }
for w in ${2:-cfg3 cfg5 cfg2}; do
  for rep in 1 2; do
    timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | tee $OUT/bench_${w}_default_$rep.json | show "$w default #$rep"
    WEBRADIO_B200_LIB=$EXP timeout 300 python bench.py --workload $w --no-cpu-baseline 2>/dev/null | tee $OUT/bench_${w}_exp_$rep.json | show "$w exp     #$rep"
  done
done
