#!/bin/bash
# GPU visit for kernel work: parity tests, then bench of each kernel variant on cfg2 and cfg3.
TAG=${1:-v2}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -40 | tee $OUT/pytest_gpu.log
for v in 1 2; do
  for w in cfg2 cfg3; do
    echo "== bench $w variant $v"
    timeout 600 python bench.py --workload $w --variant $v --no-cpu-baseline 2>$OUT/bench_${w}_v$v.err | tee $OUT/bench_${w}_v$v.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.0f MS/s  ms/step %.4f  kernel_ms %.4f audio_ms %.4f  achieved %.1f GB/s frac %.3f  e2e %.0f' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['audio_kernel_ms'], r['achieved'], r['frac'], d['e2e']['value']))"
  done
done
