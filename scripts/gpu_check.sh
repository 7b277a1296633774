#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines and ncu evidence -> gpurun_out/.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
echo "== pytest -m gpu" 
timeout 1200 python -m pytest tests -m gpu -q -x --timeout=600 2>&1 | tail -40 | tee $OUT/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench cfg2"
timeout 900 python bench.py 2>$OUT/bench_cfg2.err | tee $OUT/bench_cfg2.json
echo "== bench cfg3"
timeout 900 python bench.py --workload cfg3 --no-cpu-baseline 2>$OUT/bench_cfg3.err | tee $OUT/bench_cfg3.json
echo "== bench reference arm"
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json
echo "== ncu launch list (cfg2, short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_cfg2.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_cfg2.log 2>&1
echo "== ncu full, fused channel kernel (cfg2 and cfg3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 4 -c 2 -o $OUT/prof_chan_cfg2 -f \
  python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 4 -c 1 -o $OUT/prof_chan_cfg3 -f \
  python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg3.log 2>&1
ls -la $OUT
