#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines and ncu evidence -> gpurun_out/<tag>/.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
echo "== bench reference arm (cfg2)"
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json | cut -c1-300
echo "== bench cfg2 (default)"
timeout 900 python bench.py 2>$OUT/bench_cfg2.err | tee $OUT/bench_cfg2.json | cut -c1-300
for w in cfg1 cfg3 cfg5 cfg4; do
  echo "== bench $w"
  timeout 900 python bench.py --workload $w --no-cpu-baseline 2>$OUT/bench_$w.err | tee $OUT/bench_$w.json | cut -c1-200
done
echo "== bench cfg3 / cfg2 fed raw RTL-SDR bytes"
for w in cfg3 cfg2; do
  timeout 900 python bench.py --workload $w --input u8 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_${w}_u8.json | cut -c1-200
done
echo "== bench cfg2, v2 kernels (for the record)"
timeout 900 python bench.py --variant 2 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_cfg2_v2.json | cut -c1-200
echo "== bench cfg3, v2 kernels (for the record)"
timeout 900 python bench.py --workload cfg3 --variant 2 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_cfg3_v2.json | cut -c1-200
echo "== bench cfg2, v1 kernels (for the record)"
timeout 900 python bench.py --variant 1 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_cfg2_v1.json | cut -c1-200
# under ncu the host path hands blocks over by events (WR_HAND_IN=0): a kernel that waits for a copy
# would be timed with its wait
echo "== ncu launch list (cfg2, short)"
WR_HAND_IN=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_cfg2.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_cfg2.log 2>&1
echo "== ncu full: fused channel kernel (cfg2, cfg3), spectrum kernel (cfg4)"
WR_HAND_IN=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 6 -c 1 -o $OUT/prof_chan_cfg2 -f \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg2.log 2>&1
WR_HAND_IN=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 4 -c 1 -o $OUT/prof_chan_cfg3 -f \
  python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg3.log 2>&1
WR_HAND_IN=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:spectrum_kernel -s 3 -c 1 -o $OUT/prof_spectrum_cfg4 -f \
  python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg4.log 2>&1
echo "== block timeline of the default line (WR_TRACE / WR_TRACE_CTA), block-sized PCIe copies"
WR_TRACE=$OUT/trace_cfg2.csv WR_TRACE_CTA=$OUT/cta_cfg2.csv timeout 300 python bench.py --no-cpu-baseline --steps 500 > /dev/null 2>&1
{ echo "python bench.py --steps 500 under WR_TRACE / WR_TRACE_CTA (cfg2): per-block and per-CTA device timestamps"; echo;
  python scripts/trace_summary.py $OUT/trace_cfg2.csv 20 480; echo; python scripts/cta_summary.py $OUT/cta_cfg2.csv; } > $OUT/timeline_cfg2.txt 2>&1
[ -x build/ubench_copy ] && ./build/ubench_copy > $OUT/ubench_copy.txt 2>&1
ls -la $OUT
