#!/bin/bash
# Evidence visit for a SHORT GPU budget: same artefacts and file names as scripts/gpu_check.sh, but
# ordered by importance and written as they come, so that a visit cut off by the budget clamp still
# leaves the headline line, the parity run and the ncu passes behind (scripts/collect_profiles.py
# copies whatever exists).  Usage (under gpurun):  bash scripts/gpu_final.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
lap "bench cfg2 (default line)"
timeout 600 python bench.py 2>$OUT/bench_cfg2.err | tee $OUT/bench_cfg2.json | cut -c1-300
lap "pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
lap "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
lap "bench reference arm (cfg2)"
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json | cut -c1-300
# under ncu the host path hands blocks over by events (WR_HAND_IN=0): a kernel that waits for a copy
# would be timed with its wait
lap "ncu launch list (cfg2, short)"
WR_HAND_IN=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_cfg2.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_cfg2.log 2>&1
lap "ncu full: channel kernel, cfg2"
WR_HAND_IN=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 6 -c 1 -o $OUT/prof_chan_cfg2 -f \
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg2.log 2>&1
lap "block timeline of the default line"
WR_TRACE=$OUT/trace_cfg2.csv WR_TRACE_CTA=$OUT/cta_cfg2.csv timeout 300 python bench.py --no-cpu-baseline --steps 500 > /dev/null 2>&1
{ echo "python bench.py --steps 500 under WR_TRACE / WR_TRACE_CTA (cfg2): per-block and per-CTA device timestamps"; echo;
  python scripts/trace_summary.py $OUT/trace_cfg2.csv 20 480; echo; python scripts/cta_summary.py $OUT/cta_cfg2.csv; } > $OUT/timeline_cfg2.txt 2>&1
for w in cfg3 cfg5; do
  lap "bench $w"
  timeout 600 python bench.py --workload $w --no-cpu-baseline 2>$OUT/bench_$w.err | tee $OUT/bench_$w.json | cut -c1-200
done
lap "bench cfg2 / cfg3 fed raw RTL-SDR bytes"
for w in cfg2 cfg3; do
  timeout 600 python bench.py --workload $w --input u8 --no-cpu-baseline 2>/dev/null | tee $OUT/bench_${w}_u8.json | cut -c1-200
done
for w in cfg4 cfg1; do
  lap "bench $w"
  timeout 600 python bench.py --workload $w --no-cpu-baseline 2>$OUT/bench_$w.err | tee $OUT/bench_$w.json | cut -c1-200
done
lap "ncu full: channel kernel cfg3, spectrum kernel cfg4"
WR_HAND_IN=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 4 -c 1 -o $OUT/prof_chan_cfg3 -f \
  python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg3.log 2>&1
WR_HAND_IN=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:spectrum_kernel -s 3 -c 1 -o $OUT/prof_spectrum_cfg4 -f \
  python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_cfg4.log 2>&1
[ -x build/ubench_copy ] && ./build/ubench_copy > $OUT/ubench_copy.txt 2>&1
lap "done"
ls -la $OUT
