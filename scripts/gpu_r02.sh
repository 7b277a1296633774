#!/bin/bash
# Round-2 evidence visit: the driver's own commands first (parity tests, smoke, the default bench
# line and its reference arm), then the ncu passes of the same command -> gpurun_out/<tag>/.
# Ordered by importance and written as they come, so a visit that is cut short still leaves the
# headline behind.  Usage (under gpurun):  bash scripts/gpu_r02.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
lap() { echo "== [$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi > $OUT/nvidia-smi.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | head -20 >> $OUT/nproc.txt
lap "pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -15 | tee $OUT/pytest_gpu.log
lap "smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee $OUT/smoke.log
lap "bench (default line: cfg3 + sub-records)"
timeout 1500 python bench.py 2>$OUT/bench.err | tee $OUT/bench.json | cut -c1-300
lap "bench reference arm"
timeout 900 python bench.py --impl reference 2>$OUT/bench_ref.err | tee $OUT/bench_ref.json | cut -c1-300
# under ncu the host path hands blocks over by events (WR_HAND_IN=0): a kernel that waits for a copy
# would be timed with its wait
lap "ncu launch list of the default command (cfg3, device-resident legs only)"
WR_HAND_IN=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_cfg3.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --subs none > $OUT/ncu_launches_cfg3.log 2>&1
lap "ncu full: channel kernel cfg3 (v4)"
WR_HAND_IN=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 4 -c 1 -o $OUT/prof_chan_cfg3 -f \
  python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e --subs none > $OUT/ncu_full_cfg3.log 2>&1
lap "ncu full: spectrum kernel cfg4"
WR_HAND_IN=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:spectrum_kernel -s 3 -c 1 -o $OUT/prof_spectrum_cfg4 -f \
  python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --subs none > $OUT/ncu_full_cfg4.log 2>&1
lap "ncu full: channel kernel cfg2 (v3)"
WR_HAND_IN=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 6 -c 1 -o $OUT/prof_chan_cfg2 -f \
  python bench.py --workload cfg2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --subs none > $OUT/ncu_full_cfg2.log 2>&1
lap "done"
ls -la $OUT
