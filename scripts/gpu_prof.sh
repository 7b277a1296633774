#!/bin/bash
# Source-level ncu captures of the dominant kernels (read back with ncu -i ... --page source --csv).
TAG=${1:-prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -5 | tee $OUT/pytest_gpu.log
for w in cfg3 cfg2; do
  echo "== ncu full $w"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:chan_kernel -s 4 -c 1 -o $OUT/prof_chan_$w -f \
    python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_full_$w.log 2>&1
done
echo "== knob sweep cfg3"
for nc in 4 6 8 10; do
  echo "NC=$nc"; WR_V2_NC=$nc timeout 600 python bench.py --workload cfg3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
ls -la $OUT
