#!/bin/bash
# The default bench line (and optionally others) on the GPU box, stderr kept, one-line summaries printed.
# Usage (under gpurun): bash scripts/gpu_bench.sh <tag> [bench.py flags...]
TAG=${1:-bench}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
( time python -X faulthandler bench.py "$@" > $OUT/bench.json ) 2> $OUT/bench.err
echo "rc=$?"; tail -15 $OUT/bench.err
python - "$OUT/bench.json" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
except Exception as e:
    print("no JSON line:", e); sys.exit(0)
def show(n, r):
    e = r.get("e2e") or {}
    c = r.get("cpu_baseline") or {}
    rf = r["roofline"]
    print("%-5s value %9.0f MS/s  step %8.4f ms  kernel %8.4f ms  frac %.4f (step %.4f)  cpu %s (%s cores)  parity %s" % (
        n, r["value"], r["ms_per_step"], rf["kernel_ms"], rf["frac"], rf.get("whole_step_frac", 0),
        ("%.0f" % c["value"]) if c else None, c.get("cores"), (r.get("parity") or {}).get("bit_exact", (r.get("parity") or {}).get("max_error_over_frame_peak"))))
    if e:
        print("      e2e:", {k: round(v) for k, v in e.items() if k.endswith("value") and v})
show(d["config"]["workload"][:4], d)
for k, v in d.get("configs", {}).items():
    show(k, v)
print("clocks", d.get("clocks"), "placement", d.get("placement"), "launches", d.get("gpu_launches"))
PY
