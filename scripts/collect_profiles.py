#!/usr/bin/env python
"""Copy the evidence of one scripts/gpu_check.sh visit (gpurun_out/<tag>/) into profiles/ (tracked):
bench lines, ncu --set full summaries, the launch list and profiles/traffic.json.
Usage: python scripts/collect_profiles.py <tag> [prefix]   (prefix defaults to r01)"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else "r01"
src = os.path.join(ROOT, "gpurun_out", tag)
dst = os.path.join(ROOT, "profiles")

BENCH = {"bench_cfg1": "bench_cfg1", "bench_cfg2": "bench_cfg2", "bench_cfg3": "bench_cfg3", "bench_cfg4": "bench_cfg4",
         "bench_cfg5": "bench_cfg5", "bench_cfg2_u8": "bench_cfg2_u8", "bench_cfg3_u8": "bench_cfg3_u8",
         "bench_cfg2_v1": "bench_cfg2_v1kernels", "bench_cfg2_v2": "bench_cfg2_v2kernels",
         "bench_cfg3_v2": "bench_cfg3_v2kernels", "bench_ref": "bench_reference_cfg2",
         "bench_cfg5_u8": "bench_cfg5_u8", "bench": "bench_default", "bench_n2": "bench_default_n2",
         "bench_n8": "bench_default_n8"}
if os.path.exists(os.path.join(src, "bench_ref.json")) and prefix != "r01":
    BENCH["bench_ref"] = "bench_reference"
for a, b in BENCH.items():
    p = os.path.join(src, a + ".json")
    if os.path.exists(p) and os.path.getsize(p):
        shutil.copy(p, os.path.join(dst, f"{prefix}_{b}.json"))

for a, b in (("timeline_cfg2.txt", "timeline_cfg2.txt"), ("ubench_copy.txt", "ubench_copy.txt")):
    if os.path.exists(os.path.join(src, a)) and os.path.getsize(os.path.join(src, a)):
        shutil.copy(os.path.join(src, a), os.path.join(dst, f"{prefix}_{b}"))

if os.path.exists(os.path.join(src, "nproc.txt")):
    shutil.copy(os.path.join(src, "nproc.txt"), os.path.join(dst, f"{prefix}_host_cpu.txt"))

# launch list
LW = "cfg3" if os.path.exists(os.path.join(src, "launches_cfg3.csv")) else "cfg2"
p = os.path.join(src, f"launches_{LW}.csv")
if os.path.exists(p):
    rows = list(csv.reader(open(p)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[h]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) > mv:
            d.setdefault(r[kn], []).append(float(r[mv].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(dst, f"{prefix}_launches_{LW}.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 : python bench.py --steps 20 --warmup 3 "
                f"--no-cpu-baseline{' --no-e2e --subs none' if LW == 'cfg3' else ''} ({LW})\nper-kernel device time, ns (cold-cache, serialised under ncu: compare SHARES, "
                "not absolutes)\n\nlaunches     avg_ns     min_ns     max_ns   share  kernel\n")
        for k, v in d.items():
            f.write(f"{len(v):8d} {sum(v) / len(v):10.0f} {min(v):10.0f} {max(v):10.0f} {100 * sum(v) / tot:6.1f}%  {k[:110]}\n")
        ours = {k: sum(v) for k, v in d.items() if "wrd::" in k}
        t2 = sum(ours.values())
        f.write("\nshare of the step (this library's kernels only):\n")
        for k, v in ours.items():
            f.write(f"  {100 * v / t2:5.1f}%  {k[:110]}\n")

# ncu full summaries + traffic.json
traffic = {}
REPS = {"prof_chan_cfg2": ("cfg2", "chan_kernel"), "prof_chan_cfg3": ("cfg3", "chan_kernel"),
        "prof_spectrum_cfg4": ("cfg4", "spectrum_kernel")}
for rep, (w, kind) in REPS.items():
    p = os.path.join(src, rep + ".ncu-rep")
    if not os.path.exists(p):
        continue
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), p], text=True)
    name = {"prof_chan_cfg2": "ncu_full_chan_cfg2", "prof_chan_cfg3": "ncu_full_chan_cfg3",
            "prof_spectrum_cfg4": "ncu_full_spectrum_cfg4"}[rep]
    open(os.path.join(dst, f"{prefix}_{name}.txt"), "w").write(out)
    raw = subprocess.check_output(["ncu", "-i", p, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, val = rows[0], rows[1], rows[2]

    def get(metric):
        i = hdr.index(metric)
        x = float(val[i].replace(",", ""))
        u = units[i].lower()
        return x * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
    traffic[w] = {
        f"{kind}_dram_bytes_per_launch": int(get("dram__bytes_read.sum") + get("dram__bytes_write.sum")),
        "kernel": val[hdr.index("Kernel Name")],
        "issue_active_pct": get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warp_instructions_per_launch": int(get("smsp__inst_executed.sum")),
        "l1tex_throughput_pct": get("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        "fma_pipe_cycles_active_pct": get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
        "source": f"ncu --set full --clock-control none, one launch (profiles/{prefix}_{name}.txt)",
    }
    # where the warps wait: source-level stall samples of the same capture
    try:
        hot = subprocess.check_output([sys.executable, os.path.join(ROOT, "scripts", "ncu_hotspots.py"), p, "14"], text=True)
        open(os.path.join(dst, f"{prefix}_{name.replace('ncu_full', 'ncu_hotspots')}.txt"), "w").write(hot)
    except subprocess.CalledProcessError:
        pass
if traffic:
    json.dump(traffic, open(os.path.join(dst, "traffic.json"), "w"), indent=1)
print("profiles/ refreshed from", src)
