#!/usr/bin/env python
"""Per-phase breakdown (segments between BAR.SYNC) of an ncu source-page CSV:
   ncu -i rep --page source --csv --kernel-id ::regex:NAME:1 > src.csv ; python scripts/ncu_phases.py src.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
n = len(hdr)
data = [r for r in rows[2:] if len(r) >= n - 2 and r[0].startswith("0x")]
col = {k: hdr.index(k) for k in ["Source", "Instructions Executed", "Warp Stall Sampling (All Samples)",
                                 "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal"]}


def iv(r, k):
    try:
        return int(r[col[k]])
    except ValueError:
        return 0


tot = sum(iv(r, "Instructions Executed") for r in data)
tots = sum(iv(r, "Warp Stall Sampling (All Samples)") for r in data)
print("total warp-inst", tot, "stall samples", tots)
bars = [i for i, r in enumerate(data) if "BAR.SYNC" in r[col["Source"]]]
segs = [0] + bars + [len(data)]
for k in range(len(segs) - 1):
    seg = data[segs[k]:segs[k + 1]]
    e = sum(iv(r, "Instructions Executed") for r in seg)
    s = sum(iv(r, "Warp Stall Sampling (All Samples)") for r in seg)
    w = sum(iv(r, "L1 Wavefronts Shared") for r in seg)
    wi = sum(iv(r, "L1 Wavefronts Shared Ideal") for r in seg)
    print(f"seg {k}: {len(seg)} sass, warp-inst {e} ({100 * e / tot:.1f}%), stall samples {s} ({100 * s / max(tots, 1):.1f}%), "
          f"smem wavefronts {w} (ideal {wi})")
print("top instructions by stall samples:")
for r in sorted(data, key=lambda r: -iv(r, "Warp Stall Sampling (All Samples)"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"  {iv(r, 'Warp Stall Sampling (All Samples)'):6d} smp  {iv(r, 'Instructions Executed'):8d} inst  "
          f"{iv(r, 'L1 Wavefronts Shared'):8d} wf   {r[col['Source']].strip()[:90]}")
