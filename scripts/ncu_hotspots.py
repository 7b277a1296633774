#!/usr/bin/env python
"""Where a kernel's warps wait, from one `ncu --set full --import-source on` capture:
  python scripts/ncu_hotspots.py <file.ncu-rep> [top]
Prints (1) stall samples by reason, (2) by opcode of the instruction the warp was waiting ON (the
sampled program counter: for a scoreboard stall that is the CONSUMER of the slow result), with
executed warp-instructions next to it, and (3) the `top` instructions with the most samples and the
instruction in front of each (usually the producer).  SASS level, so it needs no source mapping."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
name = rows[0][1] if len(rows[0]) > 1 else "?"
hdr = rows[1]
data = [r for r in rows[2:] if r and r[0].startswith("0x")]
ix = {k: i for i, k in enumerate(hdr)}
reasons = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]


def iv(r, k):
    try:
        return int(r[ix[k]])
    except (ValueError, IndexError):
        return 0


def opcode(src):
    s = src.strip()
    s = re.sub(r"^@!?U?P\d+\s+", "", s)
    return s.split()[0].split(".")[0] if s else "?"


S = "Warp Stall Sampling (All Samples)"
tot_s = sum(iv(r, S) for r in data)
tot_e = sum(iv(r, "Instructions Executed") for r in data)
print(f"{name}\n{len(data)} SASS instructions, {tot_e} warp-instructions executed, {tot_s} stall samples\n")
print("stall samples by reason:")
by = sorted(((sum(iv(r, k) for r in data), k) for k in reasons), reverse=True)
for n, k in by:
    if n:
        print(f"  {100 * n / max(tot_s, 1):5.1f}%  {k}")
print("\nby opcode of the sampled instruction:   samples   share | executed   share | dominant reasons")
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in data:
    a = agg[opcode(r[ix["Source"]])]
    a[0] += iv(r, S)
    a[1] += iv(r, "Instructions Executed")
    for k in reasons:
        a[2][k] += iv(r, k)
for op, (s, e, c) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:18]:
    dom = ", ".join(f"{k[6:]} {100 * v / max(s, 1):.0f}%" for k, v in c.most_common(3) if v)
    print(f"  {op:12s} {s:9d}  {100 * s / max(tot_s, 1):5.1f}% | {e:9d} {100 * e / max(tot_e, 1):5.1f}% | {dom}")
print(f"\ntop {top} instructions by stall samples (and the instruction in front of each):")
order = sorted(range(len(data)), key=lambda i: -iv(data[i], S))[:top]
for i in order:
    r = data[i]
    c = collections.Counter({k: iv(r, k) for k in reasons})
    dom = ", ".join(f"{k[6:]} {v}" for k, v in c.most_common(2) if v)
    prev = data[i - 1][ix["Source"]].strip() if i else ""
    print(f"  {iv(r, S):7d} {100 * iv(r, S) / max(tot_s, 1):5.1f}%  {r[ix['Source']].strip()[:58]:58s} <- {prev[:44]:44s} [{dom}]")
sh = sum(iv(r, "L1 Wavefronts Shared") for r in data)
shi = sum(iv(r, "L1 Wavefronts Shared Ideal") for r in data)
if sh:
    print(f"\nshared-memory wavefronts: {sh} (ideal {shi}, x{sh / max(shi, 1):.2f})")
