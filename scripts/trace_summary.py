#!/usr/bin/env python
"""Summarise a WR_TRACE file (per-block host and device timestamps of the pipelined host path).

Usage: python scripts/trace_summary.py trace.csv [first] [last]
Prints, over blocks [first, last): the step period on the device, the channel kernel's duration and how
long its loaders waited for the tuner block, how far the next block's channel kernel overlaps this
block's demodulator kernel, and the host time spent in wr_bank_submit / wr_bank_wait.
"""
import sys
import numpy as np


def main():
    path = sys.argv[1]
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    last = int(sys.argv[3]) if len(sys.argv) > 3 else 1800
    banks, cur = [], None
    for line in open(path):
        if line.startswith("# bank"):
            cur = {"hdr": line[2:].strip(), "rows": []}
            banks.append(cur)
        elif line[0].isdigit() and cur is not None:
            cur["rows"].append([int(x) for x in line.split(",")])
    for b in banks:
        a = np.array(b["rows"], dtype=np.int64)
        if len(a) == 0:
            continue
        print(b["hdr"])
        # segments: device-resident blocks (no host timestamps) and host-path blocks, cut where the
        # launch rhythm breaks for more than 1 ms
        cs_all = a[:, 4]
        cuts = [0] + [i for i in range(1, len(a)) if (a[i, 1] == 0) != (a[i - 1, 1] == 0) or cs_all[i] - cs_all[i - 1] > 1000000] + [len(a)]
        for lo, hi in zip(cuts[:-1], cuts[1:]):
            seg = a[lo:hi]
            if len(seg) < first + 10:
                continue
            seg = seg[first:min(last, len(seg))]
            seq, hs, hd, hw, cs, ci, ce, ds, de = seg.T
            us = lambda x: float(np.mean(x)) / 1e3
            kind = "host path" if hs[0] else "device-resident"
            print("  %s, blocks %d..%d: device step %.2f us | chan %.2f us (input wait %.2f) | demod %.2f us | "
                  "chan(n+1) starts %.2f us before demod(n) ends | chan(n) end -> demod(n) start %.2f us"
                  % (kind, seq[0], seq[-1], us(np.diff(cs)), us(ce - cs), us(ci - cs), us(de - ds), us(de[:-1] - cs[1:]), us(ds - ce)))
            if hs[0]:
                print("    host: submit %.2f us/call | submit period %.2f us | wait returns %.2f us after submit"
                      % (us(hd - hs), us(np.diff(hs)), us(hw - hs)))


if __name__ == "__main__":
    main()
