#!/usr/bin/env python
"""Per-CTA timeline of the two kernels of a block from a WR_TRACE_CTA file (device-resident path).
Usage: python scripts/cta_summary.py cta.csv"""
import sys
import numpy as np


def main():
    rows = [l.strip().split(",") for l in open(sys.argv[1]) if l[0].isdigit()]
    D = {}
    for r in rows:
        D.setdefault((int(r[0]), r[1]), []).append((int(r[2]), int(r[3]), int(r[4])))
    blocks = sorted(set(k[0] for k in D))
    q = lambda x: "min %5.1f  p25 %5.1f  med %5.1f  p75 %5.1f  max %5.1f" % tuple(np.percentile(x, [0, 25, 50, 75, 100]) / 1e3)
    for b in blocks[1:3]:
        if (b, "chan") not in D or (b - 1, "demod") not in D or (b, "demod") not in D:
            continue
        ch, dm, dp = (np.array(sorted(D[k])) for k in ((b, "chan"), (b, "demod"), (b - 1, "demod")))
        t0 = ch[:, 1].min()
        print("block %d: %d channel CTAs, %d demodulator CTAs; times in us relative to the first channel CTA's start" % (b, len(ch), len(dm)))
        print("  previous block's demod CTAs: start", q(dp[:, 1] - t0))
        print("                               end  ", q(dp[:, 2] - t0))
        print("                               life ", q(dp[:, 2] - dp[:, 1]))
        print("  channel CTAs (mixers):       start", q(ch[:, 1] - t0))
        print("                               end  ", q(ch[:, 2] - t0))
        print("                               life ", q(ch[:, 2] - ch[:, 1]))
        print("  this block's demod CTAs:     start", q(dm[:, 1] - t0))
        print("                               end  ", q(dm[:, 2] - t0))


if __name__ == "__main__":
    main()
