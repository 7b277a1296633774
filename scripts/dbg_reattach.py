import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/harness'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import graphlib as G
from helpers import bits
from webradio_b200 import synth
fs, frames = 2400000, 20000
modes = ["AM", "FM", "USB"]; ifs = [50000, -250000, 400000]
which = sys.argv[1] if len(sys.argv) > 1 else "blocks"
ops = {2: [("detach", 1)], 4: [("attach", 1)], 5: [("attach", 1)], 6: [("detach", 0), ("detach", 2)], 7: [("attach", 2)]}
if len(sys.argv) > 2:
    ops = eval(sys.argv[2])
gs = []
for w in (which, "ref"):
    g = G.Graph(w, fs, frames)
    for m, f in zip(modes, ifs):
        g.add_receiver(if_hz=f, mode=m, capture=0x8)
    assert g.start(); gs.append(g)
g, r = gs
for b in range(9):
    for op, i in ops.get(b, []):
        for x in (g, r):
            getattr(x, op)(i)
    iq = synth.lattice_noise(frames, stream=5, start=b * frames)
    assert g.run(iq) and r.run(iq)
    res = []
    for i in range(3):
        a, c = g.get(i, "audio"), r.get(i, "audio")
        if a.shape != c.shape: res.append("shape"); continue
        nbad = int((bits(a) != bits(c)).sum())
        first = int(np.nonzero(bits(a) != bits(c))[0][0]) if nbad else -1
        res.append((nbad, first, float(np.max(np.abs(a - c))) if nbad else 0.0))
    print("block", b, res)
