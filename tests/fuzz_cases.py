"""Shared body of the randomised graph scenarios (see tests/test_blocks_fuzz_cpu.py): the drop-in blocks
of flavour `which` ("mock": CPU stand-in, "blocks": the CUDA library) against the reference blocks."""
import numpy as np

import graphlib as G
from helpers import bits
from webradio_b200 import synth

FS, F = 2400000, 4000
MODES = ["AM", "FM", "USB", "LSB"]
GEOMETRIES = [dict(ch_rate=240000, au_rate=48000), dict(ch_rate=0, ch_decim=50, au_rate=0, au_decim=1),
              dict(ch_rate=240000, au_rate=0, au_decim=5)]
STAGES = (("mixed", 1), ("channel", 2), ("demod", 4), ("audio", 8))


def scenario(seed, nblocks=12, which="mock"):
    rng = np.random.default_rng(seed)
    nrx = int(rng.integers(1, 6))
    caps = [int(rng.choice([0x8, 0x8, 0x8, 0xF, 0x9, 0xC, 0xA])) for _ in range(nrx)]
    plan = [(int(rng.integers(-1200000, 1200000)), MODES[int(rng.integers(0, 4))],
             GEOMETRIES[int(rng.integers(0, len(GEOMETRIES)))]) for _ in range(nrx)]
    with_spectrum = bool(rng.integers(0, 2))
    graphs = []
    for flavour in (which, "ref"):
        g = G.Graph(flavour, FS, F)
        for (f, m, geo), cap in zip(plan, caps):
            g.add_receiver(if_hz=f, mode=m, capture=cap, **geo)
        if with_spectrum:
            g.add_spectrum(512)
        assert g.start()
        graphs.append(g)
    g, r = graphs
    attached = [True] * nrx
    log = []
    try:
        for b in range(nblocks):
            for _ in range(int(rng.integers(0, 4))):
                op, i = int(rng.integers(0, 7)), int(rng.integers(0, nrx))
                if op == 0:
                    hz = int(rng.integers(-3000000, 3000000))
                    log.append(("if", b, i, hz))
                    for x in graphs:
                        x.set_if(i, hz)
                elif op == 1:
                    m = MODES[int(rng.integers(0, 4))]
                    log.append(("mode", b, i, m))
                    for x in graphs:
                        x.set_mode(i, m)
                elif op == 2:
                    w, hz = int(rng.integers(0, 2)), int(rng.choice([0, 100, 3000, 8000, 12500, 80000, 200000, 1000000]))
                    log.append(("passband", b, i, w, hz))
                    for x in graphs:
                        x.set_passband(i, w, hz)
                elif op == 3 and attached[i]:
                    log.append(("detach", b, i))
                    attached[i] = False
                    for x in graphs:
                        x.detach(i)
                elif op == 4:
                    log.append(("attach", b, i))
                    attached[i] = True
                    for x in graphs:
                        x.attach(i)
                elif op == 5 and rng.integers(0, 4) == 0:
                    log.append(("restart", b))
                    for x in graphs:
                        x.restart()
            iq = synth.lattice_noise(F, stream=seed % 97, start=b * F)
            assert g.run(iq) == r.run(iq), (seed, b, log)
            for i in range(nrx):
                for name, bit in STAGES:
                    if caps[i] & bit:
                        a, c = g.get(i, name), r.get(i, name)
                        assert a.shape == c.shape and np.array_equal(bits(a), bits(c)), \
                            f"seed {seed} block {b} rx{i} {name}: caps {caps} plan {plan} ops {log}"
            if with_spectrum and b >= 1:
                a, c = g.spectrum(512), r.spectrum(512)
                if which == "mock":
                    assert np.array_equal(bits(a), bits(c)), f"seed {seed} block {b} spectrum: ops {log}"
                else:
                    # float32 transform on the device against the reference's (stand-in) FFT: the
                    # north_star tolerance, 1e-5 of the frame's peak magnitude
                    ma, mc = 10 ** (a.astype(np.float64) / 20), 10 ** (c.astype(np.float64) / 20)
                    assert np.max(np.abs(ma - mc)) <= 1e-5 * mc.max(), f"seed {seed} block {b} spectrum: ops {log}"
    finally:
        g.close()
        r.close()
