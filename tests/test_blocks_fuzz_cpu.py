"""Randomised graph scenarios: the drop-in blocks on the CPU stand-in (tests/harness/mock_capi.cxx)
against the unmodified reference blocks, bit for bit at every tapped stage after every block, while a
seeded sequence of what the REST handlers and the glue can do is applied to both -- setIF,
setModeString, setPassband, detaching and re-attaching receivers on the live pipeline (also
attaching one that is attached), restarting the tuner.  Receivers have random geometries and random
taps between their stages, so fused banks, strict chains and both side by side are covered.
(900 further seeds ran clean when this was written.)"""
import numpy as np
import pytest

import graphlib as G
from helpers import bits
from webradio_b200 import synth

pytestmark = pytest.mark.skipif(not (G.have("mock") and G.have("ref")),
                                reason="libwr_blocks_harness_mock.so / oracle/_ref not built")

FS, F = 2400000, 4000
MODES = ["AM", "FM", "USB", "LSB"]
GEOMETRIES = [dict(ch_rate=240000, au_rate=48000), dict(ch_rate=0, ch_decim=50, au_rate=0, au_decim=1),
              dict(ch_rate=240000, au_rate=0, au_decim=5)]
STAGES = (("mixed", 1), ("channel", 2), ("demod", 4), ("audio", 8))


def scenario(seed, nblocks=12):
    rng = np.random.default_rng(seed)
    nrx = int(rng.integers(1, 6))
    caps = [int(rng.choice([0x8, 0x8, 0x8, 0xF, 0x9, 0xC, 0xA])) for _ in range(nrx)]
    plan = [(int(rng.integers(-1200000, 1200000)), MODES[int(rng.integers(0, 4))],
             GEOMETRIES[int(rng.integers(0, len(GEOMETRIES)))]) for _ in range(nrx)]
    with_spectrum = bool(rng.integers(0, 2))
    graphs = []
    for which in ("mock", "ref"):
        g = G.Graph(which, FS, F)
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
                assert np.array_equal(bits(a), bits(c)), f"seed {seed} block {b} spectrum: ops {log}"
    finally:
        g.close()
        r.close()


@pytest.mark.parametrize("seed", range(40))
def test_random_scenario(seed):
    scenario(seed)
