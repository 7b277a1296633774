"""Randomised pin of the plain-C oracle (oracle/wr_oracle.c, wro.Rx) against the unmodified reference
blocks (oracle/_ref): random sample rates and decimations, IFs, modes, designed and injected taps
(64 / 127 / 255), retunes and mode changes between blocks -- every stage bit for bit.  The GPU
parity tests lean on the oracle; this is what the oracle leans on.  (570 further seeds ran clean when
this was written.)"""
import numpy as np
import pytest

import graphlib as G
from helpers import bits
from webradio_b200 import synth

pytestmark = pytest.mark.skipif(not G.have("ref"), reason="oracle/_ref not built")

MODES = ["AM", "FM", "USB", "LSB"]


def scenario(wro, seed):
    rng = np.random.default_rng(1000 + seed)
    fs = int(rng.choice([2400000, 2048000, 10000000, 1200000]))
    d1 = int(rng.choice([8, 10, 40, 50]))
    d2 = int(rng.choice([1, 2, 4, 5]))
    n1 = int(rng.choice([64, 64, 127, 255]))
    frames = d1 * d2 * int(rng.integers(3, 40))
    if_hz = int(rng.integers(-fs // 2, fs // 2))
    mode = MODES[int(rng.integers(0, 4))]
    pb1, pb2 = int(rng.choice([12500, 80000, 200000])), int(rng.choice([3000, 8000, 20000]))
    with G.Graph("ref", fs, frames) as r:
        r.add_receiver(if_hz=if_hz, ch_passband=pb1, ch_rate=0, ch_decim=d1, mode=mode,
                       au_passband=pb2, au_rate=0, au_decim=d2, capture=0xF)
        assert r.start()
        if n1 == 64:
            t1 = wro.lowpass_design(64, pb1, fs)
        else:
            t1 = rng.standard_normal(n1).astype(np.float32) / np.float32(n1)
            r.set_taps(0, 0, t1)
        t2 = wro.lowpass_design(64, pb2, fs // d1)
        assert np.array_equal(bits(r.get_taps(0, 0)), bits(t1)) and np.array_equal(bits(r.get_taps(0, 1)), bits(t2))
        o = wro.Rx(fs, if_hz, t1, d1, mode, t2, d2)
        for b in range(int(rng.integers(2, 6))):
            if rng.integers(0, 3) == 0:
                if_hz = int(rng.integers(-fs, fs))
                r.set_if(0, if_hz)
                o.set_if(if_hz)
            if rng.integers(0, 3) == 0:
                mode = MODES[int(rng.integers(0, 4))]
                assert r.set_mode(0, mode)
                o.set_mode(mode)
            if n1 == 64 and rng.integers(0, 4) == 0:
                pb1 = int(rng.choice([0, 12500, 80000, 300000]))
                r.set_passband(0, 0, pb1)
                o.set_taps(0, wro.lowpass_design(64, pb1, fs))
            iq = synth.lattice_noise(frames, stream=seed % 89, start=b * frames)
            assert r.run(iq)
            got = o.process(iq, stages=True)
            for name in ("mixed", "channel", "demod", "audio"):
                want = r.get(0, name)
                assert want.shape == got[name].shape and np.array_equal(bits(want), bits(got[name])), \
                    f"seed {seed} block {b} {name}: fs {fs} taps {n1}/{d1} 64/{d2} mode {mode} if {if_hz}"


@pytest.mark.parametrize("seed", range(30))
def test_random_chain(wro, seed):
    scenario(wro, seed)
