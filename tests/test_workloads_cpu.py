"""bench.py measures BASELINE.json's configs: pin their geometry and the synthetic
input's properties, and check that each workload is one the reference itself accepts (integer rate
ratios, reference dspblock.cxx:119-130) by starting it on the reference blocks."""
import json
import os

import numpy as np
import pytest

import graphlib as G
from webradio_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_workloads_are_the_baseline_configs():
    W = synth.WORKLOADS
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert "64 NBFM receivers on one 2.4 MSPS tuner, 127-tap FIR, decim 50" in base[1]
    assert W["cfg2"]["desc"] in base[1]          # the default bench line names configs[1] literally
    c = W["cfg2"]
    assert (c["fs"], c["n_rx"], c["n_streams"], c["n1"], c["d1"], c["modes"]) == (2400000, 64, 1, 127, 50, "FM")
    c = W["cfg3"]
    assert (c["fs"], c["n_rx"], c["n_streams"], c["n1"], c["d1"], c["modes"]) == (2400000, 1024, 1024, 255, 50, "AM")
    c = W["cfg5"]
    assert (c["fs"], c["n_rx"], c["n_streams"], c["n1"], c["d1"], c["d2"], c["modes"]) == (10000000, 1024, 16, 127, 40, 5, "mixed")
    c = W["cfg1"]   # the shipped point: reference src/main.cxx:74-75, src/radio.cxx:78-81
    assert (c["fs"], c["frames"], c["n1"], c["d1"], c["n2"], c["d2"], c["pb1"], c["pb2"]) == \
        (2400000, 102400, 64, 10, 64, 5, 80000, 8000)
    for w in W.values():
        # a block is a whole number of audio frames, and the tuner block is the reference's where Fs is
        assert w["frames"] % (w["d1"] * w["d2"]) == 0
        assert w["n_rx"] % w["n_streams"] == 0


def test_mixed_modes_and_ifs_follow_the_survey():
    w = synth.WORKLOADS["cfg5"]
    assert list(synth.workload_modes(w)[:8]) == [0, 1, 2, 3, 0, 1, 2, 3]      # r mod 4 -> AM, FM, USB, LSB
    ifs = synth.workload_ifs(w)
    per = w["n_rx"] // w["n_streams"]
    assert ifs.shape == (w["n_rx"],)
    assert np.array_equal(ifs[:per], ifs[per:2 * per])                         # every tuner carries the same plan
    want = np.round((np.arange(per) - per / 2 + 0.5) * w["fs"] * 0.8 / per)
    assert np.array_equal(ifs[:per], want.astype(np.int32))
    assert np.abs(ifs).max() < w["fs"] // 2


def test_lattice_noise_is_the_rtlsdr_lattice_and_reproducible():
    a = synth.lattice_noise(4096, stream=3)
    b = synth.lattice_noise(4096, stream=3)
    assert np.array_equal(a, b)
    assert not np.array_equal(a, synth.lattice_noise(4096, stream=4))
    # continuing a stream block by block gives the same samples as one long block
    assert np.array_equal(np.concatenate([synth.lattice_noise(1000, stream=3), synth.lattice_noise(3096, stream=3, start=1000)]), a)
    codes = a * 128.0 + 128.0                                                  # (b - 128) / 128, rtlsdrtuner.cxx:106
    assert np.array_equal(codes, np.round(codes)) and codes.min() >= 0 and codes.max() <= 255
    assert len(np.unique(codes)) > 200


@pytest.mark.skipif(not G.have("ref"), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(synth.WORKLOADS))
def test_the_reference_accepts_every_workload(name):
    w = synth.WORKLOADS[name]
    with G.Graph("ref", w["fs"], 2 * w["d1"] * w["d2"] * 10) as g:
        g.add_receiver(if_hz=int(synth.workload_ifs(w)[0]), ch_passband=w["pb1"], ch_rate=0, ch_decim=w["d1"],
                       mode=int(synth.workload_modes(w)[0]), au_passband=w["pb2"], au_rate=0, au_decim=w["d2"])
        assert g.start()
        r = g.rates(0)
        assert r[1] == w["fs"] // w["d1"] and r[5] == w["fs"] // w["d1"] // w["d2"]
