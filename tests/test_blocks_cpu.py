"""Host logic of the DspBlock drop-in classes, no GPU needed: graph negotiation must match the
reference's (rates, decimation, integer-ratio rejection), the filter design they hold must be the
reference's, and without a CUDA device a block's process() fails (false) instead of computing on
the CPU."""
import numpy as np
import pytest

import graphlib as G
from helpers import assert_biteq

pytestmark = pytest.mark.skipif(not G.have("blocks"), reason="tests/harness/libwr_blocks_harness.so not built")

CASES = [
    dict(fs=2400000, ch_rate=240000, au_rate=48000),
    dict(fs=2048000, ch_rate=256000, au_rate=64000),
    dict(fs=2400000, ch_rate=0, ch_decim=50, au_rate=0, au_decim=1),
    dict(fs=10000000, ch_rate=250000, au_rate=50000),
]


@pytest.mark.parametrize("c", CASES)
def test_rate_negotiation_matches_reference(c):
    kw = dict(ch_rate=c["ch_rate"], ch_decim=c.get("ch_decim", 0), au_rate=c["au_rate"], au_decim=c.get("au_decim", 0))
    with G.Graph("blocks", c["fs"], 4000) as g:
        g.add_receiver(**kw)
        assert g.start()
        mine = g.rates(0)
    if G.have("ref"):
        with G.Graph("ref", c["fs"], 4000) as r:
            r.add_receiver(**kw)
            assert r.start()
            assert mine == r.rates(0)
    assert mine[0] == c["fs"] and mine[4] == 1 and mine[7] == 1


@pytest.mark.parametrize("fs,ch_rate", [(2048000, 240000), (2400000, 7), (1000, 1500)])
def test_non_integer_rates_are_rejected(fs, ch_rate):
    # reference dspblock.cxx:119-130: "Sample rates must be integer related"
    with G.Graph("blocks", fs, 4000) as g:
        g.add_receiver(ch_rate=ch_rate)
        assert not g.start()
    if G.have("ref"):
        with G.Graph("ref", fs, 4000) as r:
            r.add_receiver(ch_rate=ch_rate)
            assert not r.start()


def test_neither_rate_nor_decimation_fails():
    with G.Graph("blocks", 2400000, 4000) as g:
        g.add_receiver(ch_rate=0, ch_decim=0)
        assert not g.start()


@pytest.mark.skipif(not G.have("ref"), reason="needs oracle/_ref")
def test_designed_taps_match_reference():
    for pb in (80000, 200000, 12500):
        with G.Graph("blocks", 2400000, 4000) as g, G.Graph("ref", 2400000, 4000) as r:
            for x in (g, r):
                x.add_receiver(ch_passband=pb, au_passband=pb // 10)
                assert x.start()
            assert_biteq(g.get_taps(0, 0), r.get_taps(0, 0), f"channel taps pb={pb}")
            assert_biteq(g.get_taps(0, 1), r.get_taps(0, 1), f"audio taps pb={pb}")
            # live re-design (what a PUT /receivers/N does, receiverhandler.cxx:130-135)
            for x in (g, r):
                assert x.set_passband(0, 0, 150000) == 150000
            assert_biteq(g.get_taps(0, 0), r.get_taps(0, 0), "re-designed channel taps")


def test_mode_strings():
    with G.Graph("blocks", 2400000, 4000) as g:
        g.add_receiver(mode="AM")
        for i, m in enumerate(["AM", "FM", "USB", "LSB"]):
            assert g.set_mode(0, m) and g.get_mode(0) == i
        assert not g.set_mode(0, "CW")  # reference demodulator.cxx:47-56 returns false
        assert g.get_mode(0) == 3


def test_process_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with G.Graph("blocks", 2400000, 4000) as g:
        g.add_receiver()
        assert g.start()
        # no device -> process() returns false -> run() returns false; nothing is computed on the CPU
        assert not g.run(np.zeros(8000, np.float32))
