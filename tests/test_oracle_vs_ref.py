"""Pins the plain-C oracle port (oracle/wr_oracle.c) against the REAL reference:
oracle/_ref/libwr_ref.so = mikestir/webradio's unmodified src/dsp + src/io/spectrumsink.cxx
compiled in place (oracle/Makefile).  Bit-for-bit on every stage.  CPU only.
"""
import numpy as np
import pytest

import graphlib as G
from webradio_b200 import synth

pytestmark = pytest.mark.skipif(not G.have("ref"), reason="oracle/_ref/libwr_ref.so not built")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_biteq(a, b, what=""):
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = np.nonzero(bits(a) != bits(b))[0]
    assert bad.size == 0, f"{what}: {bad.size} of {a.size} differ, first at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}"


def test_sintable(wro):
    assert_biteq(G.ref_sintable(), wro.sintable(), "sin table")


@pytest.mark.parametrize("fs,if_hz", [(2400000, 100000), (2400000, -612345), (2048000, 1), (10000000, 4999999),
                                      (2400000, 0), (2400000, -1200000)])
def test_phase_step(wro, fs, if_hz):
    # the reference only exposes phaseStep through its effect; check the closed form instead
    step = wro.phase_step(if_hz, fs)
    assert step == int((if_hz * (1 << 31)) / fs) or step == -((-if_hz * (1 << 31)) // fs)
    if (fs, if_hz) == (2400000, 100000):
        assert step == 89478485  # SURVEY.md 8c pin


@pytest.mark.parametrize("fs,rate,passband", [(2400000, 240000, 80000), (240000, 48000, 8000),
                                              (2400000, 240000, 200000), (2400000, 48000, 12500),
                                              (2048000, 256000, 100000), (2400000, 240000, 1200000)])
def test_design_matches_reference(wro, fs, rate, passband):
    with G.Graph("ref", fs, 2 * (fs // rate) * 8) as g:
        g.add_receiver(ch_passband=passband, ch_rate=rate, au_passband=passband // 4, au_rate=0, au_decim=1)
        assert g.start()
        assert_biteq(g.get_taps(0, 0), wro.lowpass_design(64, passband, fs), "channel taps")
        assert_biteq(g.get_taps(0, 1), wro.lowpass_design(64, passband // 4, rate), "audio taps")


def test_design_dc_gain_pin(wro):
    # default filters have maxbin = 1 => coeff = Hamming/64, DC gain ~0.533
    c = wro.lowpass_design(64, 80000, 2400000)
    assert abs(float(c.astype(np.float64).sum()) - 0.5328125) < 1e-6


def run_both(wro, fs, frames, if_hz, mode, n1, d1, n2, d2, blocks, make_iq, events=None, pb1=80000, pb2=8000):
    """Runs reference graph and the port side by side; returns nothing, asserts bit equality."""
    with G.Graph("ref", fs, frames) as g:
        g.add_receiver(if_hz=if_hz, ch_passband=pb1, ch_rate=0, ch_decim=d1, mode=mode,
                       au_passband=pb2, au_rate=0, au_decim=d2)
        assert g.start()
        if n1 == 64:
            t1 = g.get_taps(0, 0)
        else:
            t1 = wro.lowpass_design(n1, pb1, fs)
            g.set_taps(0, 0, t1)
        if n2 == 64:
            t2 = g.get_taps(0, 1)
        else:
            t2 = wro.lowpass_design(n2, pb2, fs // d1)
            g.set_taps(0, 1, t2)
        rx = wro.Rx(fs, if_hz, t1, d1, mode, t2, d2)
        for b in range(blocks):
            for ev in (events or {}).get(b, []):
                if ev[0] == "if":
                    g.set_if(0, ev[1]); rx.set_if(ev[1])
                elif ev[0] == "mode":
                    assert g.set_mode(0, ev[1]); rx.set_mode(ev[1])
            iq = make_iq(b)
            assert g.run(iq)
            got = rx.process(iq, stages=True)
            for st in ("mixed", "channel", "demod", "audio"):
                assert_biteq(g.get(0, st), got[st], f"block {b} stage {st}")


@pytest.mark.parametrize("mode", ["AM", "FM", "USB", "LSB"])
def test_chain_default_point(wro, mode):
    fs, F = 2400000, 20480
    ifs, modes = [100000], [synth.MODE_NAMES.index(mode)]
    run_both(wro, fs, F, 100000, mode, 64, 10, 64, 5, 3,
             lambda b: synth.structured(F, fs, ifs, modes, start=b * F, fm_dev=50000.0))


@pytest.mark.parametrize("n1,d1,mode", [(127, 50, "FM"), (255, 50, "AM"), (127, 40, "USB"), (33, 7, "LSB"), (2, 1, "AM")])
def test_chain_injected_taps(wro, n1, d1, mode):
    fs = 2400000
    F = d1 * 64 * 2
    run_both(wro, fs, F, -333333, mode, n1, d1, 64, 1, 4,
             lambda b: synth.lattice_noise(F, stream=3, start=b * F), pb1=12500, pb2=3000)


def test_chain_retune_and_mode_change(wro):
    fs, F = 2400000, 10000
    ev = {1: [("if", -250000)], 2: [("mode", "FM")], 3: [("mode", "LSB"), ("if", 7)], 4: [("mode", "AM")]}
    run_both(wro, fs, F, 50000, "USB", 64, 10, 64, 5, 6,
             lambda b: synth.lattice_noise(F, stream=9, start=b * F), events=ev)


def test_block_shorter_than_history(wro):
    # F < ntaps-1: history spans several blocks
    fs = 2400000
    run_both(wro, fs, 40, 123456, "AM", 64, 10, 64, 2, 9,
             lambda b: synth.lattice_noise(40, stream=1, start=b * 40))


def test_fm_pins(wro):
    # first FM sample is atan2f(0,0)=0; an on-frequency carrier gives +0.25
    fs, F = 2400000, 20480
    t1 = wro.lowpass_design(64, 80000, fs)
    t2 = wro.lowpass_design(64, 8000, 240000)
    rx = wro.Rx(fs, 0, t1, 10, "FM", t2, 5)
    iq = np.zeros(2 * F, np.float32)
    iq[0::2] = 0.5
    out = rx.process(iq, stages=True)
    assert out["demod"][0] == 0.0
    assert abs(out["demod"][-1] - 0.25) < 1e-6


@pytest.mark.parametrize("n", [512, 8192])
def test_spectrum(wro, n):
    fs = 2400000
    F = 3 * n + n // 2  # leaves a partial frame to carry over
    with G.Graph("ref", fs, F) as g:
        g.add_spectrum(n)
        assert g.start()
        sp = wro.Spectrum(n)
        for b in range(3):
            iq = synth.structured(F, fs, [300000, -700000], [0, 1], start=b * F, noise_db=-40.0)
            assert g.run(iq)
            rows = sp.process(iq)
            ref = g.spectrum(n)
            assert_biteq(ref, sp.get(), f"spectrum block {b}")
            assert rows.shape[0] in (3, 4)
            assert_biteq(rows[-1], ref, "last row")


@pytest.mark.skipif(not G.have("ref_O0"), reason="oracle/_ref/libwr_ref_O0.so not built")
def test_pinned_flags_equal_the_stock_build():
    """the oracle's pinned flags (-O2 -ffp-contract=off, no -march) must give what the
    reference's stock build (-O0: configure.ac:6 pre-sets CXXFLAGS, so autoconf adds no -O2) gives,
    bit for bit -- every stage of every mode, the filter design, the spectrum."""
    fs, frames = 2400000, 20000
    gs = []
    for which in ("ref", "ref_O0"):
        g = G.Graph(which, fs, frames)
        for i, m in enumerate(["AM", "FM", "USB", "LSB"]):
            g.add_receiver(if_hz=[100000, -345678, 5, 612345][i], mode=m, capture=0xF)
        g.add_spectrum(512)
        assert g.start()
        gs.append(g)
    a, b = gs
    try:
        for i in range(4):
            assert_biteq(a.get_taps(i, 0), b.get_taps(i, 0), "channel design")
            assert_biteq(a.get_taps(i, 1), b.get_taps(i, 1), "audio design")
        for blk in range(3):
            iq = synth.structured(frames, fs, [100000, -345678], [0, 1], start=blk * frames, fm_dev=50000.0)
            assert a.run(iq) and b.run(iq)
            for i in range(4):
                for stage in ("mixed", "channel", "demod", "audio"):
                    assert_biteq(a.get(i, stage), b.get(i, stage), f"rx{i} {stage} block {blk}")
            assert_biteq(a.spectrum(512), b.spectrum(512), f"spectrum block {blk}")
    finally:
        a.close()
        b.close()
