"""Shared body of the look-back carry-over check (wr_rx_get_lookback / wr_rx_set_lookback), run on
the CUDA library by tests/test_zz_lookback_gpu.py and, to test the test, on the CPU stand-in of
tests/harness/mock_capi.cxx by tests/test_blocks_mock_cpu.py.  `make_bank()` returns a fresh
one-receiver bank object with capi.Bank's methods."""
import numpy as np

from helpers import assert_biteq, assert_fm
from webradio_b200 import synth

FS, F = 2400000, 20480
N1, D1, N2, D2 = 64, 10, 64, 5
IF_HZ = -345678
RESET_DEMOD = 4


def carry_over(make_bank, wro, taps1, taps2, step, exact_fm=True):
    """Block A through one bank; phase and look-back sample read back, handed to a FRESH bank (FIR
    histories empty, as after LowPass::deinit) that runs block B -- against the oracle's stage
    functions doing the same thing (reference downconverter.cxx:91-114, lowpass.cxx:131-162,
    demodulator.cxx:77-115)."""
    fm = wro.MODES["FM"]
    table = wro.sintable()
    A = synth.lattice_noise(F, stream=3)
    B = synth.lattice_noise(F, stream=3, start=F)

    def configure(b):
        b.set_taps(0, 0, taps1)
        b.set_taps(0, 1, taps2)
        b.set_phase_step(0, step)
        b.set_mode(0, fm)

    def check_audio(got, want, what):
        if exact_fm:
            assert_biteq(got, want, what)
        else:
            assert_fm(got, want, what, audio=True)

    # oracle: block A
    mixedA, phA = wro.mix(table, 0, step, A)
    chanA = wro.Fir(2, taps1, D1).process(mixedA)
    prev = np.zeros(2, np.float32)
    demA = wro.demod(fm, prev, chanA)
    audioA = wro.Fir(1, taps2, D2).process(demA)
    assert_biteq(prev, chanA[-2:], "the look-back sample is the block's last channel-rate frame")

    b1 = make_bank()
    configure(b1)
    assert_biteq(b1.get_lookback(0), np.zeros(2, np.float32), "fresh bank")
    check_audio(b1.process(A)[0], audioA, "block A")
    ph = b1.get_phase(0)
    lb = b1.get_lookback(0)
    b1.close()
    assert ph == phA
    assert_biteq(lb, prev, "look-back sample after block A")

    # oracle: block B on fresh filters, phase and look-back sample carried
    mixedB, _ = wro.mix(table, phA, step, B)
    chanB = wro.Fir(2, taps1, D1).process(mixedB)
    want = wro.Fir(1, taps2, D2).process(wro.demod(fm, prev.copy(), chanB))
    want0 = wro.Fir(1, taps2, D2).process(wro.demod(fm, np.zeros(2, np.float32), chanB))
    assert not np.array_equal(want, want0), "the case must be sensitive to the look-back sample"

    b2 = make_bank()
    configure(b2)
    b2.set_phase(0, ph)
    b2.set_lookback(0, lb)
    check_audio(b2.process(B)[0], want, "block B, look-back sample carried")
    b2.close()

    # a reset after the set wins; a set after the reset wins
    b3 = make_bank()
    configure(b3)
    b3.set_phase(0, ph)
    b3.set_lookback(0, lb)
    b3.reset(0, RESET_DEMOD)
    check_audio(b3.process(B)[0], want0, "block B, look-back sample reset")
    b3.close()
    b4 = make_bank()
    configure(b4)
    b4.set_phase(0, ph)
    b4.reset(0, RESET_DEMOD)
    b4.set_lookback(0, lb)
    check_audio(b4.process(B)[0], want, "block B, set after reset")
    b4.close()


def hot_reattach(which, capture):
    """Receiver::setFrontEnd(NULL) and back on a live pipeline (reference radio.cxx:109-117), blocks of
    flavour `which` against the unmodified reference blocks: the chain stops (FIR histories released,
    lowpass.cxx:118-129) and later starts again with the rates it had; the NCO phase and the FM
    look-back sample live in the objects and survive (downconverter.cxx:46, demodulator.cxx:35)."""
    import graphlib as G
    fs, frames = 2400000, 20000
    modes = ["AM", "FM", "USB"]
    ifs = [50000, -250000, 400000]
    gs = []
    for w in (which, "ref"):
        g = G.Graph(w, fs, frames)
        for m, f in zip(modes, ifs):
            g.add_receiver(if_hz=f, mode=m, capture=capture)
        assert g.start()
        gs.append(g)
    g, r = gs
    try:
        for b in range(11):
            if b == 9:
                # the whole pipeline stopped and started again (a tuner restart): every chain is
                # re-initialised at once, with all receivers attached
                for x in (g, r):
                    assert x.attach(0)
                    assert x.restart()
            if b == 2:
                for x in (g, r):
                    assert x.detach(1)
            if b == 4:
                for x in (g, r):
                    assert x.attach(1)
            if b == 5:
                # connecting a consumer that is already connected, on a live pipeline: the reference
                # starts it a second time before it notices the duplicate (dspblock.cxx:59-67);
                # init() without deinit() keeps the FIR histories (lowpass.cxx:81-129)
                for x in (g, r):
                    assert x.attach(1)
            if b == 6:
                for x in (g, r):
                    assert x.detach(0) and x.detach(2)
            if b == 7:
                for x in (g, r):
                    assert x.attach(2)
            iq = synth.lattice_noise(frames, stream=5, start=b * frames)
            assert g.run(iq) and r.run(iq)
            for i, m in enumerate(modes):
                got, want = g.get(i, "audio"), r.get(i, "audio")
                if m == "FM":
                    assert_fm(got, want, f"rx{i} block {b}", audio=True)
                else:
                    assert_biteq(got, want, f"rx{i} block {b}")
    finally:
        g.close()
        r.close()
