"""HOST logic of the DspBlock drop-in classes (webradio_b200/dsp, webradio_b200/io) on the CPU:
the same C++ blocks and graph driver as tests/test_blocks_gpu.py, linked against a CPU stand-in for
the device entry points (tests/harness/mock_capi.cxx -- test infrastructure, arithmetic by the
oracle) instead of libwebradio_b200.so, compared with the unmodified reference blocks
(oracle/_ref) through the plugin surface only.

What this pins without a GPU: receiver chains are batched into ONE bank per tuner block, triggered
by the first receiver visited; each chain gets its own slice; setIF / setModeString / setPassband
reach the bank at the next block boundary; a consumer attached between stages turns the chain
into strict per-block stage calls; NCO phase, FIR histories and the FM look-back sample carry
across blocks; SpectrumSink.  What it does NOT check is the CUDA arithmetic (tests -m gpu do).
"""
import ctypes as C
import os

import numpy as np
import pytest

import graphlib as G
from helpers import assert_biteq
from webradio_b200 import synth

pytestmark = pytest.mark.skipif(not (G.have("mock") and G.have("ref")),
                                reason="libwr_blocks_harness_mock.so / oracle/_ref not built")

FS, F = 2400000, 20000


def counters():
    lib = G.load("mock")
    for f in ("wr_mock_bank_process_calls", "wr_mock_stage_calls", "wr_mock_banks_created", "wr_mock_uploads",
              "wr_mock_upload_frames"):
        getattr(lib, f).restype = C.c_ulonglong
    lib.wr_mock_set_devices.argtypes = [C.c_int]
    lib.wr_mock_bank_device.argtypes = [C.c_uint]
    return lib


def pair(capture, modes, ifs, **kw):
    gs = []
    for which in ("mock", "ref"):
        g = G.Graph(which, FS, F)
        for m, f in zip(modes, ifs):
            g.add_receiver(if_hz=f, mode=m, capture=capture, **kw)
        assert g.start()
        gs.append(g)
    return gs


def test_fused_chains_one_bank_call_per_block():
    modes = ["AM", "FM", "USB", "LSB", "AM", "USB"]
    ifs = [100000, -345678, 0, 612345, -900000, 7]
    lib = counters()
    lib.wr_mock_reset_counters()
    g, r = pair(0x8, modes, ifs)
    try:
        for b in range(4):
            iq = synth.structured(F, FS, ifs[:4], [0, 1, 0, 0], start=b * F, fm_dev=50000.0)
            assert g.run(iq) and r.run(iq)
            for i, m in enumerate(modes):
                assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"fused rx{i} {m} block {b}")
        # six receivers, four blocks: ONE bank, ONE batched call per tuner block, no stage calls
        assert lib.wr_mock_banks_created() == 1
        assert lib.wr_mock_bank_process_calls() == 4
        assert lib.wr_mock_stage_calls() == 0
    finally:
        g.close(); r.close()


def test_strict_stage_blocks_when_every_stage_is_tapped():
    modes = ["AM", "FM", "USB", "LSB"]
    ifs = [100000, -345678, 5, 612345]
    lib = counters()
    lib.wr_mock_reset_counters()
    g, r = pair(0xF, modes, ifs)
    try:
        for b in range(3):
            iq = synth.lattice_noise(F, stream=4, start=b * F)
            assert g.run(iq) and r.run(iq)
            for i in range(4):
                for stage in ("mixed", "channel", "demod", "audio"):
                    assert_biteq(g.get(i, stage), r.get(i, stage), f"strict rx{i} {stage} b{b}")
        # a consumer sits between the stages: no bank, four stage calls per receiver per block
        assert lib.wr_mock_bank_process_calls() == 0
        assert lib.wr_mock_stage_calls() == 4 * 4 * 3
    finally:
        g.close(); r.close()


def test_mixed_graph_fused_and_strict_side_by_side():
    """Receivers 0 and 2 are plain chains (fused), receiver 1 has a tap after the mixer (strict)."""
    lib = counters()
    lib.wr_mock_reset_counters()
    gs = []
    for which in ("mock", "ref"):
        g = G.Graph(which, FS, F)
        g.add_receiver(if_hz=50000, mode="AM", capture=0x8)
        g.add_receiver(if_hz=-70000, mode="USB", capture=0x9)
        g.add_receiver(if_hz=123456, mode="FM", capture=0x8)
        assert g.start()
        gs.append(g)
    g, r = gs
    try:
        for b in range(3):
            iq = synth.lattice_noise(F, stream=9, start=b * F)
            assert g.run(iq) and r.run(iq)
            for i in range(3):
                assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"rx{i} block {b}")
            assert_biteq(g.get(1, "mixed"), r.get(1, "mixed"), f"rx1 mixed block {b}")
        assert lib.wr_mock_bank_process_calls() == 3
        assert lib.wr_mock_stage_calls() == 4 * 3
    finally:
        g.close(); r.close()


def test_setters_take_effect_at_block_boundaries():
    """setIF / setModeString / setPassband on running blocks (what the REST handlers do, reference
    src/web/receiverhandler.cxx:125-140), on fused chains -- FM included."""
    modes = ["AM", "USB", "FM"]
    ifs = [50000, -250000, 400000]
    g, r = pair(0x8, modes, ifs)
    try:
        for b in range(7):
            if b == 2:
                for x in (g, r):
                    x.set_if(0, -123456)
                    assert x.set_mode(1, "LSB")
            if b == 3:
                for x in (g, r):
                    assert x.set_passband(2, 0, 200000) == 200000
                    assert x.set_passband(2, 1, 20000) == 20000
            if b == 4:
                for x in (g, r):
                    assert x.set_mode(0, "FM")
                    x.set_if(2, 0)
            if b == 5:
                for x in (g, r):
                    assert x.set_mode(2, "AM")
                    assert not x.set_mode(1, "CW")       # rejected, mode unchanged (demodulator.cxx:47-56)
            iq = synth.lattice_noise(F, stream=5, start=b * F)
            assert g.run(iq) and r.run(iq)
            for i in range(3):
                assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"rx{i} block {b}")
    finally:
        g.close(); r.close()


@pytest.mark.parametrize("n1,d1,d2", [(127, 50, 1), (255, 50, 1), (64, 8, 4)])
def test_injected_taps_and_other_geometries(n1, d1, d2):
    """cfg2 / cfg3 / 2.048 MSPS-style geometries through the blocks: injected channel taps."""
    fs = 2048000 if d1 == 8 else FS
    frames = 20480 if d1 == 8 else F
    t1 = synth.windowed_sinc(n1, 12500 / fs)
    gs = []
    for which in ("mock", "ref"):
        g = G.Graph(which, fs, frames)
        for i in range(3):
            g.add_receiver(if_hz=100000 * i - 50000, mode=["USB", "FM", "AM"][i], ch_rate=0, ch_decim=d1,
                           au_rate=0, au_decim=d2, au_passband=3000, capture=0x8)
        assert g.start()
        gs.append(g)
    g, r = gs
    try:
        for i in range(3):
            g.set_taps(i, 0, t1)
            r.set_taps(i, 0, t1)
        for b in range(3):
            iq = synth.lattice_noise(frames, stream=6, start=b * frames)
            assert g.run(iq) and r.run(iq)
            for i in range(3):
                assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"{n1}-tap rx{i} block {b}")
    finally:
        g.close(); r.close()


def test_two_geometries_make_two_banks():
    """Receivers with different filter geometry cannot share a bank: one bank per geometry, each
    run once per block."""
    lib = counters()
    lib.wr_mock_reset_counters()
    gs = []
    for which in ("mock", "ref"):
        g = G.Graph(which, FS, F)
        g.add_receiver(if_hz=1000, mode="AM", capture=0x8)                                     # 240 k / 48 k
        g.add_receiver(if_hz=-2000, mode="USB", capture=0x8)
        g.add_receiver(if_hz=3000, mode="LSB", ch_rate=0, ch_decim=50, au_rate=0, au_decim=1, capture=0x8)
        assert g.start()
        gs.append(g)
    g, r = gs
    try:
        for b in range(2):
            iq = synth.lattice_noise(F, stream=2, start=b * F)
            assert g.run(iq) and r.run(iq)
            for i in range(3):
                assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"rx{i} block {b}")
        assert lib.wr_mock_banks_created() == 2
        assert lib.wr_mock_bank_process_calls() == 4
    finally:
        g.close(); r.close()


def test_spectrum_sink_block():
    n = 512
    g = G.Graph("mock", FS, F)
    r = G.Graph("ref", FS, F)
    try:
        for x in (g, r):
            x.add_spectrum(n)
            x.add_receiver(if_hz=0, mode="AM", capture=0x8)
            assert x.start()
        for b in range(3):
            iq = synth.structured(F, FS, [300000, -700000], [0, 1], start=b * F, noise_db=-40.0)
            assert g.run(iq) and r.run(iq)
            assert_biteq(g.spectrum(n), r.spectrum(n), f"spectrum block {b}")
            assert_biteq(g.get(0, "audio"), r.get(0, "audio"), f"audio next to the spectrum sink, block {b}")
    finally:
        g.close(); r.close()


def test_one_upload_per_tuner_block_shared_by_sink_and_banks():
    """The tuner block goes to the device ONCE per DspBlock::run of the tuner, whoever asks first: the
    SpectrumSink (connected first, as FrontEnd does, radio.cxx:126-128), two banks of different
    geometry and nobody else (strict chains take the host buffer)."""
    lib = counters()
    lib.wr_mock_reset_counters()
    g, r = G.Graph("mock", FS, F), G.Graph("ref", FS, F)
    try:
        for x in (g, r):
            x.add_spectrum(512)
            for i in range(5):
                x.add_receiver(if_hz=10000 * i, mode="FM", capture=0x8)
            x.add_receiver(if_hz=-5, mode="AM", capture=0x8, ch_rate=0, ch_decim=50, au_rate=0, au_decim=1)
            assert x.start()
        for b in range(4):
            iq = synth.lattice_noise(F, stream=9, start=b * F)
            assert g.run(iq) and r.run(iq)
            for i in range(6):
                assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"rx{i} block {b}")
            assert_biteq(g.spectrum(512), r.spectrum(512), f"spectrum block {b}")
        assert lib.wr_mock_uploads() == 4 and lib.wr_mock_upload_frames() == 4 * F
        assert lib.wr_mock_banks_created() == 2 and lib.wr_mock_bank_process_calls() == 8
    finally:
        g.close(); r.close()


def test_front_ends_are_dealt_over_the_devices():
    """Shard by tuner behind the unchanged glue: every producer (front-end) gets a device round-robin at
    first sight and its bank, its sink and its upload live there; results do not depend
    on the placement."""
    lib = counters()
    lib.wr_mock_reset_counters()
    lib.wr_mock_clear_bank_devices()
    lib.wr_mock_set_devices(4)
    gs = []
    try:
        for t in range(6):
            pairg = [G.Graph(w, FS, F) for w in ("mock", "ref")]
            for x in pairg:
                x.add_spectrum(512)
                for i in range(3):
                    x.add_receiver(if_hz=1000 * (i + 7 * t), mode=["AM", "FM", "USB"][i], capture=0x8)
                assert x.start()
            gs.append(pairg)
        for b in range(3):
            for t, (g, r) in enumerate(gs):
                iq = synth.lattice_noise(F, stream=40 + t, start=b * F)
                assert g.run(iq) and r.run(iq)
                for i in range(3):
                    assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"front-end {t} rx{i} block {b}")
                assert_biteq(g.spectrum(512), r.spectrum(512), f"front-end {t} spectrum block {b}")
        assert lib.wr_mock_banks_created() == 6
        assert [lib.wr_mock_bank_device(i) for i in range(6)] == [0, 1, 2, 3, 0, 1]
    finally:
        for g, r in gs:
            g.close(); r.close()
        lib.wr_mock_set_devices(1)


def test_profile_counters_count_frames():
    """DSPBLOCK_PROFILE bookkeeping (dspblock.cxx:186-204): every block of a fused chain still counts
    the frames it was handed, exactly as the reference's blocks do."""
    with G.Graph("mock", FS, F) as g, G.Graph("ref", FS, F) as r:
        for x in (g, r):
            x.add_receiver(capture=0x8)
            assert x.start()
            for b in range(3):
                assert x.run(synth.lattice_noise(F, stream=1, start=b * F))
        _, frames = g.profile(0)
        _, want = r.profile(0)
        assert frames == want == [3 * F, 3 * F, 3 * F // 10, 3 * F // 10]


@pytest.mark.parametrize("capture", [0x8, 0xF], ids=["fused", "strict"])
def test_hot_detach_and_reattach(capture):
    """tests/lookback_cases.py::hot_reattach on the stand-in (assert_fm is bit-exact on a glibc box)."""
    import lookback_cases as LC
    lib = counters()
    lib.wr_mock_reset_counters()
    LC.hot_reattach("mock", capture)
    if capture == 0x8:
        # fused throughout, and ONE batched call per block however the chains came and went: the
        # bank of the front-end is rebuilt (state carried) whenever its membership changes
        assert lib.wr_mock_stage_calls() == 0
        assert lib.wr_mock_bank_process_calls() == 11
        # start, rx1 leaves, rx1 back, rx0 + rx2 leave, rx2 back, restart with all three
        assert lib.wr_mock_banks_created() == 6


def test_lookback_travels_between_strict_and_fused():
    """A chain that starts strict-free (fused), is stopped and comes back keeps its discriminator
    history; checked directly on the stand-in's bank API as the blocks use it."""
    lib = G.load("mock")
    fp = C.POINTER(C.c_float)
    lib.wr_bank_create.restype = C.c_void_p
    lib.wr_bank_create.argtypes = [C.c_int] + [C.c_uint] * 7
    lib.wr_rx_set_lookback.argtypes = [C.c_void_p, C.c_uint, fp]
    lib.wr_rx_get_lookback.argtypes = [C.c_void_p, C.c_uint, fp]
    lib.wr_bank_destroy.argtypes = [C.c_void_p]
    b = lib.wr_bank_create(0, 1, 2, 1000, 64, 10, 64, 5)
    v = (C.c_float * 2)(0.25, -0.5)
    assert lib.wr_rx_set_lookback(b, 1, v) == 0
    w = (C.c_float * 2)()
    assert lib.wr_rx_get_lookback(b, 1, w) == 0 and (w[0], w[1]) == (0.25, -0.5)
    assert lib.wr_rx_get_lookback(b, 0, w) == 0 and (w[0], w[1]) == (0.0, 0.0)
    assert lib.wr_rx_get_lookback(b, 2, w) != 0
    lib.wr_bank_destroy(b)


def test_two_front_ends_and_a_large_bank():
    """Two pipelines alive at once (two front-ends): banks are keyed by their producer, so neither
    sees the other's receivers; 64 receivers on one tuner still make ONE bank call per block."""
    lib = counters()
    lib.wr_mock_reset_counters()
    ga = [G.Graph(w, FS, F) for w in ("mock", "ref")]
    gb = [G.Graph(w, FS, F) for w in ("mock", "ref")]
    try:
        for x in ga:
            for i in range(64):
                x.add_receiver(if_hz=(i - 32) * 30000 + 17, mode=["AM", "FM", "USB", "LSB"][i % 4], capture=0x8)
            assert x.start()
        for x in gb:
            for i in range(3):
                x.add_receiver(if_hz=1000 * i, mode="FM", capture=0x8)
            assert x.start()
        for b in range(3):
            ia = synth.lattice_noise(F, stream=1, start=b * F)
            ib = synth.lattice_noise(F, stream=2, start=b * F)
            for x in ga:
                assert x.run(ia)
            for x in gb:
                assert x.run(ib)
            for i in range(64):
                assert_biteq(ga[0].get(i, "audio"), ga[1].get(i, "audio"), f"front-end A rx{i} block {b}")
            for i in range(3):
                assert_biteq(gb[0].get(i, "audio"), gb[1].get(i, "audio"), f"front-end B rx{i} block {b}")
        assert lib.wr_mock_banks_created() == 2 and lib.wr_mock_bank_process_calls() == 6
    finally:
        for x in ga + gb:
            x.close()


@pytest.mark.parametrize("n,first", [(512, 0), (8192, 0), (32768, 1), (65536, 3)])
def test_spectrum_sink_sizes(n, first):
    """FFT frames shorter and LONGER than a tuner block (a frame then completes every few blocks;
    before the first one the reference reads uninitialised memory, so the comparison starts at the
    block that completes it)."""
    g = G.Graph("mock", FS, F)
    r = G.Graph("ref", FS, F)
    try:
        for x in (g, r):
            x.add_spectrum(n)
            assert x.start()
        for b in range(first + 4):
            iq = synth.structured(F, FS, [300000, -700000], [0, 1], start=b * F, noise_db=-40.0)
            assert g.run(iq) and r.run(iq)
            if b >= first:
                assert_biteq(g.spectrum(n), r.spectrum(n), f"{n}-point spectrum after block {b}")
    finally:
        g.close(); r.close()


def test_extreme_setter_values():
    """IFs at and beyond Nyquist, the int range's ends, pass-bands of 0, above the sample rate and one
    that collapses the design to all-zero taps: step arithmetic (downconverter.cxx:65,80) and design
    (lowpass.cxx:164-189) wrap and truncate exactly as the reference's."""
    modes = ["AM", "FM", "USB", "LSB"]
    ifs = [1199999, -1200000, 2400000, -7777777]
    g, r = pair(0x8, modes, ifs)
    try:
        for b in range(8):
            if b == 2:
                for x in (g, r):
                    assert x.set_passband(0, 0, 12500) == 12500 and x.set_passband(1, 1, 100) == 100
                    assert x.set_passband(2, 0, 1200000) == 1200000 and x.set_passband(3, 1, 120000) == 120000
            if b == 4:
                for x in (g, r):
                    assert x.set_if(0, 2**31 - 1) == 2**31 - 1 and x.set_if(1, -2**31) == -2**31
                    x.set_if(2, 123456789)
            if b == 6:
                for x in (g, r):
                    assert x.set_passband(0, 0, 0) == 0
                    x.set_passband(1, 0, 4000000000)
            iq = synth.lattice_noise(F, stream=5, start=b * F)
            assert g.run(iq) and r.run(iq)
            for i in range(4):
                assert_biteq(g.get_taps(i, 0), r.get_taps(i, 0), f"rx{i} channel taps block {b}")
                assert_biteq(g.get_taps(i, 1), r.get_taps(i, 1), f"rx{i} audio taps block {b}")
                assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"rx{i} block {b}")
    finally:
        g.close(); r.close()


def test_a_device_error_fails_the_block_and_nothing_else():
    """SURVEY.md 8b, errors: a failing library call must surface as process() == false -> run() ==
    false (dspblock.cxx:192-195), never as an exception or a crash, and the pipeline must be usable
    for the next block."""
    lib = counters()
    lib.wr_mock_fail_next.argtypes = [C.c_int, C.c_int]
    for capture in (0x8, 0xF):
        with G.Graph("mock", FS, F) as g:
            for i in range(3):
                g.add_receiver(if_hz=1000 * i, mode="FM", capture=capture)
            assert g.start()
            assert g.run(synth.lattice_noise(F, stream=1))
            lib.wr_mock_fail_next(1, 1)
            assert not g.run(synth.lattice_noise(F, stream=1, start=F))
            lib.wr_mock_fail_next(0, 0)
            assert g.run(synth.lattice_noise(F, stream=1, start=2 * F))
            assert g.get(2, "audio").size == F // 50


DROPIN_MOCK = os.path.join(G.ROOT, "tests", "harness", "libwr_radio_dropin_mock.so")


@pytest.mark.skipif(not os.path.exists(DROPIN_MOCK), reason="libwr_radio_dropin_mock.so not built (needs the reference tree)")
def test_reference_radio_glue_drives_the_dropin_blocks():
    """The reference's UNMODIFIED src/radio.cxx (FrontEnd / Receiver / Radio::run) compiled against
    the drop-in headers and driven as src/main.cxx drives it -- the host half of the drop-in claim,
    on the CPU stand-in; tests/test_blocks_gpu.py runs the same rig on the CUDA library."""
    L = C.CDLL(DROPIN_MOCK)
    fp = C.POINTER(C.c_float)
    L.wrr_create.restype = C.c_void_p
    L.wrr_create.argtypes = [C.c_uint, C.c_uint, C.c_uint]
    L.wrr_add_receiver.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    L.wrr_start.argtypes = [C.c_void_p]
    L.wrr_run.argtypes = [C.c_void_p, fp]
    L.wrr_retune.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.c_uint]
    L.wrr_audio.restype = C.c_long
    L.wrr_audio.argtypes = [C.c_void_p, C.c_int, fp, C.c_long]
    L.wrr_spectrum.argtypes = [C.c_void_p, fp]
    L.wrr_destroy.argtypes = [C.c_void_p]
    L.wr_mock_bank_process_calls.restype = C.c_ulonglong
    L.wr_mock_stage_calls.restype = C.c_ulonglong

    frames = 102400  # the shipped block: 204800 floats (reference src/main.cxx:75)
    rig = L.wrr_create(FS, frames, 512)
    modes = ["AM", "USB", "FM", "AM"]
    ifs = [0, 100000, -200000, 555555]
    ref = G.Graph("ref", FS, frames)
    ref.add_spectrum(512)
    for m, f in zip(modes, ifs):
        assert L.wrr_add_receiver(rig, f, m.encode()) >= 0
        ref.add_receiver(if_hz=f, mode=m, capture=0x8)  # same defaults as Receiver() (radio.cxx:78-82)
    assert L.wrr_start(rig) == 0 and ref.start()
    try:
        for b in range(3):
            if b == 2:
                assert L.wrr_retune(rig, 1, -77777, b"LSB", 100000) == 0
                ref.set_if(1, -77777); ref.set_mode(1, "LSB"); ref.set_passband(1, 0, 100000)
            iq = synth.lattice_noise(frames, stream=11, start=b * frames)
            assert L.wrr_run(rig, iq.ctypes.data_as(fp)) == 0
            assert ref.run(iq)
            for i in range(4):
                n = L.wrr_audio(rig, i, None, 0)
                assert n == frames // 50
                got = np.empty(n, np.float32)
                L.wrr_audio(rig, i, got.ctypes.data_as(fp), n)
                assert_biteq(got, ref.get(i, "audio"), f"radio.cxx drop-in rx{i} block {b}")
            db = np.empty(512, np.float32)
            assert L.wrr_spectrum(rig, db.ctypes.data_as(fp)) == 512
            assert_biteq(db, ref.spectrum(512), f"spectrum block {b}")
        # Radio::run visits four receivers per block; the first visit runs the one bank
        assert L.wr_mock_bank_process_calls() == 3 and L.wr_mock_stage_calls() == 0
    finally:
        L.wrr_destroy(rig)
        ref.close()


class MockBank:
    """capi.Bank's methods (the subset tests/lookback_cases.py uses) over the CPU stand-in."""

    def __init__(self, n_rx, max_frames, n1, d1, n2, d2):
        L = self.L = G.load("mock")
        fp, vp, u = C.POINTER(C.c_float), C.c_void_p, C.c_uint
        L.wr_bank_create.restype = vp
        L.wr_bank_create.argtypes = [C.c_int] + [u] * 7
        L.wr_bank_destroy.argtypes = [vp]
        L.wr_rx_set_taps.argtypes = [vp, u, C.c_int, fp, u]
        L.wr_rx_set_phase_step.argtypes = [vp, u, C.c_int32]
        L.wr_rx_set_mode.argtypes = [vp, u, C.c_int]
        L.wr_rx_set_phase.argtypes = [vp, u, C.c_uint32]
        L.wr_rx_get_phase.argtypes = [vp, u, C.POINTER(C.c_uint32)]
        L.wr_rx_set_lookback.argtypes = [vp, u, fp]
        L.wr_rx_get_lookback.argtypes = [vp, u, fp]
        L.wr_rx_reset.argtypes = [vp, u, u]
        L.wr_bank_process.argtypes = [vp, fp, u, fp, C.c_size_t]
        self.R, self.d1, self.d2 = n_rx, d1, d2
        self.h = L.wr_bank_create(0, 1, n_rx, max_frames, n1, d1, n2, d2)
        assert self.h

    def close(self):
        if self.h:
            self.L.wr_bank_destroy(self.h)
            self.h = None

    def set_taps(self, rx, stage, coeff):
        c = np.ascontiguousarray(coeff, np.float32)
        assert self.L.wr_rx_set_taps(self.h, rx, stage, c.ctypes.data_as(C.POINTER(C.c_float)), c.size) == 0

    def set_phase_step(self, rx, step):
        assert self.L.wr_rx_set_phase_step(self.h, rx, step) == 0

    def set_mode(self, rx, mode):
        assert self.L.wr_rx_set_mode(self.h, rx, mode) == 0

    def set_phase(self, rx, phase):
        assert self.L.wr_rx_set_phase(self.h, rx, phase) == 0

    def get_phase(self, rx):
        v = C.c_uint32(0)
        assert self.L.wr_rx_get_phase(self.h, rx, C.byref(v)) == 0
        return v.value

    def set_lookback(self, rx, iq):
        v = (C.c_float * 2)(float(iq[0]), float(iq[1]))
        assert self.L.wr_rx_set_lookback(self.h, rx, v) == 0

    def get_lookback(self, rx):
        v = (C.c_float * 2)()
        assert self.L.wr_rx_get_lookback(self.h, rx, v) == 0
        return np.array([v[0], v[1]], np.float32)

    def reset(self, rx, flags):
        assert self.L.wr_rx_reset(self.h, rx, flags) == 0

    def process(self, iq):
        a = np.ascontiguousarray(iq, np.float32)
        n = a.size // 2
        m2 = n // self.d1 // self.d2
        out = np.zeros((self.R, m2), np.float32)
        fp = C.POINTER(C.c_float)
        assert self.L.wr_bank_process(self.h, a.ctypes.data_as(fp), n, out.ctypes.data_as(fp), m2) == 0
        return out


def test_lookback_carry_over_case_on_the_stand_in(wro):
    """The body tests/test_zz_lookback_gpu.py runs on the CUDA library, dry-run here."""
    import lookback_cases as LC
    t1 = wro.lowpass_design(LC.N1, 80000, LC.FS)
    t2 = wro.lowpass_design(LC.N2, 8000, LC.FS // LC.D1)
    LC.carry_over(lambda: MockBank(1, LC.F, LC.N1, LC.D1, LC.N2, LC.D2), wro, t1, t2,
                  wro.phase_step(LC.IF_HZ, LC.FS))


def test_host_logic_under_address_sanitizer():
    """make asan-check: tests/harness/host_scenario.cxx (detach / attach / restart / teardown in both
    orders on two live pipelines) built with -fsanitize=address,undefined over the stand-in."""
    import shutil
    import subprocess
    if not shutil.which("make") or not shutil.which("g++"):
        pytest.skip("no toolchain")
    p = subprocess.run(["make", "-s", "asan-check"], cwd=G.ROOT, capture_output=True, text=True, timeout=600)
    out = p.stdout + p.stderr
    if p.returncode != 0 and not any(k in out for k in ("Sanitizer", "runtime error", "start failed", "run failed")):
        pytest.skip("sanitizer build not possible here: " + out[-300:])
    assert p.returncode == 0, out[-3000:]
    assert "scenario done" in out
    assert "AddressSanitizer" not in out and "runtime error" not in out, out[-3000:]


def test_setters_from_other_threads_under_thread_sanitizer():
    """make tsan-check: tests/harness/host_threads.cxx -- the DSP thread runs blocks while two threads
    call setIF / setModeString / setPassband / getSpectrum and the getters with no locking of their own,
    as libmicrohttpd's connection threads do in the reference (src/web/receiverhandler.cxx:113-140)."""
    import shutil
    import subprocess
    if not shutil.which("make") or not shutil.which("g++"):
        pytest.skip("no toolchain")
    p = subprocess.run(["make", "-s", "tsan-check"], cwd=G.ROOT, capture_output=True, text=True, timeout=900)
    out = p.stdout + p.stderr
    if "unexpected memory mapping" in out or "ThreadSanitizer: unsupported" in out:
        pytest.skip("ThreadSanitizer cannot run in this sandbox")
    if p.returncode != 0 and not any(k in out for k in ("Sanitizer", "start failed", "run failed")):
        pytest.skip("sanitizer build not possible here: " + out[-300:])
    assert p.returncode == 0, out[-3000:]
    assert "threads done" in out
    assert "ThreadSanitizer: data race" not in out, out[-3000:]
