"""The DspBlock drop-in classes (webradio_b200/dsp, webradio_b200/io) against the unmodified
reference blocks (oracle/_ref) on the same graphs and inputs -- through the plugin surface only:
connect / start / run / setters, exactly what the reference's Radio glue uses.

  * fused path: chains with nothing attached between stages are batched into one receiver bank;
  * strict path: a tap attached after every stage forces one kernel per block per process();
  * the reference's own src/radio.cxx (unmodified) linked against the drop-in blocks.
"""
import ctypes as C
import os

import numpy as np
import pytest

import graphlib as G
from helpers import assert_biteq, assert_fm
from webradio_b200 import synth

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (G.have("blocks") and G.have("ref")), reason="harness libraries not built")]

FS, F = 2400000, 20000


def compare_audio(mode, got, want, what):
    if mode == "FM":
        assert_fm(got, want, what, audio=True)
    else:
        assert_biteq(got, want, what)


def pair(capture, modes, ifs, **kw):
    gs = []
    for which in ("blocks", "ref"):
        g = G.Graph(which, FS, F)
        for m, f in zip(modes, ifs):
            g.add_receiver(if_hz=f, mode=m, capture=capture, **kw)
        assert g.start()
        gs.append(g)
    return gs


def test_fused_chains_match_reference():
    modes = ["AM", "FM", "USB", "LSB", "AM", "USB"]
    ifs = [100000, -345678, 0, 612345, -900000, 7]
    g, r = pair(0x8, modes, ifs)
    try:
        for b in range(4):
            iq = synth.structured(F, FS, ifs[:4], [0, 1, 0, 0], start=b * F, fm_dev=50000.0)
            assert g.run(iq) and r.run(iq)
            for i, m in enumerate(modes):
                compare_audio(m, g.get(i, "audio"), r.get(i, "audio"), f"fused rx{i} {m} block {b}")
    finally:
        g.close(); r.close()


def test_strict_stage_blocks_match_reference():
    modes = ["AM", "FM", "USB", "LSB"]
    ifs = [100000, -345678, 5, 612345]
    g, r = pair(0xF, modes, ifs)
    try:
        for b in range(3):
            iq = synth.lattice_noise(F, stream=4, start=b * F)
            assert g.run(iq) and r.run(iq)
            for i, m in enumerate(modes):
                assert_biteq(g.get(i, "mixed"), r.get(i, "mixed"), f"strict rx{i} mixed b{b}")
                assert_biteq(g.get(i, "channel"), r.get(i, "channel"), f"strict rx{i} channel b{b}")
                if m == "FM":
                    assert_fm(g.get(i, "demod"), r.get(i, "demod"), f"strict rx{i} demod b{b}")
                else:
                    assert_biteq(g.get(i, "demod"), r.get(i, "demod"), f"strict rx{i} demod b{b}")
                compare_audio(m, g.get(i, "audio"), r.get(i, "audio"), f"strict rx{i} audio b{b}")
    finally:
        g.close(); r.close()


def test_setters_take_effect_at_block_boundaries():
    """setIF / setModeString / setPassband on running blocks (what the REST handlers do,
    reference src/web/receiverhandler.cxx:125-140), on fused chains."""
    modes = ["AM", "USB", "LSB"]
    ifs = [50000, -250000, 400000]
    g, r = pair(0x8, modes, ifs)
    try:
        for b in range(6):
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
                    assert x.set_mode(0, "USB")
                    x.set_if(2, 0)
            iq = synth.lattice_noise(F, stream=5, start=b * F)
            assert g.run(iq) and r.run(iq)
            for i in range(3):
                assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"rx{i} block {b}")
    finally:
        g.close(); r.close()


def test_injected_taps_127(wro):
    """cfg2-style geometry through the blocks: 127 injected taps, decimation 50."""
    t1 = synth.windowed_sinc(127, 12500 / FS)
    gs = []
    for which in ("blocks", "ref"):
        g = G.Graph(which, FS, F)
        for i in range(3):
            g.add_receiver(if_hz=100000 * i - 50000, mode="USB", ch_rate=0, ch_decim=50, au_rate=0, au_decim=1,
                           au_passband=3000, capture=0x8)
        if which == "blocks":
            pass
        assert g.start()
        gs.append(g)
    g, r = gs
    try:
        for i in range(3):
            g.set_taps(i, 0, t1)
            r.set_taps(i, 0, t1)
        for b in range(3):
            iq = synth.lattice_noise(F, stream=6, start=b * F)
            assert g.run(iq) and r.run(iq)
            for i in range(3):
                assert_biteq(g.get(i, "audio"), r.get(i, "audio"), f"127-tap rx{i} block {b}")
    finally:
        g.close(); r.close()


def test_spectrum_sink_block():
    n = 512
    g = G.Graph("blocks", FS, F)
    r = G.Graph("ref", FS, F)
    try:
        for x in (g, r):
            x.add_spectrum(n)
            assert x.start()
        for b in range(3):
            iq = synth.structured(F, FS, [300000, -700000], [0, 1], start=b * F, noise_db=-40.0)
            assert g.run(iq) and r.run(iq)
            got, want = g.spectrum(n).astype(np.float64), r.spectrum(n).astype(np.float64)
            peak = 10 ** (want.max() / 20)
            assert np.max(np.abs(10 ** (got / 20) - 10 ** (want / 20))) <= 1e-5 * peak
    finally:
        g.close(); r.close()


DROPIN = os.path.join(G.ROOT, "tests", "harness", "libwr_radio_dropin.so")


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="libwr_radio_dropin.so not built (needs the reference tree)")
def test_reference_radio_glue_runs_on_the_dropin_blocks():
    """The reference's unmodified src/radio.cxx (FrontEnd / Receiver / Radio::run) linked against the
    CUDA-backed blocks, driven as src/main.cxx drives it, vs the reference blocks."""
    L = C.CDLL(DROPIN)
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

    frames = 102400  # the shipped block: 204800 floats (reference src/main.cxx:75)
    rig = L.wrr_create(FS, frames, 512)
    modes = ["AM", "USB", "LSB", "AM"]
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
            want = ref.spectrum(512).astype(np.float64)
            assert np.max(np.abs(10 ** (db.astype(np.float64) / 20) - 10 ** (want / 20))) <= 1e-5 * 10 ** (want.max() / 20)
    finally:
        L.wrr_destroy(rig)
        ref.close()


@pytest.mark.skipif(not os.path.exists(DROPIN), reason="libwr_radio_dropin.so not built (needs the reference tree)")
def test_sixteen_front_ends_are_sharded_by_tuner_over_the_gpus():
    """north_star: receivers shard across the GPUs of one box by assignment, behind the UNCHANGED glue.
    Sixteen FrontEnds (reference src/radio.cxx:120-156), two receivers each, one Radio::run() per block:
    every front-end's bank, spectrum sink and tuner-block upload live on the device the front-end was
    dealt (round-robin over the visible GPUs), audio bit-exact against sixteen reference graphs, spectra
    within the transform tolerance.  On a single-GPU box everything lands on device 0 and the same
    checks run."""
    import torch
    fp = C.POINTER(C.c_float)
    L = C.CDLL(DROPIN, mode=C.RTLD_LOCAL)
    L.wrr_create.restype = C.c_void_p
    L.wrr_create.argtypes = [C.c_uint, C.c_uint, C.c_uint]
    L.wrr_add_receiver.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    L.wrr_start.argtypes = [C.c_void_p]
    L.wrr_run_many.argtypes = [C.POINTER(C.c_void_p), C.c_uint, C.POINTER(C.c_void_p)]
    L.wrr_audio.restype = C.c_long
    L.wrr_audio.argtypes = [C.c_void_p, C.c_int, fp, C.c_long]
    L.wrr_spectrum.argtypes = [C.c_void_p, fp]
    L.wrr_device.argtypes = [C.c_void_p]
    L.wrr_destroy.argtypes = [C.c_void_p]
    frames, nfe = 20480, 16
    modes = ["FM", "AM"]
    rigs, refs = [], []
    try:
        for t in range(nfe):
            rig = L.wrr_create(FS, frames, 512)
            ref = G.Graph("ref", FS, frames)
            ref.add_spectrum(512)
            for i, m in enumerate(modes):
                f = 30000 * t - 200000 + 7777 * i
                assert L.wrr_add_receiver(rig, f, m.encode()) >= 0
                ref.add_receiver(if_hz=f, mode=m, capture=0x8)
            assert L.wrr_start(rig) == 0 and ref.start()
            rigs.append(rig)
            refs.append(ref)
        handles = (C.c_void_p * nfe)(*rigs)
        for b in range(3):
            blocks = [synth.structured(frames, FS, [30000 * t - 200000, 30000 * t - 200000 + 7777], [1, 0],
                                       start=b * frames, stream=t, fm_dev=50000.0) for t in range(nfe)]
            ptrs = (C.c_void_p * nfe)(*[x.ctypes.data for x in blocks])
            assert L.wrr_run_many(handles, nfe, ptrs) == 0
            for t in range(nfe):
                assert refs[t].run(blocks[t])
                for i, m in enumerate(modes):
                    n = L.wrr_audio(rigs[t], i, None, 0)
                    got = np.empty(n, np.float32)
                    L.wrr_audio(rigs[t], i, got.ctypes.data_as(fp), n)
                    want = refs[t].get(i, "audio")
                    if m == "FM":
                        assert_fm(got, want, f"front-end {t} rx{i} block {b}", audio=True)
                    else:
                        assert_biteq(got, want, f"front-end {t} rx{i} block {b}")
                db = np.empty(512, np.float32)
                assert L.wrr_spectrum(rigs[t], db.ctypes.data_as(fp)) == 512
                want = refs[t].spectrum(512).astype(np.float64)
                ma, mw = 10 ** (db.astype(np.float64) / 20), 10 ** (want / 20)
                assert np.max(np.abs(ma - mw)) <= 1e-5 * mw.max(), f"front-end {t} spectrum block {b}"
        ndev = min(torch.cuda.device_count(), int(os.environ.get("WEBRADIO_B200_DEVICES", "64")))
        if "WEBRADIO_B200_DEVICE" in os.environ:
            ndev = 1
        devs = [L.wrr_device(r) for r in rigs]
        assert len(set(devs)) == min(ndev, nfe), devs
        # round-robin: consecutive front-ends sit on consecutive devices
        assert all((devs[i + 1] - devs[i]) % ndev == 1 % ndev for i in range(nfe - 1)), devs
    finally:
        for r in rigs:
            L.wrr_destroy(r)
        for r in refs:
            r.close()
