"""Pins the plain-C oracle port against fixtures generated FROM THE REFERENCE ITSELF
(tests/golden/*.npz, made by scripts/make_golden.py out of oracle/_ref).  CPU only; this is
what keeps the oracle honest where /root/reference is not mounted.
"""
import hashlib

import numpy as np
import pytest

from helpers import CHAIN_CASES, assert_biteq, golden_events, load_golden, u8_to_iq


def test_tables(wro):
    g = load_golden("tables")
    tbl = wro.sintable()
    assert hashlib.sha256(tbl.tobytes()).digest() == bytes(g["sintable_sha256"])
    assert_biteq(tbl[g["sintable_idx"]], g["sintable_val"], "sin table samples")
    # SURVEY.md 8a/a3 sanity pins
    assert tbl[0] == 0.0 and tbl[16384] == 1.0 and tbl[32768] == np.float32(-8.742278e-08)
    for k in g.files:
        if k.startswith("design_"):
            _, fs, pb = k.split("_")
            assert_biteq(wro.lowpass_design(64, int(pb), int(fs)), g[k], k)


@pytest.mark.parametrize("name", CHAIN_CASES)
def test_chain(wro, name):
    g = load_golden(name)
    rx = wro.Rx(int(g["fs"]), int(g["if_hz"]), g["taps1"], int(g["d1"]), int(g["mode"]), g["taps2"], int(g["d2"]))
    ev = golden_events(g)
    nb = g["iq_u8"].shape[0]
    for b in range(nb):
        for kind, val in ev.get(b, []):
            rx.set_if(val) if kind == "if" else rx.set_mode(val)
        out = rx.process(u8_to_iq(g["iq_u8"][b]), stages=True)
        assert_biteq(out["channel"], g["channel"][b], f"{name} channel block {b}")
        assert_biteq(out["demod"], g["demod"][b], f"{name} demod block {b}")
        assert_biteq(out["audio"], g["audio"][b], f"{name} audio block {b}")
    assert_biteq(out["mixed"], g["mixed_last"], f"{name} mixed (last block)")


@pytest.mark.parametrize("n", [512, 8192])
def test_spectrum(wro, n):
    g = load_golden(f"spectrum_{n}")
    sp = wro.Spectrum(n)
    for b in range(g["iq"].shape[0]):
        sp.process(g["iq"][b], rows=False)
        assert_biteq(sp.get(), g["db"][b], f"spectrum {n} block {b}")


def test_spectrum_vs_float64_dft(wro):
    """The FFT boundary is 'parity unpinned' (FFTW3f is absent): check the stand-in against numpy's
    float64 FFT at the north_star tolerance instead."""
    n = 8192
    g = load_golden(f"spectrum_{n}")
    sp = wro.Spectrum(n)
    x = g["iq"][0]
    sp.process(x[: 2 * n], rows=False)
    w = wro.spectrum_window(n)
    xc = (x[0:2 * n:2] * w).astype(np.float64) + 1j * (x[1:2 * n:2] * w).astype(np.float64)
    ref = np.fft.fft(xc)
    got = sp.bins().astype(np.complex128)
    assert np.max(np.abs(np.abs(got) - np.abs(ref))) <= 1e-6 * np.max(np.abs(ref))


def test_overlap_extension(wro):
    """hop < n (cfg4's 50% overlap) must equal restarting the reference transform at each hop."""
    n, hop = 512, 256
    rng = np.random.default_rng(5)
    x = rng.uniform(-1, 1, 2 * (n + 5 * hop)).astype(np.float32)
    rows = wro.Spectrum(n, hop).process(x)
    assert rows.shape == (6, n)
    for m in range(6):
        one = wro.Spectrum(n)
        one.process(x[2 * m * hop: 2 * (m * hop + n)], rows=False)
        assert_biteq(rows[m], one.get(), f"overlap row {m}")
