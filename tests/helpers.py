"""Shared helpers for the parity tests."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CHAIN_CASES = ["chain_fm_default", "chain_am_default", "chain_usb_127_d50", "chain_am_255_d50",
               "chain_events", "chain_short_blocks"]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_biteq(a, b, what=""):
    a = np.asarray(a, np.float32).ravel()
    b = np.asarray(b, np.float32).ravel()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = np.nonzero(bits(a) != bits(b))[0]
    assert bad.size == 0, (f"{what}: {bad.size} of {a.size} differ, first at {bad[:5]}: "
                           f"{a[bad[:5]]} vs {b[bad[:5]]}")


def ulp_distance(a, b):
    """Distance in units of float32 ULP (monotone integer mapping of the bit patterns)."""
    def key(x):
        i = np.ascontiguousarray(x, dtype=np.float32).view(np.int32).astype(np.int64)
        return np.where(i < 0, np.int64(-2147483648) - i, i)
    return np.abs(key(a) - key(b))


def u8_to_iq(u8):
    """RTL-SDR sample lattice: reference src/io/rtlsdrtuner.cxx:106."""
    return ((u8.astype(np.float32) - np.float32(128.0)) / np.float32(128.0)).astype(np.float32)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_events(g):
    ev = {}
    for b, kind, val in g["events"]:
        ev.setdefault(int(b), []).append(("if" if kind == 0 else "mode", int(val)))
    return ev


_FM_EXACT = None


def fm_exact():
    """True when the C library's atan2f on this box is the glibc routine the kernels restate
    (webradio_b200/csrc/wr_atan2f.h; pinned in depth by tests/test_atan2f.py).  The FM parity checks
    are then bit-exact; on a box with another libm they fall back to the ULP tolerance."""
    global _FM_EXACT
    if _FM_EXACT is None:
        from oracle import wro
        from webradio_b200 import capi
        rng = np.random.default_rng(5)
        y = rng.uniform(-1, 1, 200000).astype(np.float32)
        x = rng.uniform(-1, 1, 200000).astype(np.float32)
        _FM_EXACT = bool(np.array_equal(bits(capi.atan2f_host(y, x)), bits(wro.libm_atan2f(y, x))))
    return _FM_EXACT


FM_MAX_ULP = 2          # fallback tolerance on the demodulated sample (foreign libm only)
FM_AUDIO_TOL = 3e-7     # fallback tolerance on FM audio (foreign libm only)


def assert_fm(got, want, what="", audio=False):
    """FM samples against the reference chain: bit-exact on a glibc box, see fm_exact()."""
    if fm_exact():
        assert_biteq(got, want, what)
        return
    got = np.asarray(got, np.float32).ravel()
    want = np.asarray(want, np.float32).ravel()
    assert got.shape == want.shape, what
    if got.size == 0:
        return
    if audio:
        assert np.max(np.abs(got - want)) <= FM_AUDIO_TOL, what
    else:
        d = ulp_distance(got, want)
        assert d.max() <= FM_MAX_ULP, f"{what}: FM demod off by {d.max()} ULP"
