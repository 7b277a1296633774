"""CPU-side checks of the C-ABI boundary: the library loads, exports every symbol that
include/webradio_b200.h declares, its host-side cold-path functions (NCO table, phase step, filter
design) match the oracle bit for bit, and -- with no GPU -- the product refuses to run rather
than fall back to a CPU path."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import assert_biteq, load_golden
from webradio_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "webradio_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    declared = header_symbols()
    assert len(declared) >= 40
    out = subprocess.check_output(["nm", "-D", "--defined-only", capi.LIB_PATH], text=True)
    exported = set(re.findall(r" T (wr_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    assert sorted(capi.SYMBOLS) == declared, "capi.SYMBOLS out of sync with the header"
    for s in declared:
        getattr(L, s)
    assert b"sm_100a" in L.wr_version()


def test_library_is_sm100a_cuda():
    out = subprocess.check_output(["cuobjdump", "-lelf", capi.LIB_PATH], text=True)
    assert "sm_100a" in out


def test_phase_step_matches_oracle(wro):
    rng = np.random.default_rng(1)
    for fs in (2400000, 2048000, 10000000, 1200000):
        for if_hz in list(rng.integers(-fs // 2, fs // 2, 50)) + [0, 1, -1, 100000, fs // 2, -fs // 2]:
            assert capi.phase_step(int(if_hz), fs) == wro.phase_step(int(if_hz), fs)
    assert capi.phase_step(100000, 2400000) == 89478485  # SURVEY.md 8c pin


def test_sintable_matches_oracle_and_golden(wro):
    t = capi.build_sintable()
    assert_biteq(t, wro.sintable(), "NCO table")
    g = load_golden("tables")
    assert_biteq(t[g["sintable_idx"]], g["sintable_val"], "NCO table vs reference samples")


@pytest.mark.parametrize("n", [64, 128, 256, 127, 255, 33])
def test_design_matches_oracle(wro, n):
    for fs, pb in [(2400000, 80000), (240000, 8000), (2400000, 12500), (2400000, 200000), (48000, 3000),
                   (2400000, 1200000), (2400000, 0)]:
        assert_biteq(capi.lowpass_design(n, pb, fs), wro.lowpass_design(n, pb, fs), f"design n={n} pb={pb} fs={fs}")


def test_design_matches_reference_golden():
    g = load_golden("tables")
    for k in g.files:
        if k.startswith("design_"):
            _, fs, pb = k.split("_")
            assert_biteq(capi.lowpass_design(64, int(pb), int(fs)), g[k], k)


def test_bad_arguments_are_reported():
    L = capi.lib()
    assert not L.wr_bank_create(0, 0, 1, 1024, 64, 10, 64, 5)
    assert b"bad geometry" in L.wr_last_error()
    assert not L.wr_spectrum_create(0, 500, 500, 1, 4096)
    assert b"power of two" in L.wr_last_error()
    assert L.wr_lowpass_design(64, 1000, 0, None) == -1


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.WrError, match="no CUDA device|no CPU fallback"):
        capi.Bank(1, 1, 1024, 64, 10, 64, 5)
    with pytest.raises(capi.WrError, match="no CUDA device|no CPU fallback"):
        capi.Stage()
    with pytest.raises(capi.WrError, match="no CUDA device|no CPU fallback"):
        capi.Spectrum(512)


# ---- the library's host cold path against the oracle on random arguments (CPU only) ----

from hypothesis import given, settings, strategies as st   # noqa: E402


@settings(max_examples=60, deadline=None, derandomize=True)
@given(if_hz=st.integers(-2**31, 2**31 - 1), fs=st.integers(1, 2**32 - 1))
def test_phase_step_random_arguments(wro, if_hz, fs):
    from webradio_b200 import capi
    assert capi.phase_step(if_hz, fs) == wro.phase_step(if_hz, fs)


@settings(max_examples=40, deadline=None, derandomize=True)
@given(log2n=st.integers(1, 10), passband=st.integers(0, 2**32 - 1), fs=st.integers(1, 2**32 - 1))
def test_lowpass_design_random_arguments(wro, log2n, passband, fs):
    """wr_lowpass_design (LowPass::init's window + LowPass::recalculate, reference
    lowpass.cxx:102-110,164-189) for every power-of-two length the reference could be compiled
    with, any pass-band and rate -- bit for bit the oracle's restatement."""
    import numpy as np
    from webradio_b200 import capi
    n = 1 << log2n
    got = capi.lowpass_design(n, passband, fs)
    want = wro.lowpass_design(n, passband, fs)
    maxbin = ((n * passband) & 0xFFFFFFFF) // fs // 2          # all-unsigned arithmetic, lowpass.cxx:167
    if maxbin <= n // 2:
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (n, passband, fs)
    else:
        # every bin set: the impulse is N at lag 0 and cancellation noise elsewhere -- the one place
        # where the result is whatever the inverse DFT's rounding leaves (FFTW3f in the reference,
        # unpinned: SURVEY.md 8c); the two restatements agree to far below a float's resolution
        assert np.max(np.abs(got.astype(np.float64) - want.astype(np.float64))) < 1e-15, (n, passband, fs)
        assert abs(float(got[n // 2]) - float(want[n // 2])) == 0.0
