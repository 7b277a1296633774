"""Pins the kernels' atan2f (webradio_b200/csrc/wr_atan2f.h, a restatement of glibc's
e_atan2f.c / s_atanf.c) against the C library installed on this box -- the routine reference
src/dsp/demodulator.cxx:97 actually calls.  The host twin is the same source compiled for the
CPU; the GPU test checks the device build against the same libm."""
import numpy as np
import pytest

import atan2f_cases as ac
from webradio_b200 import capi


def _check(y, x, got, want, what):
    ok = ac.same(got, want)
    bad = np.nonzero(~ok)[0]
    assert bad.size == 0, (f"{what}: {bad.size} of {y.size} differ from libm, e.g. y={y[bad[:3]]} x={x[bad[:3]]} "
                           f"-> {got[bad[:3]]} vs libm {want[bad[:3]]}")


@pytest.mark.parametrize("gen,n", [(ac.random_bits, 4_000_000), (ac.discriminator_like, 4_000_000)])
def test_host_twin_equals_libm_on_random_arguments(wro, gen, n):
    y, x = gen(n, 7)
    _check(y, x, capi.atan2f_host(y, x), wro.libm_atan2f(y, x), gen.__name__)


def test_host_twin_equals_libm_at_branch_thresholds_and_specials(wro):
    for y, x in (ac.threshold_sweep(), ac.specials()):
        _check(y, x, capi.atan2f_host(y, x), wro.libm_atan2f(y, x), "thresholds/specials")


def test_reference_pins_of_the_discriminator(wro):
    # on-frequency carrier: atan2f(ii > 0, 0) = pi/2 -> +0.25 after /pi/2; first sample atan2f(0, 0) = 0
    assert capi.atan2f_host([0.25], [0.0])[0] == np.float32(np.pi / 2)
    assert capi.atan2f_host([0.0], [0.0])[0] == 0.0


@pytest.mark.gpu
def test_device_atan2f_equals_libm(wro):
    st = capi.Stage()
    try:
        for name, (y, x) in {"random bits": ac.random_bits(8_000_000, 11), "discriminator": ac.discriminator_like(8_000_000, 12),
                             "thresholds": ac.threshold_sweep(), "specials": ac.specials()}.items():
            _check(y, x, st.atan2f(y, x), wro.libm_atan2f(y, x), "device " + name)
    finally:
        st.close()
