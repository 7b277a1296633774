"""The NCO table's exact compression (webradio_b200/csrc/wr_lo.h, wr_lo3.h) as the HOST verifies it
before the shared-memory kernels may use it (CPU only): the library's own table and the oracle's
must be the same 65536 floats, both compressions must reproduce every entry bit for bit, and a
table they cannot represent must be refused (the kernels then stand aside) rather than approximated."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_biteq
from webradio_b200 import capi

fp = C.POINTER(C.c_float)


def check(fn, table):
    t = None if table is None else np.ascontiguousarray(table, np.float32)
    return getattr(capi.lib(), fn)(None if t is None else t.ctypes.data_as(fp))


def test_library_table_is_the_reference_table(wro):
    t = capi.build_sintable()
    assert_biteq(t, wro.sintable(), "wr_build_sintable vs the oracle's restatement of downconverter.cxx:49-51")
    assert t[0] == 0.0 and t[16384] == 1.0 and t[32768] == np.float32(-8.742278e-08)


@pytest.mark.parametrize("fn", ["wr_lo_compress_check", "wr_lo3_compress_check"])
def test_default_table_compresses_exactly(fn):
    assert check(fn, None) == 0
    assert check(fn, capi.build_sintable()) == 0


@pytest.mark.parametrize("fn", ["wr_lo_compress_check", "wr_lo3_compress_check"])
def test_a_one_ulp_change_still_compresses_exactly(fn):
    """The corrections are exact 16-bit differences of the bit patterns: another libm's table --
    entries off by an ULP here and there -- is represented just as exactly."""
    t = capi.build_sintable()
    rng = np.random.default_rng(7)
    idx = rng.choice(65536, 4000, replace=False)
    bits = t.view(np.int32).copy()
    bits[idx] += rng.choice([-1, 1], idx.size).astype(np.int32)
    assert check(fn, bits.view(np.float32)) == 0


@pytest.mark.parametrize("fn", ["wr_lo_compress_check", "wr_lo3_compress_check"])
def test_a_table_that_does_not_fit_is_refused(fn):
    t = capi.build_sintable()
    t[12345] = 0.75          # far from sin(): the 16-bit correction cannot reach it
    assert check(fn, t) == -1
    assert check(fn, np.zeros(65536, np.float32)) in (0, -1)      # never crashes on a degenerate table
    assert check(fn, np.full(65536, np.nan, np.float32)) == -1
