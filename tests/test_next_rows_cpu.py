"""The oracle's restatements of the steps either side of the path: byte-to-sample
conversion, waterfall palette index, encoder sample format.  These rows come from reference files
that cannot be built or run here (librtlsdr, LAME, a browser), so they are pinned against values
derived by hand from the reference expressions -- "parity unpinned" beyond that, see DESIGN.md."""
import numpy as np

from helpers import u8_to_iq


def test_rtlsdr_conversion_is_the_exact_lattice(wro):
    b = np.arange(256, dtype=np.uint8)
    got = wro.rtlsdr_convert(b)
    # ((float)b - 128.0) / 128.0 is exactly representable: compare with exact rationals
    want = np.array([(int(v) - 128) / 128 for v in b], dtype=np.float64)
    assert np.array_equal(got.astype(np.float64), want)
    assert np.array_equal(got, u8_to_iq(b))
    assert got[0] == -1.0 and got[128] == 0.0 and got[255] == np.float32(127 / 128)


def test_waterfall_index_known_values(wro):
    # (dB + 50) / 25 * 255, floor, clamp (waterfall.js:92-109); non-finite -> -10000 (waterfallhandler.cxx:62-68)
    db = np.float32([-50.0, -49.95, -37.5, -25.0, -25.1, 0.0, -60.0, -np.inf, np.inf, np.nan])
    want = [0, 0, 127, 255, 253, 255, 0, 0, 0, 0]
    assert list(wro.waterfall_index(db)) == want
    # exact boundaries in double: index k starts at -50 + 25 k / 255
    k = np.arange(1, 255)
    edge = -50.0 + 25.0 * k / 255.0
    f = edge.astype(np.float32)
    idx = wro.waterfall_index(f)
    ref = np.clip(np.floor((f.astype(np.float64) + 50.0) / 25.0 * 255.0), 0, 255).astype(np.uint8)
    assert np.array_equal(idx, ref)


def test_lame_scale(wro):
    x = np.float32([0.0, -0.0, 0.5, -1.0, 1.0, 3.0517578125e-05, 1e-40])
    got = wro.lame_scale(x)
    assert np.array_equal(got, (x.astype(np.float64) * 32768.0).astype(np.float32))
    assert got[2] == 16384.0 and got[3] == -32768.0 and got[5] == 1.0
