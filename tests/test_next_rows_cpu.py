"""The oracle's restatements of the steps either side of the path: byte-to-sample
conversion, waterfall palette index, encoder sample format.  These rows come from reference files
that cannot be built or run here (librtlsdr, LAME, a browser), so they are pinned against values
derived from the reference expressions in exact rational arithmetic with hand-written IEEE rounding
(scripts/make_palette_fixture.py -> tests/golden/palette_edges.npz, lame_scale.npz): every palette
edge +-3 ULP, clamps, non-finite bins; random bit patterns, subnormals and overflow for the scale."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

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


def test_waterfall_index_against_the_rational_fixture(wro):
    """html/waterfall.js:95-101 + waterfallhandler.cxx:62-68 in exact arithmetic: 257 edges x 7 neighbours."""
    d = np.load(os.path.join(GOLDEN, "palette_edges.npz"))
    assert len(d["db"]) >= 257 * 7 and len(set(d["index"].tolist())) == 256
    assert np.array_equal(wro.waterfall_index(d["db"]), d["index"])


def test_lame_scale_against_the_rational_fixture(wro):
    """src/web/mp3encoder.cxx:66-68 in exact arithmetic, bit for bit (signed zeros, subnormals, overflow)."""
    d = np.load(os.path.join(GOLDEN, "lame_scale.npz"))
    got = wro.lame_scale(d["x"])
    assert np.array_equal(got.view(np.uint32), d["y"].view(np.uint32))
    assert np.isinf(d["y"]).sum() > 100 and (np.abs(d["x"]) < 1.2e-38).sum() > 10
