#!/usr/bin/env python
"""Golden vectors for the waterfall palette index, derived WITHOUT the oracle and
without numpy arithmetic: the reference computes the index in the browser,

    html/waterfall.js:95-101   val = (series[bin] + 50.0) / 25.0;  val = val * 255.0;
                               val = Math.floor(val);  clamp to [0, 255]
    src/web/waterfallhandler.cxx:62-68   non-finite bins are sent as -10000.0

in IEEE-754 binary64, round to nearest even -- which is what a Python float is.  Every one of the
three operations is re-derived here in exact rational arithmetic (fractions.Fraction) and rounded
to binary64 by hand (round_to_double), and the script asserts that this equals Python's own float
result, so the fixture does not lean on any one implementation.  No JavaScript engine exists in
the build image; this is as close to the browser as the parity chain gets.

Inputs: for each of the 257 palette edges -50 + 25 k / 255 the nearest float32 and its neighbours
up to +-3 ULP (the dB values are float32, spectrumsink.cxx:138), the clamps, and the non-finite
values.  Writes tests/golden/palette_edges.npz {db: float32[n], index: uint8[n]} and, derived the same way for
the encoder's sample format (src/web/mp3encoder.cxx:66-68: a binary64 product by 32768
stored to float), tests/golden/lame_scale.npz {x: float32[n], y: float32[n]}."""
import math
import os
import struct
from fractions import Fraction

import numpy as np


def round_to_double(q):
    """Correctly rounded (nearest, ties to even) binary64 of the rational q, by hand."""
    if q == 0:
        return 0.0
    sign = -1 if q < 0 else 1
    q = abs(q)
    e = q.numerator.bit_length() - q.denominator.bit_length()
    if Fraction(2) ** e > q:
        e -= 1
    assert Fraction(2) ** e <= q < Fraction(2) ** (e + 1) and -1022 <= e <= 1023
    scaled = q / Fraction(2) ** (e - 52)            # in [2^52, 2^53)
    m = scaled.numerator // scaled.denominator
    rem = scaled - m
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (m & 1)):
        m += 1
    return sign * math.ldexp(m, e - 52)


def f32_neighbours(x, n):
    b = struct.unpack("<i", struct.pack("<f", x))[0]
    out = []
    for d in range(-n, n + 1):
        bb = b + d if b >= 0 else b - d             # (negative floats: larger bit pattern = more negative)
        out.append(struct.unpack("<f", struct.pack("<i", bb))[0])
    return out


def js_index(x32):
    """x32: a float32 value held in a Python float."""
    if math.isnan(x32) or math.isinf(x32):
        x = -10000.0                                # waterfallhandler.cxx:65-68
    else:
        x = x32
    v1 = round_to_double(Fraction(x) + 50)
    assert v1 == x + 50.0
    v2 = round_to_double(Fraction(v1) / 25)
    assert v2 == v1 / 25.0
    v3 = round_to_double(Fraction(v2) * 255)
    assert v3 == v2 * 255.0
    k = math.floor(v3)
    return 0 if k < 0 else 255 if k > 255 else k


def round_to_float(q):
    """Correctly rounded binary32 of the rational q (nearest even; subnormals; overflow to inf), by hand."""
    if q == 0:
        return 0.0
    sign = -1.0 if q < 0 else 1.0
    q = abs(q)
    e = q.numerator.bit_length() - q.denominator.bit_length()
    if Fraction(2) ** e > q:
        e -= 1
    e = max(e, -126)                                # subnormals share the exponent of the smallest normal
    scaled = q / Fraction(2) ** (e - 23)
    m = scaled.numerator // scaled.denominator
    rem = scaled - m
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (m & 1)):
        m += 1
    v = math.ldexp(m, e - 23)
    return sign * (float("inf") if v > 3.4028234663852886e38 else v)


def lame_scale(x32):
    """src/web/mp3encoder.cxx:66-68: left[n] = (*ptr++) * 32768.0 -- a binary64 product stored to float."""
    if math.isnan(x32) or math.isinf(x32) or x32 == 0.0:
        return x32 * 32768.0                        # (a rational has no signed zero: -0 * 32768 = -0)
    d = round_to_double(Fraction(x32) * 32768)
    assert d == x32 * 32768.0
    f = round_to_float(Fraction(d))
    with np.errstate(over="ignore"):
        assert f == float(np.float64(d).astype(np.float32))      # (cross-check only: the value written is f)
    return f


def lame_fixture(out_dir):
    rng = np.random.default_rng(20131)
    bits = rng.integers(0, 1 << 32, 4096, dtype=np.uint64).astype(np.uint32)
    x = np.concatenate([bits.view(np.float32), np.float32([0.0, -0.0, 1.0, -1.0, 0.5, 3.0517578125e-05, 1e-40, -1e-45,
                        1.17549435e-38, 3.4028235e38, -3.4028235e38, 1.0384594e34, 1.0384593e34, 0.99999994, 0.33333334])])
    x = x[~np.isnan(x)]
    with np.errstate(over="ignore"):
        y = np.array([lame_scale(float(v)) for v in x], dtype=np.float32)
    np.savez_compressed(os.path.join(out_dir, "lame_scale.npz"), x=x, y=y)
    print(os.path.join(out_dir, "lame_scale.npz"), len(x), "values")


def main():
    vals = []
    for k in range(0, 257):
        edge = -50.0 + 25.0 * k / 255.0
        vals += f32_neighbours(struct.unpack("<f", struct.pack("<f", edge))[0], 3)
    vals += [-50.0, -25.0, 0.0, -0.0, -60.0, 12.5, 1e30, -1e30, -10000.0, 3.4028234663852886e38, -3.4028234663852886e38,
             1.401298464324817e-45, float("inf"), float("-inf"), float("nan"), -49.99999, -25.000002, -37.5]
    db = np.array(vals, dtype=np.float32)
    idx = np.array([js_index(float(v)) for v in db], dtype=np.uint8)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "palette_edges.npz")
    np.savez_compressed(out, db=db, index=idx)
    print(out, len(db), "values;", "indices seen:", len(set(idx.tolist())))
    lame_fixture(os.path.dirname(out))


if __name__ == "__main__":
    main()
