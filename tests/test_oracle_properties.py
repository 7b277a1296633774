"""Analytic properties of the oracle's restatement (CPU only): results that follow from the
reference's arithmetic alone, independent of oracle/_ref and of the golden fixtures -- a second,
size-independent pin next to tests/test_oracle_vs_ref.py and tests/test_oracle_golden.py."""
import numpy as np
import pytest

from helpers import assert_biteq

RNG = np.random.default_rng(0xB200)


def lattice(n):
    return ((RNG.integers(0, 256, 2 * n).astype(np.float32) - 128.0) / 128.0).astype(np.float32)


def test_mixer_at_zero_if_is_the_identity(wro):
    """downconverter.cxx:97-111 with phaseStep 0 and phase 0: sin index 0 -> 0.0, cos index 16384 ->
    sinf(pi/2) = 1.0, so I' = i*1 + q*0 and Q' = q*1 - i*0 reproduce the input bit for bit."""
    t = wro.sintable()
    assert t[0] == 0.0 and t[16384] == 1.0
    iq = lattice(4096)
    out, ph = wro.mix(t, 0, 0, iq)
    assert ph == 0
    assert_biteq(out, iq, "IF 0")


def test_mixer_quarter_turn_swaps_and_negates(wro):
    """A constant phase of a quarter turn (index 16384): sin = 1, cos = sinf(pi) ~ -8.7e-8 -- the pin of
    SURVEY.md 8c -- so I' = i*cos + q and Q' = q*cos - i, each product and sum rounded on its own."""
    t = wro.sintable()
    c = t[32768]
    assert c == np.float32(-8.742278e-08)
    iq = lattice(1024)
    out, _ = wro.mix(t, 16384 << 15, 0, iq)
    i, q = iq[0::2], iq[1::2]
    want = np.empty_like(iq)
    want[0::2] = (i * c).astype(np.float32) + (q * np.float32(1.0)).astype(np.float32)
    want[1::2] = (q * c).astype(np.float32) - (i * np.float32(1.0)).astype(np.float32)
    assert_biteq(out, want, "quarter turn")


def test_mixer_phase_uses_the_pre_increment_value_and_wraps_at_31_bits(wro):
    """downconverter.cxx:97-104: index from the phase BEFORE the step is added; the accumulator is
    masked to 31 bits, so a block of n frames ends at (phase0 + n*step) mod 2^31."""
    t = wro.sintable()
    step = wro.phase_step(-345678, 2400000)      # negative IF: the step is negative, the mask still applies
    assert step < 0
    n, ph0 = 5000, 0x7FFFFF00
    iq = lattice(n)
    out, ph = wro.mix(t, ph0, step, iq)
    assert ph == (ph0 + n * step) & 0x7FFFFFFF
    # frame 0 is mixed with the table entries of ph0 itself
    s, c = t[ph0 >> 15], t[((ph0 >> 15) + 16384) & 0xFFFF]
    i0, q0 = iq[0], iq[1]
    assert out[0] == np.float32(np.float32(i0 * c) + np.float32(q0 * s))
    assert out[1] == np.float32(np.float32(q0 * c) - np.float32(i0 * s))
    # two half blocks carry the phase exactly like one block
    a, pa = wro.mix(t, ph0, step, iq[:2 * 1234])
    b, pb = wro.mix(t, pa, step, iq[2 * 1234:])
    assert pb == ph
    assert_biteq(np.concatenate([a, b]), out, "split block")


@pytest.mark.parametrize("ch,n,d", [(2, 64, 10), (1, 64, 5), (2, 127, 50), (2, 255, 50), (1, 9, 3)])
def test_fir_with_a_single_unit_tap_is_a_delayed_decimator(wro, ch, n, d):
    """lowpass.cxx:151-159: coeff[N-1-j] multiplies block[k*D + j]; the block starts N-1 frames in
    the past.  A unit tap at coefficient index m therefore outputs input frame k*D - m exactly
    (zeros before the first call), all other products being +-0."""
    F = 40 * d
    x = lattice(F)[:F * ch] if ch == 2 else lattice(F)[:F]
    for m in (0, 1, n - 1):
        taps = np.zeros(n, np.float32)
        taps[m] = 1.0
        y = wro.Fir(ch, taps, d).process(x).reshape(-1, ch)
        xs = x.reshape(-1, ch)
        for k in range(F // d):
            src = k * d - m
            want = xs[src] if src >= 0 else np.zeros(ch, np.float32)
            assert np.array_equal(y[k], want), (m, k)


@pytest.mark.parametrize("blk", [500, 50])
def test_fir_history_makes_block_boundaries_invisible(wro, blk):
    """lowpass.cxx:133-141: the last N-1 frames are carried, so a stream cut into equal blocks (as
    DspSource delivers them; also blocks SHORTER than the history) gives the same output as one
    block.  Blocks of changing size are a different matter in the reference: it resizes its work
    buffer before it moves the history (lowpass.cxx:136-139), which the oracle restates
    (test_oracle_vs_ref.py) and the product deliberately does not (known differences)."""
    n, d = 127, 50
    taps = RNG.standard_normal(n).astype(np.float32)
    x = lattice(2000)
    whole = wro.Fir(2, taps, d).process(x)
    f = wro.Fir(2, taps, d)
    parts = [f.process(x[2 * a:2 * (a + blk)]) for a in range(0, 2000, blk)]
    assert_biteq(np.concatenate(parts), whole, "split stream")


def test_fir_summation_order_is_the_reference_order(wro):
    """The sum runs j = 0..N-1 over coeff[N-1-j]*block[k*D+j], every product and every partial sum
    rounded to float (no FMA, no pairwise tree): restated here with numpy float32 scalars."""
    n, d = 33, 7
    taps = RNG.standard_normal(n).astype(np.float32)
    x = RNG.standard_normal(7 * 12).astype(np.float32)
    y = wro.Fir(1, taps, d).process(x)
    block = np.concatenate([np.zeros(n - 1, np.float32), x])
    for k in range(len(y)):
        acc = np.float32(0.0)
        for j in range(n):
            acc = np.float32(acc + np.float32(taps[n - 1 - j] * block[k * d + j]))
        assert acc == y[k], k


def test_demodulators_on_exact_inputs(wro):
    """demodulator.cxx:83-112 on inputs whose results are exact: AM of 3-4-5 triangles, USB/LSB as
    plain sum and difference, FM with the real part passed as y (a constant carrier gives +0.25, the
    first sample atan2f(0,0) = 0, a quarter turn per sample gives 0 or 0.5)."""
    iq = np.array([3, 4, -6, 8, 0, 0, 5, -12], np.float32) / np.float32(16)
    prev = np.zeros(2, np.float32)
    assert_biteq(wro.demod(wro.MODES["AM"], prev, iq), np.array([5, 10, 0, 13], np.float32) / np.float32(16), "AM")
    i, q = iq[0::2], iq[1::2]
    assert_biteq(wro.demod(wro.MODES["USB"], np.zeros(2, np.float32), iq), i + q, "USB")
    assert_biteq(wro.demod(wro.MODES["LSB"], np.zeros(2, np.float32), iq), i - q, "LSB")
    carrier = np.tile(np.array([0.5, 0.0], np.float32), 8)
    prev = np.zeros(2, np.float32)
    fm = wro.demod(wro.MODES["FM"], prev, carrier)
    assert fm[0] == 0.0 and np.all(fm[1:] == np.float32(0.25))
    assert prev[0] == np.float32(0.5) and prev[1] == 0.0       # the look-back sample is carried
    # +90 degrees per sample: conj product = (0, +|z|^2) in (ii, qq) terms -> atan2f(0, +x) = 0
    rot = np.array([1, 0, 0, 1, -1, 0, 0, -1, 1, 0], np.float32)
    fm = wro.demod(wro.MODES["FM"], np.array([0, -1], np.float32), rot)
    assert np.all(fm == 0.0)
    # -90 degrees per sample: qq < 0, ii = 0 -> atan2f(+0, -x) = pi -> 0.5
    fm = wro.demod(wro.MODES["FM"], np.array([0, 1], np.float32), rot[::-1].copy().reshape(-1, 2)[:, ::-1].ravel())
    assert np.all(np.abs(fm) == np.float32(0.5))


def test_default_design_is_hamming_over_n(wro):
    """lowpass.cxx:105-110,167-189 at the shipped pass-bands: maxbin = 1, the mask has one bin, the
    inverse DFT is all ones, so coeff[n] = (0.54 - 0.46*cosf(2*pi*n/(N-1)))/N; symmetric, DC gain
    0.5328125."""
    for fs, pb in ((2400000, 80000), (240000, 8000)):
        h = wro.lowpass_design(64, pb, fs)
        assert np.allclose(h, h[::-1], rtol=0, atol=4e-9)     # cosf is rounded per entry
        n = np.arange(64)
        want = (0.54 - 0.46 * np.cos(2 * np.pi * n / 63)) / 64
        assert np.max(np.abs(h - want)) < 4e-9
        assert abs(float(np.sum(h.astype(np.float64))) - 0.5328125) < 1e-6
    # narrow pass-bands collapse to all-zero taps: maxbin = 64*12500/2400000/2 = 0 (lowpass.cxx:167)
    assert not np.any(wro.lowpass_design(64, 12500, 2400000))


def test_spectrum_of_a_bin_centred_tone(wro):
    """spectrumsink.cxx:71-74,101-121,127-140: Hamming window, forward unnormalised DFT, dB relative
    to N, fft-shift.  A unit tone on bin b peaks at output index (b + N/2) mod N with
    20*log10(0.54) dB (the window's DC gain) and Hamming's -6 dB neighbours at +-1."""
    n = 512
    for b in (5, 200, 300, 511):
        k = np.arange(n)
        tone = np.exp(2j * np.pi * b * k / n)
        iq = np.empty(2 * n, np.float32)
        iq[0::2], iq[1::2] = tone.real, tone.imag
        sp = wro.Spectrum(n)
        rows = sp.process(iq)
        assert rows.shape == (1, n)
        peak = (b + n // 2) % n
        assert int(np.argmax(rows[0])) == peak
        assert abs(rows[0][peak] - 20 * np.log10(0.54)) < 2e-2
        for nb in ((peak - 1) % n, (peak + 1) % n):
            assert abs(rows[0][nb] - 20 * np.log10(0.23)) < 5e-2
        far = rows[0][(peak + n // 2) % n]
        assert far < rows[0][peak] - 60


def test_spectrum_hop_counts_frames(wro):
    """One transform per full frame; with the overlap extension a new frame every `hop` frames once
    the first N are in; the partial frame is carried to the next block."""
    iq = lattice(512 * 3 + 100)
    assert wro.Spectrum(512).process(iq).shape[0] == 3
    sp = wro.Spectrum(512, 256)
    r1 = sp.process(iq[:2 * 700])
    r2 = sp.process(iq[2 * 700:])
    total = (512 * 3 + 100 - 512) // 256 + 1
    assert r1.shape[0] + r2.shape[0] == total
    whole = wro.Spectrum(512, 256).process(iq)
    assert_biteq(np.concatenate([r1, r2]), whole, "split spectrum stream")


# ---- randomised geometries against an independent numpy float32 restatement ----

from hypothesis import given, settings, strategies as st   # noqa: E402


def numpy_fir(taps, x, ch, d, hist):
    """lowpass.cxx:133-159 with numpy float32 vectors: for each tap j (in the reference's order) one
    rounded product and one rounded sum, over all outputs at once."""
    n = len(taps)
    block = np.concatenate([hist, x]).reshape(-1, ch)
    nout = (len(x) // ch) // d
    acc = np.zeros((nout, ch), np.float32)
    base = np.arange(nout) * d
    for j in range(n):
        acc = (acc + (np.float32(taps[n - 1 - j]) * block[base + j]).astype(np.float32)).astype(np.float32)
    return acc.ravel(), block[len(block) - (n - 1):].ravel() if n > 1 else np.zeros(0, np.float32)


@settings(max_examples=40, deadline=None, derandomize=True)
@given(n=st.integers(1, 300), d=st.integers(1, 64), ch=st.sampled_from([1, 2]), blocks=st.integers(1, 3),
       mult=st.integers(1, 12), seed=st.integers(0, 2**31 - 1))
def test_fir_random_geometry(wro, n, d, ch, blocks, mult, seed):
    rng = np.random.default_rng(seed)
    taps = rng.standard_normal(n).astype(np.float32)
    frames = d * mult
    f = wro.Fir(ch, taps, d)
    hist = np.zeros((n - 1) * ch, np.float32)
    for _ in range(blocks):
        x = ((rng.integers(0, 256, frames * ch).astype(np.float32) - 128.0) / 128.0).astype(np.float32)
        got = f.process(x)
        want, hist = numpy_fir(taps, x, ch, d, hist)
        assert_biteq(got, want, f"n={n} d={d} ch={ch} frames={frames}")


@settings(max_examples=40, deadline=None, derandomize=True)
@given(phase=st.integers(0, 2**31 - 1), step=st.integers(-2**31, 2**31 - 1), frames=st.integers(1, 3000),
       seed=st.integers(0, 2**31 - 1))
def test_mixer_random_phase_and_step(wro, phase, step, frames, seed):
    """downconverter.cxx:97-111 with numpy: index from the pre-increment phase, 31-bit wrap, four
    rounded products, two rounded sums."""
    t = wro.sintable()
    rng = np.random.default_rng(seed)
    iq = ((rng.integers(0, 256, 2 * frames).astype(np.float32) - 128.0) / 128.0).astype(np.float32)
    ph = (phase + np.arange(frames, dtype=np.int64) * step) & 0x7FFFFFFF
    si = (ph >> 15).astype(np.int64)
    ci = (si + 16384) & 0xFFFF
    s, c = t[si], t[ci]
    i, q = iq[0::2], iq[1::2]
    want = np.empty_like(iq)
    want[0::2] = ((i * c).astype(np.float32) + (q * s).astype(np.float32)).astype(np.float32)
    want[1::2] = ((q * c).astype(np.float32) - (i * s).astype(np.float32)).astype(np.float32)
    got, end = wro.mix(t, phase, step, iq)
    assert end == int((phase + frames * step) & 0x7FFFFFFF)
    assert_biteq(got, want, f"phase={phase} step={step}")


@settings(max_examples=30, deadline=None, derandomize=True)
@given(if_hz=st.integers(-2**31, 2**31 - 1), fs=st.integers(1, 2**32 - 1))
def test_phase_step_formula(wro, if_hz, fs):
    """downconverter.cxx:65,80: (int)((int64)hz * 2^31 / (int64)Fs), truncating toward zero, then
    narrowed to 32 bits."""
    q = abs(if_hz) * 2**31 // fs
    q = -q if if_hz < 0 else q
    q = (q + 2**31) % 2**32 - 2**31          # the (int) narrowing of an out-of-range int64
    assert wro.phase_step(if_hz, fs) == q
