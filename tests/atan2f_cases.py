"""Argument sets for the atan2f pinning tests (shared by the CPU and the GPU test)."""
import numpy as np

# branch thresholds of glibc's s_atanf.c / e_atan2f.c (bit patterns of |y/x|), plus 1.0 and the
# limits where the big-ratio shortcuts start
THRESHOLDS = [0x3ee00000, 0x3f300000, 0x3f980000, 0x401c0000, 0x31000000, 0x4c000000, 0x4b800000,
              0x4c800000, 0x50800000, 0x3f800000, 0x5d800000, 0x21800000]


def random_bits(n, seed):
    """Uniform over all float bit patterns (every exponent, NaNs and infinities included)."""
    rng = np.random.default_rng(seed)
    y = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    x = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    return y, x


def discriminator_like(n, seed):
    """What the FM discriminator feeds it: conjugate products of moderate magnitude, some tiny."""
    rng = np.random.default_rng(seed)
    y = rng.uniform(-1, 1, n).astype(np.float32)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    y[1::4] *= np.float32(1e-3)
    x[2::4] *= np.float32(1e-4)
    return y, x


def threshold_sweep(width=2000):
    """Ratios within +-width ULP of every branch threshold, in all four quadrants, both through
    the x == 1 shortcut and through the division."""
    ys, xs = [], []
    d = np.arange(-width, width + 1, dtype=np.int64)
    for t in THRESHOLDS:
        r = (np.int64(t) + d).astype(np.uint32).view(np.float32)
        for sx in (3.0, -3.0):
            for sy in (1.0, -1.0):
                ys.append((r * np.float32(sx) * np.float32(sy)).astype(np.float32))
                xs.append(np.full(r.size, sx, np.float32))
        ys.append(r)
        xs.append(np.ones(r.size, np.float32))
    return np.concatenate(ys), np.concatenate(xs)


def specials():
    v = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-45, -1e-45, 1e-38, 3.4e38, -3.4e38,
                  0.5, 2.0, 1e-30, 1e30], np.float32)
    y, x = np.meshgrid(v, v)
    return y.ravel().copy(), x.ravel().copy()


def same(a, b):
    """Bit-identical, NaNs of any payload counted as equal."""
    a = np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
