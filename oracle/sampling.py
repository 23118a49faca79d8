"""
Oracle restatement (NumPy) of the device samplers of csrc/sources.cu -- TEST INFRASTRUCTURE ONLY,
see oracle/__init__.py.  Philox4x32-10 keyed by (seed, stream), counter (index_lo, index_hi,
draw, epoch); uniform = (bits + 0.5) 2^-32 in float64; the transforms are the reference's
(tfrt/distributions.py:1361-1392 RandomUniformSquare, :1726-1850 the sphere family).  The raw
random bits equal the device's exactly (integer arithmetic); the transformed samples agree to the
last float32 bit except where two correctly rounded float64 libraries differ in the last place.
"""

import math

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
GOLDEN = math.pi * (1.0 + 5.0 ** 0.5)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Counter words c0..c3 (uint64 arrays holding 32-bit values), key (k0, k1) ints.  Returns
    the four output words."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def draw(seed, stream, epoch, which, first, n):
    """float64 uniforms in (0, 1) of samples first .. first + n - 1, variable `which`."""
    g = np.arange(first, first + n, dtype=np.uint64)
    w = philox4x32_10(g & MASK, g >> np.uint64(32), np.uint64(which), np.uint64(epoch), seed, stream)[0]
    return (w.astype(np.float64) + 0.5) * 2.0 ** -32


def uniform(seed, stream, epoch, which, first, n, lo, hi):
    return (lo + (hi - lo) * draw(seed, stream, epoch, which, first, n)).astype(np.float32)


def square(seed, stream, epoch, first, n, x_size, y_size):
    x = (2.0 * draw(seed, stream, epoch, 0, first, n) - 1.0) * x_size
    y = (2.0 * draw(seed, stream, epoch, 1, first, n) - 1.0) * y_size
    pts = np.stack([x, y], 1)
    return pts.astype(np.float32), (pts / max(x_size, y_size)).astype(np.float32)


def sphere(seed, stream, epoch, mode, first, n, total, angular_size, radius=1.0):
    ca = math.cos(angular_size)
    if mode < 2:
        u, v = draw(seed, stream, epoch, 0, first, n), draw(seed, stream, epoch, 1, first, n)
        c = ca + (1.0 - ca) * u if mode == 0 else np.sqrt(ca * ca + (1.0 - ca * ca) * u)
        theta = GOLDEN * v
    else:
        g = np.arange(first, first + n, dtype=np.float64)
        frac = g / (total - 1) if total > 1 else np.zeros_like(g)
        c = 1.0 + (ca - 1.0) * frac if mode == 2 else np.sqrt(1.0 + (ca * ca - 1.0) * frac)
        theta = GOLDEN * (g + 0.5)
    c = np.clip(c, -1.0, 1.0)
    theta = np.fmod(theta, 2.0 * math.pi)
    s = np.sqrt(np.maximum(0.0, 1.0 - c * c))
    pts = radius * np.stack([c, s * np.cos(theta), s * np.sin(theta)], 1)
    return pts.astype(np.float32), np.stack([np.arccos(c), theta], 1).astype(np.float32)
