"""Counter-based Gaussian noise shared by the oracle and the CUDA engine.

Oracle (test infrastructure).  The reference draws `randn(size(X))` from the
MATLAB-v5 'state' generator (run_Gaussian_demo.m:88,
SAPG_algorithm_Guassian.m:81,160), which cannot be reproduced outside MATLAB.
Parity therefore uses either an explicit noise tensor, or this generator, which
the CUDA kernel `sbd::philox_normal2` implements identically:

  Philox4x32-10 (Salmon et al. 2011), key = (seed_lo, seed_hi),
  counter = (pair_lo, pair_hi, step, stream)
      pair   = index of the PAIR of consecutive elements (column-major linear
               index >> 1) inside one N-element image
      step   = draw number (MYULA iteration counter)
      stream = chain / purpose id
  u1 = (((x1 << 32 | x0) >> 11) + 0.5) * 2^-53,  u2 likewise from (x3, x2)
  z0 = sqrt(-2 ln u1) cos(2 pi u2),  z1 = sqrt(-2 ln u1) sin(2 pi u2)
  element 2*pair gets z0, element 2*pair+1 gets z1.
"""

import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10.  Inputs are uint64 arrays/scalars holding
    32-bit values; returns four uint64 arrays of 32-bit values."""
    c0 = np.asarray(c0, dtype=np.uint64); c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64); c3 = np.asarray(c3, dtype=np.uint64)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0                     # 64-bit products of 32-bit values
        p1 = M1 * c2
        hi0 = p0 >> np.uint64(32); lo0 = p0 & MASK
        hi1 = p1 >> np.uint64(32); lo1 = p1 & MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def normal(n, seed, stream, step):
    """n standard normals (float64) for linear element indices 0..n-1."""
    npair = (n + 1) // 2
    pair = np.arange(npair, dtype=np.uint64)
    x0, x1, x2, x3 = philox4x32_10(pair & MASK, pair >> np.uint64(32),
                                   np.uint64(int(step) & 0xFFFFFFFF),
                                   np.uint64(int(stream) & 0xFFFFFFFF),
                                   int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    a = ((x1 << np.uint64(32)) | x0) >> np.uint64(11)
    b = ((x3 << np.uint64(32)) | x2) >> np.uint64(11)
    u1 = (a.astype(np.float64) + 0.5) * 2.0 ** -53
    u2 = (b.astype(np.float64) + 0.5) * 2.0 ** -53
    r = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    out = np.empty(2 * npair)
    out[0::2] = r * np.cos(ang)
    out[1::2] = r * np.sin(ang)
    return out[:n]


def randn_image(shape, seed, stream, step):
    """Noise image in MATLAB (column-major) element order: element with
    column-major linear index l gets normal()[l]."""
    n = int(shape[0]) * int(shape[1])
    return normal(n, seed, stream, step).reshape((shape[1], shape[0])).T.copy()
