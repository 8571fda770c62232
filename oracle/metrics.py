"""Monitoring scalars.  Oracle (test infrastructure).
Restates utils/l2.m, utils/MSE.m, utils/PSNR.m, utils/snr_func.m, utils/projbox.m,
utils/max_eigenval_Gaussian_Moffat.m, utils/max_eigenval_Laplace.m."""

import numpy as np


def l2(x, y):
    """utils/l2.m:1-3.  `norm` of a MATRIX is the largest singular value (Q8)."""
    d = np.asarray(x, dtype=np.float64) - np.asarray(y, dtype=np.float64)
    if d.ndim == 2 and min(d.shape) > 1:
        return float(np.linalg.norm(d, 2) ** 2)
    return float(np.linalg.norm(d.ravel()) ** 2)


def MSE(x_true, x_app):
    """utils/MSE.m:1-4 (in dB)."""
    dimX = x_true.size
    return float(10 * np.log10(np.linalg.norm(x_true - x_app, "fro") ** 2 / dimX))


def PSNR(x, y):
    """utils/PSNR.m:2-4."""
    mse = 10 * np.log10(np.max(x) ** 2)
    return float(mse - 10 * np.log10(np.linalg.norm(x.ravel() - y.ravel()) ** 2 / x.size))


def snr_func(x, y):
    """utils/snr_func.m:1-3 (norm(.,2) of a matrix = spectral norm)."""
    return float(20 * np.log10(np.linalg.norm(x, 2) / np.linalg.norm(x - y, 2)))


def projbox(x, min_x, max_x):
    """utils/projbox.m:1-3."""
    return np.minimum(np.maximum(x, min_x), max_x)


def max_eigenval(A, At, params, im_size, tol, max_iter, randn, verbose=0):
    """Power iteration on A'A.  utils/max_eigenval_Gaussian_Moffat.m:1-27 and
    utils/max_eigenval_Laplace.m:28-55 (same body; `params` is (a,b) or (b,)).
    `randn(shape)` supplies the seeded stream the reference draws from (Q21)."""
    x = randn(tuple(im_size))                       # :4
    x = x / np.linalg.norm(x.ravel())               # :5
    init_val = 1.0                                  # :6
    val = np.nan
    for _ in range(int(max_iter)):                  # :8
        y = A(x, *params)                           # :9
        x = At(y, *params)                          # :10
        val = np.linalg.norm(x.ravel())             # :11
        rel_var = abs(val - init_val) / init_val    # :12
        if rel_var < tol:                           # :16
            break
        init_val = val                              # :19
        x = x / val                                 # :20
    return float(val)
