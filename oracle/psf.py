"""Parametric 7x7 PSFs, their parameter derivatives and the pad-and-FFT step.

Oracle (test infrastructure).  numpy float64 restatement of
  utils/Gaussian_psf.m, utils/psf_gaussian.m, utils/Sum_gauss_psf.m,
  utils/diff_fftgaus_w1.m, utils/diff_fftgaus_w2.m,
  utils/moffat_psf.m, utils/psf_moffat.m, utils/sum_mof_psf.m,
  utils/diff_moffat_alpha.m, utils/diff_moffat_beta.m,
  utils/laplace_psf.m, utils/psf_laplace.m, utils/sum_lap_psf.m,
  utils/diff_laplace_b.m, utils/resize.m
Index convention: k[i, j] here is MATLAB kernel(i+1, j+1).
"""

import numpy as np

_fft2 = np.fft.fft2


def set_fft2(fn):
    """Swap the fft2 used by `resize` (bench uses scipy.fft with workers)."""
    global _fft2
    _fft2 = fn


def _axis(t):
    # Gaussian_psf.m:3-6  center=(t+1)/2; x = -t+center : t-center
    center = (t + 1) / 2.0
    return np.arange(-t + center, t - center + 0.5, 1.0)


# --------------------------------------------------------------------------
# pad + FFT                                                  utils/resize.m
# --------------------------------------------------------------------------
def resize(kernel, im_shape):
    """utils/resize.m:1-12.  `taille = length(kernel)` (largest dimension, Q2);
    the kernel is written at the TOP-LEFT corner (Q1, no centring) and the
    full complex fft2 is returned."""
    kernel = np.asarray(kernel, dtype=np.float64)
    taille = max(kernel.shape)                      # resize.m:2  length()
    full = np.zeros((int(im_shape[0]), int(im_shape[1])))   # resize.m:6
    full[:taille, :taille] = kernel                 # resize.m:8
    return _fft2(full)                              # resize.m:11


# --------------------------------------------------------------------------
# Gaussian
# --------------------------------------------------------------------------
def _gauss_uv(t, phi):
    x = _axis(t)
    # [v,u] = ndgrid(x,y): v varies along rows, u along columns  (Gaussian_psf.m:9)
    v, u = np.meshgrid(x, x, indexing="ij")
    U = u * np.cos(phi) - v * np.sin(phi)           # :11
    V = u * np.sin(phi) + v * np.cos(phi)           # :12
    return U, V


def psf_gaussian(t, w1, w2, phi):
    """utils/psf_gaussian.m:2-19 (identical to Gaussian_psf.m:2-19)."""
    U, V = _gauss_uv(t, phi)
    c = w1 ** 2 * U ** 2 + w2 ** 2 * V ** 2         # :14
    k = ((w1 * w2) / (2 * np.pi)) * np.exp(-c / 2)  # :16
    return k / k.sum()                              # :18


Gaussian_psf = psf_gaussian


def Sum_gauss_psf(t, w1, w2, phi):
    """utils/Sum_gauss_psf.m:1-28 -> (sum_psf, sum_diffw1, sum_diffw2)."""
    U, V = _gauss_uv(t, phi)
    c = w1 ** 2 * U ** 2 + w2 ** 2 * V ** 2
    e = np.exp(-c / 2)
    f = ((w1 * w2) / (2 * np.pi)) * e                       # :18
    dw2 = (w1 / (2 * np.pi)) * (1 - w2 ** 2 * V ** 2) * e   # :20
    dw1 = (w2 / (2 * np.pi)) * (1 - w1 ** 2 * U ** 2) * e   # :22
    return f.sum(), dw1.sum(), dw2.sum()


def dpsf_gaussian_w1(t, w1, w2, phi):
    """Spatial 7x7 part of utils/diff_fftgaus_w1.m:4-24 (before resize)."""
    U, V = _gauss_uv(t, phi)
    s, sd1, _ = Sum_gauss_psf(t, w1, w2, phi)               # :16
    c = w1 ** 2 * U ** 2 + w2 ** 2 * V ** 2
    e = np.exp(-c / 2)
    f = ((w1 * w2) / (2 * np.pi)) * e                       # :20
    d = (w2 / (2 * np.pi)) * (1 - w1 ** 2 * U ** 2) * e     # :22
    return (d * s - f * sd1) / (s ** 2)                     # :24


def dpsf_gaussian_w2(t, w1, w2, phi):
    """Spatial 7x7 part of utils/diff_fftgaus_w2.m:4-24."""
    U, V = _gauss_uv(t, phi)
    s, _, sd2 = Sum_gauss_psf(t, w1, w2, phi)
    c = w1 ** 2 * U ** 2 + w2 ** 2 * V ** 2
    e = np.exp(-c / 2)
    f = ((w1 * w2) / (2 * np.pi)) * e
    d = (w1 / (2 * np.pi)) * (1 - w2 ** 2 * V ** 2) * e     # :22
    return (d * s - f * sd2) / (s ** 2)


def diff_fftgaus_w1(im_size, t, w1, w2, phi):
    return resize(dpsf_gaussian_w1(t, w1, w2, phi), im_size)    # diff_fftgaus_w1.m:25


def diff_fftgaus_w2(im_size, t, w1, w2, phi):
    return resize(dpsf_gaussian_w2(t, w1, w2, phi), im_size)    # diff_fftgaus_w2.m:25


# --------------------------------------------------------------------------
# Moffat
# --------------------------------------------------------------------------
def _r2(t):
    x = _axis(t)
    return x[:, None] ** 2 + x[None, :] ** 2        # xy = X(ii)^2 + Y(jj)^2


def _moffat_raw(t, a, b):
    xy = _r2(t)
    b2 = b + 2
    return a ** 2 * (xy * a ** 2 / b + 1) ** (-b2 / 2) / (2 * np.pi)   # moffat_psf.m:16


def psf_moffat(t, a, b):
    """utils/psf_moffat.m:2-20."""
    k = _moffat_raw(t, a, b)
    return k / k.sum()


def moffat_psf(im_shape, t, a, b):
    """utils/moffat_psf.m:2-23 (spectrum)."""
    return resize(psf_moffat(t, a, b), im_shape)


def _moffat_dalpha_raw(t, a, b):
    # sum_mof_psf.m:24 / diff_moffat_alpha.m:17 -- reproduces the reference
    # formula *including* the stray 2 in the denominator (Q7).
    xy = _r2(t)
    return (2 - (((b + 2) * xy * a ** 2) / (2 * (b + xy * a ** 2)))) \
        * (1 + xy * a ** 2 / b) ** (-(b + 2) / 2) * (a / (2 * np.pi))


def _moffat_dbeta_raw(t, a, b):
    # sum_mof_psf.m:33-36 / diff_moffat_beta.m:14-18
    xy = _r2(t)
    b2 = b + 2
    cons1 = a ** 2 / (4 * np.pi)
    return (-np.log(xy * a ** 2 / b + 1) + (b2 * xy * a ** 2) / (b * (b + xy * a ** 2))) \
        * (xy * a ** 2 / b + 1) ** (-b2 / 2) * cons1


def sum_mof_psf(t, a, b):
    """utils/sum_mof_psf.m:1-40 -> (sum_psf, sum_diff_alpha, sum_diff_beta)."""
    return (_moffat_raw(t, a, b).sum(), _moffat_dalpha_raw(t, a, b).sum(),
            _moffat_dbeta_raw(t, a, b).sum())


def dpsf_moffat_alpha(t, a, b):
    """Spatial part of utils/diff_moffat_alpha.m:7-20."""
    s, sda, _ = sum_mof_psf(t, a, b)
    f = _moffat_raw(t, a, b)
    d = _moffat_dalpha_raw(t, a, b)
    return (d * s - f * sda) / s ** 2               # :18


def dpsf_moffat_beta(t, a, b):
    """Spatial part of utils/diff_moffat_beta.m:7-21."""
    s, _, sdb = sum_mof_psf(t, a, b)
    f = _moffat_raw(t, a, b)
    d = _moffat_dbeta_raw(t, a, b)
    return (d * s - f * sdb) / s ** 2               # :19


def diff_moffat_alpha(im_shape, t, a, b):
    return resize(dpsf_moffat_alpha(t, a, b), im_shape)     # :21


def diff_moffat_beta(im_shape, t, a, b):
    return resize(dpsf_moffat_beta(t, a, b), im_shape)      # :22


# --------------------------------------------------------------------------
# Laplace
# --------------------------------------------------------------------------
def _l1(t):
    x = _axis(t)
    return np.abs(x)[:, None] + np.abs(x)[None, :]


def _laplace_raw(t, b):
    return (b ** 2 / 4) * np.exp(-b * _l1(t))       # laplace_psf.m:8


def psf_laplace(t, b):
    """utils/psf_laplace.m:1-13."""
    k = _laplace_raw(t, b)
    return k / k.sum()


def laplace_psf(im_shape, t, b):
    """utils/laplace_psf.m:1-15 (spectrum)."""
    return resize(psf_laplace(t, b), im_shape)


def _laplace_db_raw(t, b):
    s = _l1(t)
    var1 = (2 * b - b ** 2 * s) / 4                 # sum_lap_psf.m:20
    var2 = np.exp(-b * s)                           # :21
    return var1 * var2


def sum_lap_psf(t, b):
    """utils/sum_lap_psf.m:1-28 -> (Sum_psf, sum_diff_b)."""
    return _laplace_raw(t, b).sum(), _laplace_db_raw(t, b).sum()


def dpsf_laplace_b(t, b):
    """Spatial part of utils/diff_laplace_b.m:2-15."""
    s, sdb = sum_lap_psf(t, b)
    f = _laplace_raw(t, b)
    d = _laplace_db_raw(t, b)
    return (d * s - f * sdb) / s ** 2               # :13


def diff_laplace_b(im_shape, t, b):
    return resize(dpsf_laplace_b(t, b), im_shape)   # :18


# --------------------------------------------------------------------------
# model-indexed access (0 gaussian, 1 moffat, 2 laplace) used by the tests
# --------------------------------------------------------------------------
GAUSSIAN, MOFFAT, LAPLACE = 0, 1, 2


def taps(model, t, psi, phi=0.0, which=0):
    """which: 0 = normalised PSF, 1 = d/dpsi[0], 2 = d/dpsi[1]."""
    if model == GAUSSIAN:
        fn = (psf_gaussian, dpsf_gaussian_w1, dpsf_gaussian_w2)[which]
        return fn(t, psi[0], psi[1], phi)
    if model == MOFFAT:
        fn = (psf_moffat, dpsf_moffat_alpha, dpsf_moffat_beta)[which]
        return fn(t, psi[0], psi[1])
    if model == LAPLACE:
        fn = (psf_laplace, dpsf_laplace_b)[which]
        return fn(t, psi[0])
    raise ValueError("unknown model")
