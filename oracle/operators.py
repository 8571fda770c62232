"""The closures the demo scripts build and hand to SAPG through `op`.

Oracle (test infrastructure).  Restates, operation by operation and UNFUSED
(every closure call rebuilds the 7x7 PSF, pads it and FFTs it again, exactly as
the MATLAB does), the setup blocks
  run_Gaussian_demo.m:34-85,122-195
  run_moffat_demo.m:34-73,122-185   (file lines; 30+ offsets in the listing)
  run_laplace_demo.m:34-70,96-153
`op` is a plain dict standing in for the MATLAB struct.
"""

import numpy as np

from . import psf as P
from . import tv
from . import metrics

_fft2 = np.fft.fft2
_ifft2 = np.fft.ifft2


def set_fft(fft2, ifft2):
    """Swap the FFT backend (bench uses scipy.fft with workers=nproc)."""
    global _fft2, _ifft2
    _fft2, _ifft2 = fft2, ifft2
    P.set_fft2(fft2)


def _fro2(a):
    return np.linalg.norm(a, "fro") ** 2


# --------------------------------------------------------------------------
# closure families
# --------------------------------------------------------------------------
def gaussian_closures(im_size, psf_size, phi):
    """run_Gaussian_demo.m:126-139."""
    h = lambda a, b: P.Gaussian_psf(psf_size, a, b, phi)                    # :126
    H_FFT = lambda a, b: P.resize(h(a, b), im_size)                         # :128
    HC_FFT = lambda a, b: np.conj(H_FFT(a, b))                              # :129
    d1 = lambda a, b: P.diff_fftgaus_w1(im_size, psf_size, a, b, phi)       # :132
    d2 = lambda a, b: P.diff_fftgaus_w2(im_size, psf_size, a, b, phi)       # :133
    A = lambda x, a, b: np.real(_ifft2(H_FFT(a, b) * _fft2(x)))             # :136
    AT = lambda x, a, b: np.real(_ifft2(HC_FFT(a, b) * _fft2(x)))           # :137
    dif_w1 = lambda x, a, b: np.real(_ifft2(d1(a, b) * _fft2(x)))           # :138
    dif_w2 = lambda x, a, b: np.real(_ifft2(d2(a, b) * _fft2(x)))           # :139
    return dict(H_FFT=H_FFT, HC_FFT=HC_FFT, A=A, AT=AT, dif=(dif_w1, dif_w2))


def moffat_closures(im_size, psf_size):
    """run_moffat_demo.m:122-138."""
    H_FFT = lambda a, b: P.moffat_psf(im_size, psf_size, a, b)              # :122
    HC_FFT = lambda a, b: np.conj(H_FFT(a, b))                              # :125
    da = lambda a, b: P.diff_moffat_alpha(im_size, psf_size, a, b)          # :128
    db = lambda a, b: P.diff_moffat_beta(im_size, psf_size, a, b)           # :131
    A = lambda x, a, b: np.real(_ifft2(H_FFT(a, b) * _fft2(x)))             # :134
    AT = lambda x, a, b: np.real(_ifft2(HC_FFT(a, b) * _fft2(x)))           # :135
    dA = lambda x, a, b: np.real(_ifft2(da(a, b) * _fft2(x)))               # :136
    dB = lambda x, a, b: np.real(_ifft2(db(a, b) * _fft2(x)))               # :137
    return dict(H_FFT=H_FFT, HC_FFT=HC_FFT, A=A, AT=AT, dif=(dA, dB))


def laplace_closures(im_size, psf_size):
    """run_laplace_demo.m:96-107."""
    H_FFT = lambda b: P.laplace_psf(im_size, psf_size, b)                   # :96
    HC_FFT = lambda b: np.conj(H_FFT(b))                                    # :99
    db = lambda b: P.diff_laplace_b(im_size, psf_size, b)                   # :102
    A = lambda x, b: np.real(_ifft2(H_FFT(b) * _fft2(x)))                   # :105
    AT = lambda x, b: np.real(_ifft2(HC_FFT(b) * _fft2(x)))                 # :106
    dB = lambda x, b: np.real(_ifft2(db(b) * _fft2(x)))                     # :107
    return dict(H_FFT=H_FFT, HC_FFT=HC_FFT, A=A, AT=AT, dif=(dB,))


def closures(model, im_size, psf_size, phi=0.0):
    if model == P.GAUSSIAN:
        return gaussian_closures(im_size, psf_size, phi)
    if model == P.MOFFAT:
        return moffat_closures(im_size, psf_size)
    if model == P.LAPLACE:
        return laplace_closures(im_size, psf_size)
    raise ValueError("unknown model")


# --------------------------------------------------------------------------
# likelihood pieces shared by the three demos
# --------------------------------------------------------------------------
def likelihood_closures(cl, y, dimX):
    """op.f / gradF / grad_* / gradF_sigma.
    run_Gaussian_demo.m:171-175, run_moffat_demo.m:163-167, run_laplace_demo.m:132-135.
    Every lambda takes (x, *psi, sigma2)."""
    A, AT, dif = cl["A"], cl["AT"], cl["dif"]
    f = lambda x, *a: _fro2(y - A(x, *a[:-1])) / (2 * a[-1])
    gradF = lambda x, *a: np.real(AT(A(x, *a[:-1]) - y, *a[:-1]) / a[-1])
    grads = tuple(
        (lambda d: (lambda x, *a: float(np.real(np.sum(np.sum(d(x, *a[:-1]) * (A(x, *a[:-1]) - y))) / a[-1]))))(d)
        for d in dif)
    gradF_sigma = lambda x, *a: _fro2(y - A(x, *a[:-1])) / (2 * a[-1] ** 2) - dimX / (2 * a[-1])
    return f, gradF, grads, gradF_sigma


def _bsnr_sigmas(Ax, dimX, bsnr, bsnr_min, bsnr_max):
    """run_Gaussian_demo.m:148-152."""
    nrm = np.linalg.norm(Ax - np.mean(np.mean(Ax, axis=0)), "fro")
    sigma = nrm / np.sqrt(dimX * 10 ** (bsnr / 10))
    sigma_min = nrm / np.sqrt(dimX * 10 ** (bsnr_min / 10))
    sigma_max = nrm / np.sqrt(dimX * 10 ** (bsnr_max / 10))
    return sigma, sigma_min, sigma_max


DEFAULTS = {
    # run_Gaussian_demo.m:34-85
    P.GAUSSIAN: dict(samples=20000, stopTol=1e-5, warmup=15000, lambdaMax=2.0, gammaFrac=0.98,
                     min_th=1e-3, max_th=1.0, min_w1=0.1, max_w1=1.0, min_w2=0.1, max_w2=1.0,
                     BSNR_max=45, BSNR_min=15, BSNR=30, th_init=0.01, w1_init=0.5, w2_init=0.3,
                     d_exp=0.8, psf_size=7, phi=0.0, w1=0.4, w2=0.3,
                     fix_w1=1, fix_w2=1, fix_sigma=0),
    # run_moffat_demo.m:34-83
    P.MOFFAT: dict(samples=20000, stopTol=1e-5, sub_sample=1, warmup=15000, lambdaMax=2.0, gammaFrac=0.98,
                   min_th=1e-3, max_th=1.0, min_alpha=1e-2, max_alpha=1.0, min_beta=0.1, max_beta=10.0,
                   BSNR_max=35, BSNR_min=18, BSNR=30, psf_size=7, th_init=0.01, alpha_init=1.0,
                   beta_init=10.0, d_exp=0.8, alpha=0.4, beta=3.5,
                   fix_alpha=0, fix_beta=0, fix_sigma=0),
    # run_laplace_demo.m:34-70
    P.LAPLACE: dict(samples=20000, stopTol=1e-5, warm_sample=1, warmup=15000, lambdaMax=0.1, gammaFrac=0.98,
                    min_th=1e-3, max_th=1.0, th_init=0.01, min_b=1e-3, max_b=1.0, b_init=0.1,
                    BSNR_max=45, BSNR_min=15, BSNR=30, psf_size=7, d_exp=0.8, b=0.3,
                    fix_b=0, fix_sigma=0),
}

C_GAUSSIAN = dict(sigma=1000.0, theta=0.01, w1=10.0, w2=10.0, lam=1.0, gam=1.0)  # run_Gaussian_demo.m:34-39


def setup_demo(model, x, randn, chambolleit=25, **overrides):
    """Everything the demo script does between reading the image and calling
    SAPG: power iteration for evMax, observation synthesis, step sizes and the
    `op` closures.  Returns (y, op) [Gaussian: (y, op, c)].

    `randn(shape)` is the seeded stream; it is consumed in the reference's
    order: first by the power iteration (Q21), then by the observation noise.
    """
    x = np.asarray(x, dtype=np.float64)
    op = dict(DEFAULTS[model])
    op.update(overrides)
    if "burnIn" not in overrides:
        op["burnIn"] = (op["samples"] * 80) // 100              # run_Gaussian_demo.m:49
    if "d_scale" not in overrides:
        op["d_scale"] = 0.01 / op["th_init"]                    # :72
    psf_size = op["psf_size"]
    dimX = x.size
    im_size = x.shape
    op["x"] = x

    if model == P.GAUSSIAN:
        if op["fix_w1"]:
            op["w1_init"] = op["w1"]                            # run_Gaussian_demo.m:102-104
        if op["fix_w2"]:
            op["w2_init"] = op["w2"]                            # :105-107
        cl = gaussian_closures(im_size, psf_size, op["phi"])
        ev_params, true_params = (1.0, 1.0), (op["w1"], op["w2"])          # :142,145
    elif model == P.MOFFAT:
        if op["fix_alpha"]:
            op["alpha_init"] = op["alpha"]                      # run_moffat_demo.m:95-98
        if op["fix_beta"]:
            op["beta_init"] = op["beta"]                        # :99-102
        cl = moffat_closures(im_size, psf_size)
        ev_params, true_params = (1.0, 5.0), (op["alpha"], op["beta"])     # :140,143
    else:
        if op["fix_b"]:
            op["b_init"] = op["b"]                              # run_laplace_demo.m:76-79
        cl = laplace_closures(im_size, psf_size)
        ev_params, true_params = (1.0,), (op["b"],)             # :110,114
    A, AT = cl["A"], cl["AT"]

    if "evMax" in overrides:       # bench only: skip the power iteration
        evMax = float(overrides["evMax"])
    else:
        evMax = metrics.max_eigenval(A, AT, ev_params, im_size, 1e-4, 1e4, randn)
    op["evMax"] = evMax

    Ax = np.real(A(x, *true_params))
    sigma, sigma_min, sigma_max = _bsnr_sigmas(Ax, dimX, op["BSNR"], op["BSNR_min"], op["BSNR_max"])
    op["sigma"] = sigma
    if op["fix_sigma"]:
        op["sigma_init"] = sigma ** 2                           # run_Gaussian_demo.m:158
    else:
        op["sigma_init"] = (sigma_min ** 2 + sigma_max ** 2) / 2    # :160
    op["sigma_min"] = sigma_min ** 2                            # :162
    op["sigma_max"] = sigma_max ** 2                            # :163

    y = Ax + sigma * randn(Ax.shape)                            # :166-168
    if model == P.LAPLACE:
        op["X0"] = y                                            # run_laplace_demo.m:127

    f, gradF, grads, gradF_sigma = likelihood_closures(cl, y, dimX)
    op["f"], op["gradF"], op["gradF_sigma"] = f, gradF, gradF_sigma
    if model == P.GAUSSIAN:
        op["grad_w1"], op["grad_w2"] = grads
    elif model == P.MOFFAT:
        op["grad_alpha"], op["grad_beta"] = grads
    else:
        op["grad_b"] = grads[0]

    Lf = lambda s2: evMax ** 2 / s2                             # run_Gaussian_demo.m:178
    if model == P.LAPLACE:
        op["Lf"] = max(Lf(sigma_min ** 2), Lf(sigma_max ** 2))  # run_laplace_demo.m:137 (Q13)
    else:
        op["Lf"] = min(Lf(sigma_min ** 2), Lf(sigma_max ** 2))  # run_Gaussian_demo.m:179
    op["lambda"] = min(5 / op["Lf"], op["lambdaMax"])           # :182
    op["gamma_max"] = 1 / (op["Lf"] + (1 / op["lambda"]))       # :183
    if model == P.LAPLACE:
        op["gamma"] = 10 * op["gammaFrac"] * op["gamma_max"]    # run_laplace_demo.m:142 (Q13)
    else:
        op["gamma"] = op["gammaFrac"] * op["gamma_max"]         # run_Gaussian_demo.m:184

    op["g"] = lambda xx: tv.TVnorm(xx)                          # :187
    op["chambolleit"] = chambolleit                             # :188
    if model == P.GAUSSIAN:
        lam = op["lambda"]
        op["proxG"] = lambda xx, theta: tv.chambolle_prox_TV_stop(
            xx, "lambda", lam * theta, "maxiter", chambolleit)[0]           # :191
    else:
        op["proxG"] = lambda xx, lam_, theta: tv.chambolle_prox_TV_stop(
            xx, "lambda", lam_ * theta, "maxiter", chambolleit)[0]          # run_moffat_demo.m:181
    g = op["g"]
    op["logPi"] = lambda xx, theta, *a: -f(xx, *a) - theta * g(xx)          # :195
    op["model"] = model
    op["y"] = y
    op["closures"] = cl
    if model == P.GAUSSIAN:
        return y, op, dict(C_GAUSSIAN)
    return y, op
