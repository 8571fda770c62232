"""The bulk / tensor-copy FFT passes (csrc/fft2.cuh, images with rows, cols >= 1024 and the 7x7 PSF) beyond what the
other files already cover (square 2048^2 / 4096^2 vs cuFFT, 4096^2 SAPG, 1024^2 blur): rectangular shapes in both
orientations, the one-pass likelihood closures, the SALSA least-squares filter mode, the tiled passes of fft.cuh kept
as an alternative (SBD_FFT_V2=0) and a PSF size that must NOT take the new passes.  Tolerances as everywhere: 1e-12
for the deterministic operators."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rel, ROOT

pytestmark = pytest.mark.gpu
PSI = {0: (0.4, 0.3), 1: (0.4, 3.5), 2: (0.3,)}


@pytest.fixture(scope="module")
def O():
    import oracle
    import scipy.fft
    w = os.cpu_count() or 1
    oracle.operators.set_fft(lambda a: scipy.fft.fft2(a, workers=w), lambda a: scipy.fft.ifft2(a, workers=w))
    yield oracle
    oracle.operators.set_fft(np.fft.fft2, np.fft.ifft2)


@pytest.mark.parametrize("shape", [(1024, 2048), (2048, 1024), (4096, 1024)])
@pytest.mark.parametrize("model", [0, 2])
def test_rectangular_blur_and_likelihood(O, shape, model):
    import sbd_b200
    rng = np.random.default_rng(shape[0] + model)
    x = rng.uniform(0, 255, shape)
    eng = sbd_b200.Engine(shape[0], shape[1], 7, model, 0.0, max_batch=2)
    cl = O.operators.closures(model, shape, 7, 0.0)
    psi = PSI[model]
    xb = np.stack([x, x[::-1, ::-1].copy()])
    got = eng.blur(xb, psi, 0)
    assert rel(got[0], cl["A"](x, *psi)) < 1e-12
    assert rel(got[1], cl["A"](xb[1], *psi)) < 1e-12
    assert rel(eng.blur(x, psi, 1), cl["AT"](x, *psi)) < 1e-12
    assert rel(eng.blur(x, psi, 2), cl["dif"][0](x, *psi)) < 1e-12
    # one-pass likelihood: forward transform + sums against the pre-rotated Y', gradient by the inverse pass
    y = cl["A"](x, *psi) + 2.0 * rng.standard_normal(shape)
    xs = np.abs(x + 3.0 * rng.standard_normal(shape))
    psi2 = {0: (0.5, 0.35), 2: (0.15,)}[model]
    s2 = 7.5
    f, gradF, grads, gsig = O.operators.likelihood_closures(cl, y, xs.size)
    got = eng.likelihood(xs, y, psi2, s2, 0.04)
    args = (*psi2, s2)
    assert abs(got["f"] - f(xs, *args)) <= 1e-12 * abs(f(xs, *args))
    assert rel(got["gradF"], gradF(xs, *args)) < 1e-12
    for i, gr in enumerate(grads):
        scale = np.sum(np.abs(cl["dif"][i](xs, *psi2) * (cl["A"](xs, *psi2) - y))) / s2
        assert abs(got[f"grad_psi{i}"] - gr(xs, *args)) <= 1e-12 * scale
    assert abs(got["gradF_sigma"] - gsig(xs, *args)) <= 1e-11 * abs(gsig(xs, *args))
    eng.close()


def test_salsa_filter_mode_1024(O, cman):
    """COL_FILTER of the new column pass (x = invLS(r) with the residual against the pre-rotated Y') inside the
    SALSA MAP loop at 1024^2, against the oracle's SALSA_v2 (SALSA/SALSA_v2.m:428-450)."""
    import sbd_b200
    n = 1024
    x = np.tile(cman, (4, 4))
    rng = np.random.default_rng(5)
    cl = O.operators.closures(0, (n, n), 7, 0.0)
    psi = (0.42, 0.31)
    y = cl["A"](x, *psi) + 2.0 * rng.standard_normal((n, n))
    tau, mu = 0.03 * 4.0, 0.003
    H = cl["H_FFT"](*psi)
    filt = 1.0 / (np.abs(H) ** 2 + mu)
    from oracle import operators as OP, salsa
    invLS = lambda z: np.real(OP._ifft2(filt * OP._fft2(z)))
    want = salsa.SALSA_v2(y, lambda z: cl["A"](z, *psi), tau, "MU", mu, "AT", lambda z: cl["AT"](z, *psi),
                            "StopCriterion", 1, "True_x", x, "ToleranceA", 1e-5, "MAXITERA", 12, "Phi", O.tv.TVnorm,
                            "TVINITIALIZATION", 1, "TViters", 10, "LS", invLS, "VERBOSE", 0)
    eng = sbd_b200.Engine(n, n, 7, 0, 0.0, max_batch=1)
    r = eng.salsa_tv(y, psi, tau, mu, maxiter=12, tolA=1e-5, tv_iters=10, x_true=x)
    assert rel(r["x"], want[0]) < 1e-9
    assert rel(r["objective"], np.ravel(want[3])[:r["n_outer"] + 1]) < 1e-10
    eng.close()


def test_psf_size_other_than_7_keeps_the_tiled_passes(O):
    """psf_size 5 at 1024^2: Horner evaluation of the PSF spectrum, tiled passes of fft.cuh (the new ones are for the
    point-symmetric 7x7 form only)."""
    import sbd_b200
    shape = (1024, 1024)
    rng = np.random.default_rng(9)
    x = rng.uniform(0, 255, shape)
    eng = sbd_b200.Engine(shape[0], shape[1], 5, 0, 0.3, max_batch=1)
    cl = O.operators.closures(0, shape, 5, 0.3)
    assert rel(eng.blur(x, PSI[0], 0), cl["A"](x, *PSI[0])) < 1e-12
    assert rel(eng.blur(x, PSI[0], 3), cl["dif"][1](x, *PSI[0])) < 1e-12
    eng.close()


def test_tiled_passes_still_agree_with_the_new_ones():
    """SBD_FFT_V2=0 (environment, read at sbd_create) selects the tiled passes of fft.cuh at every size: both paths
    must produce the same operators and the same short SAPG run to rounding (run in a subprocess: the switch is read
    once per context from the environment)."""
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
import sbd_b200
from sbd_b200 import host as H
n = 1024
rng = np.random.default_rng(1)
x = rng.uniform(0, 255, (n, n))
eng = sbd_b200.Engine(n, n, 7, 1, 0.0, max_batch=2)
out = [eng.blur(x, (0.4, 3.5), op) for op in range(4)]
y = out[0] + rng.standard_normal((n, n))
lk = eng.likelihood(np.abs(x), y, (0.5, 4.0), 6.0, 0.05)
np.savez(sys.argv[1], *out, gradF=lk["gradF"], sc=np.array([lk[k] for k in ("f", "grad_psi0", "grad_psi1", "gradF_sigma", "g")]))
''' % ROOT
    import tempfile
    res = {}
    for v2 in ("1", "0"):
        with tempfile.NamedTemporaryFile(suffix=".npz") as f:
            env = dict(os.environ, SBD_FFT_V2=v2)
            r = subprocess.run([sys.executable, "-c", code, f.name], env=env, capture_output=True, text=True, timeout=600)
            assert r.returncode == 0, r.stderr[-2000:]
            res[v2] = dict(np.load(f.name))
    for k in res["1"]:
        a, b = res["1"][k], res["0"][k]
        if k == "sc":
            assert np.allclose(a, b, rtol=1e-11, atol=0), (a, b)
        else:
            assert rel(a, b) < 1e-12, (k, rel(a, b))
