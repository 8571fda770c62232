"""GPU parity: deterministic operators (PSF builders, A/A'/dif, TV pieces) vs the
oracle.  Tolerance: relative error 1e-12 (BASELINE.json north_star), written in
each assert.  Everything goes through the C ABI (sbd_b200 -> libsbd.so)."""
import numpy as np
import pytest

from conftest import rel, kat_image

pytestmark = pytest.mark.gpu

TOL = 1e-12

PSI = {0: (0.4, 0.3), 1: (0.4, 3.5), 2: (0.3,)}


@pytest.fixture(scope="module")
def sbd():
    import sbd_b200
    return sbd_b200


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


# ---------------------------------------------------------------- PSF
@pytest.mark.parametrize("model", [0, 1, 2])
def test_psf_taps(sbd, O, model):
    eng = sbd.engine_for((32, 32), 7, model, 0.0)
    nk = 2 if model == 2 else 3
    for which in range(nk):
        got = eng.psf_taps(PSI[model], which)
        want = O.psf.taps(model, 7, PSI[model], 0.0, which)
        assert rel(got, want) < TOL, (model, which)
    assert abs(eng.psf_taps(PSI[model], 0).sum() - 1.0) < 1e-14       # Gaussian_psf.m:18


def test_psf_taps_rotated_gaussian(sbd, O):
    eng = sbd.engine_for((32, 32), 7, 0, 0.3)
    for which in range(3):
        assert rel(eng.psf_taps((0.5, 0.2), which), O.psf.taps(0, 7, (0.5, 0.2), 0.3, which)) < TOL


def test_mirror_names(sbd, O):
    assert rel(sbd.Gaussian_psf(7, 0.4, 0.3, 0.0), O.psf.Gaussian_psf(7, 0.4, 0.3, 0.0)) < TOL
    assert rel(sbd.psf_moffat(7, 0.4, 3.5), O.psf.psf_moffat(7, 0.4, 3.5)) < TOL
    assert rel(sbd.psf_laplace(7, 0.3), O.psf.psf_laplace(7, 0.3)) < TOL
    assert rel(sbd.moffat_psf((64, 32), 7, 0.4, 3.5), O.psf.moffat_psf((64, 32), 7, 0.4, 3.5)) < TOL
    assert rel(sbd.diff_laplace_b((32, 64), 7, 0.3), O.psf.diff_laplace_b((32, 64), 7, 0.3)) < TOL
    assert rel(sbd.diff_fftgaus_w1((64, 64), 7, 0.4, 0.3, 0.0), O.psf.diff_fftgaus_w1((64, 64), 7, 0.4, 0.3, 0.0)) < TOL
    assert rel(sbd.diff_moffat_alpha((64, 64), 7, 0.4, 3.5), O.psf.diff_moffat_alpha((64, 64), 7, 0.4, 3.5)) < TOL


@pytest.mark.parametrize("model", [0, 1, 2])
@pytest.mark.parametrize("shape", [(16, 16), (64, 128), (256, 256)])
def test_psf_spectrum(sbd, O, model, shape):
    eng = sbd.engine_for(shape, 7, model, 0.0)
    nk = 2 if model == 2 else 3
    for which in range(nk):
        got = eng.psf_spectrum(PSI[model], which)
        want = O.psf.resize(O.psf.taps(model, 7, PSI[model], 0.0, which), shape)
        assert rel(got, want) < TOL


# ---------------------------------------------------------------- blur operators
@pytest.mark.parametrize("model", [0, 1, 2])
@pytest.mark.parametrize("shape", [(16, 16), (32, 64), (128, 64), (256, 256), (512, 512), (1024, 1024)])
def test_blur_ops(sbd, O, model, shape):
    rng = np.random.default_rng(shape[0] * 7 + model)
    x = rng.uniform(0, 255, shape)
    cl_gpu = sbd.host._closures(model, shape, 7, 0.0)
    cl = O.operators.closures(model, shape, 7, 0.0)
    psi = PSI[model]
    assert rel(cl_gpu["A"](x, *psi), cl["A"](x, *psi)) < TOL
    assert rel(cl_gpu["AT"](x, *psi), cl["AT"](x, *psi)) < TOL
    for d_gpu, d in zip(cl_gpu["dif"], cl["dif"]):
        assert rel(d_gpu(x, *psi), d(x, *psi)) < TOL


@pytest.mark.parametrize("psf_size", [7, 5, 9])
@pytest.mark.parametrize("phi", [0.3, np.pi / 5])
def test_blur_ops_rotated_gaussian_and_other_sizes(sbd, O, psf_size, phi):
    """phi != 0: the Gaussian taps are only point-symmetric; psf_size != 7 takes the Horner path."""
    shape = (64, 128)
    rng = np.random.default_rng(psf_size)
    x = rng.uniform(0, 255, shape)
    cl_gpu = sbd.host._closures(0, shape, psf_size, phi)
    cl = O.operators.closures(0, shape, psf_size, phi)
    psi = PSI[0]
    assert rel(cl_gpu["A"](x, *psi), cl["A"](x, *psi)) < TOL
    assert rel(cl_gpu["AT"](x, *psi), cl["AT"](x, *psi)) < TOL
    for d_gpu, d in zip(cl_gpu["dif"], cl["dif"]):
        assert rel(d_gpu(x, *psi), d(x, *psi)) < TOL
    # the fused likelihood pass (residual + both PSF-parameter gradients in one column pass)
    eng = sbd.engine_for(shape, psf_size, 0, phi)
    y = cl["A"](x, *psi) + rng.standard_normal(shape)
    psi2, s2 = (0.5, 0.35), 3.0
    f, gradF, grads, gsig = O.operators.likelihood_closures(cl, y, x.size)
    got = eng.likelihood(x, y, psi2, s2, 0.0)
    args = (*psi2, s2)
    assert abs(got["f"] - f(x, *args)) <= TOL * abs(f(x, *args))
    assert rel(got["gradF"], gradF(x, *args)) < TOL
    for i, gr in enumerate(grads):
        scale = np.sum(np.abs(cl["dif"][i](x, *psi2) * (cl["A"](x, *psi2) - y))) / s2
        assert abs(got[f"grad_psi{i}"] - gr(x, *args)) <= TOL * scale


@pytest.mark.parametrize("n", [2048, 4096])
def test_blur_large_properties(sbd, n):
    """Full-size, size-independent properties: A(const)=const, adjointness,
    linearity (the oracle is too slow to be the checker at these sizes)."""
    rng = np.random.default_rng(n)
    eng = sbd.engine_for((n, n), 7, 0, 0.0)
    psi = (0.4, 0.3)
    c = np.full((n, n), 3.25)
    assert rel(eng.blur(c, psi, 0), c) < TOL                          # PSF sums to 1
    x = rng.standard_normal((n, n)); z = rng.standard_normal((n, n))
    Ax = eng.blur(x, psi, 0); ATz = eng.blur(z, psi, 1)
    lhs = float(np.vdot(Ax, z)); rhs = float(np.vdot(x, ATz))
    assert abs(lhs - rhs) <= 1e-11 * (np.linalg.norm(Ax) * np.linalg.norm(z))
    y2 = eng.blur(2.0 * x - 0.5 * z, psi, 0)
    assert rel(y2, 2.0 * Ax - 0.5 * eng.blur(z, psi, 0)) < 1e-11


def test_blur_batch(sbd, O):
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 255, (5, 64, 64))
    eng = sbd.Engine(64, 64, 7, 1, 0.0, max_batch=5)
    got = eng.blur(x, (0.4, 3.5), 0)
    cl = O.operators.closures(1, (64, 64), 7)
    for b in range(5):
        assert rel(got[b], cl["A"](x[b], 0.4, 3.5)) < TOL
    eng.close()


# ---------------------------------------------------------------- TV pieces
@pytest.mark.parametrize("shape", [(2, 2), (3, 5), (37, 53), (64, 100), (100, 64), (256, 256), (130, 70)])
def test_tvnorm_and_diffs(sbd, O, shape):
    rng = np.random.default_rng(shape[0] + shape[1])
    x = rng.uniform(0, 255, shape)
    assert abs(sbd.TVnorm(x) - O.tv.TVnorm(x)) <= TOL * O.tv.TVnorm(x)
    assert np.array_equal(sbd.diffh(x), O.tv.diffh(x))                 # exact: one subtraction
    assert np.array_equal(sbd.diffv(x), O.tv.diffv(x))


def test_tvnorm_cman(sbd, O, cman):
    want = O.tv.TVnorm(cman)
    assert abs(want - 1115956.0628163717) < 1e-6                       # SURVEY.md 8c scratch value
    assert abs(sbd.TVnorm(cman) - want) <= TOL * want


@pytest.mark.parametrize("shape", [(2, 2), (5, 5), (37, 53), (64, 100), (128, 128), (256, 256), (130, 70)])
@pytest.mark.parametrize("lam", [1e-3, 0.1, 2.0])
def test_chambolle(sbd, O, shape, lam):
    rng = np.random.default_rng(shape[0] * 3 + shape[1])
    g = rng.uniform(0, 255, shape)
    f, px, py = sbd.chambolle_prox_TV_stop(g, "lambda", lam, "maxiter", 25)
    fo, pxo, pyo, k, err = O.tv.chambolle_prox_TV_stop(g, "lambda", lam, "maxiter", 25, return_info=True)
    assert rel(f, fo) < TOL
    assert rel(px, pxo) < 1e-11 and rel(py, pyo) < 1e-11
    eng = sbd.host._tv_engine(shape)
    _, _, _, kg, errg = eng.tvprox(g, lam, 25)
    assert kg == k
    # err is a norm of differences (-ux + |grad u| px): cancellation limits its relative accuracy
    assert abs(errg - err) <= 1e-7 * err + 1e-10


def test_chambolle_cman_and_kat(sbd, O, cman):
    eng = sbd.host._tv_engine((256, 256))
    f, _, _, k, err = eng.tvprox(cman, 1.0, 25)
    assert k == 25 and abs(err - 4.412657548137149) < 1e-9
    assert abs(O.tv.TVnorm(f) - 959256.7506108371) < 1e-5
    g = kat_image(256)                                                 # early stop: k = 20
    fo, _, _, ko, erro = O.tv.chambolle_prox_TV_stop(g, "lambda", 1e-3, "maxiter", 25, return_info=True)
    f, _, _, k, err = eng.tvprox(g, 1e-3, 25)
    assert ko == 20 and k == 20
    assert abs(err - erro) <= 1e-7 * erro
    assert rel(f, fo) < TOL


def test_chambolle_constant_and_options(sbd, O):
    eng = sbd.host._tv_engine((64, 64))
    c = np.full((64, 64), 7.0)
    f, px, py, k, err = eng.tvprox(c, 0.5, 25)
    assert k == 1 and err == 0.0 and np.array_equal(f, c)
    rng = np.random.default_rng(3)
    g = rng.uniform(0, 255, (64, 64))
    dual = rng.uniform(-0.5, 0.5, (64, 128))
    got = sbd.chambolle_prox_TV_stop(g, "LAMBDA", 0.7, "MaxIter", 10, "tol", 1e-2, "tau", 0.2, "dualvars", dual)
    want = O.tv.chambolle_prox_TV_stop(g, "LAMBDA", 0.7, "MaxIter", 10, "tol", 1e-2, "tau", 0.2, "dualvars", dual)
    for a, b in zip(got, want):
        assert rel(a, b) < 1e-11
    with pytest.raises(NameError):                                     # Q4: maxiter omitted
        sbd.chambolle_prox_TV_stop(g, "lambda", 0.7)
    with pytest.raises(ValueError):                                    # wrong dual size (:102-104)
        sbd.chambolle_prox_TV_stop(g, "lambda", 0.7, "maxiter", 3, "dualvars", np.zeros((64, 64)))


def test_chambolle_batch_independent_stops(sbd, O):
    """Images of one batch stop on their own err<=tol test."""
    rng = np.random.default_rng(11)
    imgs = np.stack([kat_image(128), rng.uniform(0, 255, (128, 128)), np.full((128, 128), 2.0)])
    eng = sbd.Engine(128, 128, 1, 0, 0.0, max_batch=3)
    f, px, py, it, err = eng.tvprox(imgs, 1e-3, 25)
    for b in range(3):
        fo, _, _, ko, eo = O.tv.chambolle_prox_TV_stop(imgs[b], "lambda", 1e-3, "maxiter", 25, return_info=True)
        assert it[b] == ko
        assert rel(f[b], fo) < TOL
    assert len(set(it.tolist())) > 1
    eng.close()


@pytest.mark.parametrize("shape", [(128, 128), (130, 70)])
def test_chambolle_stop_inside_every_block_position(sbd, O, shape):
    """The last fused block writes the prox output itself; a stop test firing before, inside or exactly
    at the end of that block must still give the reference's f, sweep count and dual pair."""
    rng = np.random.default_rng(5)
    g = rng.uniform(0, 255, shape)
    eng = sbd.host._tv_engine(shape)
    for kstop in (16, 19, 20, 21, 22, 23, 24, 25):
        _, _, _, _, e_k = O.tv.chambolle_prox_TV_stop(g, "lambda", 0.3, "maxiter", kstop, "tol", 0.0, return_info=True)
        tol = e_k * (1 + 1e-9)                       # err_k <= tol < err_(k-1): stops after sweep kstop
        fo, pxo, pyo, ko, eo = O.tv.chambolle_prox_TV_stop(g, "lambda", 0.3, "maxiter", 25, "tol", tol, return_info=True)
        assert ko == kstop
        f, px, py, k, err = eng.tvprox(g, 0.3, 25, tol=tol)
        assert k == ko and rel(f, fo) < TOL and rel(px, pxo) < 1e-11 and rel(py, pyo) < 1e-11


@pytest.mark.parametrize("shape", [(64, 64), (256, 256), (130, 70)])
def test_chambolle_device_entry_without_duals(sbd, O, shape):
    """sbd_tvprox_dev (device pointers, dual pair not kept): the path the SAPG loop uses."""
    import ctypes as C
    import torch
    from sbd_b200._lib import lib
    rng = np.random.default_rng(9)
    g = rng.uniform(0, 255, (2,) + shape)
    eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=2)
    # column-major images on the device = transposed C-order tensors
    gd = torch.from_numpy(np.ascontiguousarray(g.transpose(0, 2, 1))).cuda()
    fd = torch.empty_like(gd)
    it = (C.c_int * 2)()
    er = (C.c_double * 2)()
    for maxiter in (25, 20, 7):
        rc = lib.sbd_tvprox_dev(eng._h, gd.data_ptr(), 0.8, maxiter, 1e-3, 0.249, fd.data_ptr(), it, er, 2)
        assert rc == 0, lib.sbd_last_error(eng._h)
        lib.sbd_synchronize(eng._h)
        f = fd.cpu().numpy().transpose(0, 2, 1)
        for b in range(2):
            fo, _, _, ko, eo = O.tv.chambolle_prox_TV_stop(g[b], "lambda", 0.8, "maxiter", maxiter, "tol", 1e-3, return_info=True)
            assert it[b] == ko and rel(f[b], fo) < TOL
    eng.close()


@pytest.mark.parametrize("n", [4096])
def test_chambolle_large_properties(sbd, n):
    eng = sbd.host._tv_engine((n, n))
    c = np.full((n, n), 11.0)
    f, _, _, k, err = eng.tvprox(c, 1.0, 20)
    assert k == 1 and err == 0.0 and np.array_equal(f, c)
    rng = np.random.default_rng(n)
    g = rng.uniform(0, 255, (n, n))
    f, px, py, k, err = eng.tvprox(g, 1e-9, 20)                        # prox of lambda -> 0 is the identity
    assert rel(f, g) < 1e-9
    assert np.all(px * px + py * py <= 1.0 + 1e-12)                    # dual feasibility |p| <= 1


# ---------------------------------------------------------------- likelihood closures
@pytest.mark.parametrize("model", [0, 1, 2])
def test_likelihood(sbd, O, model, cman):
    shape = (256, 256)
    rng = np.random.default_rng(model)
    cl = O.operators.closures(model, shape, 7, 0.0)
    psi_true = PSI[model]
    y = cl["A"](cman, *psi_true) + 2.0 * rng.standard_normal(shape)
    x = np.abs(cman + 3.0 * rng.standard_normal(shape))
    psi = {0: (0.5, 0.35), 1: (0.8, 6.0), 2: (0.15,)}[model]
    s2, th = 7.5, 0.04
    f, gradF, grads, gsig = O.operators.likelihood_closures(cl, y, x.size)
    eng = sbd.engine_for(shape, 7, model, 0.0)
    got = eng.likelihood(x, y, psi, s2, th)
    args = (*psi, s2)
    assert abs(got["f"] - f(x, *args)) <= TOL * abs(f(x, *args))
    assert rel(got["gradF"], gradF(x, *args)) < TOL
    for i, gr in enumerate(grads):
        want = gr(x, *args)
        # a sum of ~6e4 signed terms: compare against the scale of the sum of magnitudes
        scale = np.sum(np.abs(cl["dif"][i](x, *psi) * (cl["A"](x, *psi) - y))) / s2
        assert abs(got[f"grad_psi{i}"] - want) <= TOL * scale
    assert abs(got["gradF_sigma"] - gsig(x, *args)) <= 1e-11 * abs(gsig(x, *args))
    assert abs(got["g"] - O.tv.TVnorm(x)) <= TOL * O.tv.TVnorm(x)
    want_lp = -f(x, *args) - th * O.tv.TVnorm(x)
    assert abs(got["logPi"] - want_lp) <= TOL * abs(want_lp)
