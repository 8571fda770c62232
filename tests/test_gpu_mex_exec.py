"""The MEX gateway EXECUTED on the GPU box: mex/sbd_mex.c is linked against libsbd.so and the minimal mxArray runtime
of tests/mexrt (MATLAB / Octave are absent from this image), and every command of `mexFunction` is driven through
ctypes and compared with the goldens produced by executing the reference's .m files, or with the oracle."""
import os
import sys

import numpy as np
import pytest

from conftest import rel, GOLDEN, ROOT

pytestmark = pytest.mark.gpu


def sc(v):
    return float(np.asarray(v).ravel()[0])


@pytest.fixture(scope="module")
def mex():
    sys.path.insert(0, os.path.join(ROOT, "tests", "mexrt"))
    import runtime
    return runtime


def test_tv_psf_blur_commands(mex):
    import oracle as O
    G = dict(np.load(os.path.join(GOLDEN, "ref_operators.npz")))
    x = G["tv_x"]
    assert abs(sc(mex.call_mex("tvnorm", x)[0]) - sc(G["tvnorm"])) <= 1e-12 * sc(G["tvnorm"])
    assert np.array_equal(mex.call_mex("diff", x, 1)[0], G["diffh"])
    assert np.array_equal(mex.call_mex("diff", x, 0)[0], G["diffv"])
    f, px, py, k, err = mex.call_mex("tvprox", x, 0.5, 25, 1e-3, 0.249, nargout=5)
    assert rel(f, G["chamb1_f"]) < 1e-12 and rel(px, G["chamb1_px"]) < 1e-11 and rel(py, G["chamb1_py"]) < 1e-11
    g = G["chamb_opt_g"]; dual = G["chamb_opt_dual"]; n = g.shape[1]
    f, px, py = mex.call_mex("tvprox", g, 0.7, 10, 1e-2, 0.2, dual[:, :n], dual[:, n:], nargout=3)
    assert rel(f, G["chamb_opt_f"]) < 1e-12 and rel(py, G["chamb_opt_py"]) < 1e-11
    with pytest.raises(mex.MexError):                               # wrong dual size (chambolle_prox_TV_stop.m:102-104)
        mex.call_mex("tvprox", g, 0.7, 10, 1e-2, 0.2, dual[:, :n - 1], dual[:, n:])
    assert rel(mex.call_mex("psf", 0, 7, 0.0, [0.4, 0.3], 0)[0], G["gauss_psf"]) < 1e-12
    assert rel(mex.call_mex("psf", 1, 7, 0.0, [0.4, 3.5], 0)[0], G["moffat_psf"]) < 1e-12
    assert rel(mex.call_mex("spectrum", [16, 32], 0, 7, 0.0, [0.4, 0.3], 1)[0], G["gauss_dw1"]) < 1e-12
    assert rel(mex.call_mex("spectrum", [16, 32], 2, 7, 0.0, [0.3], 0)[0], G["laplace_H"]) < 1e-12
    rng = np.random.default_rng(1)
    xs = rng.uniform(0, 255, (64, 32))
    for model, psi in ((0, (0.4, 0.3)), (1, (0.4, 3.5)), (2, (0.3,))):
        cl = O.operators.closures(model, xs.shape, 7, 0.0)
        assert rel(mex.call_mex("blur", xs, model, 7, 0.0, list(psi), 0)[0], cl["A"](xs, *psi)) < 1e-12
        assert rel(mex.call_mex("blur", xs, model, 7, 0.0, list(psi), 1)[0], cl["AT"](xs, *psi)) < 1e-12
        assert rel(mex.call_mex("blur", xs, model, 7, 0.0, list(psi), 2)[0], cl["dif"][0](xs, *psi)) < 1e-12


def test_likelihood_command(mex, cman):
    import oracle as O
    shape = (256, 256)
    rng = np.random.default_rng(0)
    cl = O.operators.closures(0, shape, 7, 0.0)
    y = cl["A"](cman, 0.4, 0.3) + 2.0 * rng.standard_normal(shape)
    x = np.abs(cman + 3.0 * rng.standard_normal(shape))
    f, gradF, grads, gsig = O.operators.likelihood_closures(cl, y, x.size)
    s, gf = mex.call_mex("likelihood", x, y, 0, 7, 0.0, [0.5, 0.35], 7.5, 0.04, nargout=2)
    s = np.ravel(s)
    args = (0.5, 0.35, 7.5)
    assert abs(s[0] - f(x, *args)) <= 1e-12 * abs(f(x, *args))
    assert rel(gf, gradF(x, *args)) < 1e-12
    assert abs(s[3] - gsig(x, *args)) <= 1e-11 * abs(gsig(x, *args))
    assert abs(s[4] - O.tv.TVnorm(x)) <= 1e-12 * O.tv.TVnorm(x)


@pytest.mark.parametrize("name,model", [("gaussian", 0), ("moffat", 1), ("laplace", 2)])
def test_sapg_command_reproduces_the_reference_execution(mex, name, model):
    """'sapg' through mexFunction, fed the P struct that matlab/sbd_pack.m would build, against the goldens."""
    g = dict(np.load(os.path.join(GOLDEN, f"ref_sapg_{name}.npz")))
    rng = np.random.default_rng(int(sc(g["seed"])))
    shape = g["x"].shape
    rng.standard_normal(shape); rng.standard_normal(shape)
    tape = np.stack([rng.standard_normal(shape) for _ in range(5 + 15)])[:, None]
    op = {k[3:]: sc(v) for k, v in g.items() if k.startswith("op_") and np.asarray(v).size == 1}
    names = {0: ("w1", "w2"), 1: ("alpha", "beta"), 2: ("b",)}[model]
    true = {"w1": 0.4, "w2": 0.3, "alpha": 0.4, "beta": 3.5, "b": 0.3}
    init = {"w1": 0.5, "w2": 0.3, "alpha": 1.0, "beta": 10.0, "b": 0.1}
    lo = {"w1": 0.1, "w2": 0.1, "alpha": 1e-2, "beta": 0.1, "b": 1e-3}
    hi = {"w1": 1.0, "w2": 1.0, "alpha": 1.0, "beta": 10.0, "b": 1.0}
    pad = lambda d: [d[n] for n in names] + [0.0] * (2 - len(names))
    P = dict(samples=16, warmup=6, burnIn=12, n_chains=1, prox_lambda=op["lambda"], chambolle_maxiter=25,
             chambolle_tol=1e-3, chambolle_tau=0.249, th_init=0.01, min_th=1e-3, max_th=1.0,
             psi_init=pad(init), psi_min=pad(lo), psi_max=pad(hi), psi_fixed=pad(true), psi_true=pad(true),
             fix_psi=[0.0, 0.0], sigma2_init=op["sigma_init"], sigma2_min=op["sigma_min"], sigma2_max=op["sigma_max"],
             fix_sigma=0, d_scale=1.0, d_exp=0.8, seed=1)
    if model == 0:
        P.update(gam=op["gamma"], lamb=op["lambda"], c_theta=0.01, c_sigma2=1000.0, c_psi=[10.0, 10.0],
                 sigma2_fixed=op["sigma_init"], err_psf_lag=1)
    elif model == 1:
        P.update(gam=op["gamma"], lamb=op["lambda"], c_theta=0.1, c_sigma2=10000.0, c_psi=[10.0, 10000.0],
                 sigma2_fixed=op["sigma"] ** 2, err_psf_lag=0)
    else:
        P.update(gam=op["gamma"], lamb=op["lambda"], c_theta=0.01, c_sigma2=10000.0, c_psi=[100.0, 0.0],
                 sigma2_fixed=op["sigma"] ** 2, err_psf_lag=0)
    X0 = g["y"] if model == 2 else None
    xt = g["x"] if model == 2 else None
    r = mex.call_mex("sapg", g["y"], X0, xt, model, 7, 0.0, P, tape)[0]
    pairs = {0: [("thetas", "thetas"), ("psi0", "w1s"), ("psi1", "w2s"), ("sigmas", "sigmas"), ("grad_psi0", "grad_w1"),
                 ("logPiTraceX", "logPiTraceX"), ("gXTrace", "gXTrace"), ("err_psf", "err_psf"), ("X_last", "Xlast_sample")],
             1: [("thetas", "thetas"), ("psi0", "alphas"), ("psi1", "betas"), ("sigmas", "sigmas"),
                 ("logPiTraceX", "logPiTraceX"), ("X_last", "Xlast_sample"), ("X_warm", "X_warm")],
             2: [("thetas", "thetas"), ("psi0", "bs"), ("sigmas", "sigmas"), ("logPiTraceX", "logPiTraceX"),
                 ("err_sample", "err_sample"), ("X_last", "X_sample")]}[model]
    for mine, ref in pairs:
        want = np.asarray(g["res_" + ref], dtype=np.float64)
        got = np.asarray(r[mine], dtype=np.float64).reshape(want.shape)
        assert rel(got, want) < 1e-6, (mine, rel(got, want))
    eb = np.ravel(r["EB"])
    assert abs(eb[0] - sc(g["res_theta_EB" if model == 0 else "res_mean_theta"])) <= 1e-6 * abs(eb[0])
    assert sc(r["last_samp"]) == 16


def test_setup_and_map_commands(mex):
    g = dict(np.load(os.path.join(GOLDEN, "ref_sapg_gaussian.npz")))
    rng = np.random.default_rng(int(sc(g["seed"])))
    x0 = rng.standard_normal(g["x"].shape); noise = rng.standard_normal(g["x"].shape)
    v, k = mex.call_mex("max_eigenval", list(g["x"].shape), 0, 7, 0.0, [1.0, 1.0], 1e-4, 10000, x0, nargout=2)
    assert abs(sc(v) - sc(g["evMax"])) <= 1e-11 * sc(g["evMax"]) and sc(k) >= 1
    y, sg, nr = mex.call_mex("observe", g["x"], 0, 7, 0.0, [0.4, 0.3], 30, noise, nargout=3)
    assert rel(y, g["y"]) < 1e-12 and abs(sc(sg) - sc(g["op_sigma"])) <= 1e-12 * sc(g["op_sigma"])
    s = dict(np.load(os.path.join(GOLDEN, "ref_salsa_gaussian.npz")))
    th, s2 = sc(s["theta_EB"]), sc(s["sigma_EB"])
    x, obj, dist, mses, n = mex.call_mex("salsa", s["y"], 0, 7, 0.0, [sc(s["w1_EB"]), sc(s["w2_EB"])], th * s2, th / 10,
                                         500, 1e-5, 10, s["x"], nargout=5)
    assert rel(x, s["xMAP"]) < 1e-9
    k = int(sc(n))
    assert rel(np.ravel(obj)[:k + 1], np.ravel(s["objective"])) < 1e-10
