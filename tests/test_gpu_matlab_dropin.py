"""Execute the MATLAB drop-in wrappers (sbd_b200/matlab/*.m: SAPG_algorithm_*,
chambolle_prox_TV_stop, TVnorm, PSF builders ...) on the GPU box through the
MATLAB-subset interpreter, with `sbd_mex` bridged to libsbd.so - the host side
stays MATLAB source, exactly what a MATLAB/Octave user would put on the path -
and compare with the golden fixtures of the reference execution."""
import os

import numpy as np
import pytest

from conftest import rel, GOLDEN, ROOT

pytestmark = pytest.mark.gpu
MATLAB_DIR = os.path.join(ROOT, "semi-blind-image-deblurring-problems-with-tv_b200", "matlab")


def sc(v):
    return float(np.asarray(v).ravel()[0])


@pytest.fixture(scope="module", params=["mex_gateway", "python_bridge"])
def it(request):
    """`mex_gateway`: sbd_mex is the REAL mexFunction of mex/sbd_mex.c, linked against libsbd.so and run on the minimal
    mxArray runtime of tests/mexrt (MATLAB/Octave are absent); `python_bridge`: the ctypes re-implementation."""
    import sbd_b200
    from oracle.mlab.interp import Interp
    interp = Interp([MATLAB_DIR])
    if request.param == "mex_gateway":
        import sys
        sys.path.insert(0, os.path.join(ROOT, "tests", "mexrt"))
        import runtime
        interp.builtins["sbd_mex"] = runtime.make_sbd_mex_real()
    else:
        from mex_bridge import make_sbd_mex
        interp.builtins["sbd_mex"] = make_sbd_mex(sbd_b200)
    return interp


def test_operator_wrappers(it):
    from oracle.mlab.interp import M
    G = dict(np.load(os.path.join(GOLDEN, "ref_operators.npz")))
    call = lambda n, *a, nargout=1: it.call_function(n, [x if isinstance(x, (str, np.ndarray)) else M(x) for x in a], nargout)
    sz = np.array([[16.0, 32.0]])
    assert rel(call("Gaussian_psf", 7, 0.4, 0.3, 0.0)[0], G["gauss_psf"]) < 1e-12
    assert rel(call("psf_moffat", 7, 0.4, 3.5)[0], G["moffat_psf"]) < 1e-12
    assert rel(call("psf_laplace", 7, 0.3)[0], G["laplace_psf"]) < 1e-12
    assert rel(call("diff_fftgaus_w1", sz, 7, 0.4, 0.3, 0.0)[0], G["gauss_dw1"]) < 1e-12
    assert rel(call("diff_moffat_beta", sz, 7, 0.4, 3.5)[0], G["moffat_db"]) < 1e-12
    assert rel(call("laplace_psf", sz, 7, 0.3)[0], G["laplace_H"]) < 1e-12
    x = G["tv_x"]
    assert abs(sc(call("TVnorm", x)[0]) - sc(G["tvnorm"])) <= 1e-12 * sc(G["tvnorm"])
    assert np.array_equal(call("diffh", x)[0], G["diffh"]) and np.array_equal(call("diffv", x)[0], G["diffv"])
    f, px, py = call("chambolle_prox_TV_stop", x, "lambda", 0.5, "maxiter", 25, nargout=3)
    assert rel(f, G["chamb1_f"]) < 1e-12 and rel(px, G["chamb1_px"]) < 1e-11
    f, px, py = call("chambolle_prox_TV_stop", G["chamb_opt_g"], "LAMBDA", 0.7, "MaxIter", 10, "tol", 1e-2, "tau", 0.2,
                     "dualvars", G["chamb_opt_dual"], nargout=3)
    assert rel(f, G["chamb_opt_f"]) < 1e-12 and rel(py, G["chamb_opt_py"]) < 1e-11
    from oracle.mlab.interp import MatlabError
    with pytest.raises(MatlabError):                # 'maxiter' omitted -> MaxIter undefined, like the reference (Q4)
        call("chambolle_prox_TV_stop", x, "lambda", 0.5)


DRIVER = {
    "gaussian": """
op.samples = 16; op.warmup = 6; op.burnIn = 12; op.psf_size = 7; op.phi = 0;
op.min_th = 1e-3; op.max_th = 1; op.th_init = 0.01; op.d_exp = 0.8; op.d_scale = 0.01/op.th_init;
op.w1_init = 0.5; op.w2_init = 0.3; op.min_w1 = 0.1; op.max_w1 = 1; op.min_w2 = 0.1; op.max_w2 = 1;
op.w1 = 0.4; op.w2 = 0.3; op.fix_w1 = 0; op.fix_w2 = 0; op.fix_sigma = 0;
c.sigma = 1000; c.theta = 0.01; c.w1 = 10; c.w2 = 10; c.lam = 1; c.gam = 1;
[theta_EB, w1_EB, w2_EB, sigma_EB, results] = SAPG_algorithm_Guassian(y, op, c);
""",
    "moffat": """
op.samples = 16; op.warmup = 6; op.burnIn = 12; op.psf_size = 7; op.sub_sample = 1;
op.min_th = 1e-3; op.max_th = 1; op.th_init = 0.01; op.d_exp = 0.8; op.d_scale = 0.01/op.th_init;
op.alpha_init = 1; op.beta_init = 10; op.min_alpha = 1e-2; op.max_alpha = 1; op.min_beta = 0.1; op.max_beta = 10;
op.alpha = 0.4; op.beta = 3.5; op.fix_alpha = 0; op.fix_beta = 0; op.fix_sigma = 0;
[theta_EB, alpha_EB, beta_EB, sigma2_EB, results] = SAPG_algorithm_moffat(y, op);
""",
    "laplace": """
op.samples = 16; op.warmup = 6; op.burnIn = 12; op.psf_size = 7; op.warm_sample = 1;
op.min_th = 1e-3; op.max_th = 1; op.th_init = 0.01; op.d_exp = 0.8; op.d_scale = 0.01/op.th_init;
op.b_init = 0.1; op.min_b = 1e-3; op.max_b = 1; op.b = 0.3; op.fix_b = 0; op.fix_sigma = 0;
op.x = x; op.X0 = y;
[theta_EB, b_EB, sigma_EB, results] = SAPG_algorithm_laplace(y, op);
""",
}


@pytest.mark.parametrize("name", ["gaussian", "moffat", "laplace"])
def test_sapg_wrappers_reproduce_the_reference(it, name):
    from oracle.mlab.interp import MStruct, M, to_py
    g = dict(np.load(os.path.join(GOLDEN, f"ref_sapg_{name}.npz")))
    rng = np.random.default_rng(int(sc(g["seed"])))
    shape = g["x"].shape
    rng.standard_normal(shape); rng.standard_normal(shape)          # power iteration + observation noise
    tape = np.stack([rng.standard_normal(shape) for _ in range(5 + 15)])[:, None]
    op = MStruct()
    for k in ("sigma", "sigma_init", "sigma_min", "sigma_max", "lambda", "gamma"):
        op[k] = M(sc(g["op_" + k]))
    op["noise"] = tape                                              # explicit randn stream (see the wrapper)
    scope = {"y": g["y"], "x": g["x"], "op": op}
    it.run_source(DRIVER[name], scope)
    r = to_py(scope["results"])
    checked = 0
    for k, want in g.items():
        if not k.startswith("res_"):
            continue
        f = k[4:]
        if f.startswith("execTime"):
            continue
        assert f in r, f"results lack field {f}"
        got = np.asarray(r[f], dtype=np.float64); want = np.asarray(want, dtype=np.float64)
        assert got.size == want.size, (f, got.shape, want.shape)
        got = got.reshape(want.shape)
        assert np.array_equal(np.isnan(got), np.isnan(want)), f
        m = ~np.isnan(want)
        if f.startswith("tol_"):
            assert np.allclose(got[m], want[m], rtol=1e-4, atol=1e-14), f
        else:
            assert rel(got[m], want[m]) < 1e-6, (f, rel(got[m], want[m]))
        checked += 1
    assert checked >= 25
    ref_out = {"gaussian": ("theta_EB", "w1_EB", "w2_EB", "sigma_EB"), "moffat": ("theta_EB", "alpha_EB", "beta_EB", "sigma2_EB"),
               "laplace": ("theta_EB", "b_EB", "sigma_EB")}[name]
    for o in ref_out:
        assert np.isfinite(sc(scope[o]))


def test_setup_and_map_wrappers(it):
    """sbd_max_eigenval.m / sbd_observe.m / sbd_salsa_map.m executed as MATLAB source on the GPU box."""
    from oracle.mlab.interp import M
    g = dict(np.load(os.path.join(GOLDEN, "ref_sapg_gaussian.npz")))
    rng = np.random.default_rng(int(sc(g["seed"])))
    x0 = rng.standard_normal(g["x"].shape); noise = rng.standard_normal(g["x"].shape)
    scope = {"x": g["x"], "x0": x0, "noise": noise}
    it.run_source("""
im_size = size(x);
evMax = sbd_max_eigenval(0, im_size, 7, 0, [1 1], 1e-4, 1e4, x0);
[y, sigma, nrm] = sbd_observe(x, 0, 7, 0, [0.4 0.3], 30, noise);
""", scope)
    assert abs(sc(scope["evMax"]) - sc(g["evMax"])) <= 1e-11 * sc(g["evMax"])
    assert abs(sc(scope["sigma"]) - sc(g["op_sigma"])) <= 1e-12 * sc(g["op_sigma"])
    assert rel(scope["y"], g["y"]) < 1e-12
    s = dict(np.load(os.path.join(GOLDEN, "ref_salsa_gaussian.npz")))
    scope = {"y": s["y"], "x": s["x"], "theta_EB": M(sc(s["theta_EB"])), "w1_EB": M(sc(s["w1_EB"])),
             "w2_EB": M(sc(s["w2_EB"])), "sigma_EB": M(sc(s["sigma_EB"]))}
    it.run_source("""
mu = theta_EB/10;
[xMAP, objective, distance, mses, n_outer] = sbd_salsa_map(y, 0, 7, 0, [w1_EB w2_EB], theta_EB*sigma_EB, mu, 500, 1e-5, 10, x);
mse = 10*log10(norm(x-xMAP,'fro')^2 / numel(x));
""", scope)
    assert rel(scope["xMAP"], s["xMAP"]) < 1e-9
    assert rel(np.ravel(scope["objective"]), np.ravel(s["objective"])) < 1e-10
    assert abs(sc(scope["mse"]) - sc(s["mse"])) < 1e-6
