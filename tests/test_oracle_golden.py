"""Pin the numpy oracle against the golden fixtures produced by EXECUTING THE
REFERENCE'S OWN .m FILES (tests/golden/make_golden.py, MATLAB-subset interpreter
in oracle/mlab/).  No GPU needed.  Tolerances are at rounding level: both sides
are IEEE double and differ only in summation order."""
import os

import numpy as np
import pytest

from conftest import rel, GOLDEN
import oracle
from oracle import psf as P, tv, operators as OP, sapg, metrics

RT = 1e-12          # arrays: relative l2 error


def sc(v):
    return float(np.asarray(v).ravel()[0])


@pytest.fixture(scope="module")
def G():
    return dict(np.load(os.path.join(GOLDEN, "ref_operators.npz")))


def test_psf_builders(G):
    assert rel(P.Gaussian_psf(7, 0.4, 0.3, 0.0), G["gauss_psf"]) < RT
    assert rel(P.psf_gaussian(7, 0.5, 0.2, 0.3), G["gauss_psf_rot"]) < RT
    assert rel(P.resize(G["gauss_psf"], (16, 32)), G["gauss_H"]) < RT
    assert rel(P.diff_fftgaus_w1((16, 32), 7, 0.4, 0.3, 0.0), G["gauss_dw1"]) < RT
    assert rel(P.diff_fftgaus_w2((16, 32), 7, 0.4, 0.3, 0.0), G["gauss_dw2"]) < RT
    assert rel(np.array(P.Sum_gauss_psf(7, 0.4, 0.3, 0.0)), G["gauss_sums"]) < RT
    assert rel(P.psf_moffat(7, 0.4, 3.5), G["moffat_psf"]) < RT
    assert rel(P.moffat_psf((16, 32), 7, 0.4, 3.5), G["moffat_H"]) < RT
    assert rel(P.diff_moffat_alpha((16, 32), 7, 0.4, 3.5), G["moffat_da"]) < RT
    assert rel(P.diff_moffat_beta((16, 32), 7, 0.4, 3.5), G["moffat_db"]) < RT
    assert rel(np.array(P.sum_mof_psf(7, 0.4, 3.5)), G["moffat_sums"]) < RT
    assert rel(P.psf_laplace(7, 0.3), G["laplace_psf"]) < RT
    assert rel(P.laplace_psf((16, 32), 7, 0.3), G["laplace_H"]) < RT
    assert rel(P.diff_laplace_b((16, 32), 7, 0.3), G["laplace_db"]) < RT


def test_tv_pieces(G):
    x = G["tv_x"]
    assert abs(tv.TVnorm(x) - sc(G["tvnorm"])) <= RT * sc(G["tvnorm"])
    assert np.array_equal(tv.diffh(x), G["diffh"]) and np.array_equal(tv.diffv(x), G["diffv"])
    for i in range(3):
        f, px, py = tv.chambolle_prox_TV_stop(x, "lambda", sc(G[f"chamb{i}_lambda"]), "maxiter", 25)
        assert rel(f, G[f"chamb{i}_f"]) < RT and rel(px, G[f"chamb{i}_px"]) < RT and rel(py, G[f"chamb{i}_py"]) < RT
    f, px, py = tv.chambolle_prox_TV_stop(G["chamb_opt_g"], "LAMBDA", 0.7, "MaxIter", 10, "tol", 1e-2, "tau", 0.2,
                                          "dualvars", G["chamb_opt_dual"])
    assert rel(f, G["chamb_opt_f"]) < RT and rel(px, G["chamb_opt_px"]) < RT and rel(py, G["chamb_opt_py"]) < RT


def test_metrics(G):
    x = G["tv_x"]
    assert abs(metrics.l2(G["gauss_psf"], G["moffat_psf"]) - sc(G["l2"])) <= 1e-12 * sc(G["l2"])
    assert abs(metrics.MSE(x, x + 1.5) - sc(G["MSE"])) < 1e-12
    assert abs(metrics.PSNR(x, x + 1.5) - sc(G["PSNR"])) < 1e-11


MODELS = {"gaussian": 0, "moffat": 1, "laplace": 2}
OVR = {"gaussian": dict(fix_w1=0, fix_w2=0), "moffat": {}, "laplace": {}}
PROBE_PSI = {"gaussian": (0.5, 0.35), "moffat": (0.8, 6.0), "laplace": (0.15,)}


def run_oracle(name):
    g = dict(np.load(os.path.join(GOLDEN, f"ref_sapg_{name}.npz")))
    rng = np.random.default_rng(int(g["seed"]))
    randn = lambda s: rng.standard_normal(s)
    res = OP.setup_demo(MODELS[name], g["x"], randn, samples=16, warmup=6, burnIn=12, **OVR[name])
    return g, res, randn


@pytest.mark.parametrize("name", ["gaussian", "moffat", "laplace"])
def test_demo_setup_and_closures(name):
    """evMax, sigma's, lambda/gamma, y and every op.* closure of the demo scripts."""
    g, res, randn = run_oracle(name)
    y, op = res[0], res[1]
    assert abs(op["evMax"] - sc(g["evMax"])) <= 1e-12 * sc(g["evMax"])
    for k in ("sigma", "sigma_init", "sigma_min", "sigma_max", "lambda", "gamma", "Lf"):
        assert abs(op[k] - sc(g["op_" + k])) <= 1e-12 * abs(sc(g["op_" + k])), k
    assert rel(y, g["y"]) < RT
    cl = op["closures"]
    xp, psi, s2, th = g["probe_xp"], PROBE_PSI[name], 7.5, 0.04
    assert rel(cl["A"](xp, *psi), g["probe_A"]) < RT
    assert rel(cl["AT"](xp, *psi), g["probe_AT"]) < RT
    assert rel(op["gradF"](xp, *psi, s2), g["probe_gradF"]) < RT
    assert abs(op["f"](xp, *psi, s2) - sc(g["probe_f"])) <= RT * sc(g["probe_f"])
    assert abs(op["gradF_sigma"](xp, *psi, s2) - sc(g["probe_gradF_sigma"])) <= 1e-11 * abs(sc(g["probe_gradF_sigma"]))
    assert abs(op["logPi"](xp, th, *psi, s2) - sc(g["probe_logPi"])) <= RT * abs(sc(g["probe_logPi"]))
    assert abs(op["g"](xp) - sc(g["probe_g"])) <= RT * sc(g["probe_g"])
    gnames = {"gaussian": ("grad_w1", "grad_w2"), "moffat": ("grad_alpha", "grad_beta"), "laplace": ("grad_b",)}[name]
    for i, gn in enumerate(gnames):
        want = sc(g[f"probe_grad_psi{i}"])
        assert abs(op[gn](xp, *psi, s2) - want) <= 1e-9 * (abs(want) + 1.0), gn
    prox = op["proxG"](xp, th) if name == "gaussian" else op["proxG"](xp, op["lambda"], th)
    assert rel(prox, g["probe_proxG"]) < RT


@pytest.mark.parametrize("name", ["gaussian", "moffat", "laplace"])
def test_sapg_run_matches_reference_execution(name):
    """Full SAPG_algorithm_* run (warm-up + main loop): every field of `results`."""
    g, res, randn = run_oracle(name)
    if name == "gaussian":
        y, op, c = res
        out = sapg.SAPG_algorithm_Guassian(y, op, c, randn)
    elif name == "moffat":
        y, op = res
        out = sapg.SAPG_algorithm_moffat(y, op, randn)
    else:
        y, op = res
        out = sapg.SAPG_algorithm_laplace(y, op, randn)
    r = out[-1]
    checked = 0
    for k, want in g.items():
        if not k.startswith("res_"):
            continue
        f = k[4:]
        if f in ("execTimeFindParameters", "execTimeFindTheta"):
            continue
        assert f in r, f"oracle results lack field {f}"
        got = np.asarray(r[f], dtype=np.float64)
        want = np.asarray(want, dtype=np.float64)
        assert got.size == want.size, (f, got.shape, want.shape)
        got = got.reshape(want.shape) if got.shape != want.shape else got
        assert np.array_equal(np.isnan(got), np.isnan(want)), f
        m = ~np.isnan(want)
        if f.startswith("tol_"):
            assert np.allclose(got[m], want[m], rtol=1e-6, atol=1e-16), f
        else:
            assert rel(got[m], want[m]) < 1e-10, (f, rel(got[m], want[m]))
        checked += 1
    assert checked >= 25
    # every field the reference returns is also returned by the oracle (field names)
    ref_fields = {k[4:] for k in g if k.startswith("res_")}
    assert ref_fields <= set(r.keys())


def test_salsa_map_matches_reference_execution():
    """SALSA_v2 as the Gaussian demo calls it (run_Gaussian_demo.m:210-244, executed verbatim)."""
    from oracle import salsa
    g = dict(np.load(os.path.join(GOLDEN, "ref_salsa_gaussian.npz")))
    x, y = g["x"], g["y"]
    th, w1, w2, s2 = sc(g["theta_EB"]), sc(g["w1_EB"]), sc(g["w2_EB"]), sc(g["sigma_EB"])
    cl = OP.gaussian_closures(x.shape, 7, 0.0)
    A1 = lambda z: cl["A"](z, w1, w2)
    AT1 = lambda z: cl["AT"](z, w1, w2)
    mu = th / 10                                                    # run_Gaussian_demo.m:222
    F = 1.0 / (np.abs(cl["H_FFT"](w1, w2)) ** 2 + mu)               # :224
    invLS = lambda z: np.real(np.fft.ifft2(F * np.fft.fft2(z)))     # :225
    Psi = lambda z, t: tv.chambolle_prox_TV_stop(z, "lambda", t, "maxiter", 25)[0]
    out = salsa.SALSA_v2(y, A1, th * s2, "MU", mu, "AT", AT1, "StopCriterion", 1, "True_x", x, "ToleranceA", 1e-5,
                         "MAXITERA", 500, "Psi", Psi, "Phi", tv.TVnorm, "TVINITIALIZATION", 1, "TViters", 10,
                         "LS", invLS, "VERBOSE", 0)
    assert rel(out[0], g["xMAP"]) < 1e-11
    assert out[1] == sc(g["numA"]) and out[2] == sc(g["numAt"])
    assert rel(out[3], np.ravel(g["objective"])) < 1e-12
    assert rel(out[4], np.ravel(g["distance"])) < 1e-10
    assert rel(out[6], np.ravel(g["mses"])) < 1e-11
    mse = 10 * np.log10(np.linalg.norm(x - out[0], "fro") ** 2 / x.size)      # :244
    assert abs(mse - sc(g["mse"])) < 1e-9
