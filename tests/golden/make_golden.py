#!/usr/bin/env python
"""Generate the golden fixtures by EXECUTING THE REFERENCE'S OWN .m FILES.

    python tests/golden/make_golden.py            (needs /root/reference; run in the build container)

No MATLAB/Octave exists in this image, so the unmodified reference sources are
run through the MATLAB-subset interpreter in oracle/mlab/interp.py:
  * utils/*.m, SALSA/{diffh,diffv,conv2c}.m, SAPG/SAPG_algorithm_*.m as function files,
  * the closure / step-size / observation blocks of run_{Gaussian,moffat,laplace}_demo.m
    as verbatim line ranges of those scripts (everything between reading the
    image and calling SAPG; the image I/O, SALSA and plotting parts are skipped).
The only substitutions: `randn` draws from numpy's default_rng(seed) (MATLAB's
v5 'state' generator cannot be reproduced) and run lengths are shortened.
The fixtures (inputs, seeds and outputs) go to tests/golden/*.npz; the tests
compare the numpy oracle (tests/test_oracle_golden.py) and the CUDA engine
(tests/test_gpu_golden.py) against them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.mlab.interp import Interp, MStruct, M, to_py  # noqa: E402

REF = os.environ.get("SBD_REFERENCE", "/root/reference")


def lines_of(path, a, b):
    """1-based inclusive line range of a reference file, verbatim"""
    with open(os.path.join(REF, path), encoding="utf-8", errors="replace") as f:
        L = f.read().split("\n")
    return "\n".join(L[a - 1:b]) + "\n"


def new_interp(seed):
    rng = np.random.default_rng(seed)
    return Interp([os.path.join(REF, d) for d in ("utils", "SALSA", "SAPG")],
                  randn=lambda shape: rng.standard_normal(shape))


def crop(name, r0, c0, h, w):
    img = np.load(os.path.join(HERE, name)).astype(np.float64)
    return img[r0:r0 + h, c0:c0 + w].copy()


def arr(v):
    return np.asarray(v, dtype=np.complex128 if np.iscomplexobj(v) else np.float64)


# ---------------------------------------------------------------------------
def operators():
    it = new_interp(0)
    call = lambda name, *a, nargout=1: it.call_function(name, [M(x) if not isinstance(x, (str, np.ndarray)) else x for x in a], nargout)
    out = {}
    sz = np.array([[16.0, 32.0]])
    out["gauss_psf"] = call("Gaussian_psf", 7, 0.4, 0.3, 0.0)[0]
    out["gauss_psf_rot"] = call("psf_gaussian", 7, 0.5, 0.2, 0.3)[0]
    out["gauss_H"] = call("resize", out["gauss_psf"], sz)[0]
    out["gauss_dw1"] = call("diff_fftgaus_w1", sz, 7, 0.4, 0.3, 0.0)[0]
    out["gauss_dw2"] = call("diff_fftgaus_w2", sz, 7, 0.4, 0.3, 0.0)[0]
    s = call("Sum_gauss_psf", 7, 0.4, 0.3, 0.0, nargout=3)
    out["gauss_sums"] = np.array([float(v.item()) for v in s])
    out["moffat_psf"] = call("psf_moffat", 7, 0.4, 3.5)[0]
    out["moffat_H"] = call("moffat_psf", sz, 7, 0.4, 3.5)[0]
    out["moffat_da"] = call("diff_moffat_alpha", sz, 7, 0.4, 3.5)[0]
    out["moffat_db"] = call("diff_moffat_beta", sz, 7, 0.4, 3.5)[0]
    s = call("sum_mof_psf", 7, 0.4, 3.5, nargout=3)
    out["moffat_sums"] = np.array([float(v.item()) for v in s])
    out["laplace_psf"] = call("psf_laplace", 7, 0.3)[0]
    out["laplace_H"] = call("laplace_psf", sz, 7, 0.3)[0]
    out["laplace_db"] = call("diff_laplace_b", sz, 7, 0.3)[0]
    x = crop("cman_u8.npy", 100, 60, 32, 40)
    out["tv_x"] = x
    out["tvnorm"] = call("TVnorm", x)[0]
    out["diffh"] = call("diffh", x)[0]
    out["diffv"] = call("diffv", x)[0]
    for i, lam in enumerate((1e-3, 0.5, 20.0)):
        f, px, py = call("chambolle_prox_TV_stop", x, "lambda", lam, "maxiter", 25, nargout=3)
        out[f"chamb{i}_lambda"] = np.array(lam)
        out[f"chamb{i}_f"], out[f"chamb{i}_px"], out[f"chamb{i}_py"] = f, px, py
    sq = crop("cman_u8.npy", 64, 64, 32, 32)
    rng = np.random.default_rng(3)
    dual = rng.uniform(-0.5, 0.5, (32, 64))
    out["chamb_opt_g"], out["chamb_opt_dual"] = sq, dual
    f, px, py = call("chambolle_prox_TV_stop", sq, "LAMBDA", 0.7, "MaxIter", 10, "tol", 1e-2, "tau", 0.2,
                     "dualvars", dual, nargout=3)
    out["chamb_opt_f"], out["chamb_opt_px"], out["chamb_opt_py"] = f, px, py
    out["l2"] = call("l2", out["gauss_psf"], out["moffat_psf"])[0]
    out["MSE"] = call("MSE", x, x + 1.5)[0]
    out["PSNR"] = call("PSNR", x, x + 1.5)[0]
    np.savez_compressed(os.path.join(HERE, "ref_operators.npz"), **{k: arr(v) for k, v in out.items()})
    print("ref_operators.npz:", len(out), "arrays")


# ---------------------------------------------------------------------------
# script line ranges (1-based, inclusive) of the three demos
DEMO = {
    "gaussian": dict(script="run_Gaussian_demo.m", hyper=(34, 88), fix=(102, 107), setup=(123, 195), call=(199, 199),
                     img=("cman_u8.npy", 96, 64, 32, 32), overrides="fix_w1 = 0; fix_w2 = 0; op.fix_w1 = 0; op.fix_w2 = 0;\n"),
    "moffat": dict(script="run_moffat_demo.m", hyper=(34, 83), fix=(92, 105), setup=(119, 185), call=(189, 189),
                   img=("boat_u8.npy", 200, 100, 32, 64), overrides=""),
    "laplace": dict(script="run_laplace_demo.m", hyper=(34, 71), fix=None, setup=(93, 153), call=(157, 157),
                    img=("cman_u8.npy", 32, 100, 64, 32), overrides=""),
}
RUN = dict(samples=16, warmup=6, burnIn=12)


def sapg(model, seed):
    d = DEMO[model]
    it = new_interp(seed)
    sc = {}
    x = crop(*d["img"])
    it.run_source(lines_of(d["script"], *d["hyper"]), sc)                   # hyper-parameters, verbatim
    it.run_source(f"op.samples = {RUN['samples']}; op.warmup = {RUN['warmup']}; op.burnIn = {RUN['burnIn']};\n"
                  + d["overrides"], sc)
    sc["x"] = x
    if model == "gaussian":
        it.run_source("snr = 30; op.BSNR = snr;\n" + lines_of(d["script"], *d["fix"]) +
                      "dimX = numel(x); op.x = x;\n", sc)
    elif model == "moffat":
        it.run_source(lines_of(d["script"], *d["fix"]) + "snr = 30; op.BSNR = snr; op.x = x; dimX = numel(x); im_size = size(x);\n", sc)
    else:
        it.run_source("b = 0.3; op.b = b; snr = 30; op.BSNR = snr; op.x = x; dimX = numel(x); im_size = size(x);\n", sc)
    it.run_source(lines_of(d["script"], *d["setup"]), sc)                   # closures, evMax, y, step sizes: verbatim
    y = sc["y"].copy()
    op = sc["op"]
    probe = {}
    # closure probes on a fixed input (the closures are the reference's own anonymous functions)
    rng = np.random.default_rng(99)
    xp = np.abs(x + 3.0 * rng.standard_normal(x.shape))
    if model == "gaussian":
        args = [M(0.5), M(0.35)]
    elif model == "moffat":
        args = [M(0.8), M(6.0)]
    else:
        args = [M(0.15)]
    s2, th = M(7.5), M(0.04)
    probe["xp"] = xp
    probe["A"] = sc["A"]([xp] + args)[0]
    probe["AT"] = sc["AT"]([xp] + args)[0]
    probe["gradF"] = op["gradF"]([xp] + args + [s2])[0]
    probe["f"] = op["f"]([xp] + args + [s2])[0]
    probe["gradF_sigma"] = op["gradF_sigma"]([xp] + args + [s2])[0]
    probe["logPi"] = op["logPi"]([xp, th] + args + [s2])[0]
    probe["g"] = op["g"]([xp])[0]
    names = {"gaussian": ("grad_w1", "grad_w2"), "moffat": ("grad_alpha", "grad_beta"), "laplace": ("grad_b",)}[model]
    for i, nme in enumerate(names):
        probe[f"grad_psi{i}"] = op[nme]([xp] + args + [s2])[0]
    if model == "gaussian":
        probe["proxG"] = op["proxG"]([xp, th])[0]
    else:
        probe["proxG"] = op["proxG"]([xp, M(op["lambda"]), th])[0]
    # the SAPG call, verbatim
    it.run_source(lines_of(d["script"], *d["call"]), sc)
    res = to_py(sc["results"])
    out = {"x": x, "y": y, "seed": np.array(seed)}
    for k in ("evMax",):
        out[k] = arr(sc[k])
    for k in ("sigma", "sigma_init", "sigma_min", "sigma_max", "lambda", "gamma", "Lf"):
        out["op_" + k] = arr(op[k])
    for k, v in probe.items():
        out["probe_" + k] = arr(v)
    for k, v in res.items():
        if isinstance(v, (float, int, np.ndarray)):
            out["res_" + k] = arr(v)
    np.savez_compressed(os.path.join(HERE, f"ref_sapg_{model}.npz"), **out)
    print(f"ref_sapg_{model}.npz:", len(out), "arrays; last theta", np.ravel(out["res_thetas"])[-1])


def salsa(seed=7):
    """run_Gaussian_demo.m:210-244 verbatim (A1/AT1 with callcounter, mu, filter_FFT, invLS, the SALSA_v2 call
    and the MSE line), fed with fixed estimates instead of a 35 000-iteration SAPG run."""
    it = new_interp(seed)
    x = crop("cman_u8.npy", 96, 64, 32, 32)
    sc = {"x": x}
    it.run_source(lines_of("run_Gaussian_demo.m", 34, 88), sc)
    it.run_source("snr = 30; op.BSNR = snr; dimX = numel(x); op.x = x;\n", sc)
    it.run_source(lines_of("run_Gaussian_demo.m", 123, 195), sc)
    it.run_source("theta_EB = 0.045; w1_EB = 0.43; w2_EB = 0.28; sigma_EB = 4.2;\n", sc)
    it.run_source(lines_of("run_Gaussian_demo.m", 210, 244), sc)
    out = {"x": x, "y": sc["y"], "seed": np.array(seed), "theta_EB": np.array(0.045), "w1_EB": np.array(0.43),
           "w2_EB": np.array(0.28), "sigma_EB": np.array(4.2), "xMAP": sc["xMAP"], "mse": arr(sc["mse"]),
           "mu": arr(sc["mu"]), "calls": arr(it.globals.get("calls", np.zeros((1, 1))))}
    # the same call again, keeping every output of SALSA_v2
    it.run_source("[xM2, numA, numAt, objective, distance, times, mses] = SALSA_v2(y, A1, theta_EB*sigma_EB, 'MU', mu, "
                  "'AT', AT1, 'StopCriterion', 1, 'True_x', x, 'ToleranceA', tol, 'MAXITERA', outeriters, 'Psi', Psi, "
                  "'Phi', op.g, 'TVINITIALIZATION', 1, 'TViters', 10, 'LS', invLS, 'VERBOSE', 0);\n", sc)
    for k in ("objective", "distance", "mses", "numA", "numAt"):
        out[k] = arr(sc[k])
    np.savez_compressed(os.path.join(HERE, "ref_salsa_gaussian.npz"), **out)
    print("ref_salsa_gaussian.npz: outer iterations", np.ravel(out["objective"]).size - 1, "mse", float(np.ravel(out["mse"])[0]),
          "calls", float(np.ravel(out["calls"])[0]))


if __name__ == "__main__":
    salsa()
    if "--salsa-only" in sys.argv:
        sys.exit(0)
    operators()
    for i, m in enumerate(("gaussian", "moffat", "laplace")):
        sapg(m, 100 + i)
