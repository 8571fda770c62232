"""The experiment scripts of the reference mirrored on the engine (sbd_b200/demo.py): setup stage on the device,
SAPG, SALSA MAP estimate, and the `for snr` / `for i_im` batch driver with its results files."""
import os

import numpy as np
import pytest

from conftest import rel, GOLDEN

pytestmark = pytest.mark.gpu


def sc(v):
    return float(np.asarray(v).ravel()[0])


def test_setup_demo_matches_the_reference_execution():
    """demo.setup_demo (power iteration, observation, step sizes on the device) against the values the reference's
    run_Gaussian_demo.m produced when executed (tests/golden/make_golden.py): evMax, sigma, y, lambda, gamma, Lf."""
    from sbd_b200 import demo, host as H
    g = dict(np.load(os.path.join(GOLDEN, "ref_sapg_gaussian.npz")))
    rng = np.random.default_rng(int(sc(g["seed"])))
    x0 = rng.standard_normal(g["x"].shape); noise = rng.standard_normal(g["x"].shape)
    y, op, eng = demo.setup_demo(H.GAUSSIAN, g["x"], noise=noise, x0_eig=x0, samples=16, warmup=6, fix_w1=0, fix_w2=0)
    assert abs(op["evMax"] - sc(g["evMax"])) <= 1e-11 * sc(g["evMax"])
    assert rel(y, g["y"]) < 1e-12
    for k in ("sigma", "sigma_init", "sigma_min", "sigma_max", "lambda", "gamma", "Lf"):
        assert abs(op[k] - sc(g["op_" + k])) <= 1e-10 * abs(sc(g["op_" + k])), k


@pytest.mark.parametrize("model,name", [(0, "gaussian"), (1, "moffat"), (2, "laplace")])
def test_run_demo_reproduces_the_reference_sapg(model, name):
    """run_demo = the body of the experiment loops: fed the reference execution's randn stream it must give the
    reference's trajectories (goldens), then the MAP stage runs on top."""
    from sbd_b200 import demo
    g = dict(np.load(os.path.join(GOLDEN, f"ref_sapg_{name}.npz")))
    rng = np.random.default_rng(int(sc(g["seed"])))
    shape = g["x"].shape
    x0 = rng.standard_normal(shape); noise = rng.standard_normal(shape)
    tape = np.stack([rng.standard_normal(shape) for _ in range(5 + 15)])[:, None]
    kw = dict(fix_w1=0, fix_w2=0) if model == 0 else {}
    res = demo.run_demo(model, g["x"], noise=noise, x0_eig=x0, noise_sapg=tape, samples=16, warmup=6, burnIn=12,
                        map_estimate=True, use_graph=False, name="crop", **kw)
    names = {0: [("thetas", "thetas"), ("w1s", "w1s"), ("w2s", "w2s"), ("sigmas", "sigmas")],
             1: [("thetas", "thetas"), ("alphas", "alphas"), ("betas", "betas"), ("sigmas", "sigmas")],
             2: [("thetas", "thetas"), ("bs", "bs"), ("sigmas", "sigmas")]}[model]
    for mine, ref in names:
        assert rel(res[mine], np.ravel(g["res_" + ref])) < 1e-6, mine
    assert np.isfinite(res["mse"]) and res["xMAP"].shape == shape and res["salsa_outer"] >= 1


def test_run_batch_loops_and_saves(tmp_path):
    """for snr / for i_im: jobs are dealt round-robin to ranks, each result is saved like the scripts' `save`."""
    from sbd_b200 import demo, host as H
    imgs = np.load(os.path.join(GOLDEN, "images_u8.npz"))
    images = {n: imgs[n][128:192, 64:128].astype(np.float64) for n in ("barbara", "boat", "wheel")}
    kw = dict(samples=12, warmup=4, burnIn=8, map_estimate=False, use_graph=False, evMax=0.993)
    all_jobs = demo.run_batch(H.LAPLACE, images, bsnrs=(30, 20), out_dir=str(tmp_path / "all"), **kw)
    assert len(all_jobs) == 6
    r0 = demo.run_batch(H.LAPLACE, images, bsnrs=(30, 20), out_dir=str(tmp_path / "r0"), rank=0, world=2, device=0, **kw)
    r1 = demo.run_batch(H.LAPLACE, images, bsnrs=(30, 20), out_dir=str(tmp_path / "r1"), rank=1, world=2, device=0, **kw)
    assert set(r0) | set(r1) == set(all_jobs) and not (set(r0) & set(r1))
    for key, res in list(r0.items()) + list(r1.items()):
        assert np.array_equal(res["thetas"], all_jobs[key]["thetas"])          # same seed, same job -> same run
        assert np.array_equal(res["bs"], all_jobs[key]["bs"])
    files = sorted(os.listdir(tmp_path / "all"))
    assert files == sorted(f"laplace_{n}_bsnr{s}.npz" for n in images for s in (30, 20))
    saved = np.load(tmp_path / "all" / "laplace_boat_bsnr20.npz")
    assert rel(saved["thetas"], all_jobs[("boat", 20)]["thetas"]) == 0 and "X_sample" in saved.files and float(saved["op_BSNR"]) == 20
