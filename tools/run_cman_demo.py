#!/usr/bin/env python
"""BASELINE.json configs[0] on the GPU: run_Gaussian_demo.m on the 256x256 cameraman with the
reference's run lengths (warm-up 15000 + 20000 SAPG iterations, Chambolle 25), estimating
theta / sigma^2 / w1 / w2 (fix_w1 = fix_w2 = 0).  Prints one JSON line with the device time,
the estimates, and the CPU oracle's s/iteration on this host for the >=100x target.

    python tools/run_cman_demo.py [--samples 20000 --warmup 15000] [--chains 1]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=15000)
    ap.add_argument("--chains", type=int, default=1)
    ap.add_argument("--cpu-iters", type=int, default=20)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--image", default="cman", help="cman (256x256) or boat (512x512)")
    a = ap.parse_args()
    import sbd_b200
    from sbd_b200 import host as H
    x = np.load(os.path.join(ROOT, "tests", "golden", a.image + "_u8.npy")).astype(np.float64)
    n = x.shape[0]
    eng = sbd_b200.Engine(n, n, 7, H.GAUSSIAN, 0.0, a.chains, 0)
    rng = np.random.default_rng(1)
    Ax = eng.blur(x, (0.4, 0.3), H.OP_A)
    nrm = float(np.linalg.norm(Ax - Ax.mean()))
    sig = lambda b: nrm / np.sqrt(n * n * 10 ** (b / 10))
    sigma, smin, smax = sig(30), sig(15), sig(45)
    y = Ax + sigma * rng.standard_normal(x.shape)
    op = dict(samples=a.samples, warmup=a.warmup, burnIn=(a.samples * 80) // 100, psf_size=7, phi=0.0,
              min_th=1e-3, max_th=1.0, min_w1=0.1, max_w1=1.0, min_w2=0.1, max_w2=1.0, th_init=0.01,
              w1_init=0.5, w2_init=0.3, w1=0.4, w2=0.3, fix_w1=0, fix_w2=0, fix_sigma=0, d_exp=0.8, d_scale=1.0,
              sigma=sigma, sigma_init=(smin ** 2 + smax ** 2) / 2, sigma_min=smin ** 2, sigma_max=smax ** 2,
              use_graph=a.graph)
    Lf = min(0.993 ** 2 / smin ** 2, 0.993 ** 2 / smax ** 2)
    op["lambda"] = min(5 / Lf, 2.0); op["gamma"] = 0.98 / (Lf + 1 / op["lambda"])
    c = dict(sigma=1000.0, theta=0.01, w1=10.0, w2=10.0, lam=1.0, gam=1.0)
    t0 = time.perf_counter()
    th, w1, w2, s2, r = sbd_b200.SAPG_algorithm_Guassian(y, op, c, n_chains=a.chains, seed=1, engine=eng, post_mean=True)
    wall = time.perf_counter() - t0
    steps = (a.warmup - 1) + (a.samples - 1)
    import oracle
    xm = r["posteriormean"]
    out = {"config": "run_Gaussian_demo.m on cman 256x256 (BASELINE.json configs[0]), fix_w1=fix_w2=0",
           "myula_steps": steps, "chains": a.chains, "gpu_wall_s": wall, "gpu_device_s": r["execTimeFindParameters"],
           "gpu_steps_per_s": steps / r["execTimeFindParameters"],
           "theta_EB": th, "w1_EB": w1, "w2_EB": w2, "sigma2_EB": s2, "true": {"w1": 0.4, "w2": 0.3, "sigma2": sigma ** 2},
           "psnr_y": oracle.metrics.PSNR(x, y), "psnr_mmse": oracle.metrics.PSNR(x, xm),
           "chambolle_sweeps_mean": float(np.mean(r["chambolle_iters"][1:]))}
    if os.environ.get("SBD_PROFILE"):
        out["phase_ms_calls"] = {k_: (round(v_[0], 3), v_[1]) for k_, v_ in eng.phase_times().items()}
    if a.cpu_iters > 0:
        import scipy.fft
        from oracle import operators as OP
        workers = os.cpu_count() or 1
        OP.set_fft(lambda z: scipy.fft.fft2(z, workers=workers), lambda z: scipy.fft.ifft2(z, workers=workers))
        stamps = []
        rr = np.random.default_rng(2)

        def randn(shape):
            stamps.append(time.perf_counter())
            return rr.standard_normal(shape)

        y2, op2, c2 = OP.setup_demo(0, x, lambda s: rr.standard_normal(s), samples=a.cpu_iters + 1, warmup=1,
                                    burnIn=2, fix_w1=0, fix_w2=0, evMax=0.993)
        oracle.sapg.SAPG_algorithm_Guassian(y2, op2, c2, randn)
        stamps.append(time.perf_counter())
        per = float(np.median(np.diff(stamps)))
        out["cpu_s_per_iter"] = per
        out["cpu_workers"] = workers
        out["cpu_extrapolated_s"] = per * steps
        out["speedup_vs_cpu"] = per * steps / r["execTimeFindParameters"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
