#!/usr/bin/env python
"""bench.py - MYULA (SAPG main-loop) iterations per second on B200.

  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[3], the one the headline metric is quoted on):
synthetic 4096x4096 Gaussian-PSF SAPG, `--chains-per-gpu` (default 8) MYULA
chains per GPU, chains sharded over ranks (weak scaling; 8 GPUs = the 64-chain
configuration), NCCL all-gather of the per-chain stochastic gradients at every
outer iteration.  A "step" is one SAPG main-loop iteration of every chain
(SAPG_algorithm_Guassian.m:158-248: Langevin update, Chambolle prox with 25
sweeps, spectral gradients, theta/sigma^2/w1/w2 updates).  W warm-up steps are
real MYULA warm-up iterations (Guassian.m:78-91) and are not timed.

  value : chain-steps/s with y and X0 resident in HBM (device-timed main loop)
  e2e   : same metric through the host-pointer C ABI (sbd_sapg_run): y copied
          from pinned host memory and trajectories + last samples copied back,
          inside the timed region (wall clock around the call)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PSI_TRUE = (0.4, 0.3)            # run_Gaussian_demo.m:78-79
PSI_INIT = (0.5, 0.3)            # run_Gaussian_demo.m:68-69 (estimating the bandwidths, Q17)
CHAMBOLLE_K = 25                 # run_Gaussian_demo.m:188


def alg_bytes_per_chain_step(npix, k=CHAMBOLLE_K):
    """B_alg = (144 + 40 K) * P bytes per chain per main-loop step."""
    return (144 + 40 * k) * npix


def synthetic_truth(n):
    """cman (256x256) periodically tiled to n x n."""
    cm = np.load(os.path.join(ROOT, "tests", "golden", "cman_u8.npy")).astype(np.float64)
    reps = (n + 255) // 256
    return np.tile(cm, (reps, reps))[:n, :n].copy()


def gaussian_op(n, sigma, sigma_min, sigma_max, evMax, samples, warmup):
    """op / c of run_Gaussian_demo.m:34-85,177-191 with fix_w1 = fix_w2 = 0."""
    op = dict(samples=samples, warmup=warmup, burnIn=max(2, (samples * 80) // 100), psf_size=7, phi=0.0,
              min_th=1e-3, max_th=1.0, min_w1=0.1, max_w1=1.0, min_w2=0.1, max_w2=1.0,
              th_init=0.01, w1_init=PSI_INIT[0], w2_init=PSI_INIT[1], w1=PSI_TRUE[0], w2=PSI_TRUE[1],
              fix_w1=0, fix_w2=0, fix_sigma=0, d_exp=0.8, d_scale=0.01 / 0.01,
              sigma=sigma, sigma_init=(sigma_min ** 2 + sigma_max ** 2) / 2,
              sigma_min=sigma_min ** 2, sigma_max=sigma_max ** 2, chambolleit=CHAMBOLLE_K)
    Lf = min(evMax ** 2 / sigma_min ** 2, evMax ** 2 / sigma_max ** 2)
    op["lambda"] = min(5 / Lf, 2.0)
    op["gamma"] = 0.98 / (Lf + 1 / op["lambda"])
    c = dict(sigma=1000.0, theta=0.01, w1=10.0, w2=10.0, lam=1.0, gam=1.0)
    return op, c


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [s.strip() for s in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            # the busiest half of the samples = "under load"
            top = sorted(sm)[len(sm) // 2:]
            out = {"sm_mhz": float(np.median(top)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ---------------------------------------------------------------------------
# CPU arm: the reference-shaped oracle loop (24 FFTs / iteration, unfused)
# ---------------------------------------------------------------------------
def cpu_reference_rate(n_full, sample_n, iters, workers):
    """Time `iters` SAPG main-loop iterations of ONE chain of the oracle
    (oracle/sapg.py, scipy.fft with `workers` threads) on a sample_n x sample_n
    crop and scale by pixel count to the n_full x n_full workload.
    Returns (chain-steps/s at n_full, seconds per sample iteration)."""
    import scipy.fft
    import oracle
    from oracle import operators as OP
    OP.set_fft(lambda a: scipy.fft.fft2(a, workers=workers), lambda a: scipy.fft.ifft2(a, workers=workers))
    x = synthetic_truth(sample_n)
    rng = np.random.default_rng(1)
    stamps = []

    def randn(shape):
        stamps.append(time.perf_counter())
        return rng.standard_normal(shape)

    y, op, c = OP.setup_demo(0, x, lambda s: rng.standard_normal(s), samples=iters + 1, warmup=1,
                             burnIn=2, fix_w1=0, fix_w2=0, evMax=0.993)
    oracle.sapg.SAPG_algorithm_Guassian(y, op, c, randn)
    stamps.append(time.perf_counter())
    per_iter = float(np.median(np.diff(stamps)))
    scale = (n_full / sample_n) ** 2
    OP.set_fft(np.fft.fft2, np.fft.ifft2)
    return 1.0 / (per_iter * scale), per_iter


def pick_cpu_sample(n_full, budget_s, steps):
    """Largest power-of-two crop whose `steps` iterations fit the time budget
    (survey container: 0.106 s @256^2, 0.52 s @512^2, 2.7 s @1024^2, 11 s @2048^2)."""
    est = {256: 0.12, 512: 0.6, 1024: 3.0, 2048: 12.0, 4096: 50.0}
    best = 256
    for n in (256, 512, 1024, 2048, 4096):
        if n <= n_full and est[n] * steps <= budget_s:
            best = n
    return best


def run_reference(args, rank, world):
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    total = args.steps + args.warmup
    sample_n = pick_cpu_sample(args.size, 150.0, total)
    t0 = time.perf_counter()
    rate, per_iter = cpu_reference_rate(args.size, sample_n, total, workers)
    wall = time.perf_counter() - t0
    unit = "chain-steps/s"
    sample = (f"{total} SAPG main-loop iterations of one chain of the numpy/scipy oracle "
              f"(unfused, 24 FFTs + 25 Chambolle sweeps per iteration) on a {sample_n}x{sample_n} crop, "
              f"median s/iteration x {(args.size / sample_n) ** 2:.0f} (pixel ratio) -> {args.size}^2; "
              f"scipy.fft workers={workers}; MATLAB/Octave absent so maxNumCompThreads is N/A")
    line = {"impl": "reference", "metric": "MYULA chain-iterations/sec", "value": rate, "unit": unit,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_iter * (args.size / sample_n) ** 2 * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": rate, "unit": unit, "cores": workers, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "wall_s": wall}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"synthetic {args.size}x{args.size} Gaussian-PSF SAPG main loop "
                        f"(BASELINE.json configs[3]), {args.chains_per_gpu} MYULA chains per GPU sharded over "
                        f"{world} GPU(s), Chambolle K={CHAMBOLLE_K}, psf 7x7, BSNR 30, estimating theta/sigma2/w1/w2",
            "image": f"{args.size}x{args.size}", "chains_per_gpu": args.chains_per_gpu,
            "total_chains": args.chains_per_gpu * world, "parallelism": f"chains x{world}",
            "l2_policy": "working set per step (>= 9.7 GB at 4096^2 x 8 chains) far exceeds the 126 MB L2; no flush needed"
            if args.size * args.size * 8 * 9 * args.chains_per_gpu > 4 * 126e6 else "L2 flushed between runs only"}


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import sbd_b200
    from sbd_b200 import host as H
    from sbd_b200._lib import lib, sbd_traces, c_double_p

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n, nch = args.size, args.chains_per_gpu
    npix = n * n
    K, W = args.steps, args.warmup
    shard = sbd_b200.ChainShard(nch * world, world, rank)

    # ---- synthetic problem (same on every rank)
    x_true = synthetic_truth(n)
    ev = 0.993      # evMax of A'A at (w1,w2)=(1,1): what the reference's power iteration returns
    eng = sbd_b200.Engine(n, n, 7, H.GAUSSIAN, 0.0, nch, local_rank)
    shard.init_engine_comm(eng)
    Ax = eng.blur(x_true, PSI_TRUE, H.OP_A)
    nrm = float(np.linalg.norm(Ax - Ax.mean()))
    sig = lambda b: nrm / np.sqrt(npix * 10 ** (b / 10))    # run_Gaussian_demo.m:148-152
    sigma, smin, smax = sig(30), sig(15), sig(45)
    y = Ax + sigma * np.random.default_rng(2).standard_normal((n, n))
    del Ax

    def params(samples, warmup):
        op, c = gaussian_op(n, sigma, smin, smax, ev, samples, warmup)
        return H.make_params(H.GAUSSIAN, op, c, n_chains=nch, seed=1, chain_offset=shard.chain_offset,
                             total_chains=shard.total_chains)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- (1) value: y resident in HBM, device-timed main loop
    y_dev = torch.from_numpy(np.ascontiguousarray(y.T)).cuda()      # column-major image on the device
    prm = params(K + 1, W + 1)
    tr = sbd_traces()
    th = np.zeros(K + 1); tr.thetas = th.ctypes.data_as(c_double_p)
    s2 = np.zeros(K + 1); tr.sigmas = s2.ctypes.data_as(c_double_p)
    ck = np.zeros(K + 1, dtype=np.int32); tr.chambolle_iters = ck.ctypes.data_as(C.POINTER(C.c_int32))
    eng.set_profile(True)
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    rc = lib.sbd_sapg_run_dev(eng._h, y_dev.data_ptr(), None, C.byref(prm), C.byref(tr))
    if rc != 0:
        raise RuntimeError(lib.sbd_last_error(eng._h).decode())
    barrier()
    clk = clocks.stop() if clocks else None
    t_main = max_over_ranks(tr.seconds_main)
    launches = int(tr.launches_main)
    phases = eng.phase_times()
    eng.set_profile(False)
    value = K * shard.total_chains / t_main
    sweeps_executed = int(ck[1:].sum())                      # chain 0's stop behaviour (all chains share theta)

    # ---- (2) e2e: host-pointer C ABI, pinned host buffers, copies inside the timed region
    y_pin = torch.from_numpy(np.ascontiguousarray(y.T)).pin_memory()
    xl_pin = torch.empty(nch * npix, dtype=torch.float64).pin_memory()
    prm2 = params(K + 1, 1)
    tr2 = sbd_traces()
    bufs = {k: np.zeros(K + 1) for k in ("thetas", "sigmas", "psi0", "psi1", "logPiTraceX", "gXTrace")}
    for k_, b in bufs.items():
        setattr(tr2, k_, b.ctypes.data_as(c_double_p))
    tr2.X_last = C.cast(xl_pin.data_ptr(), c_double_p)
    barrier()
    t0 = time.perf_counter()
    rc = lib.sbd_sapg_run(eng._h, C.cast(y_pin.data_ptr(), c_double_p), None, None, C.byref(prm2), None, C.byref(tr2))
    if rc != 0:
        raise RuntimeError(lib.sbd_last_error(eng._h).decode())
    torch.cuda.synchronize()
    wall = max_over_ranks(time.perf_counter() - t0)
    e2e_value = K * shard.total_chains / wall
    h2d = npix * 8 / K
    d2h = (nch * npix * 8 + 6 * (K + 1) * 8) / K

    if rank != 0:
        return
    # ---- roofline of the dominant kernel (fused Chambolle sweeps), live CUDA-event timing.
    # Algorithmic bytes: 40 B per pixel per SWEEP (read g,px,py; write px,py).
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak = 6650.0; peak_src = "fallback 6.65 TB/s (B200_PROFILING.md)"
    # block plan of sbd.cu chambolle(): K = 25 -> four 4-level launches + three 3-level launches, the last of which
    # also writes the prox output; phase "chambolle_sweeps" times the 4-level launches (and their no-op redo launches) only
    a4, r4 = CHAMBOLLE_K // 4, CHAMBOLLE_K % 4
    if r4 == 1 and a4 >= 2 and os.environ.get("SBD_CHAMB_PLAN33", "1") != "0":
        n4 = a4 - 2
    else:
        n4 = a4 if r4 in (1, 3) else a4 - 1
    n_prox = int(phases["chambolle_sweeps"][1])
    launches_timed = n_prox * n4
    sweep_ms = phases["chambolle_sweeps"][0] / max(launches_timed, 1)
    sweep_bytes = 40.0 * npix * nch * 4                        # algorithmic bytes of the 4 sweeps one launch applies
    achieved = sweep_bytes / (sweep_ms * 1e-3) / 1e9 if sweep_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp) and n == 4096 and nch == 8:
        try:
            traffic = json.load(open(tp)).get("k_chamb_multi<4, 0, 3, 0, 0>")
        except Exception:
            traffic = None
    step_gbs = alg_bytes_per_chain_step(npix) * value / 1e9
    line = {
        "metric": "MYULA chain-iterations/sec", "value": value, "unit": "chain-steps/s",
        "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": t_main / K * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e_value, "unit": "chain-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "note": "one sbd_sapg_run call = K steps; y from pinned host memory in, trajectories + last samples out"},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "k_chamb_multi<4> (4 Chambolle sweeps fused per launch)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "bytes_per_launch": sweep_bytes, "ms_per_launch": sweep_ms,
                     "launches_timed": launches_timed, "sweeps_executed_chain0": sweeps_executed,
                     "note": "algorithmic bytes = 40 B/pixel/sweep x the 4 sweeps of one launch; temporal blocking keeps the 4 "
                             "sweep levels in registers, so the measured DRAM traffic per launch (ncu, `traffic`) is ~1/4 of it and "
                             "the model-based fraction exceeds 1. What bounds the kernel instead is fp64 instruction issue: ncu shows "
                             "the fp64 pipe 60 % active, ~91 % of the issue bound calibrated by tools/fp64_microbench.cu "
                             "(profiles/r01_ncu_summary.md, profiles/r01_fp64_microbench.txt, DESIGN.md section 5)"},
        "fused_step": {"alg_bytes_per_chain_step": alg_bytes_per_chain_step(npix), "achieved_gbs_per_gpu": step_gbs / world,
                       "frac_of_hbm_peak": step_gbs / world / peak,
                       "phase_ms_per_step": {("chambolle_total" if k_ == "chambolle_other" else k_): v_[0] / K
                                             for k_, v_ in phases.items()}},
        "theta_last": float(th[-1]), "sigma2_last": float(s2[-1]),
    }
    if world == 1 and not args.no_size_sweep:
        # the metric is quoted on 256^2 .. 4096^2: short runs of the same SAPG main loop at the smaller sizes
        sweep = {str(n): round(value, 2)}
        for m in (2048, 1024, 512, 256):
            if m >= n:
                continue
            sweep[str(m)] = round(quick_rate(m, nch, local_rank), 2)
        line["size_sweep_chain_steps_per_s"] = sweep
    if world == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        sample_n = 1024 if args.size >= 1024 else args.size
        iters = 4 if sample_n >= 1024 else 10
        rate, per_iter = cpu_reference_rate(args.size, sample_n, iters, workers)
        line["cpu_baseline"] = {
            "value": rate, "unit": "chain-steps/s", "cores": workers, "kind": "port",
            "sample": (f"{iters} SAPG main-loop iterations of one chain of the numpy/scipy oracle (unfused, 24 FFTs "
                       f"+ 25 Chambolle sweeps per iteration) on a {sample_n}x{sample_n} crop, {per_iter:.3f} s/iteration, "
                       f"scaled x{(args.size / sample_n) ** 2:.0f} by pixel count; scipy.fft workers={workers}; "
                       f"MATLAB/Octave absent (maxNumCompThreads N/A)")}
    print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def quick_rate(n, nch, device, steps=30, warmup=4):
    """chain-steps/s of the SAPG main loop at n x n (device-resident y, CUDA-graph replay), same model as the headline."""
    import torch
    import sbd_b200
    from sbd_b200 import host as H
    from sbd_b200._lib import lib, sbd_traces
    eng = sbd_b200.Engine(n, n, 7, H.GAUSSIAN, 0.0, nch, device)
    x = synthetic_truth(n)
    Ax = eng.blur(x, PSI_TRUE, H.OP_A)
    nrm = float(np.linalg.norm(Ax - Ax.mean()))
    sig = lambda b: nrm / np.sqrt(n * n * 10 ** (b / 10))
    y = Ax + sig(30) * np.random.default_rng(2).standard_normal((n, n))
    op, c = gaussian_op(n, sig(30), sig(15), sig(45), 0.993, steps + 1, warmup + 1)
    op["use_graph"] = 1
    prm = H.make_params(H.GAUSSIAN, op, c, n_chains=nch, seed=1)
    y_dev = torch.from_numpy(np.ascontiguousarray(y.T)).cuda()
    tr = sbd_traces()
    rc = lib.sbd_sapg_run_dev(eng._h, y_dev.data_ptr(), None, C.byref(prm), C.byref(tr))
    if rc != 0:
        raise RuntimeError(lib.sbd_last_error(eng._h).decode())
    rate = steps * nch / tr.seconds_main
    eng.close()
    return rate


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--chains-per-gpu", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-size-sweep", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
