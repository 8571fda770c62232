#!/usr/bin/env python
"""Where does the host-pointer call (sbd_sapg_run) spend its time beyond the MYULA steps?
Times the call for several run lengths, with and without the last-sample read-back and the CUDA graph.

    python tools/e2e_probe.py [--size 4096] [--chains 8]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=4096)
    ap.add_argument("--chains", type=int, default=8)
    a = ap.parse_args()
    import torch
    import bench as B
    import sbd_b200
    from sbd_b200 import host as H
    from sbd_b200._lib import lib, sbd_traces, c_double_p
    n, nch = a.size, a.chains
    npix = n * n
    eng = sbd_b200.Engine(n, n, 7, H.GAUSSIAN, 0.0, nch, 0)
    x_true = B.synthetic_truth(n)
    Ax = eng.blur(x_true, B.PSI_TRUE, H.OP_A)
    nrm = float(np.linalg.norm(Ax - Ax.mean()))
    sig = lambda b: nrm / np.sqrt(npix * 10 ** (b / 10))
    y = Ax + sig(30) * np.random.default_rng(2).standard_normal((n, n))
    y_pin = torch.from_numpy(np.ascontiguousarray(y.T)).pin_memory()
    xl_pin = torch.empty(nch * npix, dtype=torch.float64).pin_memory()
    for K in (5, 20, 40):
        for graph in (0, 1):
            for xlast in (0, 1):
                op, c = B.gaussian_op(n, sig(30), sig(15), sig(45), 0.993, K + 1, 1)
                prm = H.make_params(H.GAUSSIAN, op, c, n_chains=nch, seed=1, chain_offset=0, total_chains=nch)
                prm.use_graph = graph
                tr = sbd_traces()
                th = np.zeros(K + 1); tr.thetas = th.ctypes.data_as(c_double_p)
                if xlast:
                    tr.X_last = C.cast(xl_pin.data_ptr(), c_double_p)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rc = lib.sbd_sapg_run(eng._h, C.cast(y_pin.data_ptr(), c_double_p), None, None, C.byref(prm), None, C.byref(tr))
                assert rc == 0, lib.sbd_last_error(eng._h)
                wall = time.perf_counter() - t0
                print(f"K={K:3d} graph={graph} X_last={xlast}: wall {wall*1e3:8.1f} ms  device total {tr.seconds*1e3:8.1f} ms  "
                      f"main loop {tr.seconds_main*1e3:8.1f} ms  -> outside the device loop {1e3*(wall - tr.seconds):7.1f} ms", flush=True)


if __name__ == "__main__":
    main()
