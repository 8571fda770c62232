#!/usr/bin/env python
"""Cost of one sweep of the Chambolle prox on small images: time of sbd_tvprox_dev (device buffers, no read-back) at
maxiter 1 / 25 / 49 with tol = 0, for the cooperative kernel and the fused kernels; the slope is the cost per sweep.
    python tools/coop_sweep_cost.py [sizes...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sbd_b200
from sbd_b200._lib import lib
import bench

sizes = [int(a) for a in sys.argv[1:]] or [256, 512]
for n in sizes:
    for batch in (1, 8):
        g = torch.from_numpy(np.stack([bench.synthetic_truth(n)] * batch)).cuda()
        f = torch.empty_like(g)
        for mode in (0, 1):
            eng = sbd_b200.Engine(n, n, 1, 0, 0.0, max_batch=batch)
            eng.set_option("chamb_coop", mode)
            t = {}
            for K in (1, 25, 49):
                for rep in range(2):
                    torch.cuda.synchronize(); t0 = time.perf_counter()
                    N = 300
                    for _ in range(N):
                        rc = lib.sbd_tvprox_dev(eng._h, g.data_ptr(), 0.3, K, 0.0, 0.249, f.data_ptr(), None, None, batch)
                        assert rc == 0
                    torch.cuda.synchronize(); t[K] = (time.perf_counter() - t0) / N * 1e6
            print(f"{n}^2 x {batch} {'coop ' if mode else 'fused'}: K=1 {t[1]:7.1f} us  K=25 {t[25]:7.1f} us  K=49 {t[49]:7.1f} us   per sweep {(t[49]-t[25])/24:6.2f} us", flush=True)
            eng.close()
