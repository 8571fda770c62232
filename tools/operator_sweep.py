#!/usr/bin/env python
"""BASELINE.json configs[4]: operator sweep - fused FFT blur A / A' and the
20-iteration Chambolle TV prox on a batch of images at 256^2 .. 4096^2, device
resident (sbd_blur_dev / sbd_tvprox_dev), timed with CUDA events.  The batch of
256 images is processed in sub-batches that fit HBM.

    python tools/operator_sweep.py [--batch 256] [--sizes 256 512 1024 2048 4096]

Prints one JSON line per size: images/s, algorithmic GB/s (A or A' = 64 P bytes,
Chambolle-20 prox = 832 P bytes) and the fraction of the measured HBM peak.
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sizes", type=int, nargs="+", default=[256, 512, 1024, 2048, 4096])
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    import torch
    import sbd_b200
    from sbd_b200 import host as H
    from sbd_b200._lib import lib, c_double_p
    peak = 6554.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    for n in a.sizes:
        npix = n * n
        sub = max(1, min(a.batch, int(12e9 // (npix * 8 * 11))))       # ~11 arrays per image in flight
        eng = sbd_b200.Engine(n, n, 7, H.MOFFAT, 0.0, sub, 0)
        x = torch.rand(sub, n, n, dtype=torch.float64, device="cuda") * 255.0
        out = torch.empty_like(x)
        psi = (C.c_double * 2)(0.4, 3.5)
        nsub = (a.batch + sub - 1) // sub

        def timed(fn):
            fn()                                                        # warm-up
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(a.reps):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                lib.sbd_synchronize(eng._h)
                e0.record()
                for _ in range(nsub):
                    fn()
                lib.sbd_synchronize(eng._h)                             # the library runs on its own stream
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e-3)
            return best

        res = {"size": n, "batch": a.batch, "sub_batch": sub}
        for name, opsel in (("A", 0), ("AT", 1)):
            def run():
                rc = lib.sbd_blur_dev(eng._h, x.data_ptr(), psi, opsel, out.data_ptr(), sub)
                assert rc == 0, lib.sbd_last_error(eng._h)
            t = timed(run)
            gbs = 64.0 * npix * nsub * sub / t / 1e9
            res[name] = {"images_per_s": nsub * sub / t, "alg_GBps": gbs, "frac_hbm_peak": gbs / peak}

        def prox():
            rc = lib.sbd_tvprox_dev(eng._h, x.data_ptr(), 5.0, 20, 1e-3, 0.249, out.data_ptr(), None, None, sub)
            assert rc == 0, lib.sbd_last_error(eng._h)
        t = timed(prox)
        gbs = 832.0 * npix * nsub * sub / t / 1e9
        res["chambolle20"] = {"images_per_s": nsub * sub / t, "alg_GBps": gbs, "frac_hbm_peak": gbs / peak}
        print(json.dumps(res), flush=True)
        eng.close()
        del x, out
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
