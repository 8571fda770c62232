#!/usr/bin/env python
"""Multi-GPU check (run under torchrun, one rank per GPU): chains sharded over
ranks with the NCCL all-gather inside libsbd must give trajectories that are
BIT-IDENTICAL on every rank and to a single-GPU run of all the chains.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29512 tools/check_multigpu.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--per", type=int, default=2, help="chains per rank")
    ap.add_argument("--samples", type=int, default=30)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import sbd_b200
    from sbd_b200 import host as H
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, per = a.size, a.per
    total = per * world
    cm = np.load(os.path.join(ROOT, "tests", "golden", "cman_u8.npy")).astype(np.float64)
    x = cm[64:64 + n, 64:64 + n] if n <= 128 else np.tile(cm, ((n + 255) // 256, (n + 255) // 256))[:n, :n].copy()
    eng = sbd_b200.Engine(n, n, 7, H.MOFFAT, 0.0, total, local)
    Ax = eng.blur(x, (0.4, 3.5), H.OP_A)
    nrm = float(np.linalg.norm(Ax - Ax.mean()))
    sig = lambda b: nrm / np.sqrt(n * n * 10 ** (b / 10))
    y = Ax + sig(30) * np.random.default_rng(2).standard_normal((n, n))
    op = dict(samples=a.samples, warmup=8, burnIn=max(2, a.samples * 2 // 3), psf_size=7, min_th=1e-3, max_th=1.0, th_init=0.01,
              alpha_init=1.0, beta_init=10.0, min_alpha=1e-2, max_alpha=1.0, min_beta=0.1, max_beta=10.0,
              alpha=0.4, beta=3.5, fix_alpha=0, fix_beta=0, fix_sigma=0, d_exp=0.8, d_scale=1.0,
              sigma=sig(30), sigma_init=(sig(18) ** 2 + sig(35) ** 2) / 2, sigma_min=sig(18) ** 2, sigma_max=sig(35) ** 2)
    Lf = min(0.993 ** 2 / op["sigma_min"], 0.993 ** 2 / op["sigma_max"])
    op["lambda"] = min(5 / Lf, 2.0); op["gamma"] = 0.98 / (Lf + 1 / op["lambda"])

    shard = sbd_b200.ChainShard(total, world, rank)
    shard.init_engine_comm(eng)
    prm = H.make_params(H.MOFFAT, op, None, n_chains=per, seed=7, chain_offset=shard.chain_offset, total_chains=total)
    out = eng.sapg(y, prm)
    keys = ("thetas", "sigmas", "psi0", "psi1", "logPiTraceX")
    mine = np.stack([out[k] for k in keys])
    ok = True
    if world > 1:
        t = torch.from_numpy(mine).cuda()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        same = all(torch.equal(parts[0], p) for p in parts)
        if rank == 0:
            print("ranks bit-identical:", same)
        ok &= same
    # every rank tears its NCCL communicator down (the teardown is collective)
    sbd_b200._lib.lib.sbd_comm_destroy(eng._h)
    if rank == 0:
        prm1 = H.make_params(H.MOFFAT, op, None, n_chains=total, seed=7, chain_offset=0, total_chains=total)
        ref = eng.sapg(y, prm1)
        one = np.stack([ref[k] for k in keys])
        eq = np.array_equal(one, mine)
        print("multi-GPU == single-GPU (bitwise):", eq, " max rel diff:", float(np.max(np.abs(one - mine) / np.abs(one))))
        print("theta_EB", out["EB"][0], "alpha_EB", out["EB"][1], "beta_EB", out["EB"][2], "sigma2_EB", out["EB"][3])
        # each rank's chains differ from each other and from the other ranks' chains
        ok &= eq
        print("CHECK", "PASS" if ok else "FAIL")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
