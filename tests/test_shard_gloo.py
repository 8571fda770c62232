"""Multi-rank host logic on CPU (gloo, world_size 2): chains sharded over ranks,
per-chain sums gathered in global chain order -> every rank computes the same
trajectory as a single process running all the chains."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_chain_shard_partition():
    sys.path.insert(0, ROOT)
    from sbd_b200.shard import ChainShard
    s = ChainShard(64, 8, 3)
    assert s.n_local == 8 and s.chain_offset == 24 and list(s.local_chains()) == list(range(24, 32))
    assert s.owner(25) == 3 and s.owner(63) == 7
    with pytest.raises(ValueError):
        ChainShard(10, 4, 0)
    one = ChainShard(5, 1, 0)
    tot = one.combine([np.array([1.0, 2.0])] * 5)
    assert np.array_equal(tot, [5.0, 10.0])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from oracle import operators as OP, sapg, philox
    from sbd_b200.shard import ChainShard
    x = np.load(os.path.join(ROOT, "tests", "golden", "cman_u8.npy")).astype(np.float64)[100:132, 60:92]
    rng = np.random.default_rng(11)
    y, op, c = OP.setup_demo(0, x, lambda s: rng.standard_normal(s), samples=8, warmup=3, burnIn=5,
                             fix_w1=0, fix_w2=0, evMax=0.993)
    total = 4
    shard = ChainShard(total, world, rank)
    step = {}

    def randn_chain(local_ch, shape):               # noise is keyed by the GLOBAL chain id
        ch = shard.chain_offset + local_ch
        s = step.get(ch, 0); step[ch] = s + 1
        return philox.randn_image(shape, 5, ch, s)

    # every rank runs only its own chains; the combine hook all-gathers the per-chain sums
    out = sapg.sapg_multichain(0, y, op, c, randn_chain, shard.n_local,
                               combine=lambda vs: shard.combine(vs) * (shard.n_local / total))
    q.put((rank, out["thetas"], out["sigmas"], out["psis"]))
    dist.destroy_process_group()


def test_two_ranks_reproduce_single_process():
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    import oracle
    from oracle import operators as OP, sapg, philox
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process run of all 4 chains
    x = np.load(os.path.join(ROOT, "tests", "golden", "cman_u8.npy")).astype(np.float64)[100:132, 60:92]
    rng = np.random.default_rng(11)
    y, op, c = OP.setup_demo(0, x, lambda s: rng.standard_normal(s), samples=8, warmup=3, burnIn=5,
                             fix_w1=0, fix_w2=0, evMax=0.993)
    step = {}

    def randn_chain(ch, shape):
        s = step.get(ch, 0); step[ch] = s + 1
        return philox.randn_image(shape, 5, ch, s)

    want = sapg.sapg_multichain(0, y, op, c, randn_chain, 4)
    for rank, th, s2, psis in res:
        assert np.array_equal(th, res[0][1]) and np.array_equal(s2, res[0][2])      # ranks bit-identical
        assert np.allclose(th, want["thetas"], rtol=1e-13, atol=0)
        assert np.allclose(s2, want["sigmas"], rtol=1e-13, atol=0)
        assert np.allclose(psis, want["psis"], rtol=1e-12, atol=0)
