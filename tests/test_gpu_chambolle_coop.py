"""GPU parity of the cooperative single-launch Chambolle prox (csrc/tv_coop.cuh), the path small problems take
(rows*cols*chains <= 2^21: the reference's own 256^2 / 512^2 images).

One warp per (row, 64-pixel strip), all sweeps of a prox in one launch, a barrier between the blocks of an image per
sweep, the reference's stop test (utils/chambolle_prox_TV_stop.m:131) decided one sweep late from the previous sweep's
partial sums (the surplus sweep goes to the other buffer).  Compared with the oracle (k exact, err 1e-7, f 1e-12,
px/py 1e-11) and with the fused marching kernels of tv_multi.cuh, which share the per-pixel arithmetic.
"""
import numpy as np
import pytest

from conftest import rel
from test_gpu_chambolle_prod import natural, check

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sbd():
    import sbd_b200
    return sbd_b200


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def coop_engine(sbd, shape, batch=1):
    eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=batch)
    eng.set_option("chamb_coop", 1)
    geo = eng.geometry(batch)
    assert geo["coop_blocks_per_image"] > 0, geo
    return eng


def test_small_problems_take_the_cooperative_prox_by_default(sbd):
    for shape, batch, want in (((256, 256), 1, True), ((512, 512), 8, True), ((1024, 1024), 2, True),
                               ((1024, 1024), 8, False), ((2048, 2048), 1, False), ((255, 256), 1, False)):
        eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=batch)
        geo = eng.geometry(batch)
        assert (geo["coop_blocks_per_image"] > 0) == want, (shape, batch, geo)
        eng.set_option("chamb_seg", 16)             # an option that addresses the fused kernel switches it off
        assert eng.geometry(batch)["coop_blocks_per_image"] == 0
        eng.close()


# shape = (rows = fast axis, cols): strips of 64 pixels -> partial last strip (70, 130), one strip (8, 64), several
@pytest.mark.parametrize("shape", [(8, 8), (64, 32), (70, 33), (130, 257), (256, 256), (512, 300), (2, 5), (66, 2)])
def test_against_the_oracle(sbd, O, shape):
    eng = coop_engine(sbd, shape)
    g = natural(shape, 11)
    ks = set()
    for lam in (1e-3, 0.1, 2.0):
        for K in (25, 1, 2, 7):
            ks.add(check(eng, O, g, lam, K))
    eng.close()


def test_dualvars_warm_start(sbd, O):
    shape = (256, 256)
    eng = coop_engine(sbd, shape)
    rng = np.random.default_rng(8)
    g = natural(shape, 5)
    dual = (rng.uniform(-0.5, 0.5, shape), rng.uniform(-0.5, 0.5, shape))
    for K, lam in ((25, 0.5), (20, 2.0), (1, 0.05)):
        check(eng, O, g, lam, K, dual=dual)
    eng.close()


@pytest.mark.parametrize("shape", [(128, 256), (130, 300)])
def test_stop_at_every_sweep(sbd, O, shape):
    """The stop test firing at every sweep 1..25: the decision is taken during the NEXT sweep, which must leave the
    stopped dual pair untouched; k, err, f, px, py are the reference's."""
    g = natural(shape, 6)
    eng = coop_engine(sbd, shape)
    errs = []
    for kstop in range(1, 26):
        _, _, _, _, e_k = O.tv.chambolle_prox_TV_stop(g, "lambda", 0.3, "maxiter", kstop, "tol", 0.0, return_info=True)
        errs.append(e_k)
    hit = 0
    for kstop in range(1, 26):
        if kstop > 1 and not errs[kstop - 1] < min(errs[:kstop - 1]):
            continue                                    # err not monotone here: an earlier sweep would stop first
        tol = errs[kstop - 1] * (1 + 1e-9)
        assert check(eng, O, g, 0.3, 25, tol=tol) == kstop
        hit += 1
    assert hit >= 20
    eng.close()


def test_batch_mixed_stops(sbd, O):
    """Images of a batch are independent (a barrier per image): one stops at once, one early, one never."""
    shape, B = (128, 192), 4
    eng = coop_engine(sbd, shape, B)
    base = natural(shape, 3)
    g = np.stack([np.full(shape, 7.0), base, base * 0.05, base + 30.0])
    want = [O.tv.chambolle_prox_TV_stop(g[b], "lambda", 0.3, "maxiter", 25, "tol", 2.0, return_info=True) for b in range(B)]
    f, px, py, k, err = eng.tvprox(g, 0.3, 25, 2.0, 0.249)
    ks = [w[3] for w in want]
    assert list(k) == ks and len(set(ks)) >= 3, (list(k), ks)
    for b in range(B):
        assert rel(f[b], want[b][0]) < 1e-12 and rel(px[b], want[b][1]) < 1e-11 and rel(py[b], want[b][2]) < 1e-11
        assert abs(err[b] - want[b][4]) <= 1e-7 * want[b][4] + 1e-10
    eng.close()


@pytest.mark.parametrize("shape,batch", [((256, 256), 1), ((130, 300), 3), ((512, 512), 2)])
def test_same_dual_pair_as_the_fused_kernels(sbd, shape, batch):
    """Both paths run cm_core on the same operands: the dual pair is bit-identical, k equal, f within an ulp or two
    (the output is formed with a fused multiply-add in one and not in the other)."""
    rng = np.random.default_rng(5)
    g = np.stack([natural(shape, 20 + b) for b in range(batch)]) if batch > 1 else natural(shape, 20)
    res = {}
    for mode in (0, 1):
        eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=batch)
        eng.set_option("chamb_coop", mode)
        assert (eng.geometry(batch)["coop_blocks_per_image"] > 0) == bool(mode)
        res[mode] = eng.tvprox(g, 0.4, 25, 1e-3, 0.249)
        eng.close()
    f0, px0, py0, k0, e0 = res[0]
    f1, px1, py1, k1, e1 = res[1]
    assert np.array_equal(np.asarray(k0), np.asarray(k1))
    assert np.array_equal(px0, px1) and np.array_equal(py0, py1)
    assert rel(f1, f0) < 1e-14
    assert np.allclose(e0, e1, rtol=1e-12)


def test_partition_follows_the_total_chain_count(sbd):
    """Like the rest of the launch geometry: blocks per image / units per warp (= the order of the err_k partial sums)
    come from the total chain count of a sharded run, so 8 chains on one rank and 2 x 4 give bit-identical results."""
    shape = (256, 256)
    g = np.stack([natural(shape, 40 + b) for b in range(8)])
    eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=8)
    whole = eng.tvprox(g, 0.4, 25, 1e-3, 0.249)
    geo8 = eng.geometry(8)
    eng.close()
    eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=4)
    eng.set_option("geom_chains", 8)
    geo4 = eng.geometry(4)
    assert geo4["coop_blocks_per_image"] == geo8["coop_blocks_per_image"] > 0
    assert geo4["coop_units_per_warp"] == geo8["coop_units_per_warp"]
    for half in (0, 1):
        part = eng.tvprox(g[4 * half:4 * half + 4], 0.4, 25, 1e-3, 0.249)
        for a, b in zip(part, whole):
            assert np.array_equal(np.asarray(a), np.asarray(b)[4 * half:4 * half + 4])
    eng.close()


def test_sapg_same_trajectory_as_the_fused_kernels(sbd):
    """A SAPG run (graph replay, two streams, sweep counts recorded by the kernel itself) with the cooperative prox
    against the same run on the fused kernels: same sweep counts, trajectories equal to rounding."""
    from sbd_b200 import host as H
    import bench
    n = 128
    x = bench.synthetic_truth(n)
    out = {}
    for mode in (0, 1):
        eng = sbd.Engine(n, n, 7, H.GAUSSIAN, 0.0, 2, 0)
        eng.set_option("chamb_coop", mode)
        Ax = eng.blur(x, (0.4, 0.3), H.OP_A)
        nrm = float(np.linalg.norm(Ax - Ax.mean()))
        sig = lambda b: nrm / np.sqrt(n * n * 10 ** (b / 10))
        y = Ax + sig(30) * np.random.default_rng(2).standard_normal((n, n))
        op, c = bench.gaussian_op(n, sig(30), sig(15), sig(45), bench.EVMAX, 60, 30)
        th, w1, w2, s2, r = sbd.SAPG_algorithm_Guassian(y, op, c, n_chains=2, seed=3, engine=eng)
        out[mode] = (np.asarray(r["thetas"]), np.asarray(r["w1s"]), np.asarray(r["sigmas"]), np.asarray(r["chambolle_iters"]))
        eng.close()
    assert np.array_equal(out[0][3], out[1][3])
    for a, b in zip(out[0][:3], out[1][:3]):
        assert rel(b, a) < 1e-9
