"""GPU parity of the PRODUCTION path of the fused Chambolle kernel.

The benchmark geometry (4096^2 x 8 chains) runs `k_chamb_multi<4>` with 128-row
segments, where almost every row goes through the two-rows-per-trip fast loop
(`cm_step2` / `fast_row`, tv_multi.cuh).  Small images get 4-row segments, where
that loop executes zero trips, so the small-shape tests in test_gpu_operators.py
do not cover it.  Here the segment length is either large by itself (1024^2 ..
4096^2) or forced (`sbd_set_option("chamb_seg", ..)`) on small images, and every
result is compared with the oracle (utils/chambolle_prox_TV_stop.m:120-149):
k exact, err 1e-7, f 1e-12, px/py 1e-11.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import rel

pytestmark = pytest.mark.gpu

TOL_F, TOL_P, TOL_E = 1e-12, 1e-11, 1e-7


@pytest.fixture(scope="module")
def sbd():
    import sbd_b200
    return sbd_b200


@pytest.fixture(scope="module")
def O():
    import oracle
    return oracle


def natural(shape, seed):
    """Piecewise-smooth test image with noise (so that the sweeps neither stop at once nor all saturate)."""
    rng = np.random.default_rng(seed)
    i, j = np.meshgrid(np.arange(shape[0]), np.arange(shape[1]), indexing="ij")
    base = 120 + 80 * np.sin(i / 37.0) * np.cos(j / 23.0) + 40 * ((i // 64 + j // 48) % 2)
    return base + rng.normal(0, 12.0, shape)


def check(eng, O, g, lam, K, tol=1e-3, tau=0.249, dual=None, want_duals=True):
    if dual is None:
        want = O.tv.chambolle_prox_TV_stop(g, "lambda", lam, "maxiter", K, "tol", tol, "tau", tau, return_info=True)
    else:
        want = O.tv.chambolle_prox_TV_stop(g, "lambda", lam, "maxiter", K, "tol", tol, "tau", tau,
                                           "dualvars", np.concatenate(dual, axis=1), return_info=True)
    fo, pxo, pyo, ko, eo = want
    f, px, py, k, err = eng.tvprox(g, lam, K, tol, tau, dualvars=dual)
    assert k == ko, (k, ko)
    assert abs(err - eo) <= TOL_E * eo + 1e-10, (err, eo)
    assert rel(f, fo) < TOL_F, rel(f, fo)
    if want_duals:
        assert rel(px, pxo) < TOL_P and rel(py, pyo) < TOL_P
    return ko


# ------------------------------------------------------------------ forced long segments on small images
@pytest.mark.parametrize("shape", [(130, 512), (256, 512), (512, 300), (56, 260), (64, 129)])
@pytest.mark.parametrize("seg", [16, 128])
def test_forced_segment_small_images(sbd, O, shape, seg):
    """Edge strips x long segments x the three odd/even tail plans, at sizes the numpy oracle does in milliseconds.
    shape = (rows = fast axis, cols = marched axis)."""
    eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=1)
    eng.set_option("chamb_seg", seg)
    geo = eng.geometry(1)
    assert geo["levels"] == 4 and geo["chamb_seg"] == seg
    g = natural(shape, 3)
    rng = np.random.default_rng(4)
    ks = set()
    for lam in (1e-3, 0.1, 2.0):
        for K in (25, 20, 22, 23):
            ks.add(check(eng, O, g, lam, K))
    dual = (rng.uniform(-0.5, 0.5, shape), rng.uniform(-0.5, 0.5, shape))
    if shape[0] == shape[1]:
        check(eng, O, g, 0.7, 25, dual=dual)
    assert max(ks) >= 20
    eng.close()


def test_forced_segment_dualvars_square(sbd, O):
    """'dualvars' warm start (chambolle_prox_TV_stop.m:99-107, square only) through the fast loop."""
    shape = (256, 256)
    eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=1)
    rng = np.random.default_rng(8)
    g = natural(shape, 5)
    dual = (rng.uniform(-0.5, 0.5, shape), rng.uniform(-0.5, 0.5, shape))
    for seg in (16, 64, 128):
        eng.set_option("chamb_seg", seg)
        for K, lam in ((25, 0.5), (20, 2.0), (10, 0.05)):
            check(eng, O, g, lam, K, dual=dual)
    eng.close()


@pytest.mark.parametrize("shape,seg", [((128, 256), 32), ((130, 300), 128), ((256, 256), 16)])
def test_stop_inside_every_block_position_long_segments(sbd, O, shape, seg):
    """The stop test firing before, inside and exactly at the end of every block of the plan (4,4,4,4,3,3,3),
    with the fast loop active: the redo launch must reproduce the reference's k, f and dual pair."""
    g = natural(shape, 6)
    eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=1)
    eng.set_option("chamb_seg", seg)
    errs = []
    for kstop in range(1, 26):
        _, _, _, _, e_k = O.tv.chambolle_prox_TV_stop(g, "lambda", 0.3, "maxiter", kstop, "tol", 0.0, return_info=True)
        errs.append(e_k)
    for kstop in range(1, 26):
        if kstop > 1 and not errs[kstop - 1] < min(errs[:kstop - 1]):
            continue                                    # err not monotone here: an earlier sweep would stop first
        tol = errs[kstop - 1] * (1 + 1e-9)
        ko = check(eng, O, g, 0.3, 25, tol=tol)
        assert ko == kstop
    eng.close()


def test_batch_mixed_stops_long_segments(sbd, O):
    """Images of a batch stopping at different sweeps (one at once, one early, one never) with seg = 64."""
    shape = (256, 256)
    from conftest import kat_image
    imgs = np.stack([kat_image(256), natural(shape, 9), np.full(shape, 2.0), natural(shape, 10) * 0.01])
    eng = sbd.Engine(256, 256, 1, 0, 0.0, max_batch=4)
    eng.set_option("chamb_seg", 64)
    f, px, py, it, err = eng.tvprox(imgs, 1e-3, 25)
    for b in range(4):
        fo, pxo, pyo, ko, eo = O.tv.chambolle_prox_TV_stop(imgs[b], "lambda", 1e-3, "maxiter", 25, return_info=True)
        assert it[b] == ko and rel(f[b], fo) < TOL_F
        assert rel(px[b], pxo) < TOL_P and rel(py[b], pyo) < TOL_P
    assert len(set(it.tolist())) >= 3
    eng.close()


# ------------------------------------------------------------------ natural geometry at large sizes
@pytest.mark.parametrize("n,segmin", [(1024, 8), (2048, 16)])
def test_large_single_image(sbd, O, n, segmin):
    eng = sbd.Engine(n, n, 1, 0, 0.0, max_batch=1)
    eng.set_option("chamb_coop", 0)         # the fused kernels (one 1024^2 image would otherwise take the cooperative prox)
    geo = eng.geometry(1)
    assert geo["levels"] == 4 and geo["chamb_seg"] >= segmin and geo["coop_blocks_per_image"] == 0, geo
    g = natural((n, n), n)
    for lam, K in ((0.1, 25), (2.0, 20), (1e-3, 25)):
        check(eng, O, g, lam, K)
    eng.close()


def test_bench_geometry_4096_batch8(sbd, O):
    """The geometry of the headline benchmark: 4096^2, batch 8 -> 128-row segments, k_chamb_multi<4> fast loop,
    through the device entry the SAPG loop uses (EMIT = 2: the tail block writes f and the duals are not stored).
    The C oracle takes ~1 s per image; two of the eight images are compared, all eight must agree on k."""
    import torch
    from sbd_b200._lib import lib
    n, B = 4096, 8
    eng = sbd.Engine(n, n, 1, 0, 0.0, max_batch=B)
    geo = eng.geometry(B)
    assert geo["levels"] == 4 and geo["chamb_seg"] == 128, geo
    base = natural((n, n), 1)
    rng = np.random.default_rng(2)
    gd = torch.empty((B, n, n), dtype=torch.float64, device="cuda")          # [b][col][row] = column-major images
    hosts = {}
    for b in range(B):
        gb = base + rng.normal(0, 1.0 + b, (n, n))
        if b in (0, 5):
            hosts[b] = gb
        gd[b].copy_(torch.from_numpy(np.ascontiguousarray(gb.T)))
    fd = torch.empty_like(gd)
    it = (C.c_int * B)(); er = (C.c_double * B)()
    for lam, K in ((0.3, 25), (2.0, 20)):
        rc = lib.sbd_tvprox_dev(eng._h, gd.data_ptr(), lam, K, 1e-3, 0.249, fd.data_ptr(), it, er, B)
        assert rc == 0, lib.sbd_last_error(eng._h)
        lib.sbd_synchronize(eng._h)
        for b, gb in hosts.items():
            fo, _, _, ko, eo = O.tv.chambolle_prox_TV_stop(gb, "lambda", lam, "maxiter", K, return_info=True)
            f = fd[b].cpu().numpy().T
            assert it[b] == ko and abs(er[b] - eo) <= TOL_E * eo
            assert rel(f, fo) < TOL_F, (b, lam, rel(f, fo))
        assert len(set(list(it))) == 1
    # in-place use is refused (the tail block writes f while other blocks still read g)
    rc = lib.sbd_tvprox_dev(eng._h, gd.data_ptr(), 0.3, 25, 1e-3, 0.249, gd.data_ptr(), it, er, B)
    assert rc != 0
    del gd, fd
    # host entry with the dual pair returned (EMIT = 1), one image, same geometry forced
    eng.set_option("geom_chains", B)
    assert eng.geometry(1)["chamb_seg"] == 128
    check(eng, O, hosts[0], 0.3, 25)
    eng.close()


def test_fused_vs_single_sweep(sbd, O):
    """the fused kernel (MUFU-seeded sqrt / reciprocal, FMA) and the single-sweep fallback (IEEE sqrt and
    division) are not bitwise equal; they must agree with each other and the oracle to ~1 ulp per sweep, and on k."""
    shape = (256, 512)
    g = natural(shape, 12)
    eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=1)
    res = {}
    for T in (4, 3, 1):
        eng.set_option("chamb_levels", T)
        assert eng.geometry(1)["levels"] == T
        res[T] = eng.tvprox(g, 0.4, 25)
    for T in (4, 3):
        assert res[T][3] == res[1][3]
        assert rel(res[T][0], res[1][0]) < 1e-13 and rel(res[T][1], res[1][1]) < 1e-12
    check(eng, O, g, 0.4, 25)
    eng.close()


def test_a_production_geometry_is_exercised(sbd):
    """Guard: fails if the shapes above stop reaching the two-rows-per-trip loop (seg >= 16 with 4 levels)."""
    for shape, batch in (((2048, 2048), 1), ((4096, 4096), 8)):
        eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=batch)
        geo = eng.geometry(batch)
        assert geo["levels"] == 4 and geo["chamb_seg"] >= 16, (shape, batch, geo)
        eng.close()


def test_geometry_independent_of_sharding(sbd):
    """inside a SAPG run the segment lengths (= summation order of the TV / err_k partial sums) come from
    the TOTAL chain count, so 1 rank x 8 chains and 4 ranks x 2 chains use the same partition."""
    eng = sbd.Engine(1024, 1024, 7, 0, 0.0, max_batch=8)
    eng.set_option("geom_chains", 8)
    want = eng.geometry(8)
    for local in (1, 2, 4):
        got = eng.geometry(local)
        assert (got["chamb_seg"], got["tv_seg"]) == (want["chamb_seg"], want["tv_seg"])
    eng.set_option("geom_chains", 0)
    assert eng.geometry(2)["chamb_seg"] != want["chamb_seg"] or eng.geometry(2)["tv_seg"] != want["tv_seg"]
    eng.close()


# ------------------------------------------------------------------ sampled stop test (ERRSUB)
def _dev_prox(eng, g, lam, K, tol, want_err):
    """sbd_tvprox_dev on a batch; err == NULL lets the engine use the sampled stop test."""
    import torch
    from sbd_b200._lib import lib
    B = g.shape[0]
    gd = torch.from_numpy(np.ascontiguousarray(g.transpose(0, 2, 1))).cuda()
    fd = torch.empty_like(gd)
    it = (C.c_int * B)(); er = (C.c_double * B)()
    rc = lib.sbd_tvprox_dev(eng._h, gd.data_ptr(), lam, K, tol, 0.249, fd.data_ptr(), it, er if want_err else None, B)
    assert rc == 0, lib.sbd_last_error(eng._h)
    lib.sbd_synchronize(eng._h)
    return fd.cpu().numpy().transpose(0, 2, 1), list(it)


@pytest.mark.parametrize("shape,seg", [((128, 256), 32), ((130, 300), 128)])
def test_sampled_stop_test_is_exact(sbd, O, shape, seg):
    """ERRSUB: err_k^2 summed over a subset of the rows is a lower bound, so 'sampled sum > tol^2' is the reference's
    own decision; when it proves nothing the block is recomputed exactly.  For every stop position 1..25 (and no stop)
    the sweep count must be the oracle's and f bit-identical to the run with the full sums."""
    g = natural(shape, 21)
    gb = np.stack([g, natural(shape, 22)])
    eng = sbd.Engine(shape[0], shape[1], 1, 0, 0.0, max_batch=2)
    eng.set_option("chamb_seg", seg)
    errs = []
    for kstop in range(1, 26):
        _, _, _, _, e_k = O.tv.chambolle_prox_TV_stop(g, "lambda", 0.3, "maxiter", kstop, "tol", 0.0, return_info=True)
        errs.append(e_k)
    tols = [errs[k - 1] * (1 + 1e-9) for k in range(1, 26) if k == 1 or errs[k - 1] < min(errs[:k - 1])] + [1e-3, errs[-1] * 0.5]
    for K in (25, 22):
        for tol in tols:
            eng.set_option("chamb_errsub", 0)
            f_full, k_full = _dev_prox(eng, gb, 0.3, K, tol, want_err=False)
            eng.set_option("chamb_errsub", 1)
            f_sub, k_sub = _dev_prox(eng, gb, 0.3, K, tol, want_err=False)
            assert k_sub == k_full, (K, tol, k_sub, k_full)
            assert np.array_equal(f_sub, f_full)
            for b in range(2):
                fo, _, _, ko, _ = O.tv.chambolle_prox_TV_stop(gb[b], "lambda", 0.3, "maxiter", K, "tol", tol, return_info=True)
                assert k_sub[b] == ko and rel(f_sub[b], fo) < TOL_F
    # the entry that returns err always runs the full sums, whatever the option says
    f, px, py, k, err = eng.tvprox(g, 0.3, 25)
    _, _, _, ko, eo = O.tv.chambolle_prox_TV_stop(g, "lambda", 0.3, "maxiter", 25, return_info=True)
    assert k == ko and abs(err - eo) <= TOL_E * eo
    eng.close()


def test_sampled_stop_test_in_sapg_laplace(sbd, O, boat):
    """SAPG Laplace (lambda*theta starts at 1e-3: the stop test fires in the prox) with the sampled test forced on:
    same sweep counts and trajectories as the oracle."""
    x = boat[128:384, 128:384]
    rng = np.random.default_rng(31)
    tape = []

    def randn(shape):
        z = rng.standard_normal(shape); tape.append(z); return z

    y, op = O.operators.setup_demo(2, x, randn, samples=10, warmup=4, burnIn=6)
    del tape[:]
    th, b, s2, r = O.sapg.SAPG_algorithm_laplace(y, op, randn)
    noise = np.stack(tape)[:, None]
    eng = sbd.Engine(256, 256, 7, 2, 0.0, max_batch=1)
    eng.set_option("chamb_errsub", 1)
    eng.set_option("chamb_seg", 32)
    _, _, _, g = sbd.SAPG_algorithm_laplace(y, op, noise=noise, engine=eng)
    for k in ("thetas", "bs", "sigmas", "logPiTraceX", "gXTrace"):
        assert rel(g[k], r[k]) < 1e-6, k
    assert rel(g["X_sample"], r["X_sample"]) < 1e-6
    assert g["chambolle_iters"].min() < 25
    eng.set_option("chamb_errsub", 0)
    _, _, _, g0 = sbd.SAPG_algorithm_laplace(y, op, noise=noise, engine=eng)
    assert np.array_equal(g0["chambolle_iters"], g["chambolle_iters"]) and np.array_equal(g0["thetas"], g["thetas"])
    eng.close()


def test_set_option_rejects_unknown_names(sbd):
    from sbd_b200 import SbdError
    eng = sbd.Engine(64, 64, 1, 0, 0.0, max_batch=1)
    with pytest.raises(SbdError):
        eng.set_option("no_such_option", 1)
    for name in ("chamb_seg", "chamb_levels", "chamb_emit", "chamb_plan33", "chamb_errsub", "chamb_coop", "tv_seg", "geom_chains"):
        eng.set_option(name, -1 if name != "geom_chains" else 0)
    assert eng.geometry(1)["levels"] == 4
    eng.close()
