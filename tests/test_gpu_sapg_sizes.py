"""GPU parity of full SAPG runs AT THE SIZES BASELINE.json NAMES, against the oracle fed the same noise tape:
  configs[0]  run_Gaussian_demo.m  cameraman 256x256, Gaussian PSF, theta / sigma^2 / w1 / w2 estimated
  configs[1]  run_moffat_demo.m    boat 512x512, Moffat PSF (alpha, beta)
  configs[2]  run_laplace_demo.m   512x512 image, Laplace PSF (the Chambolle stop test fires all the time)
  configs[3]  synthetic 4096x4096 Gaussian-PSF SAPG (the geometry of the headline benchmark), two main-loop steps
Tolerance (north_star): every trajectory field and the last sample to relative error 1e-6.
The oracle runs with scipy.fft (all cores) and the plain-C TV prox (oracle/c, bit-identical to the numpy one)."""
import os

import numpy as np
import pytest

from conftest import rel

pytestmark = pytest.mark.gpu

TRAJ_TOL = 1e-6


@pytest.fixture(scope="module")
def sbd():
    import sbd_b200
    return sbd_b200


@pytest.fixture(scope="module")
def O():
    import oracle
    import scipy.fft
    w = os.cpu_count() or 1
    oracle.operators.set_fft(lambda a: scipy.fft.fft2(a, workers=w), lambda a: scipy.fft.ifft2(a, workers=w))
    yield oracle
    oracle.operators.set_fft(np.fft.fft2, np.fft.ifft2)


class NoiseTape:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)
        self.tape = []
        self.record = False

    def __call__(self, shape):
        z = self.rng.standard_normal(shape)
        if self.record:
            self.tape.append(z)
        return z


def _traj(got, want, names):
    for n in names:
        a, b = np.asarray(got[n]), np.asarray(want[n])
        assert a.shape == b.shape, (n, a.shape, b.shape)
        assert rel(a, b) < TRAJ_TOL, (n, rel(a, b))


COMMON = ["thetas", "sigmas", "logPiTraceX", "gXTrace", "logPiTrace_WU", "err_psf"]


def test_config0_gaussian_cman_256(sbd, O, cman):
    """SAPG_algorithm_Guassian.m:158-248 on the full cameraman image, fix_w1 = fix_w2 = 0."""
    tape = NoiseTape(11)
    y, op, c = O.operators.setup_demo(0, cman, tape, samples=7, warmup=4, burnIn=4, fix_w1=0, fix_w2=0)
    tape.record = True
    th, w1, w2, s2, r = O.sapg.SAPG_algorithm_Guassian(y, op, c, tape)
    noise = np.stack(tape.tape)[:, None]
    gth, gw1, gw2, gs2, g = sbd.SAPG_algorithm_Guassian(y, dict(op, use_graph=0), c, noise=noise)
    _traj(g, r, COMMON + ["w1s", "w2s", "grad_theta", "grad_w1", "grad_w2", "grad_sigma", "mean_thetas", "mean_w1s"])
    assert rel(g["Xlast_sample"], r["Xlast_sample"]) < TRAJ_TOL
    for a, b in ((gth, th), (gw1, w1), (gw2, w2), (gs2, s2)):
        assert abs(a - b) <= TRAJ_TOL * abs(b)
    # and the same run replayed from a CUDA graph (what the full 35k-step demo uses)
    _, _, _, _, g2 = sbd.SAPG_algorithm_Guassian(y, dict(op, use_graph=1), c, noise=noise)
    for k in ("thetas", "w1s", "w2s", "sigmas"):
        assert np.array_equal(g[k], g2[k])


def test_config1_moffat_boat_512(sbd, O, boat):
    """SAPG_algorithm_moffat.m:159-209 on the full boat image."""
    tape = NoiseTape(12)
    y, op = O.operators.setup_demo(1, boat, tape, samples=7, warmup=4, burnIn=4)
    tape.record = True
    th, a, b, s2, r = O.sapg.SAPG_algorithm_moffat(y, op, tape)
    noise = np.stack(tape.tape)[:, None]
    gth, ga, gb, gs2, g = sbd.SAPG_algorithm_moffat(y, op, noise=noise)
    _traj(g, r, COMMON + ["alphas", "betas", "mean_alphas", "mean_betas"])
    assert rel(g["Xlast_sample"], r["Xlast_sample"]) < TRAJ_TOL
    assert rel(g["X_warm"], r["X_warm"]) < TRAJ_TOL
    for u, v in ((gth, th), (ga, a), (gb, b), (gs2, s2)):
        assert abs(u - v) <= TRAJ_TOL * abs(v)


def test_config2_laplace_512(sbd, O, boat):
    """SAPG_algorithm_laplace.m:154-195 on a 512x512 image; lambda*theta starts at 1e-3, so the
    Chambolle stop test fires early in every prox (the redo-launch path)."""
    tape = NoiseTape(13)
    y, op = O.operators.setup_demo(2, boat, tape, samples=7, warmup=4, burnIn=4)
    tape.record = True
    th, b, s2, r = O.sapg.SAPG_algorithm_laplace(y, op, tape)
    noise = np.stack(tape.tape)[:, None]
    gth, gb, gs2, g = sbd.SAPG_algorithm_laplace(y, op, noise=noise)
    _traj(g, r, COMMON + ["bs", "err_sample", "err_warm"])
    assert rel(g["X_sample"], r["X_sample"]) < TRAJ_TOL
    assert g["chambolle_iters"].min() < 25          # the stop test fires (redo-launch path)
    for u, v in ((gth, th), (gb, b), (gs2, s2)):
        assert abs(u - v) <= TRAJ_TOL * abs(v)


def test_config3_gaussian_4096_two_steps(sbd, O, cman):
    """Two main-loop steps (plus one warm-up step) of the benchmark problem itself: 4096x4096 synthetic image
    (cameraman tiled), Gaussian PSF, K = 25 Chambolle sweeps with 128-row segments when the
    geometry is that of the 8-chain benchmark.  ~4 s per oracle iteration with the C prox."""
    n = 4096
    x = np.tile(cman, (n // 256, n // 256))
    tape = NoiseTape(14)
    y, op, c = O.operators.setup_demo(0, x, tape, samples=3, warmup=2, burnIn=2, fix_w1=0, fix_w2=0, evMax=0.993)
    tape.record = True
    th, w1, w2, s2, r = O.sapg.SAPG_algorithm_Guassian(y, op, c, tape)
    noise = np.stack(tape.tape)[:, None]
    eng = sbd.Engine(n, n, 7, 0, 0.0, max_batch=1)
    eng.set_option("geom_chains", 8)                      # segment lengths of the 8-chain benchmark geometry
    assert eng.geometry(1)["chamb_seg"] == 128
    _, _, _, _, g = sbd.SAPG_algorithm_Guassian(y, op, c, noise=noise, engine=eng)
    _traj(g, r, ["thetas", "sigmas", "w1s", "w2s", "logPiTraceX", "gXTrace", "logPiTrace_WU",
                 "grad_theta", "grad_w1", "grad_w2", "grad_sigma"])
    assert rel(g["Xlast_sample"], r["Xlast_sample"]) < TRAJ_TOL
    assert int(g["chambolle_iters"][1]) == 25
    eng.close()
