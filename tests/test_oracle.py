"""CPU tests of the oracle itself (no GPU): definitional pins derived from the
reference's own definitions and known-answer values.  The golden
fixtures produced by executing the reference .m files are checked in
tests/test_oracle_golden.py."""
import os

import numpy as np
import pytest

from conftest import rel, kat_image
import oracle
from oracle import psf as P, tv, operators as OP, sapg, philox, metrics

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_psf_normalised_and_known_taps():
    k = P.psf_gaussian(7, 0.4, 0.3, 0.0)
    assert abs(k.sum() - 1.0) < 1e-15                           # Gaussian_psf.m:18
    assert abs(k[3, 3] - 0.032059565361197) < 1e-14
    assert abs(k[0, 0] - 1.040821699694877e-02) < 1e-14
    assert abs(P.psf_moffat(7, 0.4, 3.5).sum() - 1.0) < 1e-15
    assert abs(P.psf_laplace(7, 0.3).sum() - 1.0) < 1e-15
    # w1 acts along columns (u), w2 along rows (v)            (Q19, Gaussian_psf.m:9-14)
    k2 = P.psf_gaussian(7, 0.9, 0.1, 0.0)
    assert k2[3, 0] < k2[0, 3]                                  # large w1 -> fast decay along the columns


def test_resize_is_topleft_pad_and_matches_closed_form():
    k = P.psf_moffat(7, 0.4, 3.5)
    M, N = 32, 64
    H = P.resize(k, (M, N))
    EM = np.exp(-2j * np.pi * np.outer(np.arange(M), np.arange(7)) / M)
    EN = np.exp(-2j * np.pi * np.outer(np.arange(N), np.arange(7)) / N)
    assert rel(H, EM @ k @ EN.T) < 1e-14                        # H = E h E'
    back = np.real(np.fft.ifft2(H))
    assert rel(back[:7, :7], k) < 1e-13 and np.abs(back[7:, :]).max() < 1e-15   # Q1


@pytest.mark.parametrize("model,psi", [(0, (0.4, 0.3)), (1, (0.4, 3.5)), (2, (0.3,))])
def test_derivatives_vs_finite_differences(model, psi):
    """Analytic d(normalised PSF)/d(param) vs central differences.  Moffat-alpha
    is EXPECTED to disagree: the reference has a stray 2 (Q7) which we reproduce."""
    h = 1e-6
    for i in range(len(psi)):
        up = list(psi); dn = list(psi); up[i] += h; dn[i] -= h
        fd = (P.taps(model, 7, up) - P.taps(model, 7, dn)) / (2 * h)
        an = P.taps(model, 7, psi, 0.0, which=i + 1)
        err = np.abs(fd - an).max()
        if model == 1 and i == 0:
            assert err > 1e-3                                   # reference formula, not the true derivative
        else:
            assert err < 1e-9


def test_conv2c_literal_equals_roll():
    rng = np.random.default_rng(0)
    for shape in [(5, 7), (16, 16), (33, 20)]:
        x = rng.standard_normal(shape)
        assert np.array_equal(tv.diffh(x), tv.diffh_literal(x))
        assert np.array_equal(tv.diffv(x), tv.diffv_literal(x))


def test_tvnorm_cman(cman):
    assert abs(tv.TVnorm(cman) - 1115956.0628163717) < 1e-6


def test_chambolle_known_answers(cman):
    f, px, py, k, err = tv.chambolle_prox_TV_stop(cman, "lambda", 1.0, "maxiter", 25, return_info=True)
    assert k == 25 and abs(err - 4.412657548137149) < 1e-9
    assert abs(tv.TVnorm(f) - 959256.7506108371) < 1e-5
    f, px, py, k, err = tv.chambolle_prox_TV_stop(kat_image(256), "lambda", 1e-3, "maxiter", 25, return_info=True)
    assert k == 20 and abs(err - 7.534945e-4) < 1e-9            # early stop KAT
    c = np.full((16, 16), 3.0)
    f, px, py, k, err = tv.chambolle_prox_TV_stop(c, "lambda", 0.5, "maxiter", 25, return_info=True)
    assert k == 1 and err == 0.0 and np.array_equal(f, c)
    with pytest.raises(NameError):                              # Q4
        tv.chambolle_prox_TV_stop(c, "lambda", 0.5)
    # energy decreases: 0.5||f-g||^2 + lambda TV_neumann(f) <= value at f = g
    g = np.random.default_rng(1).uniform(0, 255, (32, 32))
    lam = 5.0
    f = tv.chambolle_prox_TV_stop(g, "lambda", lam, "maxiter", 200, "tol", 1e-10)[0]
    tvn = lambda u: np.sqrt(tv.GradientIm(u)[0] ** 2 + tv.GradientIm(u)[1] ** 2).sum()
    assert 0.5 * np.sum((f - g) ** 2) + lam * tvn(f) < lam * tvn(g)


def test_blur_properties(cman):
    cl = OP.gaussian_closures(cman.shape, 7, 0.0)
    A, AT = cl["A"], cl["AT"]
    assert rel(A(np.full(cman.shape, 2.5), 0.4, 0.3), np.full(cman.shape, 2.5)) < 1e-13
    rng = np.random.default_rng(0)
    x = rng.standard_normal(cman.shape); z = rng.standard_normal(cman.shape)
    assert abs(np.vdot(A(x, 0.4, 0.3), z) - np.vdot(x, AT(z, 0.4, 0.3))) < 1e-9
    Ax = A(cman, 0.4, 0.3)
    assert abs(np.linalg.norm(Ax - Ax.mean(), "fro") - 15967.18482507) < 1e-6
    # Q1: the PSF sits at the top-left corner -> A shifts by (t-1)/2 = 3 pixels
    d = np.zeros((64, 64)); d[10, 20] = 1.0
    b = OP.gaussian_closures((64, 64), 7, 0.0)["A"](d, 0.4, 0.3)
    assert np.unravel_index(np.argmax(b), b.shape) == (13, 23)


def test_fusion_identities(cman):
    """Parseval forms the CUDA engine uses vs the reference's unfused formulas."""
    rng = np.random.default_rng(3)
    cl = OP.moffat_closures(cman.shape, 7)
    y = cl["A"](cman, 0.4, 3.5) + rng.standard_normal(cman.shape)
    x = np.abs(cman + rng.standard_normal(cman.shape))
    a, b, s2 = 0.7, 5.0, 4.0
    f, gradF, grads, gsig = OP.likelihood_closures(cl, y, x.size)
    H = cl["H_FFT"](a, b); Xh = np.fft.fft2(x); Yh = np.fft.fft2(y); Pn = x.size
    R = H * Xh - Yh
    assert abs(np.sum(np.abs(R) ** 2) / Pn / (2 * s2) - f(x, a, b, s2)) < 1e-12 * f(x, a, b, s2)
    D = P.diff_moffat_beta(cman.shape, 7, a, b)
    want = grads[1](x, a, b, s2)
    got = np.real(np.sum(np.conj(D * Xh) * R)) / Pn / s2
    assert abs(got - want) < 1e-9 * (abs(want) + 1)
    assert rel(np.real(np.fft.ifft2(np.conj(H) * R)) / s2, gradF(x, a, b, s2)) < 1e-13


def test_philox_reference_vector_and_moments():
    # Random123 known-answer test for philox4x32-10: counter = key = 0 / all ones / pi digits
    r = philox.philox4x32_10(0, 0, 0, 0, 0, 0)
    assert [int(v) for v in r] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    r = philox.philox4x32_10(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff)
    assert [int(v) for v in r] == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    r = philox.philox4x32_10(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)
    assert [int(v) for v in r] == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    z = philox.normal(200000, 7, 3, 11)
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    assert not np.allclose(philox.normal(16, 7, 3, 11), philox.normal(16, 7, 4, 11))


def test_l2_is_squared_spectral_norm():
    a = np.diag([3.0, 1.0]); b = np.zeros((2, 2))
    assert abs(metrics.l2(a, b) - 9.0) < 1e-14                  # Q8 (Frobenius would give 10)


def test_sapg_quirks_small():
    x = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "cman_u8.npy")).astype(float)[96:128, 64:96]
    rng = np.random.default_rng(5)
    randn = lambda s: rng.standard_normal(s)
    y, op, c = OP.setup_demo(0, x, randn, samples=12, warmup=4, burnIn=8, fix_w1=0, fix_w2=0)
    th, w1, w2, s2, r = sapg.SAPG_algorithm_Guassian(y, op, c, randn)
    assert r["last_samp"] == 12                                 # Q10: no early break
    assert np.all(np.isnan(r["tol_thetas"][1:8])) and np.all(np.isfinite(r["tol_thetas"][8:]))   # Q11
    assert abs(th - r["thetas"][7:].mean()) < 1e-15             # Q12: window burnIn:last inclusive
    assert r["gXTrace"][-1] == 0.0 and r["gXTrace"][0] != 0.0   # Q22: gX(ii-1) = g(X_ii)
    assert len(r["mean_thetas"]) == 4
    # multi-chain generalisation with one chain == the literal loop
    rng = np.random.default_rng(5)
    y2, op2, c2 = OP.setup_demo(0, x, randn, samples=12, warmup=4, burnIn=8, fix_w1=0, fix_w2=0)
    rng2 = np.random.default_rng(9)
    tape = [rng2.standard_normal(x.shape) for _ in range(3 + 11)]
    it1 = iter(tape); it2 = iter(tape)
    _, _, _, _, ra = sapg.SAPG_algorithm_Guassian(y2, op2, c2, lambda s: next(it1))
    rb = sapg.sapg_multichain(0, y2, op2, c2, lambda ch, s: next(it2), 1)
    for ka, kb in (("thetas", "thetas"), ("sigmas", "sigmas"), ("logPiTraceX", "logPiTraceX")):
        assert rel(rb[kb], ra[ka]) < 1e-13
    assert rel(rb["psis"][0], ra["w1s"]) < 1e-13


# ---------------------------------------------------------------- plain-C fast path of the TV oracle
@pytest.mark.parametrize("shape", [(2, 2), (5, 7), (64, 100), (131, 77), (256, 256)])
def test_c_tv_oracle_is_the_numpy_oracle(shape):
    """oracle/c/tv_oracle.c (used for >= 512^2 images) against the numpy restatement that the golden vectors pin:
    f, px, py bit-identical, k equal, err / TVnorm equal up to the order of the final sum."""
    import subprocess
    from oracle import tv
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle", "c")], check=True)
    rng = np.random.default_rng(shape[0] * 31 + shape[1])
    g = rng.uniform(0, 255, shape)
    old = tv.ACCEL
    try:
        for lam, K, tol in ((1e-3, 25, 1e-3), (0.3, 25, 1e-3), (2.0, 20, 1e-3), (0.3, 9, 0.0)):
            tv.ACCEL = False
            a = tv.chambolle_prox_TV_stop(g, "lambda", lam, "maxiter", K, "tol", tol, return_info=True)
            tv.ACCEL = True
            b = tv.chambolle_prox_TV_stop(g, "lambda", lam, "maxiter", K, "tol", tol, return_info=True)
            assert a[3] == b[3]
            for i in range(3):
                assert np.array_equal(a[i], b[i])
            assert abs(a[4] - b[4]) <= 1e-14 * max(a[4], 1e-300)
        if shape[0] == shape[1]:
            dual = rng.uniform(-0.5, 0.5, (shape[0], 2 * shape[1]))
            tv.ACCEL = False
            a = tv.chambolle_prox_TV_stop(g, "lambda", 0.7, "maxiter", 10, "tol", 1e-2, "tau", 0.2, "dualvars", dual)
            tv.ACCEL = True
            b = tv.chambolle_prox_TV_stop(g, "lambda", 0.7, "maxiter", 10, "tol", 1e-2, "tau", 0.2, "dualvars", dual)
            for u, v in zip(a, b):
                assert np.array_equal(u, v)
        tv.ACCEL = False
        t0 = tv.TVnorm(g)
        tv.ACCEL = True
        assert abs(tv.TVnorm(g) - t0) <= 1e-14 * t0
    finally:
        tv.ACCEL = old
