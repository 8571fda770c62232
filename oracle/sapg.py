"""The three SAPG drivers (MYULA warm-up + SAPG main loop), restated literally.

Oracle (test infrastructure).  Follows
  SAPG/SAPG_algorithm_Guassian.m:7-308   (function is named SAPG_algorithm_sigma, Q14)
  SAPG/SAPG_algorithm_moffat.m:7-297
  SAPG/SAPG_algorithm_laplace.m:7-268
`op` / `c` / `results` are dicts standing in for MATLAB structs; 1-based
trajectories are stored 0-based (thetas[k] == MATLAB thetas(k+1)).
`randn(shape)` replaces MATLAB's global `randn` stream.
"""

import time
import numpy as np

from . import psf as P
from .metrics import l2, MSE


def _mean_range(v, lo, hi):
    """MATLAB mean(v(lo:hi)) with 1-based inclusive bounds; empty -> NaN (Q11)."""
    if hi < lo:
        return np.nan
    return float(np.mean(v[lo - 1:hi]))


def _clip(v, lo, hi):
    return min(max(v, lo), hi)


def _tails(results, name_mean, name_tol, mean_v, tol_v):
    results[name_mean] = mean_v
    results[name_tol] = tol_v


# --------------------------------------------------------------------------
def SAPG_algorithm_Guassian(y, op, c, randn):
    """SAPG/SAPG_algorithm_Guassian.m:7-308."""
    t0 = time.perf_counter()
    op = dict(op)
    if "X0" not in op:
        op["X0"] = y                                            # :10-12
    if "stopTol" not in op:
        op["stopTol"] = -1                                      # :13-15
    dimX = op["X0"].size                                        # :16
    if "warmup" not in op:
        op["warmup"] = 100                                      # :19-21
    phi = op["phi"]; taille = op["psf_size"]                    # :24-25
    total_iter = int(op["samples"]); warmupSteps = int(op["warmup"])   # :28-29
    lamb = c["lam"] * op["lambda"]                              # :30
    gam = c["gam"] * op["gamma"]                                # :31
    min_theta, max_theta = op["min_th"], op["max_th"]           # :35-36
    w1_init, min_w1, max_w1 = op["w1_init"], op["min_w1"], op["max_w1"]    # :39-41
    w2_init, min_w2, max_w2 = op["w2_init"], op["min_w2"], op["max_w2"]    # :44-46
    sigma_init = op["sigma_init"]                               # :50
    min_sigma = min(op["sigma_min"], op["sigma_max"])           # :51
    max_sigma = max(op["sigma_min"], op["sigma_max"])           # :52
    delta = lambda i: op["d_scale"] * ((float(i) ** (-op["d_exp"])) / dimX)     # :55
    gradF, proxG, logPi = op["gradF"], op["proxG"], op["logPi"]                 # :58-60
    gradF_w1, gradF_w2, g, gradF_sigma = op["grad_w1"], op["grad_w2"], op["g"], op["gradF_sigma"]
    burnIn = int(op["burnIn"])
    results = {}

    X_wu = np.array(op["X0"], dtype=np.float64)                 # :67
    if warmupSteps > 0:
        fix_theta, fix_w1, fix_w2, fix_sigma = op["th_init"], op["w1_init"], op["w2_init"], sigma_init
        logPiTrace_WU = np.zeros(warmupSteps)                   # :74
        proxGX_wu = proxG(X_wu, fix_theta)                      # :76
        for ii in range(2, warmupSteps + 1):                    # :78
            X_wu = np.abs(X_wu + gam * (proxGX_wu - X_wu) / lamb
                          - gam * gradF(X_wu, fix_w1, fix_w2, fix_sigma)
                          + np.sqrt(2 * gam) * randn(X_wu.shape))           # :80-81
            proxGX_wu = proxG(X_wu, fix_theta)                  # :82
            logPiTrace_WU[ii - 1] = logPi(X_wu, fix_theta, fix_w1, fix_w2, fix_sigma)   # :85
        results["logPiTrace_WU"] = logPiTrace_WU                # :92

    thetas = np.zeros(total_iter); thetas[0] = op["th_init"]    # :100-101
    sigmas = np.zeros(total_iter); sigmas[0] = sigma_init       # :105-106
    w1s = np.zeros(total_iter); w1s[0] = w1_init                # :109-110
    w2s = np.zeros(total_iter); w2s[0] = w2_init                # :113-114
    tol_thetas = np.zeros(total_iter); tol_w1s = np.zeros(total_iter)       # :117-120
    tol_w2s = np.zeros(total_iter); tol_sigma = np.zeros(total_iter)
    nm = max(total_iter - burnIn, 0)
    mean_w1s = np.zeros(nm); mean_w2s = np.zeros(nm)            # :123-126
    mean_thetas = np.zeros(nm); mean_sigmas = np.zeros(nm)
    Grad_theta = np.zeros(total_iter); Grad_w1 = np.zeros(total_iter)       # :129-132 (grow)
    Grad_w2 = np.zeros(total_iter); Grad_sigma = np.zeros(total_iter)
    logPiTraceX = np.zeros(total_iter); gX = np.zeros(total_iter)           # :135-136
    logPiTraceX[0] = logPi(X_wu, thetas[0], w1s[0], w2s[0], sigmas[0])      # :137
    X = X_wu                                                    # :139
    proxGX = proxG(X, thetas[0])                                # :140
    psf_true = P.psf_gaussian(taille, op["w1"], op["w2"], phi)  # :144
    err_psf = np.zeros(total_iter)
    err_psf[0] = l2(P.psf_gaussian(taille, w1s[0], w2s[0], phi), psf_true)  # :145-146
    c_theta, c_sigma, c_w1, c_w2 = c["theta"], c["sigma"], c["w1"], c["w2"]  # :149-152

    ii = 1
    for ii in range(2, total_iter + 1):                         # :158
        k = ii - 1                                              # 0-based slot of (ii)
        Z = randn(X.shape)                                      # :160
        X = np.abs(X + gam * (proxGX - X) / lamb
                   - gam * gradF(X, w1s[k - 1], w2s[k - 1], sigmas[k - 1])
                   + np.sqrt(2 * gam) * Z)                      # :161
        proxGX = proxG(X, thetas[k - 1])                        # :162
        G_t = dimX / thetas[k - 1] - g(X)                       # :165
        thetaii = thetas[k - 1] + c_theta * delta(ii) * G_t     # :166
        thetas[k] = _clip(thetaii, min_theta, max_theta)        # :167
        G_w1 = gradF_w1(X, w1s[k - 1], w2s[k - 1], sigmas[k - 1])           # :170
        w1ii = op["w1"] if op["fix_w1"] else w1s[k - 1] - c_w1 * delta(ii) * G_w1   # :171-175
        w1s[k] = _clip(w1ii, min_w1, max_w1)                    # :176
        G_w2 = gradF_w2(X, w1s[k - 1], w2s[k - 1], sigmas[k - 1])           # :179
        w2ii = op["w2"] if op["fix_w2"] else w2s[k - 1] - c_w2 * delta(ii) * G_w2   # :180-184
        w2s[k] = _clip(w2ii, min_w2, max_w2)                    # :185
        G_s = gradF_sigma(X, w1s[k - 1], w2s[k - 1], sigmas[k - 1])         # :188
        sigmaii = op["sigma_init"] if op["fix_sigma"] else sigmas[k - 1] + c_sigma * delta(ii) * G_s  # :189-193 (Q15)
        sigmas[k] = _clip(sigmaii, min_sigma, max_sigma)        # :194
        Grad_sigma[k] = G_s; Grad_w1[k] = G_w1; Grad_w2[k] = G_w2; Grad_theta[k] = G_t   # :197-200
        err_psf[k] = l2(P.psf_gaussian(taille, w1s[k], w2s[k - 1], phi), psf_true)       # :203-204 (Q9)
        logPiTraceX[k] = logPi(X, thetas[k - 1], w1s[k - 1], w2s[k - 1], sigmas[k - 1])  # :207
        gX[k - 1] = g(X)                                        # :208 (Q22)
        for tolv, tr in ((tol_thetas, thetas), (tol_w1s, w1s), (tol_w2s, w2s), (tol_sigma, sigmas)):
            m1 = _mean_range(tr, burnIn, ii); m0 = _mean_range(tr, burnIn, ii - 1)
            tolv[k] = abs(m1 - m0) / m0                         # :218-231 (no break, Q10)
        if ii > burnIn:                                         # :236-247
            mean_thetas[ii - burnIn - 1] = _mean_range(thetas, burnIn, ii)
            mean_w1s[ii - burnIn - 1] = _mean_range(w1s, burnIn, ii)
            mean_w2s[ii - burnIn - 1] = _mean_range(w2s, burnIn, ii)
            mean_sigmas[ii - burnIn - 1] = _mean_range(sigmas, burnIn, ii)

    results["execTimeFindParameters"] = time.perf_counter() - t0            # :251
    last_samp = ii; results["last_samp"] = last_samp            # :252-253
    results["logPiTraceX"] = logPiTraceX[:last_samp]            # :254
    results["gXTrace"] = gX[:last_samp]                         # :255
    theta_EB = _mean_range(thetas, burnIn, last_samp)           # :258 (Q12)
    results.update(theta_EB=theta_EB, last_theta=thetas[last_samp - 1], thetas=thetas[:last_samp],
                   mean_thetas=mean_thetas, tol_thetas=tol_thetas)
    w1_EB = _mean_range(w1s, burnIn, last_samp)                 # :266
    results.update(w1_EB=w1_EB, last_w1=w1s[last_samp - 1], w1s=w1s[:last_samp],
                   mean_w1s=mean_w1s, tol_w1s=tol_w1s)
    w2_EB = _mean_range(w2s, burnIn, last_samp)                 # :275
    results.update(w2_EB=w2_EB, last_w2=w2s[last_samp - 1], w2s=w2s[:last_samp],
                   mean_w2s=mean_w2s, tol_w2s=tol_w2s)
    sigma_EB = _mean_range(sigmas, burnIn, last_samp)           # :284
    results.update(sigma_EB=sigma_EB, last_sigma=sigmas[last_samp - 1], sigmas=sigmas[:last_samp],
                   mean_sigmas=mean_sigmas, tol_sigma=tol_sigma)
    results.update(c_theta=c_theta, c_w1=c_w1, c_w2=c_w2, Xlast_sample=X, c_sigma=c_sigma,
                   err_psf=err_psf, grad_theta=Grad_theta, grad_w1=Grad_w1, grad_w2=Grad_w2,
                   grad_sigma=Grad_sigma, options=op)           # :296-306
    return theta_EB, w1_EB, w2_EB, sigma_EB, results


# --------------------------------------------------------------------------
def SAPG_algorithm_moffat(y, op, randn):
    """SAPG/SAPG_algorithm_moffat.m:7-297."""
    t0 = time.perf_counter()
    op = dict(op)
    if "X0" not in op:
        op["X0"] = y                                            # :10-12
    if "stopTol" not in op:
        op["stopTol"] = -1
    if "warmup" not in op:
        op["warmup"] = 100                                      # :19-21
    dimX = op["X0"].size                                        # :23
    warmupSteps = int(op["warmup"])
    sub_sample = int(op["sub_sample"])                          # :25
    lamb, gam = op["lambda"], op["gamma"]                       # :26-27
    results = dict(gamma=gam)
    results["lambda"] = lamb                                    # :28-29
    total_iter = int(op["samples"])
    min_theta, max_theta = op["min_th"], op["max_th"]
    alpha_init, min_alpha, max_alpha = op["alpha_init"], op["min_alpha"], op["max_alpha"]
    beta_init, min_beta, max_beta = op["beta_init"], op["min_beta"], op["max_beta"]
    sigma_init = op["sigma_init"]
    min_sigma = min(op["sigma_min"], op["sigma_max"]); max_sigma = max(op["sigma_min"], op["sigma_max"])
    delta = lambda i: op["d_scale"] * ((float(i) ** (-op["d_exp"])) / dimX)     # :53
    gradF, proxG, logPi = op["gradF"], op["proxG"], op["logPi"]
    gradF_alpha, gradF_beta, g, gradF_sigma = op["grad_alpha"], op["grad_beta"], op["g"], op["gradF_sigma"]
    burnIn = int(op["burnIn"])

    X_wu = np.array(op["X0"], dtype=np.float64)                 # :65
    if warmupSteps > 0:
        fix_theta, fix_alpha, fix_beta, fix_sigma = op["th_init"], alpha_init, beta_init, sigma_init
        logPiTrace_WU = np.zeros(warmupSteps)
        proxGX_wu = proxG(X_wu, lamb, fix_theta)                # :75
        for ii in range(2, warmupSteps + 1):                    # :77
            X_wu = X_wu + gam * (proxGX_wu - X_wu) / lamb - gam * gradF(X_wu, fix_alpha, fix_beta, fix_sigma) \
                + np.sqrt(2 * gam) * randn(X_wu.shape)          # :80-81
            X_wu = np.abs(X_wu)                                 # :82
            proxGX_wu = proxG(X_wu, lamb, fix_theta)            # :83
            logPiTrace_WU[ii - 1] = logPi(X_wu, fix_theta, fix_alpha, fix_beta, fix_sigma)   # :86
        results["logPiTrace_WU"] = logPiTrace_WU

    thetas = np.zeros(total_iter); thetas[0] = op["th_init"]
    sigmas = np.zeros(total_iter); sigmas[0] = sigma_init
    alphas = np.zeros(total_iter); alphas[0] = alpha_init
    betas = np.zeros(total_iter); betas[0] = beta_init
    tol_thetas = np.zeros(total_iter); tol_alphas = np.zeros(total_iter)
    tol_betas = np.zeros(total_iter); tol_sigmas = np.zeros(total_iter)
    nm = max(total_iter - burnIn, 0)
    mean_alphas = np.zeros(nm); mean_betas = np.zeros(nm); mean_thetas = np.zeros(nm); mean_sigmas = np.zeros(nm)
    logPiTraceX = np.zeros(total_iter); gX = np.zeros(total_iter)
    logPiTraceX[0] = logPi(X_wu, thetas[0], alphas[0], betas[0], sigmas[0])     # :128
    X = X_wu
    proxGX = proxG(X, lamb, thetas[0])                          # :131
    c_theta, c_alpha, c_beta, c_sigma2 = 0.1, 10.0, 10000.0, 10000.0           # :135-138
    err_psf = np.zeros(total_iter)                              # err_psf(1) is never set -> 0 (grown array)
    psf_size = op["psf_size"]
    true_psf = P.psf_moffat(psf_size, op["alpha"], op["beta"])  # :154 (loop-invariant, Q20)

    ii = 1
    for ii in range(2, total_iter + 1):                         # :141
        k = ii - 1
        g_b = np.zeros(sub_sample); g_a = np.zeros(sub_sample)  # :143-150
        g_s = np.zeros(sub_sample); g_t = np.zeros(sub_sample)
        for jj in range(1):                                     # :158  (for jj = 1:1)
            Z = randn(X.shape)                                  # :159
            X = X + gam * (proxGX - X) / lamb - gam * gradF(X, alphas[k - 1], betas[k - 1], sigmas[k - 1]) \
                + np.sqrt(2 * gam) * Z                          # :160
            X = np.abs(X)                                       # :161
            proxGX = proxG(X, lamb, thetas[k - 1])              # :163
            g_b[jj] = gradF_beta(X, alphas[k - 1], betas[k - 1], sigmas[k - 1])     # :164
            g_a[jj] = gradF_alpha(X, alphas[k - 1], betas[k - 1], sigmas[k - 1])    # :165
            g_s[jj] = gradF_sigma(X, alphas[k - 1], betas[k - 1], sigmas[k - 1])    # :166
            g_t[jj] = dimX / thetas[k - 1] - g(X)               # :167
        G_b = np.mean(g_b); G_s = np.mean(g_s); G_t = np.mean(g_t); G_a = np.mean(g_a)   # :170-173
        thetas[k] = _clip(thetas[k - 1] + c_theta * delta(ii) * G_t, min_theta, max_theta)   # :176-177
        alphaii = op["alpha"] if op["fix_alpha"] else alphas[k - 1] - c_alpha * delta(ii) * G_a  # :180-184
        alphas[k] = _clip(alphaii, min_alpha, max_alpha)
        betaii = op["beta"] if op["fix_beta"] else betas[k - 1] - c_beta * delta(ii) * G_b       # :188-192
        betas[k] = _clip(betaii, min_beta, max_beta)
        sigmaii = op["sigma"] ** 2 if op["fix_sigma"] else sigmas[k - 1] + c_sigma2 * delta(ii) * G_s   # :196-200 (Q15)
        sigmas[k] = _clip(sigmaii, min_sigma, max_sigma)
        err_psf[k] = l2(P.psf_moffat(psf_size, alphas[k], betas[k]), true_psf)      # :204-205
        logPiTraceX[k] = logPi(X, thetas[k - 1], alphas[k - 1], betas[k - 1], sigmas[k - 1])  # :208
        gX[k - 1] = g(X)                                        # :209
        for tolv, tr in ((tol_thetas, thetas), (tol_alphas, alphas), (tol_betas, betas), (tol_sigmas, sigmas)):
            m1 = _mean_range(tr, burnIn, ii); m0 = _mean_range(tr, burnIn, ii - 1)
            tolv[k] = abs(m1 - m0) / m0                         # :218-231
        if ii > burnIn:                                         # :233-242
            mean_thetas[ii - burnIn - 1] = _mean_range(thetas, burnIn, ii)
            mean_alphas[ii - burnIn - 1] = _mean_range(alphas, burnIn, ii)
            mean_betas[ii - burnIn - 1] = _mean_range(betas, burnIn, ii)
            mean_sigmas[ii - burnIn - 1] = _mean_range(sigmas, burnIn, ii)

    results["execTimeFindTheta"] = time.perf_counter() - t0     # :247
    last_samp = ii; results["last_samp"] = last_samp
    results["logPiTraceX"] = logPiTraceX[:last_samp]; results["gXTrace"] = gX[:last_samp]
    theta_EB = _mean_range(thetas, burnIn, last_samp)
    results.update(mean_theta=theta_EB, last_theta=thetas[last_samp - 1], thetas=thetas[:last_samp],
                   mean_thetas=mean_thetas, tol_thetas=tol_thetas, c_theta=c_theta)
    alpha_EB = _mean_range(alphas, burnIn, last_samp)
    results.update(alpha_EB=alpha_EB, last_alpha=alphas[last_samp - 1], alphas=alphas[:last_samp],
                   mean_alphas=mean_alphas, tol_alphas=tol_alphas, c_alpha=c_alpha)
    beta_EB = _mean_range(betas, burnIn, last_samp)
    results.update(beta_EB=beta_EB, last_beta=betas[last_samp - 1], betas=betas[:last_samp],
                   mean_betas=mean_betas, tol_betas=tol_betas, c_beta=c_beta)
    sigma2_EB = _mean_range(sigmas, burnIn, last_samp)
    results.update(sigma_EB=sigma2_EB, last_sigma=sigmas[last_samp - 1], sigmas=sigmas[:last_samp],
                   mean_sigmas=mean_sigmas, tol_sigma=tol_sigmas, c_sigma2=c_sigma2)
    results.update(Xlast_sample=X, X_warm=X_wu, options=op, err_psf=err_psf)
    return theta_EB, alpha_EB, beta_EB, sigma2_EB, results


# --------------------------------------------------------------------------
def SAPG_algorithm_laplace(y, op, randn):
    """SAPG/SAPG_algorithm_laplace.m:7-268."""
    t0 = time.perf_counter()
    op = dict(op)
    if "X0" not in op:
        op["X0"] = y
    if "stopTol" not in op:
        op["stopTol"] = -1
    if "warmup" not in op:
        op["warmup"] = 100
    dimX = op["X0"].size                                        # :25
    warm_sample = int(op["warm_sample"])                        # :26
    warmupSteps = int(op["warmup"])
    err_warm = np.zeros(max(warmupSteps, 1))
    err_warm[0] = MSE(op["X0"], op["x"])                        # :28-29
    lamb, gam = op["lambda"], op["gamma"]
    results = dict(gamma=gam)
    results["lambda"] = lamb
    total_iter = int(op["samples"])
    min_theta, max_theta = op["min_th"], op["max_th"]
    b_init, min_b, max_b = op["b_init"], op["min_b"], op["max_b"]
    sigma_init = op["sigma_init"]
    min_sigma = min(op["sigma_min"], op["sigma_max"]); max_sigma = max(op["sigma_min"], op["sigma_max"])
    delta = lambda i: op["d_scale"] * ((float(i) ** (-op["d_exp"])) / dimX)     # :57
    proxG, logPi, gradF_b = op["proxG"], op["logPi"], op["grad_b"]
    gradF, gradF_sigma, g = op["gradF"], op["gradF_sigma"], op["g"]
    burnIn = int(op["burnIn"])

    X_wu = np.array(op["X0"], dtype=np.float64)                 # :68
    if warmupSteps > 0:
        fix_theta, fix_b, fix_sigma = op["th_init"], op["b_init"], op["sigma_init"]
        logPiTrace_WU = np.zeros(warmupSteps)
        proxGX_wu = proxG(X_wu, lamb, fix_theta)                # :77
        for ii in range(2, warmupSteps + 1):                    # :79
            X_wu = X_wu + gam * (proxGX_wu - X_wu) / lamb - gam * gradF(X_wu, fix_b, fix_sigma) \
                + np.sqrt(2 * gam) * randn(X_wu.shape)          # :81-82
            X_wu = np.abs(X_wu)                                 # :83
            proxGX_wu = proxG(X_wu, lamb, fix_theta)            # :84
            logPiTrace_WU[ii - 1] = logPi(X_wu, fix_theta, fix_b, fix_sigma)    # :87
        results["logPiTrace_WU"] = logPiTrace_WU

    thetas = np.zeros(total_iter); thetas[0] = op["th_init"]
    sigmas = np.zeros(total_iter); sigmas[0] = sigma_init
    bs = np.zeros(total_iter); bs[0] = b_init
    tol_thetas = np.zeros(total_iter); tol_bs = np.zeros(total_iter); tol_sigmas = np.zeros(total_iter)
    nm = max(total_iter - burnIn, 0)
    mean_bs = np.zeros(nm); mean_thetas = np.zeros(nm); mean_sigmas = np.zeros(nm)
    err_sample = np.zeros(total_iter); err_sample[0] = MSE(X_wu, op["x"])       # :122-123
    logPiTraceX = np.zeros(total_iter); gX = np.zeros(total_iter)
    logPiTraceX[0] = logPi(X_wu, thetas[0], bs[0], sigmas[0])   # :128
    X = X_wu
    proxGX = proxG(X, lamb, thetas[0])                          # :131
    true_psf = P.psf_laplace(op["psf_size"], op["b"])           # :134
    err_psf = np.zeros(total_iter)
    err_psf[0] = l2(P.psf_laplace(op["psf_size"], bs[0]), true_psf)             # :135-136
    c_theta, c_b, c_sigma2 = 0.01, 100.0, 10000.0               # :139-141

    ii = 1
    for ii in range(2, total_iter + 1):                         # :144
        k = ii - 1
        g_b = np.zeros(warm_sample); g_s = np.zeros(warm_sample); g_t = np.zeros(warm_sample)   # :146-151
        for jj in range(1):                                     # :153
            Z = randn(X.shape)                                  # :154
            X = X + gam * (proxGX - X) / lamb - gam * gradF(X, bs[k - 1], sigmas[k - 1]) + np.sqrt(2 * gam) * Z  # :155
            X = np.abs(X)                                       # :156
            proxGX = proxG(X, lamb, thetas[k - 1])              # :157
            g_b[jj] = gradF_b(X, bs[k - 1], sigmas[k - 1])      # :159
            g_s[jj] = gradF_sigma(X, bs[k - 1], sigmas[k - 1])  # :160
            g_t[jj] = dimX / thetas[k - 1] - g(X)               # :161
        G_b = np.mean(g_b); G_s = np.mean(g_s); G_t = np.mean(g_t)          # :164-166
        thetas[k] = _clip(thetas[k - 1] + c_theta * delta(ii) * G_t, min_theta, max_theta)   # :169-170
        bii = op["b"] if op["fix_b"] else bs[k - 1] - c_b * delta(ii) * G_b                 # :173-177
        bs[k] = _clip(bii, min_b, max_b)                        # :178
        sigmaii = op["sigma"] ** 2 if op["fix_sigma"] else sigmas[k - 1] + c_sigma2 * delta(ii) * G_s   # :181-185
        sigmas[k] = _clip(sigmaii, min_sigma, max_sigma)        # :186
        err_sample[k] = MSE(X, op["x"])                         # :189
        err_psf[k] = l2(P.psf_laplace(op["psf_size"], bs[k]), true_psf)     # :190-191
        logPiTraceX[k] = logPi(X, thetas[k - 1], bs[k - 1], sigmas[k - 1])  # :194
        gX[k - 1] = g(X)                                        # :195
        for tolv, tr in ((tol_thetas, thetas), (tol_bs, bs), (tol_sigmas, sigmas)):
            m1 = _mean_range(tr, burnIn, ii); m0 = _mean_range(tr, burnIn, ii - 1)
            tolv[k] = abs(m1 - m0) / m0                         # :204-213
        if ii > burnIn:                                         # :215-223
            mean_thetas[ii - burnIn - 1] = _mean_range(thetas, burnIn, ii)
            mean_bs[ii - burnIn - 1] = _mean_range(bs, burnIn, ii)
            mean_sigmas[ii - burnIn - 1] = _mean_range(sigmas, burnIn, ii)

    results["execTimeFindTheta"] = time.perf_counter() - t0
    last_samp = ii; results["last_samp"] = last_samp
    results["logPiTraceX"] = logPiTraceX[:last_samp]; results["gXTrace"] = gX[:last_samp]
    theta_EB = _mean_range(thetas, burnIn, last_samp)
    results.update(mean_theta=theta_EB, last_theta=thetas[last_samp - 1], thetas=thetas[:last_samp],
                   mean_thetas=mean_thetas, tol_thetas=tol_thetas, c_theta=c_theta)
    b_EB = _mean_range(bs, burnIn, last_samp)
    results.update(mean_b=b_EB, last_b=bs[last_samp - 1], bs=bs[:last_samp], mean_bs=mean_bs,
                   tol_bs=tol_bs, c_b=c_b)
    sigma_EB = _mean_range(sigmas, burnIn, last_samp)
    results.update(sigma_EB=sigma_EB, last_sigma=sigmas[last_samp - 1], sigmas=sigmas[:last_samp],
                   mean_sigmas=mean_sigmas, tol_sigma=tol_sigmas, c_sigma2=c_sigma2)
    results.update(X_sample=X, X_warm=X_wu, err_warm=err_warm, err_sample=err_sample,
                   err_psf=err_psf, options=op)
    return theta_EB, b_EB, sigma_EB, results


# --------------------------------------------------------------------------
# Multi-chain generalisation (NOT in the reference).  With n_chains == 1 this
# is the literal loops above (checked in tests/test_oracle_sapg.py).  With more
# chains the per-chain stochastic gradients are averaged exactly the way the
# reference averages its size-1 mini-batch (`G_b = mean(g_b)`,
# SAPG_algorithm_moffat.m:170-173); this is the semantic the CUDA engine
# implements for chains sharded over GPUs.
# --------------------------------------------------------------------------
def sapg_multichain(model, y, op, c, randn_chain, n_chains, combine=None):
    """randn_chain(chain, shape) -> noise; combine(list_of_per_chain_vectors)
    -> summed vector (hook for the distributed all-gather test)."""
    op = dict(op)
    X0 = np.array(op.get("X0", y), dtype=np.float64)
    dimX = X0.size
    warmupSteps = int(op.get("warmup", 100))
    total_iter = int(op["samples"])
    if model == P.GAUSSIAN:
        lamb = c["lam"] * op["lambda"]; gam = c["gam"] * op["gamma"]
        names = ("w1", "w2")
        cs = dict(theta=c["theta"], sigma=c["sigma"], psi=(c["w1"], c["w2"]))
        fixed_sigma = op["sigma_init"]
        prox = lambda x, th: op["proxG"](x, th)
        grads = (op["grad_w1"], op["grad_w2"])
    elif model == P.MOFFAT:
        lamb = op["lambda"]; gam = op["gamma"]
        names = ("alpha", "beta")
        cs = dict(theta=0.1, sigma=10000.0, psi=(10.0, 10000.0))
        fixed_sigma = op["sigma"] ** 2
        prox = lambda x, th: op["proxG"](x, lamb, th)
        grads = (op["grad_alpha"], op["grad_beta"])
    else:
        lamb = op["lambda"]; gam = op["gamma"]
        names = ("b",)
        cs = dict(theta=0.01, sigma=10000.0, psi=(100.0,))
        fixed_sigma = op["sigma"] ** 2
        prox = lambda x, th: op["proxG"](x, lamb, th)
        grads = (op["grad_b"],)
    npsi = len(names)
    delta = lambda i: op["d_scale"] * ((float(i) ** (-op["d_exp"])) / dimX)
    min_sigma = min(op["sigma_min"], op["sigma_max"]); max_sigma = max(op["sigma_min"], op["sigma_max"])
    if combine is None:
        combine = lambda vs: np.sum(np.stack(vs, 0), axis=0)

    Xs = [X0.copy() for _ in range(n_chains)]
    th0 = op["th_init"]; psi0 = tuple(op[n + "_init"] for n in names); s0 = op["sigma_init"]
    Ps = [prox(X, th0) for X in Xs]
    logPi_WU = np.zeros(max(warmupSteps, 0))
    for ii in range(2, warmupSteps + 1):
        acc = []
        for ch in range(n_chains):
            X = np.abs(Xs[ch] + gam * (Ps[ch] - Xs[ch]) / lamb - gam * op["gradF"](Xs[ch], *psi0, s0)
                       + np.sqrt(2 * gam) * randn_chain(ch, X0.shape))
            Xs[ch] = X
            Ps[ch] = prox(X, th0)
            acc.append(np.array([op["logPi"](X, th0, *psi0, s0)]))
        logPi_WU[ii - 1] = combine(acc)[0] / n_chains

    thetas = np.zeros(total_iter); thetas[0] = th0
    sigmas = np.zeros(total_iter); sigmas[0] = s0
    psis = np.zeros((npsi, total_iter)); psis[:, 0] = psi0
    grad = np.zeros((2 + npsi, total_iter))
    logPiTraceX = np.zeros(total_iter); gX = np.zeros(total_iter)
    logPiTraceX[0] = combine([np.array([op["logPi"](X, th0, *psi0, s0)]) for X in Xs])[0] / n_chains
    Ps = [prox(X, thetas[0]) for X in Xs]
    for ii in range(2, total_iter + 1):
        k = ii - 1
        pk = tuple(psis[:, k - 1]); sk = sigmas[k - 1]; tk = thetas[k - 1]
        acc = []
        for ch in range(n_chains):
            Z = randn_chain(ch, X0.shape)
            X = np.abs(Xs[ch] + gam * (Ps[ch] - Xs[ch]) / lamb - gam * op["gradF"](Xs[ch], *pk, sk)
                       + np.sqrt(2 * gam) * Z)
            Xs[ch] = X
            Ps[ch] = prox(X, tk)
            gx = op["g"](X)
            v = [dimX / tk - gx] + [gr(X, *pk, sk) for gr in grads] + \
                [op["gradF_sigma"](X, *pk, sk), op["logPi"](X, tk, *pk, sk), gx]
            acc.append(np.array(v))
        tot = combine(acc) / n_chains
        G_t, G_psi, G_s = tot[0], tot[1:1 + npsi], tot[1 + npsi]
        logPiTraceX[k] = tot[2 + npsi]; gX[k - 1] = tot[3 + npsi]
        thetas[k] = _clip(tk + cs["theta"] * delta(ii) * G_t, op["min_th"], op["max_th"])
        for p, n in enumerate(names):
            v = op[n] if op["fix_" + n] else psis[p, k - 1] - cs["psi"][p] * delta(ii) * G_psi[p]
            psis[p, k] = _clip(v, op["min_" + n], op["max_" + n])
        v = fixed_sigma if op["fix_sigma"] else sk + cs["sigma"] * delta(ii) * G_s
        sigmas[k] = _clip(v, min_sigma, max_sigma)
        grad[0, k] = G_t; grad[1:1 + npsi, k] = G_psi; grad[1 + npsi, k] = G_s
    return dict(thetas=thetas, sigmas=sigmas, psis=psis, grad=grad, logPiTraceX=logPiTraceX,
                gXTrace=gX, logPiTrace_WU=logPi_WU, X=Xs)
