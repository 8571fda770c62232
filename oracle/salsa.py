"""SALSA (ADMM) MAP deblurring with TV as the demos configure it - the
post-SAPG stage.  Oracle (test infrastructure): numpy
restatement of SALSA/SALSA_v2.m:156-494 for the option set the demo scripts
pass (run_Gaussian_demo.m:210-242): 'MU', 'AT', 'StopCriterion', 'True_x',
'ToleranceA', 'MAXITERA', 'Psi', 'Phi', 'TVINITIALIZATION', 'TViters', 'LS',
'VERBOSE' (+ 'INITIALIZATION' 0 / 2)."""
import numpy as np

from . import tv


def SALSA_v2(y, A, tau, *varargin):
    """-> x, numA, numAt, objective, distance, times, mses   (SALSA_v2.m:156-157)"""
    stopCriterion = 1; compute_mse = 0; maxiter = 10000; init = 0; AT = None          # :169-173
    mu = 1e-3; tolA = 0.001; isTV = 0; TViters = 5; invLS = None; psi = None; phi = None
    true = None
    numA = 0; numAt = 0
    if len(varargin) % 2 == 1:
        raise ValueError("Optional parameters should always go by pairs")           # :192-193
    for i in range(0, len(varargin) - 1, 2):
        name = str(varargin[i]).upper(); val = varargin[i + 1]
        if name == "PSI": psi = val
        elif name == "PHI": phi = val
        elif name == "TVINITIALIZATION": isTV = val
        elif name == "TVITERS": TViters = int(val)
        elif name == "MU": mu = float(val)
        elif name == "STOPCRITERION": stopCriterion = int(val)
        elif name == "TOLERANCEA": tolA = float(val)
        elif name == "MAXITERA": maxiter = int(val)
        elif name == "INITIALIZATION": init = val
        elif name == "TRUE_X": compute_mse = 1; true = val
        elif name == "AT": AT = val
        elif name == "VERBOSE": pass
        elif name == "LS": invLS = val
        else:
            raise ValueError(f"Unrecognized option: '{varargin[i]}'")                # :239
    if stopCriterion not in (1, 2, 3):
        raise ValueError("Unknown stopping criterion")                              # :245-247
    if AT is None:
        raise ValueError("The function handle for transpose of A is missing")       # :261-263
    ATy = AT(y); numAt += 1                                                         # :287-288
    if invLS is None:
        raise ValueError("(A^T A + mu I)^(-1) must be specified as a function handle.")   # :294-296
    if not isTV:
        raise NotImplementedError("only the TVINITIALIZATION path of the demos is restated")
    phi = lambda x: tv.TVnorm(x)                                                    # :354-358
    if isinstance(init, (int, float)) and init == 0:
        x = AT(np.zeros(y.shape))                                                   # :368
    elif isinstance(init, (int, float)) and init == 2:
        x = ATy                                                                     # :372
    else:
        raise NotImplementedError("INITIALIZATION option not restated")
    PTx = x
    u = PTx; bu = 0 * u                                                             # :391-392
    threshold = tau / mu                                                            # :393
    criterion = [1.0]
    resid = y - A(x); numA += 1                                                     # :398-399
    prev_f = 0.5 * float(np.vdot(resid, resid)) + tau * phi(u)                      # :400
    times = [0.0]; objective = [prev_f]; mses = []; distance = []
    if compute_mse:
        mses.append(float(np.sum(np.sum((x - true) ** 2)) / x.size))                # :414
    pux = 0 * u; puy = 0 * u                                                        # :418-419
    for outer in range(1, maxiter + 1):                                             # :422
        xprev = x
        u, pux, puy = tv.chambolle_prox_TV_stop(np.real(PTx - bu), "lambda", threshold, "maxiter", TViters,
                                                "dualvars", np.concatenate([pux, puy], axis=1))   # :428
        r = ATy + mu * (u + bu)                                                     # :433
        x = invLS(r)                                                                # :435
        PTx = x
        bu = bu + (u - PTx)                                                         # :439
        resid = y - A(x); numA += 1
        objective.append(0.5 * float(np.vdot(resid, resid)) + tau * phi(u))          # :443
        if compute_mse:
            err = x - true
            mses.append(float(np.vdot(err, err)) / x.size)                          # :447
        distance.append(float(np.linalg.norm((PTx - u).ravel()) /
                              np.sqrt(np.linalg.norm(PTx.ravel()) ** 2 + np.linalg.norm(u.ravel()) ** 2)))   # :450
        if outer > 1:
            if stopCriterion == 1:
                criterion.append(abs(objective[outer] - objective[outer - 1]) / objective[outer - 1])   # :457
            elif stopCriterion == 2:
                criterion.append(abs(np.linalg.norm((x - xprev).ravel()) / np.linalg.norm(x.ravel())))
            else:
                criterion.append(objective[outer])
            if criterion[outer - 1] < tolA:                                         # :470
                times.append(0.0)
                break
        else:
            pass
        times.append(0.0)
    return x, numA, numAt, np.array(objective), np.array(distance), np.array(times), np.array(mses)
