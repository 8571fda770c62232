"""`sbd_mex` for the MATLAB-subset interpreter: the same command set as
mex/sbd_mex.c, forwarded to libsbd.so through the ctypes binding.  It lets the
tests EXECUTE the MATLAB drop-in wrappers (sbd_b200/matlab/*.m) on a GPU box
that has neither MATLAB nor Octave.  Test infrastructure only."""
import numpy as np

from oracle.mlab.interp import MStruct, M, MatlabError


def _s(v):
    return float(np.asarray(v).ravel()[0])


def _vec(v):
    return np.asarray(v, dtype=np.float64).ravel()


def make_sbd_mex(sbd):
    from sbd_b200 import host as H
    from sbd_b200._lib import sbd_params

    def psi_of(a):
        v = _vec(a)
        return tuple(v[:2]) if v.size else (0.0,)

    def sbd_mex(args, nargout=1):
        cmd = args[0]
        a = args[1:]
        if cmd == "tvnorm":
            return [M(sbd.TVnorm(np.asarray(a[0])))]
        if cmd == "diff":
            return [sbd.diffh(a[0]) if int(_s(a[1])) == 1 else sbd.diffv(a[0])]
        if cmd == "tvprox":
            g = np.asarray(a[0], dtype=np.float64)
            dual = None
            if len(a) > 6 and np.asarray(a[5]).size:
                dual = (np.asarray(a[5]), np.asarray(a[6]))
            f, px, py, k, err = H._tv_engine(g.shape).tvprox(g, _s(a[1]), int(_s(a[2])), _s(a[3]), _s(a[4]), dual)
            return [f, px, py, M(float(k)), M(err)][:max(nargout, 1)]
        if cmd == "psf":
            model, t, phi = int(_s(a[0])), int(_s(a[1])), _s(a[2])
            return [H._taps(model, t, psi_of(a[3]), phi, int(_s(a[4])))]
        if cmd == "spectrum":
            sz = _vec(a[0]).astype(int)
            model, t, phi = int(_s(a[1])), int(_s(a[2])), _s(a[3])
            return [H.engine_for((sz[0], sz[1]), t, model, phi).psf_spectrum(psi_of(a[4]), int(_s(a[5])))]
        if cmd == "blur":
            x = np.asarray(a[0], dtype=np.float64)
            model, t, phi = int(_s(a[1])), int(_s(a[2])), _s(a[3])
            return [H.engine_for(x.shape, t, model, phi).blur(x, psi_of(a[4]), int(_s(a[5])))]
        if cmd == "sapg":
            y = np.asarray(a[0], dtype=np.float64)
            X0 = np.asarray(a[1]) if np.asarray(a[1]).size else None
            xt = np.asarray(a[2]) if np.asarray(a[2]).size else None
            model, t, phi, P = int(_s(a[3])), int(_s(a[4])), _s(a[5]), a[6]
            noise = np.asarray(a[7]) if len(a) > 7 and np.asarray(a[7]).size else None
            p = sbd_params()
            for f in ("samples", "warmup", "burnIn", "n_chains", "chambolle_maxiter", "fix_sigma", "err_psf_lag"):
                setattr(p, f, int(_s(P[f])))
            for f in ("gam", "lamb", "prox_lambda", "chambolle_tol", "chambolle_tau", "th_init", "min_th", "max_th",
                      "c_theta", "sigma2_init", "sigma2_min", "sigma2_max", "c_sigma2", "sigma2_fixed", "d_scale", "d_exp"):
                setattr(p, f, _s(P[f]))
            for f in ("psi_init", "psi_min", "psi_max", "c_psi", "psi_fixed", "psi_true"):
                v = _vec(P[f])
                for i in range(2):
                    getattr(p, f)[i] = v[i] if i < v.size else 0.0
            v = _vec(P["fix_psi"])
            for i in range(2):
                p.fix_psi[i] = int(v[i] != 0) if i < v.size else 0
            p.seed = int(_s(P["seed"])); p.chain_offset = 0; p.total_chains = p.n_chains
            p.use_graph = int(_s(P["use_graph"])) if "use_graph" in P else -1
            eng = H.engine_for(y.shape, t, model, phi, max_batch=p.n_chains)
            if noise is not None:           # MATLAB side hands [rows, cols*draws]; the tests pass a 4-D numpy tape
                noise = np.asarray(noise)
            o = eng.sapg(y, p, X0=X0, x_true=xt, noise=noise)
            r = MStruct()
            row = lambda z: np.asarray(z, dtype=np.float64).reshape(1, -1)
            for f in ("logPiTrace_WU", "thetas", "sigmas", "psi0", "psi1", "grad_theta", "grad_psi0", "grad_psi1",
                      "grad_sigma", "logPiTraceX", "gXTrace", "err_psf", "err_sample", "tol_theta", "tol_psi0",
                      "tol_psi1", "tol_sigma", "mean_theta", "mean_psi0", "mean_psi1", "mean_sigma"):
                r[f] = row(o[f])
            r["X_last"] = o["X_last"][0]; r["X_warm"] = o["X_warm"][0]
            r["EB"] = row(o["EB"]); r["err_warm0"] = M(o["err_warm0"]); r["seconds"] = M(o["seconds"])
            r["last_samp"] = M(float(o["last_samp"]))
            return [r]
        if cmd == "max_eigenval":
            sz = _vec(a[0]).astype(int)
            model, t, phi = int(_s(a[1])), int(_s(a[2])), _s(a[3])
            x0 = np.asarray(a[7]) if len(a) > 7 and np.asarray(a[7]).size else None
            v, k = H.engine_for((sz[0], sz[1]), t, model, phi).max_eigenval(psi_of(a[4]), _s(a[5]), int(_s(a[6])), x0=x0)
            return [M(v), M(float(k))][:max(nargout, 1)]
        if cmd == "observe":
            x = np.asarray(a[0], dtype=np.float64)
            model, t, phi = int(_s(a[1])), int(_s(a[2])), _s(a[3])
            noise = np.asarray(a[6]) if len(a) > 6 and np.asarray(a[6]).size else None
            y, sg, nr = H.engine_for(x.shape, t, model, phi).observe(x, psi_of(a[4]), _s(a[5]), noise=noise)
            return [y, M(sg), M(nr)][:max(nargout, 1)]
        if cmd == "salsa":
            y = np.asarray(a[0], dtype=np.float64)
            model, t, phi = int(_s(a[1])), int(_s(a[2])), _s(a[3])
            xt = np.asarray(a[10]) if len(a) > 10 and np.asarray(a[10]).size else None
            maxiter = int(_s(a[7]))
            r = H.engine_for(y.shape, t, model, phi).salsa_tv(y, psi_of(a[4]), _s(a[5]), _s(a[6]), maxiter, _s(a[8]),
                                                             int(_s(a[9])), x_true=xt)
            pad = lambda v, n: np.concatenate([np.ravel(v), np.zeros(n - np.size(v))]).reshape(1, -1)
            return [r["x"], pad(r["objective"], maxiter + 1), pad(r["distance"], maxiter), pad(r["mses"], maxiter + 1),
                    M(float(r["n_outer"]))][:max(nargout, 1)]
        raise MatlabError(f"sbd_mex: unknown command {cmd}")
    return sbd_mex
