"""Total-variation pieces: TVnorm (periodic backward differences) and the
Chambolle dual-projection prox with the reference's stop test.

Oracle (test infrastructure).  numpy float64 restatement of
  utils/TVnorm.m, SALSA/diffh.m, SALSA/diffv.m, SALSA/conv2c.m,
  utils/chambolle_prox_TV_stop.m
"""

import ctypes as _C
import os as _os

import numpy as np

# --------------------------------------------------------------------------
# optional plain-C fast path for LARGE images (oracle/c/tv_oracle.c, built by
# `make -C oracle/c` / __graft_entry__.build()).  Same operations in the same
# floating-point order as the numpy code below, element for element; only the
# order of the err / TV sums differs (tests/test_oracle.py pins one against the
# other).  ACCEL: "auto" = use it for images of >= ACCEL_MIN_PIXELS pixels when
# the library is there, False = numpy only, True = always (raises if missing).
# --------------------------------------------------------------------------
ACCEL = "auto"
ACCEL_MIN_PIXELS = 512 * 512
_clib = None


def _c():
    global _clib
    if _clib is None:
        path = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "c", "_build", "liboracle_tv.so")
        if not _os.path.exists(path):
            _clib = False
            return _clib
        lib = _C.CDLL(path)
        dp = _C.POINTER(_C.c_double)
        lib.oc_chambolle.restype = _C.c_int
        lib.oc_chambolle.argtypes = [dp, _C.c_long, _C.c_long, _C.c_double, _C.c_int, _C.c_double, _C.c_double,
                                     dp, dp, dp, dp]
        lib.oc_tvnorm.restype = _C.c_double
        lib.oc_tvnorm.argtypes = [dp, _C.c_long, _C.c_long]
        lib.oc_set_threads.restype = _C.c_int
        lib.oc_set_threads.argtypes = [_C.c_int]
        _clib = lib
    return _clib


def set_threads(n):
    """Threads of the plain-C loops (torchrun exports OMP_NUM_THREADS=1); returns the number in effect (0: no C lib)."""
    lib = _c()
    return int(lib.oc_set_threads(int(n))) if lib else 0


def _use_c(npix):
    if ACCEL is False:
        return False
    lib = _c()
    if ACCEL is True:
        if not lib:
            raise RuntimeError("oracle/c/_build/liboracle_tv.so is not built (make -C oracle/c)")
        return True
    return bool(lib) and npix >= ACCEL_MIN_PIXELS


def _dp(a):
    return a.ctypes.data_as(_C.POINTER(_C.c_double))


# --------------------------------------------------------------------------
# SALSA/conv2c.m  (literal: wrap-around padding + conv2 'valid')
# --------------------------------------------------------------------------
def _wraparound(x, m):
    """SALSA/conv2c.m:7-50."""
    mx, nx = x.shape
    mm, nm = m.shape
    if mm > mx or nm > nx:
        raise ValueError("Mask does not fit inside array")      # :14-16
    mo = (1 + mm) // 2; no = (1 + nm) // 2                      # :18
    ml = mo - 1; nl = no - 1                                    # :19
    mr = mm - mo; nr = nm - no                                  # :20
    me = mx - ml + 1; ne = nx - nl + 1                          # :21
    mt = mx + ml; nt = nx + nl                                  # :22
    my = mx + mm - 1; ny = nx + nm - 1                          # :23
    y = np.zeros((my, ny))
    # 1-based inclusive MATLAB ranges a:b  ->  python slices [a-1:b]
    y[mo - 1:mt, no - 1:nt] = x                                 # :26
    if ml > 0:
        y[0:ml, no - 1:nt] = x[me - 1:mx, :]                    # :28
        if nl > 0:
            y[0:ml, 0:nl] = x[me - 1:mx, ne - 1:nx]             # :30
        if nr > 0:
            y[0:ml, nt:ny] = x[me - 1:mx, 0:nr]                 # :33
    if mr > 0:
        y[mt:my, no - 1:nt] = x[0:mr, :]                        # :37
        if nl > 0:
            y[mt:my, 0:nl] = x[0:mr, ne - 1:nx]                 # :39
        if nr > 0:
            y[mt:my, nt:ny] = x[0:mr, 0:nr]                     # :42
    if nl > 0:
        y[mo - 1:mt, 0:nl] = x[:, ne - 1:nx]                    # :46
    if nr > 0:
        y[mo - 1:mt, nt:ny] = x[:, 0:nr]                        # :49
    return y


def conv2c_literal(x, h):
    """SALSA/conv2c.m:1-4 using scipy's conv2 equivalent."""
    from scipy.signal import convolve2d
    h = np.atleast_2d(np.asarray(h, dtype=np.float64))
    return convolve2d(_wraparound(np.asarray(x, dtype=np.float64), h), h, mode="valid")


def diffh_literal(x):
    return conv2c_literal(x, np.array([[0.0, 1.0, -1.0]]))      # diffh.m:2-3


def diffv_literal(x):
    return conv2c_literal(x, np.array([[0.0, 1.0, -1.0]]).T)    # diffv.m:2-3


# Closed forms (tests check them bit-for-bit vs the literal)
def diffh(x):
    """x(i,j) - x(i,j-1) with periodic wrap  (SALSA/diffh.m:1-3)."""
    return x - np.roll(x, 1, axis=1)


def diffv(x):
    """x(i,j) - x(i-1,j) with periodic wrap  (SALSA/diffv.m:1-3)."""
    return x - np.roll(x, 1, axis=0)


def TVnorm(x):
    """utils/TVnorm.m:1-2."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 2 and _use_c(x.size):
        xc = np.ascontiguousarray(x)
        return float(_c().oc_tvnorm(_dp(xc), xc.shape[0], xc.shape[1]))
    return float(np.sum(np.sum(np.sqrt(diffh(x) ** 2 + diffv(x) ** 2), axis=0)))


# --------------------------------------------------------------------------
# utils/chambolle_prox_TV_stop.m
# --------------------------------------------------------------------------
def DivergenceIm(p1, p2):
    """chambolle_prox_TV_stop.m:152-159.  Last row/col is -p(end) (Q3)."""
    z = p2[:, 1:-1] - p2[:, :-2]                                # :153
    v = np.concatenate([p2[:, :1], z, -p2[:, -1:]], axis=1)     # :154
    z = p1[1:-1, :] - p1[:-2, :]                                # :156
    u = np.concatenate([p1[:1, :], z, -p1[-1:, :]], axis=0)     # :157
    return v + u                                                # :159


def GradientIm(u):
    """chambolle_prox_TV_stop.m:161-166 (forward differences, zero last row/col)."""
    z = u[1:, :] - u[:-1, :]
    dux = np.concatenate([z, np.zeros((1, z.shape[1]))], axis=0)    # :163
    z = u[:, 1:] - u[:, :-1]
    duy = np.concatenate([z, np.zeros((z.shape[0], 1))], axis=1)    # :166
    return dux, duy


def chambolle_prox_TV_stop(g, *varargin, return_info=False):
    """[f, px, py] = chambolle_prox_TV_stop(g, 'lambda', l, 'maxiter', K, ...)
    utils/chambolle_prox_TV_stop.m:1-150.  Option names are case-insensitive
    (:88).  Omitting 'maxiter' raises, like the reference's undefined `MaxIter`
    (Q4, :80/:96/:131)."""
    g = np.asarray(g, dtype=np.float64)
    px = np.zeros(g.shape)                                      # :68
    py = np.zeros(g.shape)                                      # :69
    k = 0                                                       # :71
    tau = 0.249; tol = 1e-3; lam = 1.0; verbose = 0             # :77-81
    MaxIter = None                                              # (default is mis-named `maxiter`, :80)
    for i in range(0, len(varargin) - 1, 2):                    # :87
        name = str(varargin[i]).upper()
        val = varargin[i + 1]
        if name == "LAMBDA":
            lam = float(val)
        elif name == "VERBOSE":
            verbose = val
        elif name == "TOL":
            tol = float(val)
        elif name == "MAXITER":
            MaxIter = int(val)
        elif name == "TAU":
            tau = float(val)
        elif name == "DUALVARS":                                # :99-107
            M, N = g.shape
            val = np.asarray(val, dtype=np.float64)
            Maux, Naux = val.shape
            if M != Maux or Naux != 2 * N:
                raise ValueError("Wrong size of the dual variables")
            py = val[:, M:].copy()                              # :106 (splits at M: square only, Q5)
            px = val[:, :M].copy()                              # :107
    if MaxIter is not None and MaxIter >= 1 and g.ndim == 2 and min(g.shape) >= 2 and _use_c(g.size):
        gc = np.ascontiguousarray(g)
        px = np.ascontiguousarray(px, dtype=np.float64).copy(); py = np.ascontiguousarray(py, dtype=np.float64).copy()
        f = np.empty_like(gc)
        e = _C.c_double()
        k = _c().oc_chambolle(_dp(gc), gc.shape[0], gc.shape[1], lam, MaxIter, tol, tau, _dp(px), _dp(py), _dp(f),
                              _C.byref(e))
        if k < 0:
            raise MemoryError("oc_chambolle")
        if return_info:
            return f, px, py, k, float(e.value)
        return f, px, py
    err = np.nan
    while True:                                                 # :120
        k += 1
        divp = DivergenceIm(px, py)                             # :123
        u = divp - g / lam                                      # :124
        upx, upy = GradientIm(u)                                # :126
        tmp = np.sqrt(upx ** 2 + upy ** 2)                      # :127
        err = np.sum((-upx + tmp * px) ** 2 + (-upy + tmp * py) ** 2) ** 0.5   # :128
        px = (px + tau * upx) / (1 + tau * tmp)                 # :129
        py = (py + tau * upy) / (1 + tau * tmp)                 # :130
        if MaxIter is None:
            raise NameError("Undefined function or variable 'MaxIter'")  # Q4
        if not ((k < MaxIter) and (err > tol)):                 # :131
            break
    f = g - lam * DivergenceIm(px, py)                          # :149
    if return_info:
        return f, px, py, k, float(err)
    return f, px, py
