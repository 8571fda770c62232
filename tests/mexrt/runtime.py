"""Drives the REAL MEX gateway (mex/sbd_mex.c) without MATLAB: the gateway is compiled together with the minimal
mxArray runtime of stub_mx.c, linked against libsbd.so, and `mexFunction` is called through ctypes.
Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "semi-blind-image-deblurring-problems-with-tv_b200")
OUT = os.path.join(HERE, "_build", "libsbd_mex_test.so")


class MexError(RuntimeError):
    def __init__(self, ident, msg):
        super().__init__(f"{ident}: {msg}")
        self.ident, self.msg = ident, msg


def build():
    src = [os.path.join(PKG, "mex", "sbd_mex.c"), os.path.join(HERE, "stub_mx.c")]
    libdir = os.path.join(PKG, "lib")
    deps = src + [os.path.join(ROOT, "include", "sbd.h"), os.path.join(libdir, "libsbd.so")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["gcc", "-std=c99", "-O1", "-Wall", "-shared", "-fPIC", "-o", OUT] + src +
                   ["-I" + os.path.join(PKG, "mex", "stub"), "-I" + os.path.join(ROOT, "include"),
                    "-L" + libdir, "-lsbd", "-Wl,-rpath," + libdir], check=True)
    return OUT


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp = C.c_void_p
        L.mxCreateDoubleMatrix.restype = vp; L.mxCreateDoubleMatrix.argtypes = [C.c_size_t, C.c_size_t, C.c_int]
        L.mxGetPr.restype = C.POINTER(C.c_double); L.mxGetPr.argtypes = [vp]
        L.mxGetPi.restype = C.POINTER(C.c_double); L.mxGetPi.argtypes = [vp]
        L.mxGetM.restype = C.c_size_t; L.mxGetM.argtypes = [vp]
        L.mxGetN.restype = C.c_size_t; L.mxGetN.argtypes = [vp]
        L.stub_string.restype = vp; L.stub_string.argtypes = [C.c_char_p]
        L.stub_struct.restype = vp; L.stub_struct.argtypes = [C.c_int]
        L.stub_add_field.argtypes = [vp, C.c_char_p, vp]
        L.stub_kind.argtypes = [vp]
        L.stub_nfields.argtypes = [vp]
        L.stub_field_name.restype = C.c_char_p; L.stub_field_name.argtypes = [vp, C.c_int]
        L.stub_field_value.restype = vp; L.stub_field_value.argtypes = [vp, C.c_int]
        L.stub_last_error.restype = C.c_char_p
        L.stub_last_error_id.restype = C.c_char_p
        L.stub_free.argtypes = [vp]
        L.stub_call.argtypes = [C.c_int, C.POINTER(vp), C.c_int, C.POINTER(vp)]
        _lib = L
    return _lib


def to_mx(v):
    """python value -> mxArray*.  str -> char row; dict / mapping -> 1x1 struct; numbers / arrays -> double matrix
    (column-major).  A 4-D array [draws, chains, rows, cols] becomes rows x (cols*chains*draws): images back to back,
    which is what MATLAB's randn(rows, cols*...) tape would be."""
    L = lib()
    if isinstance(v, str):
        return L.stub_string(v.encode())
    if hasattr(v, "keys"):
        keys = list(v.keys())
        s = L.stub_struct(len(keys))
        for k in keys:
            L.stub_add_field(s, str(k).encode(), to_mx(v[k]))
        return s
    if v is None:
        v = np.zeros((0, 0))
    a = np.asarray(v)
    cplx = np.iscomplexobj(a)
    a = a.astype(np.complex128 if cplx else np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(1, -1)
    elif a.ndim == 4:
        d, ch, r, c = a.shape
        a = a.transpose(2, 0, 1, 3).reshape(r, d * ch * c)      # [r, (d, ch, c)]: column index = ((d*ch)+ch)*c + c
    m, n = a.shape
    mx = L.mxCreateDoubleMatrix(m, n, 1 if cplx else 0)
    if m * n:
        col = np.asfortranarray(a)
        C.memmove(L.mxGetPr(mx), np.ascontiguousarray(col.real.T).ctypes.data, m * n * 8)
        if cplx:
            C.memmove(L.mxGetPi(mx), np.ascontiguousarray(col.imag.T).ctypes.data, m * n * 8)
    return mx


def from_mx(mx):
    L = lib()
    kind = L.stub_kind(mx)
    if kind == 2:
        return {L.stub_field_name(mx, f).decode(): from_mx(L.stub_field_value(mx, f)) for f in range(L.stub_nfields(mx))
                if L.stub_field_value(mx, f)}
    if kind != 0:
        raise TypeError("unexpected mxArray kind")
    m, n = L.mxGetM(mx), L.mxGetN(mx)
    if m * n == 0:
        return np.zeros((m, n))
    re = np.ctypeslib.as_array(L.mxGetPr(mx), shape=(n, m)).T.copy()
    pi = L.mxGetPi(mx)
    if pi:
        return re + 1j * np.ctypeslib.as_array(pi, shape=(n, m)).T
    return re


def call_mex(*args, nargout=1):
    """out = sbd_mex(args...) through the real mexFunction.  Returns a list of `max(nargout, 1)` python values."""
    L = lib()
    nrhs = len(args)
    prhs = (C.c_void_p * max(nrhs, 1))(*[to_mx(a) for a in args])
    nl = max(nargout, 1)
    plhs = (C.c_void_p * 8)()
    rc = L.stub_call(nargout, plhs, nrhs, prhs)
    for i in range(nrhs):
        L.stub_free(prhs[i])
    if rc:
        raise MexError(L.stub_last_error_id().decode(), L.stub_last_error().decode())
    out = []
    for i in range(nl):
        out.append(from_mx(plhs[i]) if plhs[i] else None)
    for i in range(8):
        if plhs[i]:
            L.stub_free(plhs[i])
    return out


def make_sbd_mex_real():
    """`sbd_mex` builtin for the MATLAB-subset interpreter that goes through the REAL gateway (mexFunction of
    mex/sbd_mex.c) instead of a Python re-implementation."""
    from oracle.mlab.interp import MStruct, MatlabError

    def conv(v):
        if isinstance(v, dict):
            s = MStruct()
            for k, x in v.items():
                s[k] = conv(x)
            return s
        return v

    def sbd_mex(args, nargout=1):
        try:
            out = call_mex(*args, nargout=nargout)
        except MexError as e:
            raise MatlabError(str(e))
        return [conv(o) for o in out][:max(nargout, 1)]
    return sbd_mex
