"""CPU checks of the MATLAB side of the boundary: the MEX gateway compiles
against a stand-in mex.h and binds only symbols that include/sbd.h declares;
the drop-in .m wrappers exist for every reference entry point of the path."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "semi-blind-image-deblurring-problems-with-tv_b200")


def test_gateway_compiles_and_uses_only_declared_symbols(tmp_path):
    src = os.path.join(PKG, "mex", "sbd_mex.c")
    obj = tmp_path / "sbd_mex.o"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-c", "-I" + os.path.join(PKG, "mex", "stub"),
                    "-I" + os.path.join(ROOT, "include"), src, "-o", str(obj)], check=True)
    undefined = subprocess.run(["nm", "-u", str(obj)], capture_output=True, text=True, check=True).stdout
    used = set(re.findall(r"\b(sbd_[a-z0-9_]+)\b", undefined))
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "sbd.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(sbd_[a-z0-9_]+)\s*\(", hdr))
    assert used and used <= declared, used - declared
    assert {"sbd_sapg_run", "sbd_tvprox", "sbd_tvnorm", "sbd_blur", "sbd_psf_spectrum"} <= used


def test_wrappers_cover_the_reference_entry_points():
    have = {f[:-2] for f in os.listdir(os.path.join(PKG, "matlab")) if f.endswith(".m")}
    need = {"SAPG_algorithm_Guassian", "SAPG_algorithm_moffat", "SAPG_algorithm_laplace", "chambolle_prox_TV_stop",
            "TVnorm", "diffh", "diffv", "Gaussian_psf", "psf_gaussian", "psf_moffat", "psf_laplace", "moffat_psf",
            "laplace_psf", "diff_fftgaus_w1", "diff_fftgaus_w2", "diff_moffat_alpha", "diff_moffat_beta",
            "diff_laplace_b"}
    assert need <= have, need - have
    # signatures identical to the reference's
    sig = lambda n: re.search(r"^function\s+(.*)$", open(os.path.join(PKG, "matlab", n + ".m")).read(), flags=re.M).group(1)
    norm = lambda s: re.sub(r"\s+", "", s)
    assert norm(sig("SAPG_algorithm_Guassian")) == norm("[theta_EB, w1_EB, w2_EB, sigma_EB, results] = SAPG_algorithm_Guassian(y, op, c)")
    assert norm(sig("SAPG_algorithm_moffat")) == norm("[theta_EB, alpha_EB, beta_EB, sigma2_EB, results] = SAPG_algorithm_moffat(y, op)")
    assert norm(sig("SAPG_algorithm_laplace")) == norm("[theta_EB, b_EB, sigma_EB, results] = SAPG_algorithm_laplace(y, op)")
    assert norm(sig("chambolle_prox_TV_stop")) == norm("[f, px, py] = chambolle_prox_TV_stop(g, varargin)")


def test_wrappers_parse_as_matlab():
    import sys
    sys.path.insert(0, ROOT)
    from oracle.mlab.interp import Parser, tokenize
    d = os.path.join(PKG, "matlab")
    for f in sorted(os.listdir(d)):
        if f.endswith(".m"):
            Parser(tokenize(open(os.path.join(d, f)).read())).parse_file()


def test_gateway_executes_on_the_stub_runtime():
    """mex/sbd_mex.c linked with libsbd.so and tests/mexrt/stub_mx.c: mexFunction really runs.  Without a GPU the
    commands that need a context must raise through mexErrMsgIdAndTxt with libsbd's own message (no CPU fallback)."""
    import sys
    import numpy as np
    import pytest
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "mexrt"))
    import __graft_entry__ as g
    g.build()
    import runtime
    with pytest.raises(runtime.MexError) as e:
        runtime.call_mex("no_such_command")
    assert e.value.ident == "sbd:usage"
    with pytest.raises(runtime.MexError) as e:
        runtime.call_mex(np.ones((2, 2)))                       # first argument must be the command string
    assert e.value.ident == "sbd:usage"
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(runtime.MexError) as e:
            runtime.call_mex("tvnorm", np.ones((8, 8)))
        assert e.value.ident == "sbd:error" and "no CPU fallback" in e.value.msg
    # value marshalling of the runtime itself (column-major, complex split, struct fields)
    a = np.arange(6.0).reshape(2, 3)
    assert np.array_equal(runtime.from_mx(runtime.to_mx(a)), a)
    z = a + 1j * a[::-1]
    assert np.array_equal(runtime.from_mx(runtime.to_mx(z)), z)
    s = runtime.from_mx(runtime.to_mx({"u": 3.0, "v": a}))
    assert s["u"].item() == 3.0 and np.array_equal(s["v"], a)
