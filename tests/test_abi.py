"""CPU tests of the drop-in boundary: libsbd.so loads, exports every symbol that
include/sbd.h declares, the ctypes mirrors of the POD structs have the C layout,
and compute entry points fail loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "include", "sbd.h")
PKG = "semi-blind-image-deblurring-problems-with-tv_b200"


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import sbd_b200
    return sbd_b200


def header_functions():
    src = open(HDR).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbd_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from sbd_b200._lib import load_library, SIGNATURES, LIB_PATH
    lib = load_library()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sbd.h but not exported by {LIB_PATH}"
        assert n in SIGNATURES, f"{n} has no ctypes prototype in sbd_b200/_lib.py"
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\sT\s+(sbd_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert lib.sbd_version() == 100


def test_struct_layout_matches_c(built, tmp_path):
    """Compile a C probe against include/sbd.h and compare sizeof/offsetof with ctypes."""
    from sbd_b200._lib import sbd_params, sbd_traces
    fields_p = [f[0] for f in sbd_params._fields_]
    fields_t = [f[0] for f in sbd_traces._fields_]
    probe = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HDR}"', 'int main(void){',
             'printf("%zu %zu\\n", sizeof(sbd_params), sizeof(sbd_traces));']
    for f in fields_p:
        probe.append(f'printf("p.{f} %zu\\n", offsetof(sbd_params, {f}));')
    for f in fields_t:
        probe.append(f'printf("t.{f} %zu\\n", offsetof(sbd_traces, {f}));')
    probe.append('return 0;}')
    src = tmp_path / "probe.c"
    src.write_text("\n".join(probe))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    sp, st = map(int, lines[0].split())
    assert sp == C.sizeof(sbd_params) and st == C.sizeof(sbd_traces)
    for ln in lines[1:]:
        if not ln.strip():
            continue
        name, off = ln.split()
        kind, f = name.split(".")
        cls = sbd_params if kind == "p" else sbd_traces
        assert getattr(cls, f).offset == int(off), name


def test_no_cpu_fallback_without_gpu(built):
    """Without a B200 every compute entry point must raise (SBD_E_NODEVICE),
    never silently compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the -m gpu tests")
    from sbd_b200 import SbdError
    with pytest.raises(SbdError) as e:
        built.TVnorm(np.zeros((8, 8)))
    assert e.value.code == -3 and "no CPU fallback" in e.value.msg
    with pytest.raises(SbdError):
        built.chambolle_prox_TV_stop(np.zeros((8, 8)), "lambda", 1.0, "maxiter", 3)
    with pytest.raises(SbdError):
        built.Engine(64, 64)


def test_product_never_imports_the_oracle():
    """The shipped package must not import, call or link anything under oracle/."""
    pkg = os.path.join(ROOT, "semi-blind-image-deblurring-problems-with-tv_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".m")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dirpath, f)
    code = ("import sys; sys.path.insert(0, %r); import sbd_b200; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)


def test_host_mirror_argument_errors(built):
    """Errors raised before the library is reached mirror the reference's behaviour."""
    g = np.zeros((8, 8))
    with pytest.raises(NameError):                  # 'maxiter' omitted: MaxIter undefined (Q4)
        built.chambolle_prox_TV_stop(g, "lambda", 1.0)
    with pytest.raises(ValueError):                 # dual variables of the wrong size (:102-104)
        built.chambolle_prox_TV_stop(g, "lambda", 1.0, "maxiter", 2, "dualvars", np.zeros((8, 8)))
    with pytest.raises(KeyError):                   # Laplace needs op.x (laplace.m:28)
        built.SAPG_algorithm_laplace(g, {"samples": 3})


def test_make_params_follows_the_reference_constants(built):
    from sbd_b200 import host as H
    op = dict(samples=20, warmup=5, burnIn=16, psf_size=7, phi=0.0, min_th=1e-3, max_th=1.0, th_init=0.01,
              d_exp=0.8, d_scale=1.0, sigma=2.0, sigma_init=60.0, sigma_min=0.1, sigma_max=120.0, fix_sigma=0)
    op["lambda"] = 2.0
    op["gamma"] = 1.9
    g = dict(op, w1_init=0.5, w2_init=0.3, min_w1=0.1, max_w1=1.0, min_w2=0.1, max_w2=1.0, w1=0.4, w2=0.3,
             fix_w1=0, fix_w2=1)
    c = dict(sigma=1000.0, theta=0.01, w1=10.0, w2=10.0, lam=0.5, gam=2.0)
    p = H.make_params(H.GAUSSIAN, g, c)
    assert p.lamb == 1.0 and p.gam == 3.8 and p.prox_lambda == 2.0          # Guassian.m:30-31 vs demo:191 (Q14)
    assert p.c_theta == 0.01 and p.c_sigma2 == 1000.0 and list(p.c_psi) == [10.0, 10.0]
    assert p.sigma2_fixed == 60.0 and p.err_psf_lag == 1 and list(p.fix_psi) == [0, 1]   # Q15, Q9
    m = dict(op, alpha_init=1.0, beta_init=10.0, min_alpha=1e-2, max_alpha=1.0, min_beta=0.1, max_beta=10.0,
             alpha=0.4, beta=3.5, fix_alpha=0, fix_beta=0)
    p = H.make_params(H.MOFFAT, m)
    assert (p.c_theta, list(p.c_psi), p.c_sigma2) == (0.1, [10.0, 10000.0], 10000.0)     # moffat.m:135-138
    assert p.sigma2_fixed == 4.0 and p.err_psf_lag == 0
    l = dict(op, b_init=0.1, min_b=1e-3, max_b=1.0, b=0.3, fix_b=0)
    p = H.make_params(H.LAPLACE, l)
    assert (p.c_theta, p.c_psi[0], p.c_sigma2) == (0.01, 100.0, 10000.0)                # laplace.m:139-141
    assert p.chambolle_maxiter == 25 and p.chambolle_tol == 1e-3 and p.chambolle_tau == 0.249


def test_every_option_is_documented():
    """Each name sbd_set_option accepts is described in include/sbd.h and (the tuning ones) in DESIGN.md's knob table;
    SBD_N_GEOM agrees between the header, the ctypes mirror and the geometry keys of host.py."""
    import re
    src = open(os.path.join(ROOT, PKG, "csrc", "sbd.cu")).read()
    body = src[src.index("int sbd_set_option("):]
    body = body[:body.index("int sbd_get_geometry(")]
    names = re.findall(r'n == "(\w+)"', body)
    assert len(names) >= 9 and "chamb_coop" in names
    header = open(os.path.join(ROOT, "include", "sbd.h")).read()
    design = open(os.path.join(ROOT, "DESIGN.md")).read()
    for n in names:
        assert f'"{n}"' in header, n
        assert f'"{n}"' in design, n
    ngeom = int(re.search(r"#define SBD_N_GEOM (\d+)", header).group(1))
    lib_py = open(os.path.join(ROOT, PKG, "_lib.py")).read()
    assert int(re.search(r"SBD_N_GEOM = (\d+)", lib_py).group(1)) == ngeom
    host_py = open(os.path.join(ROOT, PKG, "host.py")).read()
    keys = re.search(r"keys = \(([^)]*)\)", host_py).group(1)
    assert len(re.findall(r'"\w+"', keys)) == ngeom
