"""The SASS-level claims of DESIGN.md, checked on the built library (cuobjdump, no GPU needed): the bulk / tensor copy
instructions of the large-image FFT passes are really there, and the steady-state loop of the fused Chambolle kernel has
the instruction mix the issue model of bench.py / profiles/roofline_traffic.json is built on."""
import json
import os
import re
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "semi-blind-image-deblurring-problems-with-tv_b200", "lib", "libsbd.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not available")


@pytest.fixture(scope="module")
def sass():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    g.build()
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs = {}
    for fn in re.split(r"\n\s*Function : ", out)[1:]:
        funcs[fn.split("\n", 1)[0].strip()] = fn
    return funcs


def one(funcs, *parts):
    hits = [v for k, v in funcs.items() if all(p in k for p in parts)]
    assert len(hits) == 1, (parts, len(hits))
    return hits[0]


def test_blackwell_copy_engines_are_used(sass):
    for mode in (0, 1, 2, 3, 4):                    # every mode of the column pass: bulk copy + mbarrier
        body = one(sass, f"k_cols2ILi4096ELi{mode}E")
        assert "UBLKCP" in body and "SYNCS.ARRIVE.TRANS64" in body and "TRYWAIT" in body
    assert one(sass, "k_cols2ILi4096ELi1E").count("UBLKCP") == 2          # X^ column, then the Y' column
    fwd = one(sass, "k_rows2_fwdILi4096E")
    assert fwd.count("UTMASTG.3D") == 8 and "FENCE.VIEW.ASYNC" in fwd     # eight 256 x 2 boxes
    inv = one(sass, "k_rows2_invILi4096E")
    assert inv.count("UTMALDG.3D") == 8 and "TRYWAIT" in inv
    # nothing on this path is a dense contraction: no tensor-core instruction anywhere in the library
    assert not any(re.search(r"\b(UTC\w*MMA|HMMA|DMMA|IMMA)\b", b) for b in sass.values())


def test_chambolle_loop_mix_matches_the_issue_model(sass):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_loop_mix as M
    model = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["issue_model"]
    body = one(sass, "k_chamb_multiILi4ELb0ELi3ELb0ELi0ELb1E")
    ins = [(int(m.group(1), 16), m.group(3), m.group(4))
           for m in re.finditer(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*([^;]*);", body)]
    loops = []
    for addr, op, args in ins:
        if op.startswith("BRA"):
            t = re.search(r"0x([0-9a-f]+)", args)
            if t and int(t.group(1), 16) <= addr:
                loops.append((int(t.group(1), 16), addr))
    best = None
    for lo, hi in loops:
        ops = [o for a, o, _ in ins if lo <= a <= hi]
        cls = [M.classify(o) for o in ops]
        if cls.count("ctrl") <= 3 and cls.count("mufu") == 32:          # the branch-free two-row trip
            best = cls
    assert best is not None
    fp64 = best.count("fp64")
    assert fp64 == model["fp64_per_trip"] == 340                        # 21.25 fp64 operations per pixel and sweep
    assert best.count("shfl") == 32
    assert abs((len(best) - fp64) - model["other_per_trip"]) <= 12, (len(best), fp64)
    # the variant with the full err sums carries 64 more fp64 instructions per trip
    full = one(sass, "k_chamb_multiILi4ELb0ELi3ELb0ELi0ELb0E")
    assert full.count("DFMA") > body.count("DFMA")


def test_cooperative_prox_kernel(sass):
    """tv_coop.cuh: the per-sweep barrier is one gpu-scope fence + one reduction per block and an acquire poll; the
    operands written by other SMs are read past L1; the level step keeps the MUFU-seeded root / reciprocal (two of each
    for the two pixels of a lane, once); nothing spills under the two-blocks-per-SM bound."""
    body = one(sass, "k_chamb_coop")
    assert body.count("MEMBAR.ALL.GPU") == 1 and body.count("REDG.E.ADD.STRONG.GPU") == 1
    assert "LDG.E.STRONG.GPU" in body and "CCTL.IVALL" in body                       # the polling load of the barrier
    assert len(re.findall(r"LDG\.E\.(64|128)\.STRONG\.GPU", body)) >= 10                # dual pair read from L2
    assert body.count("MUFU.RSQ64H") >= 2 and body.count("MUFU.RCP64H") >= 2       # + the two sqrt of err_k and 1 / lambda
    assert not re.search(r"\b(LDL|STL)\b", body)
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    m = re.search(r"k_chamb_coop[^\n]*\n\s*REG:(\d+)", res)
    assert m and int(m.group(1)) <= 128
