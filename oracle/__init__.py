"""CPU oracle for the SAPG / MYULA semi-blind deblurring hot path.

TEST INFRASTRUCTURE ONLY.  This package is a float64 numpy restatement of the
reference's MATLAB algorithm (charles-kmc/Semi-blind-image-deblurring-problems-with-TV,
mounted read-only at /root/reference while building).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it - and only as the checker, never as the thing that is
shipped or measured.  The product path (`sbd_b200`, `libsbd.so`) never imports
anything from here and fails loudly if the CUDA library is missing.

PARITY PIN STATUS
-----------------
The reference ships no tests, golden vectors or recorded outputs
 and neither MATLAB nor Octave exists in the build image,
so the oracle cannot be pinned against a MATLAB run.  It is pinned instead
against fixtures produced by *executing the reference's own unmodified .m
sources* through the small MATLAB-subset interpreter in `oracle/mlab/`
(`tests/golden/make_golden.py` is the generating script; the fixtures are in
`tests/golden/*.npz`).  Where the interpreter is not used the header of the
module says "parity unpinned".

Conventions: a numpy array `a[i, j]` is MATLAB's `a(i+1, j+1)`; every function
cites the reference file:line it follows.
"""

from . import psf, tv, operators, sapg, philox, metrics  # noqa: F401
