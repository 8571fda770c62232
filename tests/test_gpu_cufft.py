"""cuFFT as a correctness oracle for the hand-written FFT passes (north_star: "cuFFT is used only as a correctness
oracle"; BASELINE.json configs[4] "vs cuFFT oracle").

A / A' / dif_* are rebuilt from cuFFT's D2Z / Z2D (torch.fft.rfft2 / irfft2 on the GPU call cufftExecD2Z / Z2D) and
the oracle's 7x7 PSF taps, and compared with sbd_blur_dev at the sizes the numpy oracle is too slow for:
2048^2 and 4096^2, batch 8, the three PSF families, relative error 1e-12.  Test-only: libsbd.so does not link cuFFT
(checked below)."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from conftest import rel

pytestmark = pytest.mark.gpu

PSI = {0: (0.4, 0.3), 1: (0.4, 3.5), 2: (0.3,)}


def cufft_blur(x, taps, conj=False):
    """real(ifft2(fft2(pad(taps)) .* fft2(x))) per image with cuFFT (run_Gaussian_demo.m:136-137, utils/resize.m)."""
    import torch
    b, n0, n1 = x.shape
    pad = torch.zeros((n0, n1), dtype=torch.float64, device=x.device)
    t = taps.shape[0]
    pad[:t, :t] = torch.from_numpy(np.ascontiguousarray(taps)).to(x.device)      # top-left corner (resize.m:1-12)
    H = torch.fft.rfft2(pad)
    if conj:
        H = torch.conj(H)
    return torch.fft.irfft2(torch.fft.rfft2(x) * H, s=(n0, n1))


@pytest.mark.parametrize("n", [2048, 4096])
@pytest.mark.parametrize("model", [0, 1, 2])
def test_blur_vs_cufft(n, model):
    import torch
    import oracle as O
    import sbd_b200
    from sbd_b200._lib import lib, c_double_p
    B = 8
    eng = sbd_b200.Engine(n, n, 7, model, 0.0, max_batch=B)
    g = torch.Generator(device="cuda").manual_seed(100 * model + n)
    # images as the library sees them: [b][col][row]; torch sees x_t[b] = image^T, and the 2-D transform of a
    # transposed image is the transposed transform, so cuFFT is applied to x_t with transposed taps
    x_t = torch.rand((B, n, n), dtype=torch.float64, device="cuda", generator=g) * 255.0
    out = torch.empty_like(x_t)
    psi = np.zeros(2); psi[:len(PSI[model])] = PSI[model]
    nops = 3 if model == 2 else 4
    for op in range(nops):
        which = {0: 0, 1: 0, 2: 1, 3: 2}[op]
        taps = O.psf.taps(model, 7, PSI[model], 0.0, which)                   # taps[i, j], i = row
        want = cufft_blur(x_t, taps.T, conj=(op == 1))
        rc = lib.sbd_blur_dev(eng._h, x_t.data_ptr(), psi.ctypes.data_as(c_double_p), op, out.data_ptr(), B)
        assert rc == 0, lib.sbd_last_error(eng._h)
        lib.sbd_synchronize(eng._h)
        num = torch.linalg.vector_norm(out - want).item()
        den = torch.linalg.vector_norm(want).item()
        assert num / den < 1e-12, (n, model, op, num / den)
        # per image too (a single bad image would hide in the batch norm)
        for b in (0, B - 1):
            e = (torch.linalg.vector_norm(out[b] - want[b]) / torch.linalg.vector_norm(want[b])).item()
            assert e < 1e-12, (n, model, op, b, e)
        del want
    eng.close()


def test_cufft_agrees_with_numpy_oracle_small():
    """Pins the cuFFT construction above to the numpy oracle at a size both can do."""
    import torch
    import oracle as O
    n = 256
    rng = np.random.default_rng(0)
    x = rng.uniform(0, 255, (n, n))
    for model in (0, 1, 2):
        cl = O.operators.closures(model, (n, n), 7, 0.0)
        taps = O.psf.taps(model, 7, PSI[model], 0.0, 0)
        xt = torch.from_numpy(np.ascontiguousarray(x.T))[None].cuda()
        got = cufft_blur(xt, taps.T)[0].cpu().numpy().T
        assert rel(got, cl["A"](x, *PSI[model])) < 1e-12
        got = cufft_blur(xt, taps.T, conj=True)[0].cpu().numpy().T
        assert rel(got, cl["AT"](x, *PSI[model])) < 1e-12


def test_libsbd_does_not_link_cufft():
    from sbd_b200._lib import LIB_PATH
    out = subprocess.run(["ldd", LIB_PATH], capture_output=True, text=True).stdout
    assert "cufft" not in out.lower()
    nm = subprocess.run(["nm", "-D", LIB_PATH], capture_output=True, text=True).stdout
    assert "cufft" not in nm.lower()
