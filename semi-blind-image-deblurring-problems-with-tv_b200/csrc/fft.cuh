// fft.cuh - shared-memory Stockham FFT (fp64, power-of-two 16..4096) and the
// 2-D real transforms built from it, fused with the Fourier-domain PSF work.
//
// This replaces MATLAB's fft2/ifft2 inside the closures
//   A  = real(ifft2(H_FFT .* fft2(x)))           run_Gaussian_demo.m:136
//   AT = real(ifft2(conj(H_FFT) .* fft2(x)))     run_Gaussian_demo.m:137
//   dif_w = real(ifft2(dH_FFT .* fft2(x)))       run_Gaussian_demo.m:138-139
// and the reductions of op.f / op.grad_w1 / op.grad_w2 / op.gradF_sigma
// (run_Gaussian_demo.m:171-175) through Parseval (DESIGN.md "Fusion identities").
//
// 2-D real FFT of an nx x ny image (nx = fast axis):
//   rows pass  : two real lines per complex length-nx FFT (z = a + i b), split
//                into the two half spectra -> spec[q][k], k = 0..nx/2
//   column pass: C adjacent bins k per block, length-ny complex FFT along q.
#pragma once
#include "common.cuh"
#include "psf.cuh"

namespace sbd {

constexpr int FFT_ITER = 4;                 // radix-4 butterflies held in registers per thread

template <int N> struct ILog2 { static constexpr int v = 1 + ILog2<N / 2>::v; };
template <> struct ILog2<1> { static constexpr int v = 0; };

// (x,y) * (-i) forward, * (+i) inverse
template <bool INV>
__device__ __forceinline__ double2 mul_mi(double2 a) {
    return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
template <bool INV>
__device__ __forceinline__ double2 twid(const double2* __restrict__ tw, int idx) {
    double2 w = __ldg(tw + idx);
    if (INV) w.y = -w.y;
    return w;
}

// One radix-4 Stockham stage over B lines of length N stored at s[b*ld + n].
// Every thread of the block must call this (contains __syncthreads()).
template <int N, int NS, bool INV>
__device__ __forceinline__ void stockham_r4(double2* __restrict__ s, int ld, int B,
                                            const double2* __restrict__ tw) {
    constexpr int Q = N / 4;
    const int total = B * Q;
    double2 r[FFT_ITER][4];
    int dst[FFT_ITER];
#pragma unroll
    for (int it = 0; it < FFT_ITER; ++it) {
        const int idx = threadIdx.x + it * blockDim.x;
        dst[it] = -1;
        if (idx < total) {
            const int b = idx / Q, j = idx - b * Q;
            const int k = j & (NS - 1);
            const double2* L = s + b * ld;
            double2 a0 = L[j], a1 = L[j + Q], a2 = L[j + 2 * Q], a3 = L[j + 3 * Q];
            if (NS > 1) {
                constexpr int STEP = N / (4 * NS);
                a1 = cmul(a1, twid<INV>(tw, k * STEP));
                a2 = cmul(a2, twid<INV>(tw, 2 * k * STEP));
                a3 = cmul(a3, twid<INV>(tw, 3 * k * STEP));
            }
            const double2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
            const double2 t2 = cadd(a1, a3), t3 = mul_mi<INV>(csub(a1, a3));
            r[it][0] = cadd(t0, t2);
            r[it][1] = cadd(t1, t3);
            r[it][2] = csub(t0, t2);
            r[it][3] = csub(t1, t3);
            dst[it] = b * ld + ((j - k) << 2) + k;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < FFT_ITER; ++it) {
        if (dst[it] >= 0) {
            double2* D = s + dst[it];
            D[0] = r[it][0]; D[NS] = r[it][1]; D[2 * NS] = r[it][2]; D[3 * NS] = r[it][3];
        }
    }
    __syncthreads();
}

// Final radix-2 stage (NS = N/2) when log2(N) is odd.
template <int N, bool INV>
__device__ __forceinline__ void stockham_r2_last(double2* __restrict__ s, int ld, int B,
                                                 const double2* __restrict__ tw) {
    constexpr int H = N / 2;
    const int total = B * H;
    double2 r[2 * FFT_ITER][2];
    int dst[2 * FFT_ITER];
#pragma unroll
    for (int it = 0; it < 2 * FFT_ITER; ++it) {
        const int idx = threadIdx.x + it * blockDim.x;
        dst[it] = -1;
        if (idx < total) {
            const int b = idx / H, j = idx - b * H;
            const double2* L = s + b * ld;
            const double2 a0 = L[j];
            const double2 a1 = cmul(L[j + H], twid<INV>(tw, j));
            r[it][0] = cadd(a0, a1);
            r[it][1] = csub(a0, a1);
            dst[it] = b * ld + j;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < 2 * FFT_ITER; ++it) {
        if (dst[it] >= 0) { s[dst[it]] = r[it][0]; s[dst[it] + H] = r[it][1]; }
    }
    __syncthreads();
}

template <int N, int NS, bool INV>
__device__ __forceinline__ void stockham_stages(double2* s, int ld, int B, const double2* tw) {
    if constexpr (NS * 4 <= N) {
        stockham_r4<N, NS, INV>(s, ld, B, tw);
        stockham_stages<N, NS * 4, INV>(s, ld, B, tw);
    } else if constexpr (NS * 2 == N) {
        stockham_r2_last<N, INV>(s, ld, B, tw);
    }
}

// In-place FFT of B lines (natural order in, natural order out).  The caller
// must __syncthreads() after filling `s`; on return the data is synchronised.
// Threads required: B*N/(4*FFT_ITER) (at least 32).
template <int N, bool INV>
__device__ __forceinline__ void fft_lines(double2* s, int ld, int B, const double2* tw) {
    stockham_stages<N, 1, INV>(s, ld, B, tw);
}

// ---------------------------------------------------------------------------
// rows pass, forward: real lines -> half spectra
// grid = (ny/2/LP, batch), dynamic smem = LP*N*16 bytes
// ---------------------------------------------------------------------------
template <int N>
__global__ void k_rows_fwd(const double* __restrict__ x, double2* __restrict__ spec, int sp, int LP,
                           size_t img_stride, size_t spec_stride, const double2* __restrict__ tw) {
    extern __shared__ double2 fsm[];
    const int img = blockIdx.y;
    const int line0 = 2 * blockIdx.x * LP;
    const double* xi = x + (size_t)img * img_stride + (size_t)line0 * N;
    double2* so = spec + (size_t)img * spec_stride + (size_t)line0 * sp;
    for (int e = threadIdx.x; e < LP * N; e += blockDim.x) {
        const int b = e / N, n = e - b * N;
        const double* p = xi + (size_t)(2 * b) * N + n;
        fsm[e] = make_double2(__ldg(p), __ldg(p + N));
    }
    __syncthreads();
    fft_lines<N, false>(fsm, N, LP, tw);
    constexpr int HB = N / 2 + 1;
    for (int e = threadIdx.x; e < LP * HB; e += blockDim.x) {
        const int b = e / HB, k = e - b * HB;
        const double2 zk = fsm[b * N + k], zm = fsm[b * N + ((N - k) & (N - 1))];
        so[(size_t)(2 * b) * sp + k] = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
        so[(size_t)(2 * b + 1) * sp + k] = make_double2(0.5 * (zk.y + zm.y), 0.5 * (zm.x - zk.x));
    }
}

// rows pass, inverse: half spectra -> real lines (unnormalised; the 1/(nx*ny)
// factor is folded into the spectral multiply of the column pass)
template <int N>
__global__ void k_rows_inv(const double2* __restrict__ spec, double* __restrict__ out, int sp, int LP,
                           size_t img_stride, size_t spec_stride, const double2* __restrict__ tw) {
    extern __shared__ double2 fsm[];
    const int img = blockIdx.y;
    const int line0 = 2 * blockIdx.x * LP;
    const double2* si = spec + (size_t)img * spec_stride + (size_t)line0 * sp;
    double* xo = out + (size_t)img * img_stride + (size_t)line0 * N;
    constexpr int HB = N / 2 + 1;
    for (int e = threadIdx.x; e < LP * HB; e += blockDim.x) {
        const int b = e / HB, k = e - b * HB;
        const double2 A = __ldg(si + (size_t)(2 * b) * sp + k);
        const double2 Bv = __ldg(si + (size_t)(2 * b + 1) * sp + k);
        if (k == 0 || k == N / 2) {
            fsm[b * N + k] = make_double2(A.x, Bv.x);       // C2R: imaginary parts of the real bins dropped
        } else {
            fsm[b * N + k] = make_double2(A.x - Bv.y, A.y + Bv.x);
            fsm[b * N + N - k] = make_double2(A.x + Bv.y, Bv.x - A.y);
        }
    }
    __syncthreads();
    fft_lines<N, true>(fsm, N, LP, tw);
    for (int e = threadIdx.x; e < LP * N; e += blockDim.x) {
        const int b = e / N, n = e - b * N;
        const double2 z = fsm[e];
        double* p = xo + (size_t)(2 * b) * N + n;
        p[0] = z.x;
        p[N] = z.y;
    }
}

// ---------------------------------------------------------------------------
// column pass.  grid = (ceil(nk/C), batch); dynamic smem = C*N*16 bytes.
// ---------------------------------------------------------------------------
enum ColMode {
    COL_FWD = 0,            // spectrum of y (no PSF work)
    COL_FWD_REDUCE = 1,     // X^ = fft; store; rss, c0, c1 against Y^ with the CURRENT parameters
    COL_MUL_INV = 2,        // G^ = conj(H)(H X^ - Y^) * inv_scale ; inverse fft ; store
    COL_OP = 3              // fft ; multiply by the selected kernel / (nx*ny) ; inverse fft ; store
};

struct ColArgs {
    const double2* in;      // input half spectrum (rows-pass output or X^)
    double2* out;           // output
    const double2* yhat;    // Y^ (shared by all images of the batch)
    const double2* coef;    // PSF column coefficients [3][nk][MAXT]
    const double2* tw;      // twiddles of size N (this pass)
    const Control* ctl;
    double* partials;       // [batch][ntiles][4]
    unsigned int* counters; // [batch]
    double* stats;          // [batch][NSTAT] (rss -> 1, c0 -> 2, c1 -> 3), already divided by nx*ny
    size_t spec_stride;
    int sp, nk, nxfull, t, npsi, C, opsel;
    double opscale;
};

template <int N, int MODE>
__global__ void k_cols(const ColArgs a) {
    extern __shared__ double2 fsm[];
    __shared__ double2 coefS[8][3][MAXT];
    __shared__ double redS[3 * 32];
    const int img = blockIdx.y;
    const int C = a.C;
    const int k0 = blockIdx.x * C;
    const double2* in = a.in + (size_t)img * a.spec_stride;
    double2* out = a.out + (size_t)img * a.spec_stride;

    if (MODE != COL_FWD) {
        for (int e = threadIdx.x; e < C * 3 * a.t; e += blockDim.x) {
            const int j = e % a.t, m = (e / a.t) % 3, c = e / (3 * a.t);
            const int k = min(k0 + c, a.nk - 1);
            coefS[c][m][j] = __ldg(a.coef + ((size_t)m * a.nk + k) * MAXT + j);
        }
        __syncthreads();
    }

    // ---- load (+ spectral multiply for MUL_INV)
    for (int e = threadIdx.x; e < C * N; e += blockDim.x) {
        const int c = e % C, q = e / C;
        const int k = k0 + c;
        double2 v = make_double2(0.0, 0.0);
        if (k < a.nk) {
            v = __ldg(in + (size_t)q * a.sp + k);
            if (MODE == COL_MUL_INV) {
                const double2 w = __ldg(a.tw + q);
                const double2 H = psf_horner(coefS[c][0], a.t, w);
                const double2 yv = __ldg(a.yhat + (size_t)q * a.sp + k);
                const double2 R = csub(cmul(H, v), yv);                 // H X^ - Y^
                const double2 G = cmulc(R, H);                          // conj(H) R
                const double sc = a.ctl->inv_scale;
                v = make_double2(G.x * sc, G.y * sc);
            }
        }
        fsm[c * N + q] = v;
    }
    __syncthreads();

    if (MODE == COL_MUL_INV) {
        fft_lines<N, true>(fsm, N, C, a.tw);
    } else {
        fft_lines<N, false>(fsm, N, C, a.tw);
    }

    if (MODE == COL_OP) {
        for (int e = threadIdx.x; e < C * N; e += blockDim.x) {
            const int c = e % C, q = e / C;
            const double2 w = __ldg(a.tw + q);
            const int m = (a.opsel == SBD_OP_A || a.opsel == SBD_OP_AT) ? 0 : (a.opsel == SBD_OP_D0 ? 1 : 2);
            double2 K = psf_horner(coefS[c][m], a.t, w);
            if (a.opsel == SBD_OP_AT) K.y = -K.y;
            const double2 v = cmul(K, fsm[c * N + q]);
            fsm[c * N + q] = make_double2(v.x * a.opscale, v.y * a.opscale);
        }
        __syncthreads();
        fft_lines<N, true>(fsm, N, C, a.tw);
    }

    // ---- store (+ reductions for FWD_REDUCE)
    double acc[3] = {0.0, 0.0, 0.0};
    for (int e = threadIdx.x; e < C * N; e += blockDim.x) {
        const int c = e % C, q = e / C;
        const int k = k0 + c;
        if (k < a.nk) {
            const double2 v = fsm[c * N + q];
            out[(size_t)q * a.sp + k] = v;
            if (MODE == COL_FWD_REDUCE) {
                const double2 w = __ldg(a.tw + q);
                const double2 yv = __ldg(a.yhat + (size_t)q * a.sp + k);
                const double2 H = psf_horner(coefS[c][0], a.t, w);
                const double2 R = csub(cmul(H, v), yv);
                // Hermitian weights of the half spectrum: interior bins count twice
                const double wt = (k == 0 || 2 * k == a.nxfull) ? 1.0 : 2.0;
                acc[0] += wt * (R.x * R.x + R.y * R.y);
                const double2 T0 = cmul(psf_horner(coefS[c][1], a.t, w), v);
                acc[1] += wt * (T0.x * R.x + T0.y * R.y);               // Re conj(D0 X^) R
                if (a.npsi > 1) {
                    const double2 T1 = cmul(psf_horner(coefS[c][2], a.t, w), v);
                    acc[2] += wt * (T1.x * R.x + T1.y * R.y);
                }
            }
        }
    }
    if (MODE == COL_FWD_REDUCE) {
        block_sum<3>(acc, redS);
        const unsigned int ntiles = gridDim.x;
        double* part = a.partials + (size_t)img * ntiles * 4;
        if (threadIdx.x == 0) {
            part[blockIdx.x * 4 + 0] = acc[0];
            part[blockIdx.x * 4 + 1] = acc[1];
            part[blockIdx.x * 4 + 2] = acc[2];
        }
        if (last_block_ticket(a.counters + img, ntiles)) {
            if (threadIdx.x < 32) {
                const double invP = 1.0 / ((double)a.nxfull * (double)N);
                const double s0 = warp_sum_partials(part + 0, (int)ntiles, 4);
                const double s1 = warp_sum_partials(part + 1, (int)ntiles, 4);
                const double s2 = warp_sum_partials(part + 2, (int)ntiles, 4);
                if (threadIdx.x == 0) {
                    double* st = a.stats + (size_t)img * NSTAT;
                    st[1] = s0 * invP; st[2] = s1 * invP; st[3] = s2 * invP;
                }
            }
        }
    }
}

}  // namespace sbd
