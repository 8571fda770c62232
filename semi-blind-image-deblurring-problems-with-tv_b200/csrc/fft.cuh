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
// 1-D engine: mixed-radix Stockham (radices 16/8/4), at most three stages per
// size; every thread owns 16 complex values per stage.  The first stage reads
// its operands straight from global memory and the last stage writes its results
// straight to global memory, so a 4096-point transform makes only two trips
// through shared memory (one per stage boundary).  Twiddles come from a table
// in global memory (L1-resident): fp64 issue slots are the scarce resource here.
//
// 2-D real FFT of an nx x ny image (nx = fast axis):
//   rows pass  : two real lines per complex length-nx FFT (z = a + i b), split
//                into the two half spectra
//   column pass: C adjacent bins k per block, length-ny complex FFT along q.
// Half-spectrum layout ("tile-major"): spec[tile][q][c], tile = k / C, c = k % C,
// so that the column pass streams one contiguous ny*C*16-byte chunk per block;
// the strided side of the transpose lands on the rows pass, whose 32..128-byte
// pieces of neighbouring lines are combined by the L2.
#pragma once
#include "common.cuh"
#include "psf.cuh"

namespace sbd {

// ---------------------------------------------------------------------------
// plans: radices per size (product = N), first radix = shared-memory pad period
// ---------------------------------------------------------------------------
template <int N> struct FftPlan;
template <> struct FftPlan<4096> { static constexpr int NST = 3, R0 = 16, R1 = 16, R2 = 16; };
template <> struct FftPlan<2048> { static constexpr int NST = 3, R0 = 16, R1 = 16, R2 = 8; };
template <> struct FftPlan<1024> { static constexpr int NST = 3, R0 = 16, R1 = 16, R2 = 4; };
template <> struct FftPlan<512>  { static constexpr int NST = 3, R0 = 8,  R1 = 8,  R2 = 8; };
template <> struct FftPlan<256>  { static constexpr int NST = 2, R0 = 16, R1 = 16, R2 = 1; };
template <> struct FftPlan<128>  { static constexpr int NST = 2, R0 = 16, R1 = 8,  R2 = 1; };
template <> struct FftPlan<64>   { static constexpr int NST = 2, R0 = 8,  R1 = 8,  R2 = 1; };
template <> struct FftPlan<32>   { static constexpr int NST = 2, R0 = 8,  R1 = 4,  R2 = 1; };
template <> struct FftPlan<16>   { static constexpr int NST = 2, R0 = 4,  R1 = 4,  R2 = 1; };

// padded position inside a line: one spare element every R0 positions keeps the
// stride-R0 writes of the first stage off the same shared-memory banks
template <int N>
__device__ __forceinline__ int fft_pad(int p) { return p + p / FftPlan<N>::R0; }
template <int N>
constexpr int fft_line_elems() { return N + N / FftPlan<N>::R0; }

// ---------------------------------------------------------------------------
// small DFTs in registers (natural order in, natural order out)
// ---------------------------------------------------------------------------
template <bool INV>
__device__ __forceinline__ double2 mul_mi(double2 a) {          // a * (-i) forward, a * (+i) inverse
    return INV ? make_double2(-a.y, a.x) : make_double2(a.y, -a.x);
}
template <bool INV>
__device__ __forceinline__ double2 twid(const double2* __restrict__ tw, int idx) {
    double2 w = __ldg(tw + idx);
    if (INV) w.y = -w.y;
    return w;
}
template <bool INV>
__device__ __forceinline__ void dft4(double2& a0, double2& a1, double2& a2, double2& a3) {
    const double2 t0 = cadd(a0, a2), t1 = csub(a0, a2);
    const double2 t2 = cadd(a1, a3), t3 = mul_mi<INV>(csub(a1, a3));
    a0 = cadd(t0, t2); a1 = cadd(t1, t3); a2 = csub(t0, t2); a3 = csub(t1, t3);
}
// multiply by W_16^m (forward) or its conjugate (inverse), m = 0..9
template <bool INV, int M>
__device__ __forceinline__ double2 mul_w16(double2 a) {
    if (M == 0) return a;
    if (M == 4) return mul_mi<INV>(a);
    if (M == 8) return make_double2(-a.x, -a.y);
    constexpr double C1 = 0.92387953251128675613, S1 = 0.38268343236508977173, H = 0.70710678118654752440;
    constexpr double c = (M == 1) ? C1 : (M == 2) ? H : (M == 3) ? S1 : (M == 6) ? -H : (M == 9) ? -C1 : 0.0;
    constexpr double s = (M == 1) ? S1 : (M == 2) ? H : (M == 3) ? C1 : (M == 6) ? H : (M == 9) ? -S1 : 0.0;
    // forward twiddle = c - i s
    const double si = INV ? -s : s;
    return make_double2(fma(a.x, c, a.y * si), fma(a.y, c, -a.x * si));
}

template <int R, bool INV> struct Dft;
template <bool INV> struct Dft<4, INV> {
    static __device__ __forceinline__ void run(double2* v) { dft4<INV>(v[0], v[1], v[2], v[3]); }
};
template <bool INV> struct Dft<8, INV> {
    // n = 2 n1 + n2 (R1 = 4 over n1, R2 = 2 over n2), k = k1 + 4 k2
    static __device__ __forceinline__ void run(double2* v) {
        double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
        double2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
        dft4<INV>(e0, e1, e2, e3);
        dft4<INV>(o0, o1, o2, o3);
        o1 = mul_w16<INV, 2>(o1); o2 = mul_w16<INV, 4>(o2); o3 = mul_w16<INV, 6>(o3);
        v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
        v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
        v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
        v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
    }
};
template <bool INV> struct Dft<16, INV> {
    // n = 4 n1 + n2, k = k1 + 4 k2: DFT4 over n1, twiddle W16^(n2 k1), DFT4 over n2
    static __device__ __forceinline__ void run(double2* v) {
        double2 y[4][4];
#pragma unroll
        for (int n2 = 0; n2 < 4; ++n2) {
            y[n2][0] = v[n2]; y[n2][1] = v[4 + n2]; y[n2][2] = v[8 + n2]; y[n2][3] = v[12 + n2];
            dft4<INV>(y[n2][0], y[n2][1], y[n2][2], y[n2][3]);
        }
        y[1][1] = mul_w16<INV, 1>(y[1][1]); y[1][2] = mul_w16<INV, 2>(y[1][2]); y[1][3] = mul_w16<INV, 3>(y[1][3]);
        y[2][1] = mul_w16<INV, 2>(y[2][1]); y[2][2] = mul_w16<INV, 4>(y[2][2]); y[2][3] = mul_w16<INV, 6>(y[2][3]);
        y[3][1] = mul_w16<INV, 3>(y[3][1]); y[3][2] = mul_w16<INV, 6>(y[3][2]); y[3][3] = mul_w16<INV, 9>(y[3][3]);
#pragma unroll
        for (int k1 = 0; k1 < 4; ++k1) {
            dft4<INV>(y[0][k1], y[1][k1], y[2][k1], y[3][k1]);
            v[k1] = y[0][k1]; v[k1 + 4] = y[1][k1]; v[k1 + 8] = y[2][k1]; v[k1 + 12] = y[3][k1];
        }
    }
};

// ---------------------------------------------------------------------------
// one Stockham stage for the 16 values a thread owns.
//   tl  : thread index inside the line, 0 .. N/16-1
//   src(p) -> double2 : operand at (unpadded) position p of the line
//   dst(p, v)         : result for position p
// The 16/R butterflies of a thread are jb = tl + m*N/16.
// ---------------------------------------------------------------------------
template <int N, int R, int NS, bool INV, class Src>
__device__ __forceinline__ void stage_load(int tl, double2 (&v)[16], Src src, const double2* __restrict__ tw) {
    constexpr int M = 16 / R, TL = N / 16, Q = N / R;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int jb = tl + m * TL;
#pragma unroll
        for (int r = 0; r < R; ++r) v[m * R + r] = src(jb + r * Q);
        if (NS > 1) {
            // twiddles w^r, w = W_N^(k*STEP): only the powers 1, 2, 4, 8 are fetched from the table, the
            // rest are products of at most three of them (<= 4 ulp).  Fetching all R-1 costs several
            // times more LSU wavefronts than moving the data itself (scattered 16-byte table reads).
            const int k = jb & (NS - 1);
            constexpr int STEP = N / (NS * R);
            double2 w[R];
            w[1] = twid<INV>(tw, k * STEP);
            if (R >= 4) { w[2] = twid<INV>(tw, 2 * k * STEP); w[3] = cmul(w[1], w[2]); }
            if (R >= 8) {
                w[4] = twid<INV>(tw, 4 * k * STEP);
                w[5] = cmul(w[1], w[4]); w[6] = cmul(w[2], w[4]); w[7] = cmul(w[3], w[4]);
            }
            if (R >= 16) {
                w[8] = twid<INV>(tw, 8 * k * STEP);
#pragma unroll
                for (int r = 1; r < 8; ++r) w[8 + r] = cmul(w[r], w[8]);
            }
#pragma unroll
            for (int r = 1; r < R; ++r) v[m * R + r] = cmul(v[m * R + r], w[r]);
        }
        Dft<R, INV>::run(&v[m * R]);
    }
}
template <int N, int R, int NS, class Dst>
__device__ __forceinline__ void stage_store(int tl, const double2 (&v)[16], Dst dst) {
    constexpr int M = 16 / R, TL = N / 16;
#pragma unroll
    for (int m = 0; m < M; ++m) {
        const int jb = tl + m * TL;
        const int k = jb & (NS - 1);
        const int base = (jb - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) dst(base + r * NS, v[m * R + r]);
    }
}

// the operand fetch of a first stage on its own (same order as stage_load<N, R, 1>)
template <int N, int R, class Src>
__device__ __forceinline__ void stage_fetch(int tl, double2 (&v)[16], Src src) {
    constexpr int M = 16 / R, TL = N / 16, Q = N / R;
#pragma unroll
    for (int m = 0; m < M; ++m) {
#pragma unroll
        for (int r = 0; r < R; ++r) v[m * R + r] = src(tl + m * TL + r * Q);
    }
}

// Full transform of the line(s) a block holds.  src0 feeds the first stage, dstL
// receives the last stage; the stage boundaries go through the shared-memory
// line `s` (padded, element stride `cs`, i.e. position p lives at s[fft_pad(p)*cs]).
// `active` threads own data; every thread of the block must call this.
template <int N, bool INV, bool SYNC_LAST, class Src, class Dst>
__device__ __forceinline__ void fft_run(bool active, int tl, double2* __restrict__ s, int cs,
                                        const double2* __restrict__ tw, Src src0, Dst dstL) {
    // SYNC_LAST: the last stage writes into the same shared-memory line it reads (in place)
    using P = FftPlan<N>;
    double2 v[16];
    auto smem_src = [&](int p) { return s[fft_pad<N>(p) * cs]; };
    auto smem_dst = [&](int p, double2 x) { s[fft_pad<N>(p) * cs] = x; };
    if (active) {
        stage_load<N, P::R0, 1, INV>(tl, v, src0, tw);
        stage_store<N, P::R0, 1>(tl, v, smem_dst);
    }
    __syncthreads();
    if constexpr (P::NST == 3) {
        if (active) stage_load<N, P::R1, P::R0, INV>(tl, v, smem_src, tw);
        __syncthreads();
        if (active) stage_store<N, P::R1, P::R0>(tl, v, smem_dst);
        __syncthreads();
        if (active) stage_load<N, P::R2, P::R0 * P::R1, INV>(tl, v, smem_src, tw);
        if (SYNC_LAST) __syncthreads();
        if (active) stage_store<N, P::R2, P::R0 * P::R1>(tl, v, dstL);
    } else {
        if (active) stage_load<N, P::R1, P::R0, INV>(tl, v, smem_src, tw);
        if (SYNC_LAST) __syncthreads();
        if (active) stage_store<N, P::R1, P::R0>(tl, v, dstL);
    }
}

// Same transform, but the caller consumes the 16 results of the last stage itself: `last(v)` receives the
// registers of the last stage_load, v[m*R + r] being the result for position jb + r*(N/R), jb = tl + m*N/16.
template <int N, bool INV, class Src, class Last>
__device__ __forceinline__ void fft_run_last(bool active, int tl, double2* __restrict__ s, int cs,
                                             const double2* __restrict__ tw, Src src0, Last last) {
    using P = FftPlan<N>;
    double2 v[16];
    auto smem_src = [&](int p) { return s[fft_pad<N>(p) * cs]; };
    auto smem_dst = [&](int p, double2 x) { s[fft_pad<N>(p) * cs] = x; };
    if (active) {
        stage_load<N, P::R0, 1, INV>(tl, v, src0, tw);
        stage_store<N, P::R0, 1>(tl, v, smem_dst);
    }
    __syncthreads();
    if constexpr (P::NST == 3) {
        if (active) stage_load<N, P::R1, P::R0, INV>(tl, v, smem_src, tw);
        __syncthreads();
        if (active) stage_store<N, P::R1, P::R0>(tl, v, smem_dst);
        __syncthreads();
        if (active) { stage_load<N, P::R2, P::R0 * P::R1, INV>(tl, v, smem_src, tw); last(v); }
    } else {
        if (active) { stage_load<N, P::R1, P::R0, INV>(tl, v, smem_src, tw); last(v); }
    }
}

// ---------------------------------------------------------------------------
// half-spectrum addressing (tile-major)
// ---------------------------------------------------------------------------
struct SpecGeom {
    int C, logC, ny;            // bins per tile, log2(C), number of lines
};
__device__ __forceinline__ size_t spec_idx(const SpecGeom& g, int k, int q) {
    return ((size_t)(k >> g.logC) * g.ny + q) * g.C + (k & (g.C - 1));
}

// ---------------------------------------------------------------------------
// L2 prefetch of what the block `pf` positions later in launch order will read.
// The passes hold one to three blocks per SM and a block's first stage waits for
// all its data (DRAM ~2000 cycles under load here), so the block that takes over
// the SM one wave later should find its input in L2.  pf = resident blocks.
// Measured at 4096^2 x 8: likelihood column pass 1.22 -> 1.10 ms, forward rows pass 0.58 -> 0.55 ms.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void l2_prefetch(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// contiguous region, one prefetch per 128-byte line
__device__ __forceinline__ void l2_prefetch_span(const void* p, size_t bytes) {
    const char* b = reinterpret_cast<const char*>(p);
    for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)blockDim.x * 128) l2_prefetch(b + o);
}

// ---------------------------------------------------------------------------
// rows pass, forward: real lines -> half spectra.  One block = LP line pairs.
// grid = (ny/2/LP, batch), block = max(32, LP*N/16), smem = LP*fft_line_elems<N>()*16
// ---------------------------------------------------------------------------
template <int N>
__global__ void k_rows_fwd(const double* __restrict__ x, double2* __restrict__ spec, SpecGeom sg, int LP,
                           size_t img_stride, size_t spec_stride, const double2* __restrict__ tw, int pf) {
    extern __shared__ double2 fsm[];
    constexpr int TL = N / 16, LE = fft_line_elems<N>();
    const int img = blockIdx.y;
    const int line0 = 2 * blockIdx.x * LP;
    {
        const unsigned int nxt = blockIdx.y * gridDim.x + blockIdx.x + pf;
        if (pf > 0 && nxt < gridDim.x * gridDim.y) {
            const unsigned int i2 = nxt / gridDim.x, b2 = nxt - i2 * gridDim.x;
            l2_prefetch_span(x + (size_t)i2 * img_stride + (size_t)(2 * b2 * LP) * N, (size_t)2 * LP * N * sizeof(double));
        }
    }
    const bool active = threadIdx.x < LP * TL;
    const int b = threadIdx.x / TL, tl = threadIdx.x - b * TL;
    const double* xa = x + (size_t)img * img_stride + (size_t)(line0 + 2 * b) * N;
    double2* line = fsm + (size_t)b * LE;
    // every stage lands in shared memory: the split below needs Z[k] and Z[N-k] together
    auto src = [&](int p) { return make_double2(__ldg(xa + p), __ldg(xa + N + p)); };
    auto dst = [&](int p, double2 v) { line[fft_pad<N>(p)] = v; };
    fft_run<N, false, true>(active, tl, line, 1, tw, src, dst);
    __syncthreads();
    double2* so = spec + (size_t)img * spec_stride;
    constexpr int HB = N / 2 + 1;
    for (int e = threadIdx.x; e < LP * HB; e += blockDim.x) {
        const int bb = e / HB, k = e - bb * HB;
        const double2* L = fsm + (size_t)bb * LE;
        const double2 zk = L[fft_pad<N>(k)], zm = L[fft_pad<N>((N - k) & (N - 1))];
        const int ja = line0 + 2 * bb;
        so[spec_idx(sg, k, ja)] = make_double2(0.5 * (zk.x + zm.x), 0.5 * (zk.y - zm.y));
        so[spec_idx(sg, k, ja + 1)] = make_double2(0.5 * (zk.y + zm.y), 0.5 * (zm.x - zk.x));
    }
}

// rows pass, inverse: half spectra -> real lines (unnormalised; the 1/(nx*ny)
// factor is folded into the spectral multiply of the column pass)
// (no prefetch here: its registers cost this pass the third resident block, 0.87 -> 0.68 ms at 4096^2 x 8)
template <int N>
__global__ void __launch_bounds__(256, 3)
k_rows_inv(const double2* __restrict__ spec, double* __restrict__ out, SpecGeom sg, int LP,
           size_t img_stride, size_t spec_stride, const double2* __restrict__ tw) {
    extern __shared__ double2 fsm[];
    constexpr int TL = N / 16, LE = fft_line_elems<N>();
    const int img = blockIdx.y;
    const int line0 = 2 * blockIdx.x * LP;
    const double2* si = spec + (size_t)img * spec_stride;
    constexpr int HB = N / 2 + 1;
    for (int e = threadIdx.x; e < LP * HB; e += blockDim.x) {
        const int bb = e / HB, k = e - bb * HB;
        double2* L = fsm + (size_t)bb * LE;
        const int ja = line0 + 2 * bb;
        const double2 A = __ldg(si + spec_idx(sg, k, ja));
        const double2 Bv = __ldg(si + spec_idx(sg, k, ja + 1));
        if (k == 0 || k == N / 2) {
            L[fft_pad<N>(k)] = make_double2(A.x, Bv.x);     // C2R: imaginary parts of the real bins dropped
        } else {
            L[fft_pad<N>(k)] = make_double2(A.x - Bv.y, A.y + Bv.x);
            L[fft_pad<N>(N - k)] = make_double2(A.x + Bv.y, Bv.x - A.y);
        }
    }
    __syncthreads();
    const bool active = threadIdx.x < LP * TL;
    const int b = threadIdx.x / TL, tl = threadIdx.x - b * TL;
    double2* line = fsm + (size_t)b * LE;
    double* xo = out + (size_t)img * img_stride + (size_t)(line0 + 2 * b) * N;
    // the first stage reads the packed spectrum from shared memory; it must finish
    // reading before the in-place writes of the same stage start
    using P = FftPlan<N>;
    double2 v[16];
    auto smem_src = [&](int p) { return line[fft_pad<N>(p)]; };
    auto smem_dst = [&](int p, double2 z) { line[fft_pad<N>(p)] = z; };
    auto gdst = [&](int p, double2 z) { xo[p] = z.x; xo[N + p] = z.y; };
    if (active) stage_load<N, P::R0, 1, true>(tl, v, smem_src, tw);
    __syncthreads();
    if (active) stage_store<N, P::R0, 1>(tl, v, smem_dst);
    __syncthreads();
    if constexpr (P::NST == 3) {
        if (active) stage_load<N, P::R1, P::R0, true>(tl, v, smem_src, tw);
        __syncthreads();
        if (active) stage_store<N, P::R1, P::R0>(tl, v, smem_dst);
        __syncthreads();
        if (active) {
            stage_load<N, P::R2, P::R0 * P::R1, true>(tl, v, smem_src, tw);
            stage_store<N, P::R2, P::R0 * P::R1>(tl, v, gdst);
        }
    } else {
        if (active) {
            stage_load<N, P::R1, P::R0, true>(tl, v, smem_src, tw);
            stage_store<N, P::R1, P::R0>(tl, v, gdst);
        }
    }
}

// ---------------------------------------------------------------------------
// column pass.  grid = (batch, ntiles*nsub); block = max(32, C*N/16);
// dynamic smem = C*fft_line_elems<N>()*16 bytes.  Thread -> (tl, c), c fastest.
// ---------------------------------------------------------------------------
enum ColMode {
    COL_FWD = 0,            // spectrum of y; SYM: stored PRE-ROTATED, Y' = conj((wk w)^3) Y^ (see k_cols)
    COL_FWD_REDUCE = 1,     // X^ = fft; store; rss, c0, c1 against Y^ with the CURRENT parameters
    COL_MUL_INV = 2,        // G^ = conj(H)(H X^ - Y^) * inv_scale ; inverse fft ; store
    COL_OP = 3,             // fft ; multiply by the selected kernel / (nx*ny) ; inverse fft ; store
    COL_FILTER = 4          // fft ; X^ = R^ / (|H|^2 + mu) ; rss = sum |Y^ - H X^|^2 ; inverse fft ; store   (SALSA invLS)
};

struct ColArgs {
    const double2* in;      // input half spectrum (rows-pass output or X^)
    double2* out;           // output
    const double2* yhat;    // Y^ (shared by all images of the batch)
    const double2* coef;    // PSF column coefficients [3][nk][MAXT]
    const double2* tw;      // twiddles of size N (this pass)
    const double2* tw_x;    // twiddles of the rows pass (size nxfull): wk of the point-symmetric PSF form
    const Control* ctl;
    double* partials;       // [batch][ntiles][4]
    unsigned int* counters; // [batch]
    double* stats;          // [batch][NSTAT] (rss -> 1, c0 -> 2, c1 -> 3), already divided by nx*ny
    size_t spec_stride;
    int nk, nxfull, t, npsi, C, logC, opsel;   // C = columns per block
    int LC, nsub, ntiles;                      // layout tile width, blocks per tile, tiles per image
    int pf;                                    // L2 prefetch distance in blocks (0 = off)
    double opscale;
    double mu;              // COL_FILTER: the ADMM penalty in 1/(|H|^2 + mu)   (run_Gaussian_demo.m:224)
};

// SYM: psf_size == 7, point-symmetric PSF evaluation (psf_sym3); otherwise the Horner form.  A compile-time
// switch: with both forms in one kernel the 16-times unrolled epilogue doubles to 160 KB of code.
//
// SYM and the phase.  Every kernel of the family is K^ = ph * S with ph = (wk w)^3 of unit modulus, the same for
// the PSF and its derivatives and independent of the parameters, and S REAL.  The likelihood passes only ever
// see K^ next to Y^ in |H X^ - Y^|^2, Re conj(D X^)(H X^ - Y^) and conj(H)(H X^ - Y^), all of which are unchanged
// when Y^ is replaced by Y' = conj(ph) Y^ and every K^ by its real factor:
//     |sh X^ - Y'|^2,   s_d Re conj(X^)(sh X^ - Y'),   sh (sh X^ - Y').
// So Y^ is stored pre-rotated (COL_FWD, once per run) and the per-step passes do real multiplications only; the
// real factors along the 16 bins a thread owns come from psf_seq (psf.cuh).  COL_OP (the A / A' / dif_* operator
// entry) is the only mode that needs the true phase.
template <int N, int MODE, bool SYM>
__global__ void __launch_bounds__(512) k_cols(const ColArgs a) {
    extern __shared__ double2 fsm[];
    __shared__ double2 coefS[8][3][MAXT];
    __shared__ double redS[3 * 32];
    constexpr int TL = N / 16;
    // grid = (batch, blocks per image): the images of a batch are neighbours in launch order, so the
    // Y^ tile they all read comes from DRAM once and from L2 for the rest of the batch
    const int img = blockIdx.x, bx = blockIdx.y;
    const int C = a.C, LC = a.LC;
    const int tile = bx / a.nsub, sub = bx - tile * a.nsub;
    const int k0 = tile * LC + sub * C;
    {
        const unsigned int nxt = blockIdx.y * gridDim.x + blockIdx.x + a.pf;
        if (a.pf > 0 && nxt < gridDim.x * gridDim.y) {
            const unsigned int b2 = nxt / gridDim.x, i2 = nxt - b2 * gridDim.x;
            const unsigned int t2 = b2 / a.nsub, s2 = b2 - t2 * a.nsub;
            const size_t off2 = (size_t)t2 * N * LC + (size_t)s2 * C;
            const double2* p2 = a.in + (size_t)i2 * a.spec_stride + off2;
            const bool yh2 = (MODE == COL_MUL_INV || MODE == COL_FWD_REDUCE || MODE == COL_FILTER) && i2 == 0;   // once per batch
            if (C == LC) {
                l2_prefetch_span(p2, (size_t)N * C * sizeof(double2));
                if (yh2) l2_prefetch_span(a.yhat + off2, (size_t)N * C * sizeof(double2));
            } else {
                for (int q = threadIdx.x; q < N; q += blockDim.x) {
                    l2_prefetch(p2 + (size_t)q * LC);
                    if (yh2) l2_prefetch(a.yhat + off2 + (size_t)q * LC);
                }
            }
        }
    }
    const size_t tile_off = (size_t)img * a.spec_stride + (size_t)tile * N * LC + (size_t)sub * C;
    const double2* in = a.in + tile_off;
    double2* out = a.out + tile_off;
    const double2* yh = a.yhat + (size_t)tile * N * LC + (size_t)sub * C;

    if (MODE != COL_FWD || SYM) {
        for (int e = threadIdx.x; e < C * 3 * a.t; e += blockDim.x) {
            const int j = e % a.t, m = (e / a.t) % 3, c = e / (3 * a.t);
            const int k = min(k0 + c, a.nk - 1);
            const double2* cf = a.coef + ((size_t)m * a.nk + k) * MAXT;
            if (SYM) {                               // point-symmetric form, see psf_sym3
                const double2 wk = __ldg(a.tw_x + k);
                const double2 wk3 = cmul(cmul(wk, wk), wk);
                if (j <= 3) coefS[c][m][j] = psf_sym3_coef(cf, make_double2(wk3.x, -wk3.y), j);
                else if (j == 4) coefS[c][m][j] = wk3;
            } else {
                coefS[c][m][j] = __ldg(cf + j);
            }
        }
    }
    constexpr bool sym = SYM;
    const bool active = threadIdx.x < C * TL;
    const int c = threadIdx.x & (C - 1), tl = threadIdx.x >> a.logC;
    // The coefficients are first needed behind a stage barrier, except by the gradient pass, whose first
    // stage multiplies by H: there the raw operands are fetched BEFORE the block waits for the handful of
    // threads that build the coefficients (ncu: that barrier alone was 13 % of the pass's stall samples; taking it
    // out of the way is free but gains nothing measurable - the wait moves to the first use of the operands).
    double2 raw[16];
    if (MODE == COL_MUL_INV) {
        if (active) stage_fetch<N, FftPlan<N>::R0>(tl, raw, [&](int q) { return __ldg(in + (size_t)q * LC + c); });
        __syncthreads();
        if (SYM && active) {
            // G^ = sc * sh * (sh X^ - Y'): real factors along the R0 operands of each first-stage butterfly
            constexpr int R = FftPlan<N>::R0, M = 16 / R, Q = N / R;
            const double sc = a.ctl->inv_scale;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int jb = tl + m * TL;
                const PsfSeqK kh = psf_seq_prep(coefS[c][0], psf_seq_w(__ldg(a.tw + jb)));
                static_for<0, R / 4>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    double sh[4];
                    psf_seq_eval4<R, r>(kh, sh);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = r + i * (R / 4);
                        const double2 yv = __ldg(yh + (size_t)(jb + rr * Q) * LC + c);
                        const double2 v = raw[m * R + rr];
                        const double t = sh[i] * sc;
                        raw[m * R + rr] = make_double2(t * fma(sh[i], v.x, -yv.x), t * fma(sh[i], v.y, -yv.y));
                    }
                });
            }
        }
    }

    const int k = k0 + c;
    const bool kin = k < a.nk;                       // the last tile may be partly empty
    double2* line = fsm + c;                         // element stride C between positions
    double acc[3] = {0.0, 0.0, 0.0};
    auto kern = [&](int m, double2 w) {              // K^[k, q] of kernel m (0: h, 1: dh/dpsi0, 2: dh/dpsi1)
        if (sym) {
            const PsfW3 p = psf_w3(w);
            const double2 ph = cmul(coefS[c][0][4], p.w3);         // (wk w)^3
            const double sv = psf_sym3(coefS[c][m], p);
            return make_double2(sv * ph.x, sv * ph.y);
        }
        return psf_horner(coefS[c][m], a.t, w);
    };

    int nraw = 0;                                    // stage_load asks for the operands in stage_fetch order
    auto gsrc = [&](int q) {
        double2 v;
        if (MODE == COL_MUL_INV) v = raw[nraw++];
        else v = __ldg(in + (size_t)q * LC + c);
        if (MODE == COL_MUL_INV && !SYM) {
            const double2 w = __ldg(a.tw + q);
            const double2 H = kern(0, w);
            const double2 yv = __ldg(yh + (size_t)q * LC + c);
            const double2 R = csub(cmul(H, v), yv);                 // H X^ - Y^
            const double2 G = cmulc(R, H);                          // conj(H) R
            const double sc = a.ctl->inv_scale;
            v = make_double2(G.x * sc, G.y * sc);
        }
        return v;
    };
    auto gdst = [&](int q, double2 v) {
        if (!kin) return;
        if (MODE == COL_FWD && SYM) {                // Y' = conj((wk w)^3) Y^
            const PsfW3 p = psf_w3(__ldg(a.tw + q));
            v = cmulc(v, cmul(coefS[c][0][4], p.w3));
        }
        out[(size_t)q * LC + c] = v;
        if (MODE == COL_FWD_REDUCE) {
            const double2 w = __ldg(a.tw + q);
            const double2 yv = __ldg(yh + (size_t)q * LC + c);
            // Hermitian weights of the half spectrum: interior bins count twice
            const double wt = (k == 0 || 2 * k == a.nxfull) ? 1.0 : 2.0;
            double2 R, T0, T1 = make_double2(0.0, 0.0);
            if (sym) {
                // (only reached when the block epilogue below is not used) Y^ is pre-rotated: real factors only
                const PsfW3 p = psf_w3(w);
                const double sh = psf_sym3(coefS[c][0], p), s0 = psf_sym3(coefS[c][1], p);
                R = make_double2(fma(sh, v.x, -yv.x), fma(sh, v.y, -yv.y));
                T0 = make_double2(s0 * v.x, s0 * v.y);
                if (a.npsi > 1) { const double s1 = psf_sym3(coefS[c][2], p); T1 = make_double2(s1 * v.x, s1 * v.y); }
            } else {
                R = csub(cmul(psf_horner(coefS[c][0], a.t, w), v), yv);
                T0 = cmul(psf_horner(coefS[c][1], a.t, w), v);
                if (a.npsi > 1) T1 = cmul(psf_horner(coefS[c][2], a.t, w), v);
            }
            acc[0] += wt * (R.x * R.x + R.y * R.y);
            acc[1] += wt * (T0.x * R.x + T0.y * R.y);               // Re conj(D0 X^) R
            if (a.npsi > 1) acc[2] += wt * (T1.x * R.x + T1.y * R.y);
        }
    };

    if (MODE == COL_OP || MODE == COL_FILTER) {
        // forward transform ending in shared memory, multiply, inverse transform from shared memory
        auto sdst = [&](int q, double2 v) {
            if (MODE == COL_FILTER) {
                const double2 w = __ldg(a.tw + q);
                // SYM: Y^ is stored pre-rotated, |Y^ - H X^| = |Y' - sh X^| with the real factor sh of H
                const double2 H = sym ? make_double2(psf_sym3(coefS[c][0], psf_w3(w)), 0.0) : kern(0, w);
                const double F = 1.0 / ((H.x * H.x + H.y * H.y) + a.mu);       // filter_FFT, demo:224
                const double2 X = make_double2(v.x * F, v.y * F);
                if (kin) {
                    const double2 yv = __ldg(yh + (size_t)q * LC + c);
                    const double2 R = csub(yv, cmul(H, X));                       // Y^ - H X^
                    const double wt = (k == 0 || 2 * k == a.nxfull) ? 1.0 : 2.0;
                    acc[0] += wt * (R.x * R.x + R.y * R.y);
                }
                line[fft_pad<N>(q) * C] = make_double2(X.x * a.opscale, X.y * a.opscale);
                return;
            }
            const double2 w = __ldg(a.tw + q);
            const int m = (a.opsel == SBD_OP_A || a.opsel == SBD_OP_AT) ? 0 : (a.opsel == SBD_OP_D0 ? 1 : 2);
            double2 K = kern(m, w);
            if (a.opsel == SBD_OP_AT) K.y = -K.y;
            const double2 z = cmul(K, v);
            line[fft_pad<N>(q) * C] = make_double2(z.x * a.opscale, z.y * a.opscale);
        };
        auto ssrc = [&](int q) { return line[fft_pad<N>(q) * C]; };
        using P = FftPlan<N>;
        // the last forward stage writes in place: all its reads are done (stage_load precedes
        // stage_store inside fft_run only for the fused case), so run it split by hand
        double2 v[16];
        auto smem_src = [&](int p) { return line[fft_pad<N>(p) * C]; };
        auto smem_dst = [&](int p, double2 x) { line[fft_pad<N>(p) * C] = x; };
        if (active) { stage_load<N, P::R0, 1, false>(tl, v, gsrc, a.tw); stage_store<N, P::R0, 1>(tl, v, smem_dst); }
        __syncthreads();
        if constexpr (P::NST == 3) {
            if (active) stage_load<N, P::R1, P::R0, false>(tl, v, smem_src, a.tw);
            __syncthreads();
            if (active) stage_store<N, P::R1, P::R0>(tl, v, smem_dst);
            __syncthreads();
            if (active) stage_load<N, P::R2, P::R0 * P::R1, false>(tl, v, smem_src, a.tw);
            __syncthreads();
            if (active) stage_store<N, P::R2, P::R0 * P::R1>(tl, v, sdst);
        } else {
            if (active) stage_load<N, P::R1, P::R0, false>(tl, v, smem_src, a.tw);
            __syncthreads();
            if (active) stage_store<N, P::R1, P::R0>(tl, v, sdst);
        }
        __syncthreads();
        // inverse: first stage reads shared memory and writes it in place
        if (active) stage_load<N, P::R0, 1, true>(tl, v, ssrc, a.tw);
        __syncthreads();
        if (active) stage_store<N, P::R0, 1>(tl, v, smem_dst);
        __syncthreads();
        if constexpr (P::NST == 3) {
            if (active) stage_load<N, P::R1, P::R0, true>(tl, v, smem_src, a.tw);
            __syncthreads();
            if (active) stage_store<N, P::R1, P::R0>(tl, v, smem_dst);
            __syncthreads();
            if (active) {
                stage_load<N, P::R2, P::R0 * P::R1, true>(tl, v, smem_src, a.tw);
                stage_store<N, P::R2, P::R0 * P::R1>(tl, v, gdst);
            }
        } else {
            if (active) {
                stage_load<N, P::R1, P::R0, true>(tl, v, smem_src, a.tw);
                stage_store<N, P::R1, P::R0>(tl, v, gdst);
            }
        }
    } else if (MODE == COL_MUL_INV) {
        fft_run<N, true, false>(active, tl, line, C, a.tw, gsrc, gdst);
    } else if (MODE == COL_FWD_REDUCE && SYM) {
        // forward transform whose last stage is followed, butterfly by butterfly, by the likelihood sums
        //   rss += |sh X^ - Y'|^2,  c_d += s_d Re conj(X^)(sh X^ - Y')      (Hermitian weight applied once at the end)
        using P = FftPlan<N>;
        constexpr int RL = (P::NST == 3) ? P::R2 : P::R1, NSL = N / RL, ML = 16 / RL;
        auto last = [&](const double2 (&v)[16]) {
#pragma unroll
            for (int m = 0; m < ML; ++m) {
                const int jb = tl + m * TL;
                const PsfSeqW w = psf_seq_w(__ldg(a.tw + jb));
                const PsfSeqK kh = psf_seq_prep(coefS[c][0], w), kd0 = psf_seq_prep(coefS[c][1], w);
                PsfSeqK kd1 = kd0;
                if (a.npsi > 1) kd1 = psf_seq_prep(coefS[c][2], w);
                static_for<0, RL / 4>([&](auto rc) {
                    constexpr int r = decltype(rc)::value;
                    double sh[4], s0[4], s1[4];
                    psf_seq_eval4<RL, r>(kh, sh);
                    psf_seq_eval4<RL, r>(kd0, s0);
                    if (a.npsi > 1) psf_seq_eval4<RL, r>(kd1, s1);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rr = r + i * (RL / 4);
                        const size_t o = (size_t)(jb + rr * NSL) * LC + c;
                        const double2 x = v[m * RL + rr];
                        if (kin) {
                            const double2 yv = __ldg(yh + o);
                            out[o] = x;
                            const double rx = fma(sh[i], x.x, -yv.x), ry = fma(sh[i], x.y, -yv.y);
                            acc[0] = fma(rx, rx, fma(ry, ry, acc[0]));
                            const double d = fma(x.x, rx, x.y * ry);            // Re conj(X^) R'
                            acc[1] = fma(s0[i], d, acc[1]);
                            if (a.npsi > 1) acc[2] = fma(s1[i], d, acc[2]);
                        }
                    }
                });
            }
        };
        fft_run_last<N, false>(active, tl, line, C, a.tw, gsrc, last);
        const double wt = (k == 0 || 2 * k == a.nxfull) ? 1.0 : 2.0;    // Hermitian weights of the half spectrum
        acc[0] *= wt; acc[1] *= wt; acc[2] *= wt;
    } else {
        fft_run<N, false, false>(active, tl, line, C, a.tw, gsrc, gdst);
    }

    if (MODE == COL_FWD_REDUCE || MODE == COL_FILTER) {
        block_sum<3>(acc, redS);
        const unsigned int ntiles = gridDim.y;
        double* part = a.partials + (size_t)img * ntiles * 4;
        if (threadIdx.x == 0) {
            part[bx * 4 + 0] = acc[0];
            part[bx * 4 + 1] = acc[1];
            part[bx * 4 + 2] = acc[2];
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
