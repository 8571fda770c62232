// fft2.cuh - the 2-D real FFT passes for LARGE images (rows, cols in {1024, 2048, 4096}, psf_size 7), built on the
// Blackwell asynchronous copy engines.  Same arithmetic as fft.cuh (same Stockham stages, twiddles and spectral
// epilogues); what changes is how the data moves.
//
// Layout: column-contiguous half spectrum  spec[img][k][q]  (k = bin of the fast axis, q = 0..ny-1), i.e. the
// tile-major layout of fft.cuh with tile width C = 1.
//
// Column pass (k_cols2): one block = ONE column = ny*16 bytes (64 KB at 4096), brought into shared memory by a single
//   cp.async.bulk (SASS UBLKCP) that completes on an mbarrier; the Stockham stages run in place in a DENSE buffer with
//   an XOR swizzle instead of padding (swz below), so a block needs exactly ny*16 bytes and several blocks share an
//   SM (two at 4096, where the register file is the limit): while one waits for its column the other computes.
//   (The padded two-column tiles of fft.cuh are 139 KB: one block per SM, load / butterflies / store serialised -
//   ncu r01: 1.0 ms, 35 % of the HBM peak.)
//   A 2-CTA cluster exchanging halves through distributed shared memory was considered and dropped: DSMEM moves
//   ~17-21 B/clk/SM (B300_MICROARCH.md), the exchange of half a column per CTA would cost as much as its HBM traffic.
// Rows passes (k_rows2_fwd / k_rows2_inv): one block = one pair of real lines = one complex FFT of length nx.  The
//   transposed side of the pass - 32-byte pieces (two q's of one k), one per column - is not issued by the warps any
//   more: the block assembles its [k][2] slab in shared memory and the TMA moves it as 256 x 2 boxes of a 3-D tensor
//   map (cp.async.bulk.tensor.3d, SASS UTMASTG / UTMALDG).  tools/proto/tma_transpose.cu measures that path alone at
//   5.8 TB/s (store) / 4.9 TB/s (load) for exactly this access pattern; issued as LSU stores it costs 2049 wavefronts
//   per block on top of the shared-memory traffic of the FFT itself (ncu r01: LSU 81 % busy).
#pragma once
#include <cuda.h>
#include "fft.cuh"

namespace sbd {

// position p of a line -> slot in the dense buffer.  Conflict-free for every access of the radix-16 first stage
// (writes p = 16*jb + r: the low three bits become (r ^ jb) & 7 for eight consecutive jb) and neutral for the
// others (their eight lanes share (p >> 4) & 7).  Sizes >= 1024 only (first radix 16).
__device__ __forceinline__ int swz(int p) { return p ^ ((p >> 4) & 7); }

// ---- PTX: mbarrier, bulk copy, tensor copy -----------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
// contiguous global -> shared, completes `bytes` on the mbarrier (bytes: multiple of 16, 16-byte aligned both sides)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, unsigned long long* b) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, int c0, int c1, int c2, const void* src) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (before a TMA store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

constexpr int TMA_KBOX = 256;       // rows (bins k) per box of the rows-pass tensor copies; box = 256 x 2 complex
// Slab [k][2] of the rows passes as the TMA lays it out with CU_TENSOR_MAP_SWIZZLE_32B: the two 16-byte halves of
// row k are exchanged when bit 7 of the row's byte offset is set, i.e. slot(k, w) = 2k + (w ^ ((k >> 2) & 1)).
// With it eight consecutive k (one shared-memory phase of a 16-byte access) hit eight different bank groups
// (plain [k][2]: lanes i and i+4 collide, two wavefronts per phase).
__device__ __forceinline__ int slab(int k, int w) { return 2 * k + (w ^ ((k >> 2) & 1)); }

// ---------------------------------------------------------------------------
// rows pass, forward.  grid = (ny/2, batch), block = N/16, dynamic smem = N*16 bytes (+ static)
// tm: 3-D tensor map of the output spectrum as doubles: dims {2*ny, nk, batch}, box {4, 256, 1}
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(N / 16, (N == 4096) ? 3 : 1)
k_rows2_fwd(const double* __restrict__ x, double2* __restrict__ spec, const __grid_constant__ CUtensorMap tm, int ny,
            size_t img_stride, size_t spec_stride, const double2* __restrict__ tw) {
    extern __shared__ __align__(1024) double2 buf[];
    using P = FftPlan<N>;
    static_assert(P::NST == 3 && P::R0 == 16, "v2 passes: 1024 <= N <= 4096");
    constexpr int TL = N / 16;
    const int img = blockIdx.y, tl = threadIdx.x;
    const int q0 = 2 * blockIdx.x;                                  // the two lines of this block: q0, q0 + 1
    const double* xa = x + (size_t)img * img_stride + (size_t)q0 * N;
    auto gsrc = [&](int p) { return make_double2(__ldg(xa + p), __ldg(xa + N + p)); };      // z = a + i b
    auto ssrc = [&](int p) { return buf[swz(p)]; };
    auto sdst = [&](int p, double2 v) { buf[swz(p)] = v; };
    double2 v[16];
    stage_load<N, P::R0, 1, false>(tl, v, gsrc, tw);
    stage_store<N, P::R0, 1>(tl, v, sdst);
    __syncthreads();
    stage_load<N, P::R1, P::R0, false>(tl, v, ssrc, tw);
    __syncthreads();
    stage_store<N, P::R1, P::R0>(tl, v, sdst);
    __syncthreads();
    stage_load<N, P::R2, P::R0 * P::R1, false>(tl, v, ssrc, tw);
    __syncthreads();
    // The results of the last stage are Z[jb + 256 r] (jb = tl + m*TL, r < R2).  The split below pairs Z[k] with
    // Z[N-k]: a thread's own k < N/2 are exactly its results with r < R2/2, so those stay in registers and only the
    // upper half (r >= R2/2: the Z[N-k] of other threads, and Z[N/2]) goes through shared memory.
    constexpr int RL = P::R2, ML = 16 / RL;
#pragma unroll
    for (int m = 0; m < ML; ++m)
#pragma unroll
        for (int r = RL / 2; r < RL; ++r) buf[swz(tl + m * TL + r * (N / RL))] = v[m * RL + r];
    __syncthreads();
    // split Z into the half spectra A (line q0) and B (line q0 + 1):  A[k] = (Z[k] + conj Z[N-k]) / 2,
    // B[k] = (Z[k] - conj Z[N-k]) / (2i); thread t takes k = t + j*TL, j = 0..7 (k < N/2), all reads first
    double2 zk[8], zm[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = tl + j * TL;
        zk[j] = v[(j % ML) * RL + j / ML];                          // = Z[k]: k = tl + (j % ML)*TL + 256*(j / ML)
        zm[j] = (k == 0) ? zk[j] : buf[swz(N - k)];                 // Z[N-0] = Z[0]
    }
    double2 zn = make_double2(0.0, 0.0);
    if (tl == 0) zn = buf[swz(N / 2)];
    __syncthreads();
    // slab [k][2] (32 bytes per k), dense: the source of the tensor stores
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = tl + j * TL;
        buf[slab(k, 0)] = make_double2(0.5 * (zk[j].x + zm[j].x), 0.5 * (zk[j].y - zm[j].y));
        buf[slab(k, 1)] = make_double2(0.5 * (zk[j].y + zm[j].y), 0.5 * (zm[j].x - zk[j].x));
    }
    fence_async_smem();
    __syncthreads();
    if (tl == 0) {
#pragma unroll
        for (int bx = 0; bx < N / 2 / TMA_KBOX; ++bx)
            tma_store_3d(&tm, 2 * q0, bx * TMA_KBOX, img, buf + 2 * bx * TMA_KBOX);
        // Nyquist bin k = N/2 (Z[N/2] pairs with itself): 32 bytes, plain stores
        double2* so = spec + (size_t)img * spec_stride + (size_t)(N / 2) * ny + q0;
        so[0] = make_double2(zn.x, 0.0);
        so[1] = make_double2(zn.y, 0.0);
        tma_store_commit_wait();
    }
}

// ---------------------------------------------------------------------------
// rows pass, inverse (unnormalised).  Same geometry; tm = tensor map of the INPUT spectrum.
// ---------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(N / 16, (N == 4096) ? 3 : 1)
k_rows2_inv(const double2* __restrict__ spec, double* __restrict__ out, const __grid_constant__ CUtensorMap tm, int ny,
            size_t img_stride, size_t spec_stride, const double2* __restrict__ tw) {
    extern __shared__ __align__(1024) double2 buf[];
    __shared__ __align__(8) unsigned long long mbar;
    using P = FftPlan<N>;
    static_assert(P::NST == 3 && P::R0 == 16, "v2 passes: 1024 <= N <= 4096");
    constexpr int TL = N / 16;
    const int img = blockIdx.y, tl = threadIdx.x;
    const int q0 = 2 * blockIdx.x;
    double2 an = make_double2(0.0, 0.0), bn = an;
    if (tl == 0) {
        mbar_init(&mbar, 1);
        mbar_expect_tx(&mbar, (uint32_t)(N / 2) * 32u);
#pragma unroll
        for (int bx = 0; bx < N / 2 / TMA_KBOX; ++bx)
            tma_load_3d(buf + 2 * bx * TMA_KBOX, &tm, 2 * q0, bx * TMA_KBOX, img, &mbar);
        const double2* si = spec + (size_t)img * spec_stride + (size_t)(N / 2) * ny + q0;       // Nyquist bin
        an = __ldg(si); bn = __ldg(si + 1);
    }
    __syncthreads();
    mbar_wait(&mbar, 0);
    // slab [k][2] -> Z[k] = A[k] + i B[k],  Z[N-k] = conj(A[k]) + i conj(B[k]);  all reads first (in place)
    double2 A[8], B[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = tl + j * TL;
        A[j] = buf[slab(k, 0)];
        B[j] = buf[slab(k, 1)];
    }
    __syncthreads();
    // The first stage (radix 16) of thread t reads Z[t + r*N/16], r = 0..15: r < 8 are exactly the Z[k] this thread
    // forms from its own slab entries - they stay in registers; only Z[N-k] (and Z[N/2]) go through shared memory.
    double2 raw[16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = tl + j * TL;
        if (k == 0) {
            raw[j] = make_double2(A[j].x, B[j].x);                  // C2R: imaginary parts of the real bins dropped
        } else {
            raw[j] = make_double2(A[j].x - B[j].y, A[j].y + B[j].x);
            buf[swz(N - k)] = make_double2(A[j].x + B[j].y, B[j].x - A[j].y);
        }
    }
    if (tl == 0) buf[swz(N / 2)] = make_double2(an.x, bn.x);
    __syncthreads();
    double* xo = out + (size_t)img * img_stride + (size_t)q0 * N;
    auto ssrc = [&](int p) { return buf[swz(p)]; };
    auto sdst = [&](int p, double2 z) { buf[swz(p)] = z; };
    auto gdst = [&](int p, double2 z) { xo[p] = z.x; xo[N + p] = z.y; };
    double2 v[16];
#pragma unroll
    for (int r = 8; r < 16; ++r) raw[r] = buf[swz(tl + r * TL)];
    int nraw = 0;
    auto rsrc = [&](int) { return raw[nraw++]; };                    // stage_load asks for tl + r*N/16, r = 0..15, in order
    stage_load<N, P::R0, 1, true>(tl, v, rsrc, tw);
    __syncthreads();
    stage_store<N, P::R0, 1>(tl, v, sdst);
    __syncthreads();
    stage_load<N, P::R1, P::R0, true>(tl, v, ssrc, tw);
    __syncthreads();
    stage_store<N, P::R1, P::R0>(tl, v, sdst);
    __syncthreads();
    stage_load<N, P::R2, P::R0 * P::R1, true>(tl, v, ssrc, tw);
    stage_store<N, P::R2, P::R0 * P::R1>(tl, v, gdst);
}

// ---------------------------------------------------------------------------
// column pass.  grid = (batch, nk), block = N/16, dynamic smem = N*16 bytes.  Modes as in fft.cuh (SYM form only).
// ---------------------------------------------------------------------------
// Registers: the 16 complex values a thread owns are 64 registers by themselves, so the register file holds two
// columns per SM (2 x 256 threads x 128 registers), not three; the third block's worth of shared memory stays free.
template <int N, int MODE>
__global__ void __launch_bounds__(N / 16, (N == 4096) ? 2 : 1) k_cols2(const ColArgs a) {
    extern __shared__ __align__(1024) double2 buf[];
    __shared__ double2 coefS[3][MAXT];
    __shared__ double redS[3 * 32];
    __shared__ __align__(8) unsigned long long mbar;
    using P = FftPlan<N>;
    static_assert(P::NST == 3 && P::R0 == 16, "v2 passes: 1024 <= N <= 4096");
    constexpr int TL = N / 16;
    const int img = blockIdx.x, k = blockIdx.y, tl = threadIdx.x;
    const size_t col = (size_t)img * a.spec_stride + (size_t)k * N;
    const double2* in = a.in + col;
    double2* out = a.out + col;
    const double2* yh = a.yhat + (size_t)k * N;
    if (tl == 0) {
        mbar_init(&mbar, 1);
        mbar_expect_tx(&mbar, (uint32_t)N * 16u);
        bulk_g2s(buf, in, (uint32_t)N * 16u, &mbar);
    }
    // the batch shares Y': the first image's block pulls the column of the NEXT bin into L2 ahead of its use
    // (measured: asking for the block's OWN Y' column in L1 here instead makes the pass 8 % slower)
    if (img == 0 && k + 1 < a.nk && (MODE == COL_MUL_INV || MODE == COL_FWD_REDUCE || MODE == COL_FILTER))
        l2_prefetch_span(a.yhat + (size_t)(k + 1) * N, (size_t)N * sizeof(double2));
    for (int e = tl; e < 3 * 7; e += TL) {                            // point-symmetric coefficients of this bin
        const int j = e % 7, m = e / 7;
        const double2* cf = a.coef + ((size_t)m * a.nk + k) * MAXT;
        const double2 wk = __ldg(a.tw_x + k);
        const double2 wk3 = cmul(cmul(wk, wk), wk);
        if (j <= 3) coefS[m][j] = psf_sym3_coef(cf, make_double2(wk3.x, -wk3.y), j);
        else if (j == 4) coefS[m][j] = wk3;
    }
    __syncthreads();
    mbar_wait(&mbar, 0);

    auto ssrc = [&](int p) { return buf[swz(p)]; };
    auto sdst = [&](int p, double2 v) { buf[swz(p)] = v; };
    auto kern = [&](int m, double2 w) {              // K^[k, q] of kernel m, with its phase (COL_OP only)
        const PsfW3 p = psf_w3(w);
        const double2 ph = cmul(coefS[0][4], p.w3);
        const double sv = psf_sym3(coefS[m], p);
        return make_double2(sv * ph.x, sv * ph.y);
    };
    double acc[3] = {0.0, 0.0, 0.0};
    double2 v[16];

    // ---- first stage: operands in natural order straight from the bulk copy (linear slots, not swizzled)
    constexpr int R0 = P::R0, M0 = 16 / R0, Q0 = N / R0;
    double2 raw[16];
    stage_fetch<N, R0>(tl, raw, [&](int q) { return buf[q]; });
    if (MODE == COL_MUL_INV) {
        // G^ = sc * sh * (sh X^ - Y'): real factors along the R0 operands of each first-stage butterfly
        const double sc = a.ctl->inv_scale;
#pragma unroll
        for (int m = 0; m < M0; ++m) {
            const int jb = tl + m * TL;
            const PsfSeqK kh = psf_seq_prep(coefS[0], psf_seq_w(__ldg(a.tw + jb)));
            static_for<0, R0 / 4>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                double sh[4];
                psf_seq_eval4<R0, r>(kh, sh);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = r + i * (R0 / 4);
                    const double2 yv = __ldg(yh + jb + rr * Q0);
                    const double2 x = raw[m * R0 + rr];
                    const double t = sh[i] * sc;
                    raw[m * R0 + rr] = make_double2(t * fma(sh[i], x.x, -yv.x), t * fma(sh[i], x.y, -yv.y));
                }
            });
        }
    }
    int nraw = 0;
    auto rsrc = [&](int) { return raw[nraw++]; };                    // stage_load asks in stage_fetch order
    constexpr bool INV1 = (MODE == COL_MUL_INV);
    stage_load<N, R0, 1, INV1>(tl, v, rsrc, a.tw);
    __syncthreads();                                                  // every thread has read its raw operands
    stage_store<N, R0, 1>(tl, v, sdst);
    __syncthreads();
    stage_load<N, P::R1, R0, INV1>(tl, v, ssrc, a.tw);
    __syncthreads();
    stage_store<N, P::R1, R0>(tl, v, sdst);
    __syncthreads();
    constexpr int RL = P::R2, NSL = N / RL, ML = 16 / RL;
    if (MODE == COL_FWD_REDUCE) {
        // the operands of the last stage go to registers, which frees the buffer: the Y' column of this bin is
        // bulk-copied into it while the last twiddles and butterflies run, and the likelihood sums then read Y'
        // from shared memory (16 dependent global loads per thread at the very end of the block were the
        // top stall of this pass: ncu long_scoreboard 4.4 warps per issue cycle)
        stage_fetch<N, RL>(tl, raw, ssrc);
        fence_async_smem();                          // order this thread's earlier generic accesses before the async write
        __syncthreads();
        if (tl == 0) {
            mbar_expect_tx(&mbar, (uint32_t)N * 16u);
            bulk_g2s(buf, yh, (uint32_t)N * 16u, &mbar);
        }
        nraw = 0;
        stage_load<N, RL, R0 * P::R1, INV1>(tl, v, rsrc, a.tw);
        mbar_wait(&mbar, 1);
    } else {
        stage_load<N, RL, R0 * P::R1, INV1>(tl, v, ssrc, a.tw);
    }

    if (MODE == COL_MUL_INV) {
        stage_store<N, RL, NSL>(tl, v, [&](int q, double2 z) { out[q] = z; });
    } else if (MODE == COL_FWD) {
        // spectrum of y, stored pre-rotated: Y' = conj((wk w)^3) Y^
        stage_store<N, RL, NSL>(tl, v, [&](int q, double2 z) {
            const PsfW3 p = psf_w3(__ldg(a.tw + q));
            out[q] = cmulc(z, cmul(coefS[0][4], p.w3));
        });
    } else if (MODE == COL_FWD_REDUCE) {
        //   rss += |sh X^ - Y'|^2,  c_d += s_d Re conj(X^)(sh X^ - Y')      (Hermitian weight applied once at the end)
#pragma unroll
        for (int m = 0; m < ML; ++m) {
            const int jb = tl + m * TL;
            const PsfSeqW w = psf_seq_w(__ldg(a.tw + jb));
            const PsfSeqK kh = psf_seq_prep(coefS[0], w), kd0 = psf_seq_prep(coefS[1], w);
            PsfSeqK kd1 = kd0;
            if (a.npsi > 1) kd1 = psf_seq_prep(coefS[2], w);
            static_for<0, RL / 4>([&](auto rc) {
                constexpr int r = decltype(rc)::value;
                double sh[4], s0[4], s1[4];
                psf_seq_eval4<RL, r>(kh, sh);
                psf_seq_eval4<RL, r>(kd0, s0);
                if (a.npsi > 1) psf_seq_eval4<RL, r>(kd1, s1);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = r + i * (RL / 4);
                    const int q = jb + rr * NSL;
                    const double2 x = v[m * RL + rr];
                    const double2 yv = buf[q];                              // Y' column, natural order
                    out[q] = x;
                    const double rx = fma(sh[i], x.x, -yv.x), ry = fma(sh[i], x.y, -yv.y);
                    acc[0] = fma(rx, rx, fma(ry, ry, acc[0]));
                    const double d = fma(x.x, rx, x.y * ry);                // Re conj(X^) R'
                    acc[1] = fma(s0[i], d, acc[1]);
                    if (a.npsi > 1) acc[2] = fma(s1[i], d, acc[2]);
                }
            });
        }
        const double wt = (k == 0 || 2 * k == a.nxfull) ? 1.0 : 2.0;
        acc[0] *= wt; acc[1] *= wt; acc[2] *= wt;
    } else {
        // COL_OP / COL_FILTER: the forward transform ends in shared memory with the spectral multiply applied, then the
        // inverse transform runs from there
        __syncthreads();
        stage_store<N, RL, NSL>(tl, v, [&](int q, double2 z) {
            const double2 w = __ldg(a.tw + q);
            if (MODE == COL_FILTER) {
                // x = invLS(r): X^ = R^ / (|H|^2 + mu) (run_Gaussian_demo.m:224); rss = |Y' - sh X^|^2 (Y' pre-rotated)
                const double sh = psf_sym3(coefS[0], psf_w3(w));
                const double F = 1.0 / (sh * sh + a.mu);
                const double2 X = make_double2(z.x * F, z.y * F);
                const double2 yv = __ldg(yh + q);
                const double rx = yv.x - sh * X.x, ry = yv.y - sh * X.y;
                const double wt = (k == 0 || 2 * k == a.nxfull) ? 1.0 : 2.0;
                acc[0] += wt * (rx * rx + ry * ry);
                buf[swz(q)] = make_double2(X.x * a.opscale, X.y * a.opscale);
            } else {
                const int m = (a.opsel == SBD_OP_A || a.opsel == SBD_OP_AT) ? 0 : (a.opsel == SBD_OP_D0 ? 1 : 2);
                double2 K = kern(m, w);
                if (a.opsel == SBD_OP_AT) K.y = -K.y;
                const double2 y = cmul(K, z);
                buf[swz(q)] = make_double2(y.x * a.opscale, y.y * a.opscale);
            }
        });
        __syncthreads();
        stage_load<N, R0, 1, true>(tl, v, ssrc, a.tw);
        __syncthreads();
        stage_store<N, R0, 1>(tl, v, sdst);
        __syncthreads();
        stage_load<N, P::R1, R0, true>(tl, v, ssrc, a.tw);
        __syncthreads();
        stage_store<N, P::R1, R0>(tl, v, sdst);
        __syncthreads();
        stage_load<N, P::R2, R0 * P::R1, true>(tl, v, ssrc, a.tw);
        stage_store<N, RL, NSL>(tl, v, [&](int q, double2 z) { out[q] = z; });
    }

    if (MODE == COL_FWD_REDUCE || MODE == COL_FILTER) {
        block_sum<3>(acc, redS);
        const unsigned int ncolumns = gridDim.y;
        double* part = a.partials + (size_t)img * ncolumns * 4;
        if (tl == 0) {
            part[k * 4 + 0] = acc[0];
            part[k * 4 + 1] = acc[1];
            part[k * 4 + 2] = acc[2];
        }
        if (last_block_ticket(a.counters + img, ncolumns)) {
            if (tl < 32) {
                const double invP = 1.0 / ((double)a.nxfull * (double)N);
                const double s0 = warp_sum_partials(part + 0, (int)ncolumns, 4);
                const double s1 = warp_sum_partials(part + 1, (int)ncolumns, 4);
                const double s2 = warp_sum_partials(part + 2, (int)ncolumns, 4);
                if (tl == 0) {
                    double* st = a.stats + (size_t)img * NSTAT;
                    st[1] = s0 * invP; st[2] = s1 * invP; st[3] = s2 * invP;
                }
            }
        }
    }
}

}  // namespace sbd
