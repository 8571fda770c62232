// tv_coop.cuh - the Chambolle prox of SMALL problems as one cooperative launch.
//
// utils/chambolle_prox_TV_stop.m:120-149 (sweeps, stop test, f = g - lambda div p), same arithmetic per pixel as the
// fused kernel of tv_multi.cuh (cm_core: MUFU-seeded square root / reciprocal, ~1 ulp).
//
// Why: on the reference's own images (256^2, 512^2, one chain) the marching kernels of tv_multi.cuh are
// latency-bound - a warp walks its rows one after the other (4-row segments plus 8 halo rows at 256^2), 7 launches
// of ~14 us each, and the prox is ~105 of the 130 us of a MYULA step.  Here every (row, 64-pixel strip) is a unit of
// its own warp, all units of a sweep run side by side, and the sweeps of one prox are separated by a barrier between
// the blocks of the SAME image (images are independent) instead of a kernel boundary: one launch per prox, no halo
// rows, the reference's stop test decided by every block from the same partial sums in the same order.
// Launched with the cooperative attribute (all blocks co-resident), so the spin barrier cannot deadlock.
#pragma once
#include "tv_multi.cuh"

namespace sbd {

// -DSBD_CC_TIMING (tools/proto/coop_bench.cu): thread 0 of block 0 adds up clock64 intervals of the phases of a sweep.
// Measured on B200, 256^2, one image, cycles per sweep: operands of the sweep arrive (L2, written by other SMs) 1625,
// level step 707, stores issued 178, block sum (incl. waiting for the block's slowest warp) 1111, release (MEMBAR.GPU
// + count) 988, wait for the other blocks 854, closing barrier 358 = 5800 cycles, 2.9 us.  Variants that lost:
// per-warp arrival without the block sum (eight fences per block instead of one: 82 vs 78 us per prox), relaxed
// polling instead of acquire (no change: the L1 invalidation is not what delays the operands), 3 or 4 resident blocks
// per SM (80 / 82 us), 4 / 16 / 32 warps per block (84 / 94 / 228 us), the operands of a warp's next unit fetched
// while the current one is worked on (upw > 1; 512^2: 125 vs 116 us - 128 registers, and the other resident warps
// already cover the wait).
#ifdef SBD_CC_TIMING
__device__ long long cc_timing[8];
#define CC_T(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long t__ = clock64(); cc_timing[i] += t__ - cc_t0; cc_t0 = t__; } } while (0)
// the same, but the clock is read only once `val` is available
#define CC_TD(i, val) do { if (blockIdx.x == 0 && threadIdx.x == 0) { long long t__; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t__) : "d"(val) : "memory"); cc_timing[i] += t__ - cc_t0; cc_t0 = t__; } } while (0)
#else
#define CC_T(i) do { } while (0)
#define CC_TD(i, val) do { } while (0)
#endif

#ifndef SBD_CC_WARPS
#define SBD_CC_WARPS 8
#endif
constexpr int CC_WARPS = SBD_CC_WARPS;
constexpr int CC_THREADS = CC_WARPS * 32;

__device__ __forceinline__ unsigned int cc_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void cc_red_release(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double2 cc_ld2(const double* p) { return __ldcg(reinterpret_cast<const double2*>(p)); }

struct CcRow { double px[2], py[2], u[2]; };

// grid.x = batch * bpi (block b of image z = blockIdx.x / bpi works on units [b*CC_WARPS*upw, ...) of that image),
// unit u = (row j = u / nstrips, strip u % nstrips), lane -> pixels i, i+1 (nx even).
// partials: 2 slots x gridDim.x doubles; bar: one counter per image, zero at entry (k_chamb_reset).
__global__ void __launch_bounds__(CC_THREADS, 2)
k_chamb_coop(const double* __restrict__ g, double* px0, double* py0, double* px1, double* py1, double* __restrict__ f,
             int nx, int ny, size_t img_stride, int bpi, int upw, int zero_start, int maxiter_host,
             const Control* __restrict__ ctl, ChambState* __restrict__ st, double* partials, unsigned int* bar,
             Control* ctl_rw, int* trace, int ntrace) {
    __shared__ double sm[32];
    __shared__ double s_err;
    const int z = blockIdx.x / bpi, b = blockIdx.x - z * bpi;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nstrips = (nx + 63) >> 6;
    const int units = ny * nstrips;
    const int u0 = (b * CC_WARPS + warp) * upw;
    const double lambda = ctl->prox_lambda_run, tau = ctl->tau, tol = ctl->tol;
    const double invlam = 1.0 / lambda;
    const int maxiter = max(1, min(maxiter_host, ctl->maxiter));
    const size_t off = (size_t)z * img_stride;
    g += off; f += off; px0 += off; py0 += off; px1 += off; py1 += off;
    const CmK K = cm_consts();
    unsigned int* mybar = bar + z;

    int k = 0;
    double err = 0.0;
#ifdef SBD_CC_TIMING
    long long cc_t0 = clock64();
#endif
    for (int s = 1; s <= maxiter; ++s) {
        const double* pxi = ((s - 1) & 1) ? px1 : px0;
        const double* pyi = ((s - 1) & 1) ? py1 : py0;
        double* pxo = (s & 1) ? px1 : px0;
        double* pyo = (s & 1) ? py1 : py0;
        const bool zero = zero_start && s == 1;              // p^0 = 0 is not read (chambolle_prox_TV_stop.m:68-69)
        double acc[1] = {0.0};
        // err of the PREVIOUS sweep: its partial sums are complete (barrier s-1); warp 0 asks for them now and adds them
        // up after its own units, so the stop test costs no round trip of its own.  If it fires, this sweep was one too
        // many - it wrote the other buffer, p^(s-1) is untouched.
        double psum = 0.0;
        if (warp == 0 && s > 1) {
            const double* pp = partials + (size_t)((s - 1) & 1) * gridDim.x + (size_t)z * bpi;
            for (int q = lane; q < bpi; q += 32) psum += __ldcg(pp + q);
        }
        for (int m = 0; m < upw; ++m) {
            const int u = u0 + m;
            if (u >= units) break;                           // warp-uniform
            const int j = u / nstrips, sx = u - j * nstrips;
            const int i = sx * 64 + 2 * lane;
            const bool on = i < nx;
            const bool last1 = on && (i + 2 >= nx);          // the pair's second pixel is the last column
            const bool edgeR = on && !last1 && lane == 31;   // the pixel right of the pair belongs to another warp
            const bool edgeL = on && lane == 0 && i > 0;
            const bool up = j > 0, down = j + 1 < ny;
            const size_t r = (size_t)j * nx + i;
            double2 P = make_double2(0.0, 0.0), Q = P, Qu = P, Pn = P, Qn = P, G = P, Gn = P;
            double pxL = 0.0, pxnL = 0.0, pxR = 0.0, pyR = 0.0, pyuR = 0.0, gR = 0.0;
            if (on) {
                G = __ldg(reinterpret_cast<const double2*>(g + r));
                if (down) Gn = __ldg(reinterpret_cast<const double2*>(g + r + nx));
                if (edgeR) gR = __ldg(g + r + 2);
                if (!zero) {
                    P = cc_ld2(pxi + r); Q = cc_ld2(pyi + r);
                    if (up) Qu = cc_ld2(pyi + r - nx);
                    if (down) { Pn = cc_ld2(pxi + r + nx); Qn = cc_ld2(pyi + r + nx); }
                    if (edgeL) { pxL = __ldcg(pxi + r - 1); if (down) pxnL = __ldcg(pxi + r + nx - 1); }
                    if (edgeR) { pxR = __ldcg(pxi + r + 2); pyR = __ldcg(pyi + r + 2); if (up) pyuR = __ldcg(pyi + r - nx + 2); }
                }
            }
            CC_TD(5, P.x + Q.y + Qu.x + Pn.y + Qn.x + G.x + Gn.y);        // operands have arrived
            // u on row j (:152-159, :124)
            CcRow h;
            h.px[0] = P.x; h.px[1] = P.y; h.py[0] = Q.x; h.py[1] = Q.y;
            {
                double l = shfl_up_d(P.y, 1);
                if (lane == 0) l = pxL;
                const double ux0 = P.x - l;
                const double ux1 = last1 ? -P.y : P.y - P.x;
                const double uy0 = down ? Q.x - Qu.x : -Q.x;
                const double uy1 = down ? Q.y - Qu.y : -Q.y;
                h.u[0] = (uy0 + ux0) - __dmul_rn(G.x, invlam);
                h.u[1] = (uy1 + ux1) - __dmul_rn(G.y, invlam);
            }
            // u on row j+1; the row below the image does not exist: upy = 0 (:165-166)
            double un[2];
            if (down) {
                double l = shfl_up_d(Pn.y, 1);
                if (lane == 0) l = pxnL;
                const bool lastrow = j + 2 >= ny;
                const double ux0 = Pn.x - l;
                const double ux1 = last1 ? -Pn.y : Pn.y - Pn.x;
                const double uy0 = lastrow ? -Qn.x : Qn.x - Q.x;
                const double uy1 = lastrow ? -Qn.y : Qn.y - Q.y;
                un[0] = (uy0 + ux0) - __dmul_rn(Gn.x, invlam);
                un[1] = (uy1 + ux1) - __dmul_rn(Gn.y, invlam);
            } else {
                un[0] = h.u[0]; un[1] = h.u[1];
            }
            // u of the pixel right of the pair
            double ur = shfl_down_d(h.u[0], 1);
            if (edgeR) {
                const double uy = down ? pyR - pyuR : -pyR;
                ur = (uy + (pxR - P.y)) - __dmul_rn(gR, invlam);
            }
            double upx[2] = {h.u[1] - h.u[0], last1 ? 0.0 : ur - h.u[1]};               // :162-163
            double opx[2], opy[2], ex[2], ey[2];
            cm_core<2>(upx, un, h, tau, opx, opy, ex, ey, K);
            const double e = fma(ex[1], ex[1], fma(ey[1], ey[1], fma(ex[0], ex[0], ey[0] * ey[0])));      // :128
            CC_TD(6, e + opx[1] + opy[0]);                                     // level step done
            if (on) {
                acc[0] += e;
                *reinterpret_cast<double2*>(pxo + r) = make_double2(opx[0], opx[1]);
                *reinterpret_cast<double2*>(pyo + r) = make_double2(opy[0], opy[1]);
            }
        }
        CC_T(0);                                             // stores issued
        if (warp == 0 && s > 1) {
            const double tot = warp_sum(psum);               // same order in every block of the image (warp_sum_partials)
            if (lane == 0) s_err = sqrt(tot);                                                           // :128
        }
        block_sum<1>(acc, sm);                               // (its barriers also publish s_err)
        if (s > 1) {
            err = s_err;                                                                                // err_(s-1)
            if (!(err > tol)) break;                                                                    // :131, k = s-1
        }
        CC_T(1);                                             // block sum
        double* part = partials + (size_t)(s & 1) * gridDim.x + (size_t)z * bpi;
        if (threadIdx.x == 0) {
            part[b] = acc[0];
            // barrier between the blocks of this image: the release orders the block's stores (which happen before it
            // through the __syncthreads inside block_sum and this thread's program order) before the count
            cc_red_release(mybar, 1u);
            CC_T(2);                                         // release (fence + count)
            const unsigned int target = (unsigned int)s * (unsigned int)bpi;
            while (cc_ld_acquire(mybar) < target) { }
            CC_T(3);                                         // wait for the other blocks
        }
        __syncthreads();
        CC_T(4);
        k = s;                                                                                          // :121
    }
    if (k == maxiter) {
        // the loop ran out: err of the last sweep (the value the caller may read; no decision hangs on it)
        if (warp == 0) {
            const double tot = warp_sum_partials(partials + (size_t)(k & 1) * gridDim.x + (size_t)z * bpi, bpi, 1);
            if (lane == 0) s_err = sqrt(tot);
        }
        __syncthreads();
        err = s_err;
    }

    // f = g - lambda div p^k (:149); p^k is complete in buffer k & 1 (barrier k)
    {
        const double* px = (k & 1) ? px1 : px0;
        const double* py = (k & 1) ? py1 : py0;
        for (int m = 0; m < upw; ++m) {
            const int u = u0 + m;
            if (u >= units) break;
            const int j = u / nstrips, sx = u - j * nstrips;
            const int i = sx * 64 + 2 * lane;
            const bool on = i < nx;
            const bool last1 = on && (i + 2 >= nx);
            const size_t r = (size_t)j * nx + i;
            double2 P = make_double2(0.0, 0.0), Q = P, Qu = P, G = P;
            double pxL = 0.0;
            if (on) {
                G = __ldg(reinterpret_cast<const double2*>(g + r));
                P = cc_ld2(px + r); Q = cc_ld2(py + r);
                if (j > 0) Qu = cc_ld2(py + r - nx);
                if (lane == 0 && i > 0) pxL = __ldcg(px + r - 1);
            }
            double l = shfl_up_d(P.y, 1);
            if (lane == 0) l = pxL;
            const double ux0 = P.x - l;
            const double ux1 = last1 ? -P.y : P.y - P.x;
            const double uy0 = (j + 1 < ny) ? Q.x - Qu.x : -Q.x;
            const double uy1 = (j + 1 < ny) ? Q.y - Qu.y : -Q.y;
            if (on)
                *reinterpret_cast<double2*>(f + r) = make_double2(G.x - lambda * (uy0 + ux0), G.y - lambda * (uy1 + ux1));
        }
    }
    if (b == 0 && threadIdx.x == 0) {
        ChambState& S = st[z];
        S.k = k; S.err = err; S.done = 1; S.buf = k & 1; S.emitted = 1; S.redo = 0;
        if (trace && z == 0) {                      // k_chamb_record folded in: sweeps of chain 0 -> SAPG `chambolle_iters`
            const int slot = ctl_rw->prox_count;
            if (slot >= 0 && slot < ntrace) trace[slot] = k;
            ctl_rw->prox_count = slot + 1;
        }
    }
}

}  // namespace sbd
