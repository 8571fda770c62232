// tv_multi.cuh - T Chambolle sweeps fused in one kernel (temporal blocking).
//
// Same arithmetic as k_chamb_sweep (utils/chambolle_prox_TV_stop.m:121-131), but
// a warp carries T sweep levels in registers while it marches down the rows:
// level s holds one row of (p^s, u^s, g/lambda); when the row below arrives it
// updates its held row to p^(s+1) and hands it to level s+1.  HBM traffic per
// T sweeps is one read of (g, px, py) and one write of (px, py): 40 B/pixel
// instead of 40*T.  Laterally each level loses one pixel of validity per side,
// so a 64-pixel strip produces 64 - 2*HL output pixels (HL = T rounded up to
// even); vertically a segment reads T extra rows above and below.
//
// The stop test of the reference (err_k <= tol, :128,:131) is evaluated after
// the kernel for each of the T sweeps in order (speculation): if it fires at
// sweep s < T the block is re-run from the same input with s levels ("redo"
// launch, a no-op otherwise), so the result is exactly the reference's.
#pragma once
#include "common.cuh"
#include "tv.cuh"

namespace sbd {

__device__ __forceinline__ double fast_rsqrt_seed(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
}
__device__ __forceinline__ double fast_rcp_seed(double a) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    return y;
}
// sqrt(a) for a >= 0 (a == 0 -> exactly 0): MUFU seed + two Goldschmidt steps (~1 ulp)
__device__ __forceinline__ double fast_sqrt(double a) {
    const double y = fast_rsqrt_seed(a + 1e-300);
    double g = a * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    return g;
}
// 1/d for d >= 1: MUFU seed + two Newton steps (~1 ulp)
__device__ __forceinline__ double fast_rcp(double d) {
    double r = fast_rcp_seed(d);
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    return r;
}

struct CmPk { double px[2], py[2], g[2]; };             // a row travelling up the levels
struct CmLv { double px[2], py[2], u[2], g[2]; };       // the row a level holds

struct CmLane {         // lane constants
    int i;              // fast-axis index of this lane's first pixel
    bool in0, in1;      // pixel inside the image
    bool last0, last1;  // pixel is i == nx-1
    bool central;       // both pixels in the strip's output region (and inside the image)
};

// One level step.  `p` comes in as the row received by the level and leaves as
// the row it emits.  GEN = generic path (row flags honoured), EDGE = strip
// touches a lateral image border.
template <bool EDGE, bool GEN>
__device__ __forceinline__ void cm_step(CmLv& h, CmPk& p, const CmLane& L, double tau,
                                        double& err, bool virt, bool last, bool acc) {
    double un[2];
    if (GEN && virt) {
        un[0] = h.u[0]; un[1] = h.u[1];                 // row ny does not exist: upy = 0 (:165-166)
    } else {
        const double pxl = shfl_up_d(p.px[1], 1);
        double ux0 = p.px[0] - pxl, ux1 = p.px[1] - p.px[0];                 // :156-157
        if (EDGE) {
            if (L.last0) ux0 = -p.px[0];
            if (L.last1) ux1 = -p.px[1];
        }
        double uy0 = p.py[0] - h.py[0], uy1 = p.py[1] - h.py[1];            // :153-154
        if (GEN && last) { uy0 = -p.py[0]; uy1 = -p.py[1]; }
        un[0] = (uy0 + ux0) - p.g[0];                                        // :159, :124
        un[1] = (uy1 + ux1) - p.g[1];
    }
    const double ur = shfl_down_d(h.u[0], 1);
    double upx[2] = {h.u[1] - h.u[0], ur - h.u[1]};                          // :162-163
    if (EDGE) {
        if (L.last0) upx[0] = 0.0;
        if (L.last1) upx[1] = 0.0;
    }
    double e = 0.0;
    CmPk o;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        const double upy = un[v] - h.u[v];
        const double s2 = fma(upx[v], upx[v], upy * upy);
        const double tmp = fast_sqrt(s2);                                    // :127
        const double rinv = fast_rcp(fma(tau, tmp, 1.0));
        const double ex = fma(tmp, h.px[v], -upx[v]), ey = fma(tmp, h.py[v], -upy);
        e += fma(ex, ex, ey * ey);                                           // :128
        o.px[v] = fma(tau, upx[v], h.px[v]) * rinv;                          // :129
        o.py[v] = fma(tau, upy, h.py[v]) * rinv;                             // :130
        o.g[v] = h.g[v];
    }
    if (EDGE) {
        if (!L.in0) { o.px[0] = 0.0; o.py[0] = 0.0; }
        if (!L.in1) { o.px[1] = 0.0; o.py[1] = 0.0; }
    }
    if (GEN) err = L.central ? e : 0.0;     // generic path: caller decides whether the row counts
    else err += L.central ? e : 0.0;
#pragma unroll
    for (int v = 0; v < 2; ++v) { h.px[v] = p.px[v]; h.py[v] = p.py[v]; h.u[v] = un[v]; h.g[v] = p.g[v]; }
    p = o;
}

template <bool EDGE>
__device__ __forceinline__ void cm_load(CmPk& p, const double* __restrict__ g, const double* __restrict__ px,
                                        const double* __restrict__ py, size_t off, const CmLane& L, double invlam) {
    if (!EDGE) {
        const double2 a = __ldg(reinterpret_cast<const double2*>(px + off));
        const double2 b = __ldg(reinterpret_cast<const double2*>(py + off));
        const double2 c = __ldg(reinterpret_cast<const double2*>(g + off));
        p.px[0] = a.x; p.px[1] = a.y; p.py[0] = b.x; p.py[1] = b.y;
        p.g[0] = c.x * invlam; p.g[1] = c.y * invlam;
    } else {
        p.px[0] = L.in0 ? __ldg(px + off) : 0.0;      p.px[1] = L.in1 ? __ldg(px + off + 1) : 0.0;
        p.py[0] = L.in0 ? __ldg(py + off) : 0.0;      p.py[1] = L.in1 ? __ldg(py + off + 1) : 0.0;
        p.g[0] = L.in0 ? __ldg(g + off) * invlam : 0.0; p.g[1] = L.in1 ? __ldg(g + off + 1) * invlam : 0.0;
    }
}

template <bool EDGE>
__device__ __forceinline__ void cm_store(const CmPk& p, double* __restrict__ pxo, double* __restrict__ pyo,
                                         size_t off, const CmLane& L) {
    if (!L.central) return;
    if (!EDGE) {
        *reinterpret_cast<double2*>(pxo + off) = make_double2(p.px[0], p.px[1]);
        *reinterpret_cast<double2*>(pyo + off) = make_double2(p.py[0], p.py[1]);
    } else {
        if (L.in0) { pxo[off] = p.px[0]; pyo[off] = p.py[0]; }
        if (L.in1) { pxo[off + 1] = p.px[1]; pyo[off + 1] = p.py[1]; }
    }
}

// One generic step of the march (row r arrives): honours the row flags.
template <int T, bool EDGE>
__device__ __forceinline__ void cm_generic_iter(int r, CmLv (&h)[T], CmPk& nxt, double (&err)[T],
                                                const double* __restrict__ g, const double* __restrict__ pxi,
                                                const double* __restrict__ pyi, double* __restrict__ pxo,
                                                double* __restrict__ pyo, int nx, int ny, int j0, int jlast,
                                                int r0, int rend, long long ibase, const CmLane& L,
                                                double invlam, double tau, int nlev) {
    CmPk p = nxt;
    if (r + 1 <= min(rend, ny - 1))
        cm_load<EDGE>(nxt, g, pxi, pyi, (size_t)((long long)(r + 1) * nx + ibase), L, invlam);
    bool live = true;
#pragma unroll
    for (int s = 0; s < T; ++s) {
        const int jj = r - s;                       // row this level receives
        if (live && jj >= r0 && jj <= ny) {
            const bool virt = (jj == ny), last = (jj == ny - 1), first = (jj == r0);
            const int hr = jj - 1;                  // row being updated by this level
            double e = 0.0;
            // first row of a level: nothing held yet -> only u of that row is formed (with the
            // "row above" = 0, exact at the image top); the emitted row is meaningless
            cm_step<EDGE, true>(h[s], p, L, tau, e, virt, last, true);
            if (first) {
                live = false;
            } else {
                if (hr >= j0 && hr <= jlast) err[s] += e;
                if (s + 1 == nlev) {
                    if (hr >= j0 && hr <= jlast)
                        cm_store<EDGE>(p, pxo, pyo, (size_t)((long long)hr * nx + ibase), L);
                    live = false;
                }
            }
        } else if (jj < r0) {
            live = false;                           // level (and all above) not started yet
        }
        // jj > ny: this level is finished; a higher one may still be working
    }
}

// The march of one warp.  nlev <= T levels are applied.
template <int T, bool EDGE>
__device__ __forceinline__ void cm_march(const double* __restrict__ g, const double* __restrict__ pxi,
                                         const double* __restrict__ pyi, double* __restrict__ pxo,
                                         double* __restrict__ pyo, int nx, int ny, int j0, int j1,
                                         const CmLane& L, double invlam, double tau, int nlev,
                                         double (&err)[T]) {
    CmLv h[T];
#pragma unroll
    for (int s = 0; s < T; ++s) {
#pragma unroll
        for (int v = 0; v < 2; ++v) { h[s].px[v] = 0.0; h[s].py[v] = 0.0; h[s].u[v] = 0.0; h[s].g[v] = 0.0; }
    }
    const int r0 = max(j0 - nlev, 0);
    const int jlast = min(j1 - 1, ny - 1);          // last output row of this segment
    const int rend = jlast + nlev;                  // last step
    const long long ibase = L.i;
    CmPk nxt;
    cm_load<EDGE>(nxt, g, pxi, pyi, (size_t)((long long)r0 * nx + ibase), L, invlam);

    int r = r0;
    if (nlev == T) {
        const int fast_lo = j0 + T, fast_hi = min(j1, ny - 2);
        for (; r < min(fast_lo, rend + 1); ++r)
            cm_generic_iter<T, EDGE>(r, h, nxt, err, g, pxi, pyi, pxo, pyo, nx, ny, j0, jlast, r0, rend, ibase, L, invlam, tau, nlev);
        for (; r <= fast_hi; ++r) {                 // steady state: every level live, no row flags
            CmPk p = nxt;
            cm_load<EDGE>(nxt, g, pxi, pyi, (size_t)((long long)(r + 1) * nx + ibase), L, invlam);
#pragma unroll
            for (int s = 0; s < T; ++s) cm_step<EDGE, false>(h[s], p, L, tau, err[s], false, false, true);
            cm_store<EDGE>(p, pxo, pyo, (size_t)((long long)(r - T) * nx + ibase), L);
        }
    }
    for (; r <= rend; ++r)
        cm_generic_iter<T, EDGE>(r, h, nxt, err, g, pxi, pyi, pxo, pyo, nx, ny, j0, jlast, r0, rend, ibase, L, invlam, tau, nlev);
}

// grid = (ceil(nstrips / TV_WARPS), nsegs, batch); block = TV_THREADS.
// redo == 0: main launch of a block of T sweeps; redo == 1: re-run with the
// number of levels the stop test asked for (no-op unless st.redo != 0).
template <int T>
__global__ void __launch_bounds__(TV_THREADS)
k_chamb_multi(const double* __restrict__ g, const double* __restrict__ pxi, const double* __restrict__ pyi,
              double* __restrict__ pxo, double* __restrict__ pyo, int nx, int ny, int seg, int nstrips,
              size_t img_stride, const Control* __restrict__ ctl, ChambState* __restrict__ st,
              double* __restrict__ partials, int redo) {
    constexpr int HL = (T + 1) & ~1;
    constexpr int WO = 64 - 2 * HL;
    __shared__ double sm[T * 32];
    const int img = blockIdx.z;
    ChambState* S = st + img;
    int nlev;
    if (redo) {
        nlev = S->redo;
        if (nlev == 0) return;
    } else {
        if (S->done) return;
        nlev = min(T, ctl->maxiter - S->k);
    }
    const double lambda = ctl->prox_lambda_theta, tau = ctl->tau;
    const double invlam = 1.0 / lambda;
    const size_t off = (size_t)img * img_stride;
    g += off; pxi += off; pyi += off; pxo += off; pyo += off;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x * TV_WARPS + warp;
    double err[T];
#pragma unroll
    for (int s = 0; s < T; ++s) err[s] = 0.0;
    if (strip < nstrips) {
        CmLane L;
        const int i0 = strip * WO - HL;             // first loaded pixel (may be < 0)
        L.i = i0 + 2 * lane;
        L.in0 = L.i >= 0 && L.i < nx; L.in1 = L.i + 1 >= 0 && L.i + 1 < nx;
        L.last0 = L.i == nx - 1; L.last1 = L.i + 1 == nx - 1;
        const bool cen = (2 * lane >= HL) && (2 * lane + 1 < 64 - HL);
        L.central = cen && (L.in0 || L.in1);
        const int j0 = blockIdx.y * seg, j1 = min(j0 + seg, ny);
        const bool edge = (i0 < 0) || (i0 + 64 > nx);
        if (edge) cm_march<T, true>(g, pxi, pyi, pxo, pyo, nx, ny, j0, j1, L, invlam, tau, nlev, err);
        else      cm_march<T, false>(g, pxi, pyi, pxo, pyo, nx, ny, j0, j1, L, invlam, tau, nlev, err);
    }

    block_sum<T>(err, sm);
    const unsigned int nparts = gridDim.x * gridDim.y;
    double* part = partials + (size_t)img * T * nparts;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < T; ++s) part[(size_t)s * nparts + blockIdx.y * gridDim.x + blockIdx.x] = err[s];
    }
    if (last_block_ticket(&S->counter, nparts)) {
        if (threadIdx.x < 32) {
            double e[T];
#pragma unroll
            for (int s = 0; s < T; ++s) e[s] = sqrt(warp_sum_partials(part + (size_t)s * nparts, (int)nparts, 1));   // :128
            if (threadIdx.x == 0) {
                const int k0 = S->k;
                if (redo) {                         // the stop sweep was decided by the main launch
                    S->k = k0 + nlev; S->done = 1; S->redo = 0; S->buf ^= 1;
                } else {
                    int stop = 0;
                    double estop = 0.0, elast = 0.0;
#pragma unroll
                    for (int s = 1; s <= T; ++s) {
                        if (s <= nlev) {
                            const bool cont = (k0 + s < ctl->maxiter) && (e[s - 1] > ctl->tol);   // :131
                            if (!cont && stop == 0) { stop = s; estop = e[s - 1]; }
                            elast = e[s - 1];
                        }
                    }
                    if (stop == 0) {                // all nlev sweeps continue
                        S->k = k0 + nlev; S->err = elast; S->buf ^= 1;
                    } else if (stop == nlev) {      // stops exactly at the end of this block
                        S->k = k0 + nlev; S->err = estop; S->done = 1; S->buf ^= 1;
                    } else {                        // stopped inside the block: redo with `stop` levels
                        S->err = estop; S->redo = stop;
                    }
                }
            }
        }
    }
}

}  // namespace sbd
