// tv.cuh - TV kernels: TVnorm, diffh/diffv, one Chambolle dual sweep, prox output.
//
// Reference semantics restated here:
//   utils/TVnorm.m:1-2, SALSA/diffh.m:1-3, SALSA/diffv.m:1-3 (periodic backward differences)
//   utils/chambolle_prox_TV_stop.m:120-131 (sweep), :149 (output), :152-159 (DivergenceIm,
//   last row/col is -p(end): Q3), :161-166 (GradientIm, forward differences, zero last row/col)
//
// Access pattern ("marching warps"): a warp owns a strip of 32*V consecutive
// fast-axis pixels and walks down `seg` slow-axis rows.  Each array element is
// loaded from HBM exactly once per sweep: the row above stays in registers, the
// lateral neighbours come from warp shuffles (lanes 0 / 31 fetch the one pixel
// outside the strip themselves; those hit L1/L2).
#pragma once
#include "common.cuh"

namespace sbd {

constexpr int TV_WARPS = 4;                 // warps per block
constexpr int TV_THREADS = TV_WARPS * 32;

template <int V>
__device__ __forceinline__ void ld_row(const double* __restrict__ p, bool ok, double (&o)[V]);
template <>
__device__ __forceinline__ void ld_row<1>(const double* __restrict__ p, bool ok, double (&o)[1]) {
    o[0] = ok ? __ldg(p) : 0.0;
}
template <>
__device__ __forceinline__ void ld_row<2>(const double* __restrict__ p, bool ok, double (&o)[2]) {
    if (ok) {
        const double2 t = __ldg(reinterpret_cast<const double2*>(p));
        o[0] = t.x; o[1] = t.y;
    } else {
        o[0] = 0.0; o[1] = 0.0;
    }
}
template <int V>
__device__ __forceinline__ void st_row(double* __restrict__ p, bool ok, const double (&o)[V]);
template <>
__device__ __forceinline__ void st_row<1>(double* __restrict__ p, bool ok, const double (&o)[1]) {
    if (ok) *p = o[0];
}
template <>
__device__ __forceinline__ void st_row<2>(double* __restrict__ p, bool ok, const double (&o)[2]) {
    if (ok) *reinterpret_cast<double2*>(p) = make_double2(o[0], o[1]);
}

struct StripGeom {
    int lane, i0, i, j0, j1, iend;
    bool warp_on, on;       // strip inside the image / this lane's pixels inside (V=2: nx even)
};
template <int V>
__device__ __forceinline__ StripGeom strip_geom(int nx, int ny, int seg) {
    StripGeom s;
    s.lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    s.i0 = (blockIdx.x * TV_WARPS + warp) * 32 * V;
    s.i = s.i0 + s.lane * V;
    s.iend = s.i0 + 32 * V;
    s.j0 = blockIdx.y * seg;
    s.j1 = min(s.j0 + seg, ny);
    s.warp_on = s.i0 < nx;
    s.on = s.i < nx;
    return s;
}

// ---------------------------------------------------------------------------
// TVnorm(x) = sum sqrt(diffh(x)^2 + diffv(x)^2)                 utils/TVnorm.m:2
// out[img*out_stride] receives the total (written by the last block to finish).
// ---------------------------------------------------------------------------
template <int V>
__global__ void __launch_bounds__(TV_THREADS)
k_tvnorm(const double* __restrict__ x, int nx, int ny, int seg, size_t img_stride,
         double* __restrict__ partials, unsigned int* __restrict__ counters,
         double* __restrict__ out, int out_stride) {
    __shared__ double sm[32];
    const StripGeom s = strip_geom<V>(nx, ny, seg);
    const int img = blockIdx.z;
    const double* xi = x + (size_t)img * img_stride;
    double acc[1] = {0.0};
    if (s.warp_on) {
        const int jp = (s.j0 == 0) ? ny - 1 : s.j0 - 1;             // periodic wrap (conv2c.m:27-44)
        const int il = (s.i0 == 0) ? nx - 1 : s.i0 - 1;
        double prev[V];
        ld_row<V>(xi + (size_t)jp * nx + s.i, s.on, prev);
        auto row = [&](const double (&cur)[V], double left0) {
            double left = shfl_up_d(cur[V - 1], 1);
            if (s.lane == 0) left = left0;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const double dv = cur[v] - (v == 0 ? left : cur[v - 1]);    // diffv: x(i,j)-x(i-1,j)
                const double dh = cur[v] - prev[v];                           // diffh: x(i,j)-x(i,j-1)
                if (s.i + v < nx) acc[0] += sqrt(dh * dh + dv * dv);
                prev[v] = cur[v];
            }
        };
        // four rows of loads in flight per warp (the sum keeps its row order)
        constexpr int U = 4;
        int j = s.j0;
        for (; j + U <= s.j1; j += U) {
            double cur[U][V], l0[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                ld_row<V>(xi + (size_t)(j + u) * nx + s.i, s.on, cur[u]);
                l0[u] = (s.lane == 0) ? __ldg(xi + (size_t)(j + u) * nx + il) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) row(cur[u], l0[u]);
        }
        for (; j < s.j1; ++j) {
            double cur[V];
            ld_row<V>(xi + (size_t)j * nx + s.i, s.on, cur);
            row(cur, (s.lane == 0) ? __ldg(xi + (size_t)j * nx + il) : 0.0);
        }
    }
    block_sum<1>(acc, sm);
    const unsigned int nparts = gridDim.x * gridDim.y;
    double* part = partials + (size_t)img * nparts;
    if (threadIdx.x == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = acc[0];
    if (last_block_ticket(counters + img, nparts)) {
        if (threadIdx.x < 32) {
            const double tot = warp_sum_partials(part, (int)nparts, 1);
            if (threadIdx.x == 0) out[(size_t)img * out_stride] = tot;
        }
    }
}

// diffh (axis=1) / diffv (axis=0), SALSA/diffh.m, diffv.m
__global__ void k_diff(const double* __restrict__ x, double* __restrict__ out, int nx, int ny,
                       int axis, size_t total) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t P = (size_t)nx * ny;
    const size_t base = (idx / P) * P;
    const size_t r = idx - base;
    const int i = (int)(r % nx), j = (int)(r / nx);
    size_t nb;
    if (axis == 0) nb = base + (size_t)j * nx + (i == 0 ? nx - 1 : i - 1);
    else           nb = base + (size_t)(j == 0 ? ny - 1 : j - 1) * nx + i;
    out[idx] = x[idx] - x[nb];
}

// ---------------------------------------------------------------------------
// One Chambolle sweep:  p <- (p + tau*grad(u)) / (1 + tau*|grad(u)|),
// u = div(p) - g/lambda, plus err_k (uses the OLD p).       chambolle_prox_TV_stop.m:121-131
// ---------------------------------------------------------------------------
__device__ __forceinline__ double chamb_u(double px, double pxl, double py, double pyp, double gv,
                                          double invlam, int i, int j, int nx, int ny) {
    // DivergenceIm, chambolle_prox_TV_stop.m:152-159
    const double ux = (i == 0) ? px : ((i == nx - 1) ? -px : px - pxl);     // :156-157
    const double uy = (j == 0) ? py : ((j == ny - 1) ? -py : py - pyp);     // :153-154
    return (uy + ux) - gv * invlam;                                          // :159, :124
}

template <int V>
__global__ void __launch_bounds__(TV_THREADS)
k_chamb_sweep(const double* __restrict__ g, const double* __restrict__ pxi,
              const double* __restrict__ pyi, double* __restrict__ pxo, double* __restrict__ pyo,
              int nx, int ny, int seg, size_t img_stride, const Control* __restrict__ ctl,
              ChambState* __restrict__ st, double* __restrict__ partials) {
    __shared__ double sm[32];
    const int img = blockIdx.z;
    if (st[img].done) return;                       // this image already met the stop test (:131)
    const StripGeom s = strip_geom<V>(nx, ny, seg);
    const double lambda = ctl->prox_lambda_run, tau = ctl->tau;
    const double invlam = 1.0 / lambda;
    const size_t off = (size_t)img * img_stride;
    g += off; pxi += off; pyi += off; pxo += off; pyo += off;
    double acc[1] = {0.0};

    if (s.warp_on) {
        const bool haveR = s.iend < nx;             // a pixel right of the strip exists
        const bool laneR = haveR && s.lane == 31;
        const bool laneL = (s.i0 > 0) && s.lane == 0;
        const int ix = laneL ? s.i0 - 1 : s.iend;   // the one outside pixel lanes 0 / 31 fetch
        double pxc[V], pyc[V], uc[V], gv[V], pyp[V];
        double uEc = 0.0, pyEc = 0.0;

        // ---- prologue: u on row j0
        {
            const int j = s.j0;
            const size_t r = (size_t)j * nx;
            if (j > 0) ld_row<V>(pyi + r - nx + s.i, s.on, pyp);
            else {
#pragma unroll
                for (int v = 0; v < V; ++v) pyp[v] = 0.0;
            }
            ld_row<V>(pxi + r + s.i, s.on, pxc);
            ld_row<V>(pyi + r + s.i, s.on, pyc);
            ld_row<V>(g + r + s.i, s.on, gv);
            const double pxX = (laneL || laneR) ? __ldg(pxi + r + ix) : 0.0;
            double pxl = shfl_up_d(pxc[V - 1], 1);
            if (s.lane == 0) pxl = pxX;
#pragma unroll
            for (int v = 0; v < V; ++v)
                uc[v] = chamb_u(pxc[v], v == 0 ? pxl : pxc[v - 1], pyc[v], pyp[v], gv[v], invlam,
                                s.i + v, j, nx, ny);
            if (laneR) {
                pyEc = __ldg(pyi + r + s.iend);
                const double pyEp = (j > 0) ? __ldg(pyi + r - nx + s.iend) : 0.0;
                const double gE = __ldg(g + r + s.iend);
                uEc = chamb_u(pxX, pxc[V - 1], pyEc, pyEp, gE, invlam, s.iend, j, nx, ny);
            }
        }

        for (int j = s.j0; j < s.j1; ++j) {
            double pxn[V], pyn[V], un[V];
            double uEn = 0.0, pyEn = 0.0;
            const bool more = (j + 1 < ny);
            if (more) {
                const size_t r = (size_t)(j + 1) * nx;
                ld_row<V>(pxi + r + s.i, s.on, pxn);
                ld_row<V>(pyi + r + s.i, s.on, pyn);
                ld_row<V>(g + r + s.i, s.on, gv);
                const double pxX = (laneL || laneR) ? __ldg(pxi + r + ix) : 0.0;
                double pxl = shfl_up_d(pxn[V - 1], 1);
                if (s.lane == 0) pxl = pxX;
#pragma unroll
                for (int v = 0; v < V; ++v)
                    un[v] = chamb_u(pxn[v], v == 0 ? pxl : pxn[v - 1], pyn[v], pyc[v], gv[v], invlam,
                                    s.i + v, j + 1, nx, ny);
                if (laneR) {
                    pyEn = __ldg(pyi + r + s.iend);
                    const double gE = __ldg(g + r + s.iend);
                    uEn = chamb_u(pxX, pxn[V - 1], pyEn, pyEc, gE, invlam, s.iend, j + 1, nx, ny);
                }
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) { pxn[v] = 0.0; pyn[v] = 0.0; un[v] = 0.0; }
            }
            double uright = shfl_down_d(uc[0], 1);
            if (s.lane == 31) uright = uEc;
            double po[V], qo[V];
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int iv = s.i + v;
                const double ur = (v < V - 1) ? uc[(v + 1) % V] : uright;
                const double upx = (iv < nx - 1) ? ur - uc[v] : 0.0;        // GradientIm :162-163
                const double upy = more ? un[v] - uc[v] : 0.0;              // :165-166
                const double tmp = sqrt(upx * upx + upy * upy);             // :127
                const double ex = -upx + tmp * pxc[v], ey = -upy + tmp * pyc[v];
                if (iv < nx) acc[0] += ex * ex + ey * ey;                   // :128
                const double rinv = 1.0 / (1.0 + tau * tmp);
                po[v] = (pxc[v] + tau * upx) * rinv;                        // :129
                qo[v] = (pyc[v] + tau * upy) * rinv;                        // :130
            }
            st_row<V>(pxo + (size_t)j * nx + s.i, s.on, po);
            st_row<V>(pyo + (size_t)j * nx + s.i, s.on, qo);
#pragma unroll
            for (int v = 0; v < V; ++v) { pxc[v] = pxn[v]; pyc[v] = pyn[v]; uc[v] = un[v]; }
            uEc = uEn; pyEc = pyEn;
        }
    }

    block_sum<1>(acc, sm);
    const unsigned int nparts = gridDim.x * gridDim.y;
    double* part = partials + (size_t)img * nparts;
    if (threadIdx.x == 0) part[blockIdx.y * gridDim.x + blockIdx.x] = acc[0];
    if (last_block_ticket(&st[img].counter, nparts)) {
        if (threadIdx.x < 32) {
            const double tot = warp_sum_partials(part, (int)nparts, 1);
            if (threadIdx.x == 0) {
                const double err = sqrt(tot);                               // :128  (...)^0.5
                const int k = st[img].k + 1;                                // :121
                st[img].k = k;
                st[img].buf ^= 1;
                st[img].err = err;
                st[img].done = !((k < ctl->maxiter) && (err > ctl->tol));   // :131
            }
        }
    }
}

// f = g - lambda * DivergenceIm(px, py)                  chambolle_prox_TV_stop.m:149
// The dual pair lives in buffer st.buf of the ping-pong pair.
template <int V>
__global__ void __launch_bounds__(TV_THREADS)
k_chamb_out(const double* __restrict__ g, const double* __restrict__ px0,
            const double* __restrict__ py0, const double* __restrict__ px1,
            const double* __restrict__ py1, double* __restrict__ f, int nx, int ny, int seg,
            size_t img_stride, const Control* __restrict__ ctl, const ChambState* __restrict__ st) {
    const int img = blockIdx.z;
    if (st[img].emitted) return;                    // the last fused block wrote f already (tv_multi.cuh)
    const StripGeom s = strip_geom<V>(nx, ny, seg);
    if (!s.warp_on) return;
    const double lambda = ctl->prox_lambda_run;
    const bool odd = st[img].buf != 0;
    const size_t off = (size_t)img * img_stride;
    const double* px = (odd ? px1 : px0) + off;
    const double* py = (odd ? py1 : py0) + off;
    g += off; f += off;
    const bool laneL = (s.i0 > 0) && s.lane == 0;
    double pyp[V], pxc[V], pyc[V], gv[V], o[V];
    if (s.j0 > 0) ld_row<V>(py + (size_t)(s.j0 - 1) * nx + s.i, s.on, pyp);
    else {
#pragma unroll
        for (int v = 0; v < V; ++v) pyp[v] = 0.0;
    }
    for (int j = s.j0; j < s.j1; ++j) {
        const size_t r = (size_t)j * nx;
        ld_row<V>(px + r + s.i, s.on, pxc);
        ld_row<V>(py + r + s.i, s.on, pyc);
        ld_row<V>(g + r + s.i, s.on, gv);
        double pxl = shfl_up_d(pxc[V - 1], 1);
        if (s.lane == 0) pxl = laneL ? __ldg(px + r + s.i0 - 1) : 0.0;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int iv = s.i + v;
            const double pl = (v == 0) ? pxl : pxc[(v + V - 1) % V];
            const double ux = (iv == 0) ? pxc[v] : ((iv == nx - 1) ? -pxc[v] : pxc[v] - pl);
            const double uy = (j == 0) ? pyc[v] : ((j == ny - 1) ? -pyc[v] : pyc[v] - pyp[v]);
            o[v] = gv[v] - lambda * (uy + ux);
            pyp[v] = pyc[v];
        }
        st_row<V>(f + r + s.i, s.on, o);
    }
}

// start of a prox: clears the per-image state and takes the snapshot of lambda*theta the sweeps work from (the
// scalar update of the same iteration may change ctl->prox_lambda_theta while the sweeps are still running)
__global__ void k_chamb_reset(ChambState* st, int n, Control* ctl, unsigned int* bar) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) ctl->prox_lambda_run = ctl->prox_lambda_theta;
    if (i < n) bar[i] = 0u;                         // barrier counters of the cooperative kernel (tv_coop.cuh)
    if (i < n) { st[i].k = 0; st[i].done = 0; st[i].err = 0.0; st[i].counter = 0u; st[i].redo = 0; st[i].buf = 0; st[i].emitted = 0; }
}

// end of a main-loop prox: sweeps executed by chain 0 -> trace slot (SAPG `chambolle_iters`)
__global__ void k_chamb_record(const ChambState* st, Control* ctl, int* trace, int n) {
    const int k = ctl->prox_count;
    if (k >= 0 && k < n) trace[k] = st[0].k;
    ctl->prox_count = k + 1;
}

}  // namespace sbd
