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
// launch, a no-op otherwise), so the sweep count k and the dual pair are the
// reference's.  Arithmetic tolerance: sqrt and the reciprocal come from MUFU
// seeds plus one third-order step (<= 1 ulp each, cm_core) and FMA contraction
// is on, so p differs from an IEEE sqrt/divide evaluation (k_chamb_sweep, the
// T = 1 fallback) by ~1 ulp per sweep; the two paths are not bitwise equal and
// an err_k within ~1e-15 relative of tol may decide the stop test differently
// (tests/test_gpu_chambolle_prod.py::test_fused_vs_single_sweep).
//
// ERRSUB (sampled stop test): err_k is only ever COMPARED with tol, and err_k^2 is a sum of non-negative terms, so
// the sum over any SUBSET of the pixels is a lower bound: "subset sum > tol^2" proves err_k > tol and the sweep
// continues - exactly the reference's decision.  The steady-state loop then forms no err term at all (340 instead of
// 404 fp64 instructions per trip); the subset is the rows of the generic path at the two ends of every segment.  The
// dual pair does not depend on err, so p and f are bit-identical.  When the subset does NOT prove it (subset sum <= tol^2: the test is
// about to fire, or the image is tiny), the block is recomputed with the full sum by the "exact" launch (redo == 2,
// a no-op otherwise), which decides as before.  Used where the caller does not ask for the value of err (the SAPG
// loop, sbd_tvprox_dev with err == NULL); the entry points that return err always run the full sum.
//
// EMIT: the block that is planned to be the last one also forms the prox output
// f = g - lambda*div p (chambolle_prox_TV_stop.m:134) from the rows its top level
// emits, instead of a separate pass over (g, px, py).  div p needs p one pixel to
// the left and one row above the emitted row, so this needs one pixel of lateral
// validity to spare (HL > nlev: blocks with an odd number of levels have it) and the march starts
// one row earlier.  EMIT = 2 additionally skips the store of the dual pair (the
// SAPG loop starts every prox from zero and never reads it back).
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
// sqrt and reciprocal of a level step (cm_core), ~1 ulp each, from MUFU seeds and one third-order step:
//   y = (1+dl)/sqrt(a) (seed), g0 = a*y, r = 1 - g0*y = -(2 dl + dl^2), 1/(1+dl) = 1 + r/2 + 3/8 r^2 + O(r^3)
//     => sqrt(a) = g0 + g0*(r*(0.5 + 0.375 r)),   |rel err| ~ 0.3 |r|^3 < 2^-60   (a == 0 -> exactly 0)
//   d = 1 + tau*sqrt(a), rs = seed(1/d), e = 1 - d*rs  =>  1/d = rs + rs*(e + e^2) + O(e^3)
// (an IEEE sqrt plus an IEEE division cost ~3x the fp64 instructions)

struct CmPk { double px[2], py[2], g[2]; };             // a row travelling up the levels
struct CmLv { double px[2], py[2], u[2], g[2]; };       // the row a level holds

// sqrt seed of a >= 0 with a == 0 mapped to a tiny normal (a * y is then exactly 0): MUFU.RSQ64H only
// reads the high word, so the clamp is one integer instruction instead of an fp64 add
__device__ __forceinline__ double fast_rsqrt_seed_nz(double a) {
    const int hi = max(__double2hiint(a), 0x00100000);
    return fast_rsqrt_seed(__hiloint2double(hi, __double2loint(a)));
}
// The MUFU seeds only define the high word (the PTX instruction zeroes the low one with an extra move).
// A seed does not care about its low word, so it is taken from a value that is dead anyway; ptxas can
// then write the MUFU result next to it and drop the move.
__device__ __forceinline__ double seed_with_low_of(double seed, double dead) {
    return __hiloint2double(__double2hiint(seed), __double2loint(dead));
}

// The arithmetic of one level step for the two pixels of a lane, given upx (:162-163) and the new u
// (:159).  Written stage by stage (all of stage k for both pixels before stage k+1): ptxas keeps roughly
// this order, and back-to-back dependent fp64 instructions of ONE chain each wait out the pipe latency.
// On B200 an fp64 instruction holds the SMSP's issue port for two cycles and every other instruction
// for one (tools/fp64_microbench.cu, MIX lines), so the kernel is issue-bound and every instruction
// counts: the square root / reciprocal refinement is written with the fewest operations that reach
// ~1 ulp (derivation above).
// Variants of the level-step arithmetic, A/B-measured with tools/proto/chamb_bench.cu on B200 (4096^2 x 8 chains,
// 128-row segments, ms per 4-level launch; gpurun_out/r02_chamb_ab*.log):
//   baseline (integer clamp in front of the rsqrt seed, seed of 1/d from the refined d)        1.532
//   SBD_CM_BIAS   1e-300 folded into the first square instead of the clamp (-16 integer ops)   1.445   <- default
//   SBD_CM_REUSE  p - (tau/d)(|grad u| p - grad u), reusing the err terms (-16 fp64 ops)       1.657 (1.636 with BIAS)
//   SBD_CM_EARLYD reciprocal seed from the unrefined root (+16 fp64, shorter dependency chain) 1.489 (with BIAS)
//   SBD_CM_OPX0   numerators multiplied by the seed beside the residual (+16 fp64)             1.510 (with BIAS)
// i.e. the loop follows its instruction count (EARLYD / OPX0 shorten the chain and lose), except where ptxas
// answers a shorter source with more register moves (REUSE: 108 -> 127 MOVs per trip).
#ifndef SBD_CM_BIAS
#define SBD_CM_BIAS 1
#endif
// pixels per lane whose err terms the two rows of a trip accumulate under ERRSUB (2 / 1 / 0).  Measured with
// tools/proto/chamb_bench.cu (ms per 4-level launch, 4096^2 x 8, 128-row segments; full sums: 1.446):
//   (1,0) 1.56-1.62   (1,1) 1.55   (2,1) 1.41   (2,0) 1.361   (0,2) 1.365   (0,0) 1.35  <- default
// (one pixel of a lane's pair breaks the stage-by-stage order of the two pixel chains and is SLOWER than the full
// sums; dropping a whole row's terms gains).  With (0,0) the steady-state loop forms no err term at all and the
// lower bound is the sum over the rows of the generic path at the segment ends (full terms, ~14 rows per segment):
// at 4096^2 x 32 chains and 512-row segments 5.65 -> 5.04 ms.
#ifndef SBD_CM_KREG
#define SBD_CM_KREG 1
#endif
#ifndef SBD_CM_EMA
#define SBD_CM_EMA 0
#endif
#ifndef SBD_CM_EMB
#define SBD_CM_EMB 0
#endif
#ifndef SBD_CM_REUSE
#define SBD_CM_REUSE 0
#endif
#ifndef SBD_CM_EARLYD
#define SBD_CM_EARLYD 0
#endif
#ifndef SBD_CM_OPX0
#define SBD_CM_OPX0 0
#endif
// The two fp64 constants of the level step that have no short immediate form.  ptxas rebuilds a literal with two MOVs at
// every use inside the loop (16 MOVs per trip); read once through an opaque asm they stay in registers (SBD_CM_KREG).
struct CmK { double c375, tiny; };
__device__ __forceinline__ CmK cm_consts() {
    CmK k;
#if SBD_CM_KREG
    asm volatile("mov.f64 %0, 0d3FD8000000000000;" : "=d"(k.c375));
    asm volatile("mov.f64 %0, 0d01A56E1FC2F8F359;" : "=d"(k.tiny));
#else
    k.c375 = 0.375; k.tiny = 1e-300;
#endif
    return k;
}
__device__ __forceinline__ CmK cm_consts_plain() { CmK k; k.c375 = 0.375; k.tiny = 1e-300; return k; }
// ERRMODE: which pixels' err terms are formed - 2: both, 1: pixel 0 only, 0: none (ex / ey are then left untouched)
template <int ERRMODE = 2, class Lv>
__device__ __forceinline__ void cm_core(const double (&upx)[2], const double (&un)[2], const Lv& h, double tau,
                                        double (&opx)[2], double (&opy)[2], double (&ex)[2], double (&ey)[2],
                                        const CmK& K) {
    double upy[2], s2[2], y[2], g[2], rs[2], r[2], t[2], d[2], ee[2];
    const double mtau = -tau;
#define CM_V _Pragma("unroll") for (int v = 0; v < 2; ++v)
    CM_V upy[v] = un[v] - h.u[v];
    double sq[2];
    // + 1e-300: |grad u|^2 = 0 (flat areas) stays a normal number for the MUFU seed (a * y is then 1e-150 instead of
    // exactly 0, which changes nothing that is representable: d = 1, p unchanged, err += 1e-300 |p|^2); for any
    // other value the bias is below half an ulp.  It replaces an integer clamp of the seed's high word.
#if SBD_CM_BIAS
    CM_V sq[v] = fma(upy[v], upy[v], K.tiny);
    CM_V s2[v] = fma(upx[v], upx[v], sq[v]);
    CM_V y[v] = seed_with_low_of(fast_rsqrt_seed(s2[v]), sq[v]);
#else
    CM_V sq[v] = upy[v] * upy[v];
    CM_V s2[v] = fma(upx[v], upx[v], sq[v]);
    CM_V y[v] = seed_with_low_of(fast_rsqrt_seed_nz(s2[v]), sq[v]);
#endif
    CM_V g[v] = s2[v] * y[v];
#if SBD_CM_EARLYD
    // reciprocal seed from the UNREFINED root (relative error 2^-23, the same as the seed's own): the MUFU latency
    // runs under the refinement of the root instead of behind it; the residual below uses the refined d
    double d0[2];
    CM_V d0[v] = fma(tau, g[v], 1.0);
    CM_V rs[v] = seed_with_low_of(fast_rcp_seed(d0[v]), sq[v]);
#endif
    CM_V r[v] = fma(-g[v], y[v], 1.0);
    CM_V t[v] = fma(r[v], K.c375, 0.5);
    CM_V t[v] = r[v] * t[v];                                                 // r + 1.5 r^2
    CM_V g[v] = fma(g[v], t[v], g[v]);                                       // :127
    CM_V d[v] = fma(tau, g[v], 1.0);
#if !SBD_CM_EARLYD
    CM_V rs[v] = seed_with_low_of(fast_rcp_seed(d[v]), t[v]);                                        // seed from the final d: no early estimate to compute
#endif
#pragma unroll
    for (int v = 0; v < ERRMODE; ++v) {
        ex[v] = fma(g[v], h.px[v], -upx[v]);                                 // :128
        ey[v] = fma(g[v], h.py[v], -upy[v]);
    }
    CM_V ee[v] = fma(-d[v], rs[v], 1.0);
#if SBD_CM_REUSE
    static_assert(ERRMODE == 2, "SBD_CM_REUSE needs the err terms of both pixels");
    CM_V ee[v] = fma(ee[v], ee[v], ee[v]);
    CM_V rs[v] = fma(rs[v], ee[v], rs[v]);                                   // 1 / (1 + tau |grad u|)
    // :129-130  (p + tau grad u) / d  =  p - (tau / d) (|grad u| p - grad u): the err terms are reused, one fma each
    CM_V rs[v] = rs[v] * mtau;
    CM_V opx[v] = fma(rs[v], ex[v], h.px[v]);
    CM_V opy[v] = fma(rs[v], ey[v], h.py[v]);
#elif SBD_CM_OPX0
    // (p + tau grad u) * rs0 * (1 + e + e^2): the product with the seed runs beside the residual, one fma closes
    CM_V opx[v] = fma(tau, upx[v], h.px[v]);
    CM_V opy[v] = fma(tau, upy[v], h.py[v]);
    CM_V opx[v] = opx[v] * rs[v];
    CM_V opy[v] = opy[v] * rs[v];
    CM_V ee[v] = fma(ee[v], ee[v], ee[v]);
    CM_V opx[v] = fma(opx[v], ee[v], opx[v]);                                // :129
    CM_V opy[v] = fma(opy[v], ee[v], opy[v]);                                // :130
    (void)mtau;
#else
    CM_V opx[v] = fma(tau, upx[v], h.px[v]);
    CM_V opy[v] = fma(tau, upy[v], h.py[v]);
    CM_V ee[v] = fma(ee[v], ee[v], ee[v]);
    CM_V rs[v] = fma(rs[v], ee[v], rs[v]);
    CM_V opx[v] = opx[v] * rs[v];                                            // :129
    CM_V opy[v] = opy[v] * rs[v];                                            // :130
    (void)mtau;
#endif
#undef CM_V
}


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
    CmPk o;
    double ex[2], ey[2];
    cm_core(upx, un, h, tau, o.px, o.py, ex, ey, cm_consts_plain());
#pragma unroll
    for (int v = 0; v < 2; ++v) o.g[v] = h.g[v];
    const double e = fma(ex[1], ex[1], fma(ey[1], ey[1], fma(ex[0], ex[0], ey[0] * ey[0])));      // :128
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

// Fast-path level step without register rotation (interior rows, all levels live).  `ho` is the
// level's old state; `hn` its new state, whose px, py, g were already written by the level below (or by
// the load) and whose u is formed here; the row the level emits goes straight into the new state of
// the level above (`up`).  A trip of two rows alternates two state sets, so no value is ever copied
// (the rotating form spent one instruction in six on register moves).  err is not masked here.
template <bool EDGE, int ERRMODE = 2>
__device__ __forceinline__ void cm_step2(const CmLv& ho, CmLv& hn, CmLv& up, const CmLane& L, double tau, double& err,
                                         const CmK& K) {
    const double pxl = shfl_up_d(hn.px[1], 1);
    double ux0 = hn.px[0] - pxl, ux1 = hn.px[1] - hn.px[0];                 // :156-157
    if (EDGE) {
        if (L.last0) ux0 = -hn.px[0];
        if (L.last1) ux1 = -hn.px[1];
    }
    const double uy0 = hn.py[0] - ho.py[0], uy1 = hn.py[1] - ho.py[1];     // :153-154
    hn.u[0] = (uy0 + ux0) - hn.g[0];                                        // :159, :124
    hn.u[1] = (uy1 + ux1) - hn.g[1];
    const double ur = shfl_down_d(ho.u[0], 1);
    double upx[2] = {ho.u[1] - ho.u[0], ur - ho.u[1]};                      // :162-163
    if (EDGE) {
        if (L.last0) upx[0] = 0.0;
        if (L.last1) upx[1] = 0.0;
    }
    double ex[2], ey[2];
    cm_core<ERRMODE>(upx, hn.u, ho, tau, up.px, up.py, ex, ey, K);
    up.g[0] = ho.g[0]; up.g[1] = ho.g[1];
    if (EDGE) {
        if (!L.in0) { up.px[0] = 0.0; up.py[0] = 0.0; }
        if (!L.in1) { up.px[1] = 0.0; up.py[1] = 0.0; }
    }
    // :128, the four squares go straight into the level's accumulator (one fp64 op per square)
    if (ERRMODE == 2) err = fma(ex[1], ex[1], fma(ey[1], ey[1], fma(ex[0], ex[0], fma(ey[0], ey[0], err))));
    else if (ERRMODE == 1) err = fma(ex[0], ex[0], fma(ey[0], ey[0], err));
}

// ZERO: the incoming dual pair is identically zero (chambolle_prox_TV_stop.m:68-69) and is not loaded
template <bool EDGE, bool ZERO>
__device__ __forceinline__ void cm_load(CmPk& p, const double* __restrict__ g, const double* __restrict__ px,
                                        const double* __restrict__ py, size_t off, const CmLane& L) {
    // raw values only: nothing here may depend on the loaded data, or the prefetch would stall
    // 64-bit loads on purpose: a 128-bit load needs an aligned register quad, and the copies the
    // register allocator then inserts right behind the load stall on it (ncu: long_scoreboard on a MOV)
    if (ZERO) {
        p.px[0] = 0.0; p.px[1] = 0.0; p.py[0] = 0.0; p.py[1] = 0.0;
        p.g[0] = (!EDGE || L.in0) ? __ldg(g + off) : 0.0;
        p.g[1] = (!EDGE || L.in1) ? __ldg(g + off + 1) : 0.0;
    } else {
        p.px[0] = (!EDGE || L.in0) ? __ldg(px + off) : 0.0;   p.px[1] = (!EDGE || L.in1) ? __ldg(px + off + 1) : 0.0;
        p.py[0] = (!EDGE || L.in0) ? __ldg(py + off) : 0.0;   p.py[1] = (!EDGE || L.in1) ? __ldg(py + off + 1) : 0.0;
        p.g[0] = (!EDGE || L.in0) ? __ldg(g + off) : 0.0;     p.g[1] = (!EDGE || L.in1) ? __ldg(g + off + 1) : 0.0;
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

// prox output of an emitted (final) row; prevpy = py of the row above (0 above the image)
template <bool EDGE, int EMIT>
__device__ __forceinline__ void cm_emit(const CmPk& p, double (&prevpy)[2], double* __restrict__ f,
                                        double* __restrict__ pxo, double* __restrict__ pyo, size_t off,
                                        const CmLane& L, double lambda, bool inseg, bool lastrow) {
    const double pxl = shfl_up_d(p.px[1], 1);
    double ux0 = p.px[0] - pxl, ux1 = p.px[1] - p.px[0];
    if (EDGE) {
        if (L.last0) ux0 = -p.px[0];
        if (L.last1) ux1 = -p.px[1];
    }
    double uy0 = p.py[0] - prevpy[0], uy1 = p.py[1] - prevpy[1];
    if (lastrow) { uy0 = -p.py[0]; uy1 = -p.py[1]; }
    prevpy[0] = p.py[0]; prevpy[1] = p.py[1];
    if (!inseg || !L.central) return;
    // p.g carries g / lambda
    const double f0 = fma(-lambda, uy0 + ux0, p.g[0] * lambda), f1 = fma(-lambda, uy1 + ux1, p.g[1] * lambda);
    if (!EDGE) {
        *reinterpret_cast<double2*>(f + off) = make_double2(f0, f1);
        if (EMIT == 1) {
            *reinterpret_cast<double2*>(pxo + off) = make_double2(p.px[0], p.px[1]);
            *reinterpret_cast<double2*>(pyo + off) = make_double2(p.py[0], p.py[1]);
        }
    } else {
        if (L.in0) { f[off] = f0; if (EMIT == 1) { pxo[off] = p.px[0]; pyo[off] = p.py[0]; } }
        if (L.in1) { f[off + 1] = f1; if (EMIT == 1) { pxo[off + 1] = p.px[1]; pyo[off + 1] = p.py[1]; } }
    }
}

// Schedule.  PIPE = false: in iteration r level s receives row r - s, which level
// s-1 emitted earlier in the same iteration (levels run bottom-up, one dependent
// chain).  PIPE = true: the levels are software-pipelined - level s consumes the
// row level s-1 emitted in the PREVIOUS iteration (row r - 2s), the levels run
// top-down and the T level steps of one iteration are independent (more ILP,
// more registers; measured slower on B200 because ptxas keeps the level steps sequential anyway and
// occupancy drops - only PIPE = false is instantiated).  inbox[s] is the row waiting for level s.
__device__ __forceinline__ void cm_take(CmPk& dst, const CmPk& raw, double invlam) {
    dst.px[0] = raw.px[0]; dst.px[1] = raw.px[1]; dst.py[0] = raw.py[0]; dst.py[1] = raw.py[1];
    dst.g[0] = raw.g[0] * invlam; dst.g[1] = raw.g[1] * invlam;           // g / lambda (:124)
}

// One generic iteration of the march: honours all row flags, any nlev <= T.
template <int T, bool EDGE, bool PIPE, bool ZERO, int EMIT>
__device__ __forceinline__ void cm_generic_iter(int r, CmLv (&h)[T], CmPk (&inbox)[T], CmPk& nxt, double (&err)[T],
                                                const double* __restrict__ g, const double* __restrict__ pxi,
                                                const double* __restrict__ pyi, double* __restrict__ pxo,
                                                double* __restrict__ pyo, int nx, int ny, int j0, int jlast,
                                                int r0, long long ibase, const CmLane& L,
                                                double invlam, double tau, int nlev,
                                                double* __restrict__ f, double lambda, double (&prevpy)[2]) {
    constexpr int D = PIPE ? 2 : 1;
    CmPk loc[T];                                    // PIPE = false: rows only travel within the iteration
#define CM_BOX(s_) (PIPE ? inbox[s_] : loc[s_])
#pragma unroll
    for (int q = 0; q < T; ++q) {
#pragma unroll
        for (int v = 0; v < 2; ++v) { loc[q].px[v] = 0.0; loc[q].py[v] = 0.0; loc[q].g[v] = 0.0; }
    }
    cm_take(CM_BOX(0), nxt, invlam);
    if (r + 1 <= min(jlast + nlev, ny - 1))
        cm_load<EDGE, ZERO>(nxt, g, pxi, pyi, (size_t)((long long)(r + 1) * nx + ibase), L);
#pragma unroll
    for (int q = 0; q < T; ++q) {
        const int s = PIPE ? T - 1 - q : q;
        const int jj = r - D * s;                   // row this level receives
        const int need = jlast + nlev - s;          // last row this level has to receive
        if (s < nlev && jj >= r0 && jj <= min(need, ny)) {
            const bool virt = (jj == ny), last = (jj == ny - 1), first = (jj == r0);
            const int hr = jj - 1;                  // row being updated by this level
            double e = 0.0;
            CmPk p = CM_BOX(s);
            // first row of a level: nothing held yet -> only u of that row is formed (with the
            // "row above" = 0, exact at the image top); the emitted row is meaningless
            cm_step<EDGE, true>(h[s], p, L, tau, e, virt, last, true);
            if (!first) {
                const bool inseg = hr >= j0 && hr <= jlast;
                if (inseg) err[s] += e;
                if (s + 1 == nlev) {
                    if (EMIT) cm_emit<EDGE, EMIT>(p, prevpy, f, pxo, pyo, (size_t)((long long)hr * nx + ibase), L, lambda, inseg, hr == ny - 1);
                    else if (inseg) cm_store<EDGE>(p, pxo, pyo, (size_t)((long long)hr * nx + ibase), L);
                } else if (s + 1 < T) {
                    CM_BOX(s + 1) = p;
                }
            }
        }
    }
}

// L2 prefetch of the rows the march will load CM_PF rows from now.  ptxas schedules the first use of a
// loaded register a fixed ~180 instructions behind the load whatever the source order says, which
// covers an L2 hit but not the ~2000 cycles a DRAM access takes here (ncu: 24 % of the samples were
// long_scoreboard on those two uses); the prefetch turns the loads into L2 hits without registers.
constexpr int CM_PF = 3;            // distances 1..8 measure the same, 16 is slower
template <bool EDGE, bool ZERO>
__device__ __forceinline__ void cm_prefetch(const double* __restrict__ g, const double* __restrict__ px,
                                            const double* __restrict__ py, size_t off, const CmLane& L) {
    if (EDGE && !L.in0) return;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(g + off));
    if (!ZERO) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(px + off));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(py + off));
    }
}

// The march of one warp.  nlev <= T levels are applied.
template <int T, bool EDGE, bool PIPE, bool ZERO, int EMIT, bool ERRSUB = false>
__device__ __forceinline__ void cm_march(const double* __restrict__ g, const double* __restrict__ pxi,
                                         const double* __restrict__ pyi, double* __restrict__ pxo,
                                         double* __restrict__ pyo, int nx, int ny, int j0, int j1,
                                         const CmLane& L, double invlam, double tau, int nlev,
                                         double (&err)[T], double* __restrict__ f, double lambda) {
    constexpr int D = PIPE ? 2 : 1;
    CmLv h[T];
    CmPk inbox[T];
#pragma unroll
    for (int s = 0; s < T; ++s) {
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            h[s].px[v] = 0.0; h[s].py[v] = 0.0; h[s].u[v] = 0.0; h[s].g[v] = 0.0;
            inbox[s].px[v] = 0.0; inbox[s].py[v] = 0.0; inbox[s].g[v] = 0.0;
        }
    }
    const int r0 = max(j0 - nlev - (EMIT ? 1 : 0), 0);     // EMIT: the final row above the segment is needed too
    double prevpy[2] = {0.0, 0.0};
    const int jlast = min(j1 - 1, ny - 1);          // last output row of this segment
    const int rend = jlast + D * (nlev - 1) + 1;    // last iteration (level nlev-1 receives row jlast+1)
    const long long ibase = L.i;
    CmPk nxt;
    cm_load<EDGE, ZERO>(nxt, g, pxi, pyi, (size_t)((long long)r0 * nx + ibase), L);

    int r = r0;
    if (nlev == T) {
        // steady state: every level live, interior rows only, every updated row inside the segment
        const int fast_lo = j0 + D * (T - 1) + 1, fast_hi = min(j1, ny - 2);
        for (; r < min(fast_lo, rend + 1); ++r)
            cm_generic_iter<T, EDGE, PIPE, ZERO, EMIT>(r, h, inbox, nxt, err, g, pxi, pyi, pxo, pyo, nx, ny, j0, jlast, r0, ibase, L, invlam, tau, nlev, f, lambda, prevpy);
        // Two rows per trip over two level-state sets: the first row of a trip reads `h` and builds `hb`,
        // the second reads `hb` and rebuilds `h` (cm_step2), so the loop carries no register rotation.
        // The next row is loaded straight into the level-0 state that has just been consumed (three
        // quarters of a row ahead of its use; the L2 prefetch runs CM_PF rows ahead of that), and the
        // load / prefetch / store addresses are running pointers.
        if (!PIPE && r <= fast_hi) {
            CmLv hb[T];
#pragma unroll
            for (int v = 0; v < 2; ++v) { hb[0].px[v] = nxt.px[v]; hb[0].py[v] = nxt.py[v]; hb[0].g[v] = nxt.g[v]; }   // raw row r
            const long long o1 = (long long)(r + 1) * nx + ibase, os = (long long)(r - T) * nx + ibase;
            const double *gl = g + o1, *pxl = pxi + o1, *pyl = pyi + o1;        // next row to load
            double *pxs = pxo + os, *pys = pyo + os, *fs = EMIT ? f + os : nullptr;   // next row to store
            const size_t pfo = (size_t)CM_PF * nx;
            const CmK KK = cm_consts();
            auto fast_row = [&](int rr, CmLv (&ho)[T], CmLv (&hn)[T], auto emc) {
                constexpr int EM = decltype(emc)::value;                        // err terms of this row: 2 / 1 / 0 pixels
                hn[0].g[0] *= invlam; hn[0].g[1] *= invlam;                     // g / lambda (:124)
                CmLv top;
                cm_step2<EDGE, EM>(ho[0], hn[0], T > 1 ? hn[T > 1 ? 1 : 0] : top, L, tau, err[0], KK);
                // ho[0] is dead from here on: row rr + 1 lands in it
                if (rr + 1 + CM_PF < ny) cm_prefetch<EDGE, ZERO>(gl, pxl, pyl, pfo, L);
                {
                    CmPk ld;
                    cm_load<EDGE, ZERO>(ld, gl, pxl, pyl, 0, L);
#pragma unroll
                    for (int v = 0; v < 2; ++v) { ho[0].px[v] = ld.px[v]; ho[0].py[v] = ld.py[v]; ho[0].g[v] = ld.g[v]; }
                }
                gl += nx; pxl += nx; pyl += nx;
#pragma unroll
                for (int s = 1; s < T; ++s) {
                    if (s + 1 < T) cm_step2<EDGE, EM>(ho[s], hn[s], hn[s + 1 < T ? s + 1 : s], L, tau, err[s], KK);
                    else cm_step2<EDGE, EM>(ho[s], hn[s], top, L, tau, err[s], KK);
                }
                CmPk p;
#pragma unroll
                for (int v = 0; v < 2; ++v) { p.px[v] = top.px[v]; p.py[v] = top.py[v]; p.g[v] = top.g[v]; }
                if (EMIT) cm_emit<EDGE, EMIT>(p, prevpy, fs, pxs, pys, 0, L, lambda, true, false);
                else cm_store<EDGE>(p, pxs, pys, 0, L);
                pxs += nx; pys += nx;
                if (EMIT) fs += nx;
            };
            // (four rows per trip would halve the register moves ptxas leaves at the back edge, but the loop
            // then outgrows the instruction cache close to the SMSP: ncu shows no_instruction stalls and no
            // net gain)
            // ERRSUB: SBD_CM_EMA / SBD_CM_EMB pixels of the two rows (any subset is a valid lower bound of err_k^2)
            constexpr int EMA = ERRSUB ? SBD_CM_EMA : 2, EMB = ERRSUB ? SBD_CM_EMB : 2;
            for (; r + 1 <= fast_hi; r += 2) {
                fast_row(r, h, hb, std::integral_constant<int, EMA>{});
                fast_row(r + 1, hb, h, std::integral_constant<int, EMB>{});
            }
#pragma unroll
            for (int v = 0; v < 2; ++v) { nxt.px[v] = hb[0].px[v]; nxt.py[v] = hb[0].py[v]; nxt.g[v] = hb[0].g[v]; }   // raw row r
            // the fast rows add their err terms unmasked: lanes outside the output region hold 0
            if (!L.central) {
#pragma unroll
                for (int s = 0; s < T; ++s) err[s] = 0.0;
            }
        }
    }
    for (; r <= rend; ++r)
        cm_generic_iter<T, EDGE, PIPE, ZERO, EMIT>(r, h, inbox, nxt, err, g, pxi, pyi, pxo, pyo, nx, ny, j0, jlast, r0, ibase, L, invlam, tau, nlev, f, lambda, prevpy);
}

// grid = (ceil(nstrips / TV_WARPS), nsegs, batch); block = TV_THREADS.
// redo == 0: main launch of a block of T sweeps; redo == 1: re-run with the
// number of levels the stop test asked for (no-op unless st.redo > 0);
// redo == 2: "exact" launch behind an ERRSUB main launch - recomputes the block with the full err sums and decides
// (no-op unless the sampled sums left the decision open, st.redo == -1).
template <int T, bool PIPE, int MINB, bool ZERO, int EMIT, bool ERRSUB = false>
__global__ void __launch_bounds__(TV_THREADS, MINB)
k_chamb_multi(const double* __restrict__ g, const double* __restrict__ pxi, const double* __restrict__ pyi,
              double* __restrict__ pxo, double* __restrict__ pyo, int nx, int ny, int seg, int nstrips,
              size_t img_stride, const Control* __restrict__ ctl, ChambState* __restrict__ st,
              double* __restrict__ partials, int redo, double* __restrict__ f) {
    static_assert(EMIT == 0 || ((T + 1) & ~1) > T, "EMIT needs a spare pixel of lateral validity");
    // Programmatic dependent launch (the launches of a prox follow each other on one stream, most of them no-ops or a
    // few microseconds long on small images): wait here for the previous kernel of the stream to complete and flush
    // before touching anything it wrote, then let the next launch be staged behind this one.  Both instructions are
    // no-ops when the kernel was launched without the programmatic-serialization attribute.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    constexpr int HL = (T + 1) & ~1;
    constexpr int WO = 64 - 2 * HL;
    __shared__ double sm[T * 32];
    const int img = blockIdx.z;
    ChambState* S = st + img;
    int nlev;
    if (redo == 1) {
        nlev = S->redo;
        if (nlev <= 0) return;
    } else if (redo == 2) {
        if (S->redo != -1) return;
        nlev = min(T, ctl->maxiter - S->k);
    } else {
        if (S->done) return;
        nlev = min(T, ctl->maxiter - S->k);
    }
    const double lambda = ctl->prox_lambda_run, tau = ctl->tau;
    const double invlam = 1.0 / lambda;
    const size_t off = (size_t)img * img_stride;
    g += off; pxi += off; pyi += off; pxo += off; pyo += off;
    if (EMIT) f += off;

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
        if (edge) cm_march<T, true, PIPE, ZERO, EMIT, ERRSUB>(g, pxi, pyi, pxo, pyo, nx, ny, j0, j1, L, invlam, tau, nlev, err, f, lambda);
        else      cm_march<T, false, PIPE, ZERO, EMIT, ERRSUB>(g, pxi, pyi, pxo, pyo, nx, ny, j0, j1, L, invlam, tau, nlev, err, f, lambda);
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
                if (redo == 1) {                    // the stop sweep was decided by the main launch
                    S->k = k0 + nlev; S->done = 1; S->redo = 0; S->buf ^= 1;
                } else {
                    int stop = 0;
                    bool open = false;              // ERRSUB: a sampled sum <= tol^2 proves nothing
                    double estop = 0.0, elast = 0.0;
#pragma unroll
                    for (int s = 1; s <= T; ++s) {
                        if (s <= nlev && stop == 0 && !open) {
                            const bool more = k0 + s < ctl->maxiter;
                            const bool cont = more && (e[s - 1] > ctl->tol);                      // :131
                            if (ERRSUB && more && !cont) open = true;                             // decided by the exact launch
                            else if (!cont) { stop = s; estop = e[s - 1]; }
                            elast = e[s - 1];
                        }
                    }
                    if (ERRSUB && open) {
                        S->redo = -1;               // nothing advances: the exact launch recomputes this block
                    } else if (stop == 0) {         // all nlev sweeps continue
                        S->k = k0 + nlev; S->err = elast; S->buf ^= 1; S->redo = 0;
                    } else if (stop == nlev) {      // stops exactly at the end of this block
                        S->k = k0 + nlev; S->err = estop; S->done = 1; S->buf ^= 1; S->redo = 0;
                        if (EMIT) S->emitted = 1;   // f written above is the prox output
                    } else {                        // stopped inside the block: redo with `stop` levels
                        S->err = estop; S->redo = stop;
                    }
                }
            }
        }
    }
}

}  // namespace sbd
