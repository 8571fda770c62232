// common.cuh - shared declarations for libsbd (sm_100a, fp64).
//
// Device layout (DESIGN.md "Data layout in HBM"):
//   an image is the MATLAB column-major rows x cols buffer taken as is:
//   element (i, j) [MATLAB (i+1, j+1)] lives at  j*nx + i  with nx = rows
//   (fast axis, "x") and ny = cols (slow axis, "y").
//   a half spectrum holds bins k = 0..nx/2 of the fast axis for all ny slow
//   indices:  spec[q*sp + k]  (double2 = re, im),  sp = pitch >= nx/2+1.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <math.h>

#include "../../include/sbd.h"

namespace sbd {

constexpr int MAXT = 15;            // largest supported psf_size
constexpr int NSTAT = 6;            // per-chain statistics: tv, rss, c0, c1, sqerr, spare

// ---------------------------------------------------------------------------
// device-resident control block: every scalar the iteration needs lives here so
// that a MYULA iteration can be replayed (CUDA graph) without host involvement.
// ---------------------------------------------------------------------------
struct Control {
    // current parameters
    double theta, sigma2, psi[2];
    // derived, refreshed by the scalar kernel
    double prox_lambda_theta;   // lambda passed to Chambolle = prox_lambda * theta
    double inv_scale;           // 1 / (sigma2 * nx * ny): folded into the inverse spectral multiply
    // Chambolle options
    double tau, tol;
    int maxiter;
    // counters
    int ii;                     // MATLAB loop index of the NEXT iteration (2..)
    unsigned int draw;          // Philox step counter (number of noise images drawn so far per chain)
    int post_n;                 // samples accumulated in the posterior mean
    int phase;                  // 0 warm-up, 1 main
    // The prox of an iteration runs on its own stream next to the analysis of the new sample and the scalar update
    // (sbd.cu): it works from a SNAPSHOT of lambda*theta taken by k_chamb_reset, and counts its own calls
    double prox_lambda_run;     // lambda*theta of the prox in flight
    int prox_count;             // slot of the chamb_k trace the next main-loop prox fills
    int pad_;
};

struct ChambState {             // one per image / chain
    int k;                      // sweeps executed
    int done;                   // stop flag (k >= maxiter || err <= tol)
    double err;                 // err of the last executed sweep
    unsigned int counter;       // last-block-done ticket
    int redo;                   // fused kernel: number of levels the redo launch must apply (0 = none)
    int buf;                    // which buffer of the ping-pong pair holds the current dual pair
    int emitted;                // fused kernel: the prox output f was written by the final sweep block
};

struct SapgConst {              // constants of a run (device copy)
    int model, t, npsi;
    double phi;
    double gam, lamb, sq2gam, prox_lambda;
    double min_th, max_th, c_theta;
    double psi_min[2], psi_max[2], c_psi[2], psi_fixed[2];
    int fix_psi[2];
    double sigma2_min, sigma2_max, c_sigma2, sigma2_fixed;
    int fix_sigma;
    double dimX;
    int n_local, n_total;       // chains on this rank / over all ranks
    int burnIn, samples, warmup;
    int has_xtrue;
};

struct Traces {                 // device trace arrays (length samples unless noted)
    double *logPiWU;            // [warmup]
    double *thetas, *sigmas, *psi0, *psi1;
    double *g_theta, *g_psi0, *g_psi1, *g_sigma;
    double *logPi, *gX, *sqerr;
    int *chamb_k;
    const double* delta;        // [samples+1]  delta(ii), host-precomputed (same libm pow as the oracle)
};

// ---------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------
struct Error {
    int code;
    std::string msg;
};

#define SBD_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            char b__[512];                                                               \
            snprintf(b__, sizeof b__, "%s:%d: %s -> %s", __FILE__, __LINE__, #call,      \
                     cudaGetErrorString(e__));                                           \
            throw sbd::Error{e__ == cudaErrorMemoryAllocation ? SBD_E_NOMEM : SBD_E_CUDA, b__}; \
        }                                                                                \
    } while (0)

#define SBD_REQUIRE(cond, code, text)                                                    \
    do {                                                                                 \
        if (!(cond)) throw sbd::Error{(code), std::string(text)};                        \
    } while (0)

// ---------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {   // a * conj(b)
    return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {   // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}

__device__ __forceinline__ double shfl_down_d(double v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_up_d(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
__device__ __forceinline__ double shfl_xor_d(double v, int d) { return __shfl_xor_sync(0xffffffffu, v, d); }

// warp sum, fixed xor tree -> bitwise deterministic
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

// Block sum of NV values per thread.  Result valid in thread 0.  `sm` must hold
// NV * 32 doubles.  Fixed order (warp tree, then warps in index order).
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sm) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sm[i * 32 + warp] = v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            double s = 0.0;
            for (int w = 0; w < nwarp; ++w) s += sm[i * 32 + w];
            v[i] = s;
        }
    }
}

// "last block done" ticket (CUDA threadFenceReduction pattern).  Thread 0 has
// already written this block's partials.  Returns true in ALL threads of the
// last block to arrive; that block may then read every block's partials.
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, unsigned int nblocks) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        s_last = (t == nblocks - 1u);
        if (s_last) *counter = 0u;      // re-arm for the next launch
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last != 0;
}

// Deterministic sum of n partials at stride `stride` by the calling warp
// (lane-strided sequential sums, then the fixed xor tree).
__device__ __forceinline__ double warp_sum_partials(const double* p, int n, int stride) {
    const int lane = threadIdx.x & 31;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += __ldcg(p + (size_t)i * stride);
    return warp_sum(s);
}

}  // namespace sbd
