// philox.cuh - counter-based Gaussian noise, replaces `randn(size(X))`
// (SAPG/SAPG_algorithm_Guassian.m:81,160).  Definition shared with
// oracle/philox.py (see its header for the exact counter/key layout).
#pragma once
#include "common.cuh"

namespace sbd {

__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3,
                                              uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// two standard normals for the element pair `pair` of draw `step` on `stream`
__device__ __forceinline__ double2 philox_normal2(uint64_t seed, uint32_t stream, uint32_t step, uint64_t pair) {
    uint32_t c0 = (uint32_t)pair, c1 = (uint32_t)(pair >> 32), c2 = step, c3 = stream;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = (((uint64_t)c1 << 32) | c0) >> 11;
    const uint64_t b = (((uint64_t)c3 << 32) | c2) >> 11;
    const double u1 = ((double)a + 0.5) * 0x1p-53;
    const double u2 = ((double)b + 0.5) * 0x1p-53;
    const double r = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);         // = sincos(2 pi u2) to rounding, without the argument reduction of a general angle
    return make_double2(r * c, r * s);
}

}  // namespace sbd
