// Prototype / microbenchmark (not part of the library): the transposed side of the rows pass done by the TMA.
// A block owns one line pair (q = 2b, 2b+1) and all nk = N/2+1 bins k; the half spectrum is stored column-major
// spec[k][q] (16-byte complex elements), so the block's data is a [nk x 2] sub-block with 32-byte rows.
//   k_store: smem staging [k][2] -> global with cp.async.bulk.tensor.2d (9 boxes of 256 x 2), measures write BW
//   k_load : global -> smem with cp.async.bulk.tensor.2d + mbarrier, checks and measures read BW
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tma_transpose tma_transpose.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int N = 4096, NK = N / 2 + 1, KBOX = 256, NBOX = (NK + KBOX - 1) / KBOX;

__global__ void __launch_bounds__(256) k_store(const __grid_constant__ CUtensorMap tmap, int nimg) {
    extern __shared__ __align__(128) double2 stg[];          // [NBOX*KBOX][2]
    const int b = blockIdx.x, img = blockIdx.y;
    for (int k = threadIdx.x; k < NBOX * KBOX; k += blockDim.x) {
        stg[2 * k] = make_double2((double)k, (double)(2 * b) + 0.001 * img);
        stg[2 * k + 1] = make_double2((double)k, (double)(2 * b + 1) + 0.001 * img);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int bx = 0; bx < NBOX; ++bx) {
            const int c0 = 4 * b;                                  // doubles along q: 2 q's x (re, im)
            const int c1 = img * NK + bx * KBOX;                   // row = k (images stacked along k)
            // rows beyond this image's nk would spill into the next image: clip by issuing a shorter box? the last
            // box only has 1 valid row (k = 2048); use a second tensor map with box height 1 for it
            if (bx < NBOX - 1)
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];"
                             :: "l"(&tmap), "r"(c0), "r"(c1), "r"(smem_u32(stg + 2 * bx * KBOX)) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    // Nyquist row by plain stores
    if (threadIdx.x < 2) {
        double2* g = nullptr; (void)g;
    }
}

__global__ void __launch_bounds__(256) k_load(const __grid_constant__ CUtensorMap tmap, int nimg, unsigned long long* bad) {
    extern __shared__ __align__(128) double2 stg[];
    __shared__ __align__(8) unsigned long long mbar;
    const int b = blockIdx.x, img = blockIdx.y;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (NBOX - 1) * KBOX * 32;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(bytes) : "memory");
        for (int bx = 0; bx < NBOX - 1; ++bx) {
            const int c0 = 4 * b, c1 = img * NK + bx * KBOX;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         :: "r"(smem_u32(stg + 2 * bx * KBOX)), "l"(&tmap), "r"(c0), "r"(c1), "r"(smem_u32(&mbar)) : "memory");
        }
    }
    // wait (parity 0)
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
        }
    }
    unsigned long long nb = 0;
    for (int k = threadIdx.x; k < (NBOX - 1) * KBOX; k += blockDim.x) {
        const double2 a = stg[2 * k], c = stg[2 * k + 1];
        if (a.x != (double)k || a.y != (double)(2 * b) + 0.001 * img) nb++;
        if (c.x != (double)k || c.y != (double)(2 * b + 1) + 0.001 * img) nb++;
    }
    if (nb) atomicAdd(bad, nb);
}

// reference: the same transposed store with plain 32-byte vector stores (what an LSU-based rows pass does)
__global__ void __launch_bounds__(256) k_store_lsu(double2* __restrict__ spec, int nimg) {
    const int b = blockIdx.x, img = blockIdx.y;
    double2* base = spec + (size_t)img * NK * N;
    for (int k = threadIdx.x; k < (NBOX - 1) * KBOX; k += blockDim.x) {
        double2* p = base + (size_t)k * N + 2 * b;
        const double2 v0 = make_double2((double)k, (double)(2 * b) + 0.001 * img), v1 = make_double2((double)k, (double)(2 * b + 1) + 0.001 * img);
        asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" :: "l"(p), "d"(v0.x), "d"(v0.y), "d"(v1.x), "d"(v1.y) : "memory");
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int nimg = 8;
    double2* spec;
    const size_t elems = (size_t)nimg * NK * N;
    CK(cudaMalloc(&spec, elems * sizeof(double2)));
    CK(cudaMemset(spec, 0, elems * sizeof(double2)));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
    if (!fn) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)2 * N, (cuuint64_t)nimg * NK};           // doubles along q, rows = (img, k)
    const cuuint64_t gstr[1] = {(cuuint64_t)N * 16};
    const cuuint32_t box[2] = {4, KBOX};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, spec, gdim, gstr, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const size_t smem = (size_t)NBOX * KBOX * 32;
    CK(cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_load, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned long long* bad; CK(cudaMalloc(&bad, 8)); CK(cudaMemset(bad, 0, 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid(N / 2, nimg);
    const double bytes = (double)nimg * (NBOX - 1) * KBOX * N * 16;
    for (int rep = 0; rep < 3; ++rep) {
        float ms;
        cudaEventRecord(e0); k_store<<<grid, 256, smem>>>(tmap, nimg); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("tma store : %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(e0); k_load<<<grid, 256, smem>>>(tmap, nimg, bad); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("tma load  : %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
        cudaEventRecord(e0); k_store_lsu<<<grid, 256>>>(spec, nimg); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); printf("lsu store : %.3f ms  %.0f GB/s\n", ms, bytes / ms / 1e6);
    }
    unsigned long long hb; CK(cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost));
    printf("mismatches after tma store + tma load: %llu\n", hb);
    CK(cudaGetLastError());
    return hb != 0;
}
