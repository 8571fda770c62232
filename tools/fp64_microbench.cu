// fp64 pipe microbenchmark (B200): DFMA / DADD / DMUL throughput per SM and dependent latency.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void k_tp(double* out, int iters, double a, double b) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) v[i] = fma(v[i], a, b);
            else if (OP == 1) v[i] = v[i] + a;
            else v[i] = v[i] * a;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// fp64 mixed with NI independent integer instructions per DFMA: does a non-fp64 instruction cost the
// fp64 stream issue time?  (the fused Chambolle loop has 0.57 non-fp64 instructions per fp64 one)
template <int ILP, int NI>
__global__ void k_mix(double* out, int iters, double a, double b, unsigned int m) {
    double v[ILP];
    unsigned int w[8];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = threadIdx.x * 7u + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            v[i] = fma(v[i], a, b);
#pragma unroll
            for (int j = 0; j < NI; ++j) w[(i * NI + j) & 7] = (w[(i * NI + j) & 7] ^ m) + 0x9e3779b9u;   // LOP3 + IADD -> 2 instr
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    unsigned int x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (double)x;
}
template <int ILP, int NI>
void run_mix(int warps_per_sm) {
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int threads = 128, blocks = p.multiProcessorCount * warps_per_sm / 4;
    double* out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    const int iters = 20000;
    k_mix<ILP, NI><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9, 0x5bd1e995u);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_mix<ILP, NI><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9, 0x5bd1e995u);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double ops = (double)blocks * threads * iters * ILP;
    printf("MIX    ILP=%d int-ops/DFMA=%d warps/SM=%2d : %.3f ms  %.1f DFMA lanes/clk/SM\n", ILP, 2 * NI, warps_per_sm, ms,
           ops / (ms * 1e-3) / (clk * 1e3) / p.multiProcessorCount);
    cudaFree(out);
}

template <int ILP, int OP>
void run(const char* name, int warps_per_sm) {
    int dev; cudaGetDevice(&dev);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    const int threads = 128, blocks = p.multiProcessorCount * warps_per_sm / 4;
    double* out; cudaMalloc(&out, sizeof(double) * threads * blocks);
    const int iters = 20000;
    k_tp<ILP, OP><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k_tp<ILP, OP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const double ops = (double)blocks * threads * iters * ILP;
    const double per_clk_sm = ops / (ms * 1e-3) / (clk * 1e3) / p.multiProcessorCount;
    printf("%-6s ILP=%d warps/SM=%2d : %.3f ms  %.1f Gop/s  %.1f lanes/clk/SM (at %d MHz nominal)\n", name, ILP,
           warps_per_sm, ms, ops / ms * 1e-6, per_clk_sm, clk / 1000);
    cudaFree(out);
}

int main() {
    run<1, 0>("DFMA", 4);  run<1, 0>("DFMA", 8);  run<1, 0>("DFMA", 16); run<1, 0>("DFMA", 32); run<1, 0>("DFMA", 64);
    run<2, 0>("DFMA", 8);  run<2, 0>("DFMA", 12); run<2, 0>("DFMA", 16); run<2, 0>("DFMA", 20); run<2, 0>("DFMA", 24); run<2, 0>("DFMA", 32);
    run<3, 0>("DFMA", 12); run<4, 0>("DFMA", 12); run<6, 0>("DFMA", 8); run<6, 0>("DFMA", 12);
    run<4, 0>("DFMA", 4);  run<4, 0>("DFMA", 8);  run<4, 0>("DFMA", 16); run<8, 0>("DFMA", 16); run<8, 0>("DFMA", 32);
    run<8, 1>("DADD", 16); run<8, 2>("DMUL", 16);
    run_mix<8, 0>(12); run_mix<8, 1>(12); run_mix<8, 2>(12); run_mix<8, 4>(12);
    run_mix<4, 1>(12); run_mix<2, 1>(12);
    return 0;
}
