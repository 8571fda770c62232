// Microbenchmark (not part of the library): the cooperative Chambolle prox kernel alone on one n x n image, with the
// per-phase clock64 breakdown of a sweep (block 0, thread 0) when built with -DSBD_CC_TIMING.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -DSBD_CC_TIMING -I../../semi-blind-image-deblurring-problems-with-tv_b200/csrc -o coop_bench coop_bench.cu
#include "tv_coop.cuh"
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace sbd;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 256, K = argc > 2 ? atoi(argv[2]) : 25, upw = argc > 3 ? atoi(argv[3]) : 1, reps = 200;
    const size_t npix = (size_t)n * n;
    double *g, *px0, *py0, *px1, *py1, *f, *part;
    for (double** p : {&g, &px0, &py0, &px1, &py1, &f}) CK(cudaMalloc(p, npix * 8));
    std::vector<double> h(npix);
    for (size_t i = 0; i < npix; ++i) {
        const int x = (int)(i % n), y = (int)(i / n);
        h[i] = 120.0 + 80.0 * sin(x / 37.0) * cos(y / 23.0) + 40.0 * (((x / 64) + (y / 48)) % 2) + 9.0 * ((double)((x * 7919 + y * 104729) % 1000) / 1000.0 - 0.5);
    }
    CK(cudaMemcpy(g, h.data(), npix * 8, cudaMemcpyHostToDevice));
    Control hc; memset(&hc, 0, sizeof hc);
    hc.prox_lambda_theta = 0.3; hc.prox_lambda_run = 0.3; hc.tau = 0.249; hc.tol = 0.0; hc.maxiter = K;
    Control* ctl; CK(cudaMalloc(&ctl, sizeof hc)); CK(cudaMemcpy(ctl, &hc, sizeof hc, cudaMemcpyHostToDevice));
    ChambState* st; CK(cudaMalloc(&st, sizeof(ChambState))); CK(cudaMemset(st, 0, sizeof(ChambState)));
    unsigned int* bar; CK(cudaMalloc(&bar, 4));
    const int nstrips = (n + 63) / 64, units = n * nstrips, bpi = (units + CC_WARPS * upw - 1) / (CC_WARPS * upw);
    CK(cudaMalloc(&part, sizeof(double) * 2 * bpi));
    int nx = n, ny = n, zero_start = 1;
    int* trace = nullptr; int ntrace = 0;
    void* args[] = {&g, &px0, &py0, &px1, &py1, &f, &nx, &ny, (void*)&npix, (void*)&bpi, (void*)&upw, &zero_start, (void*)&K, &ctl, &st, &part, &bar, &ctl, &trace, &ntrace};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&]() {
        CK(cudaMemsetAsync(bar, 0, 4));
        CK(cudaLaunchCooperativeKernel((void*)k_chamb_coop, dim3(bpi), dim3(CC_THREADS), args, 0, 0));
    };
    for (int i = 0; i < 5; ++i) run();
    CK(cudaDeviceSynchronize());
#ifdef SBD_CC_TIMING
    long long z8[8] = {0}; CK(cudaMemcpyToSymbol(cc_timing, z8, sizeof z8));
#endif
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) run();
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    ChambState hs; CK(cudaMemcpy(&hs, st, sizeof hs, cudaMemcpyDeviceToHost));
    printf("n %d K %d upw %d blocks %d: %.2f us per prox (incl. memset + launch), k=%d err=%.6e\n", n, K, upw, bpi, 1e3 * ms / reps, hs.k, hs.err);
#ifdef SBD_CC_TIMING
    CK(cudaMemcpyFromSymbol(z8, cc_timing, sizeof z8));
    const char* nm[7] = {"stores issued", "block sum", "release (fence + count)", "wait for other blocks", "closing __syncthreads", "operands arrived", "level step done"};
    for (int i : {5, 6, 0, 1, 2, 3, 4}) printf("   %-24s %8.0f cycles per sweep\n", nm[i], (double)z8[i] / ((double)reps * K));
#endif
    return 0;
}
