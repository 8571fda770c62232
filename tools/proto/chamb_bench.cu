// Microbenchmark (not part of the library): the 4-level fused Chambolle kernel alone, at the benchmark geometry
// (4096^2, `batch` chains, 128-row segments), for A/B-ing variants of tv_multi.cuh built with -D switches.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -I../../semi-blind-image-deblurring-problems-with-tv_b200/csrc -o chamb_bench chamb_bench.cu
#include "tv_multi.cuh"
#ifndef HARNESS_ERRSUB
#define HARNESS_ERRSUB false
#endif
#ifndef HARNESS_T
#define HARNESS_T 4
#endif
#ifndef HARNESS_MINB
#define HARNESS_MINB 3
#endif
#include <cstdio>
#include <cstdlib>
#include <vector>
using namespace sbd;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char** argv) {
    const int n = 4096, batch = argc > 1 ? atoi(argv[1]) : 8, seg = argc > 2 ? atoi(argv[2]) : 128, reps = argc > 3 ? atoi(argv[3]) : 20;
    const size_t npix = (size_t)n * n, tot = npix * batch;
    double *g, *px0, *py0, *px1, *py1, *part;
    CK(cudaMalloc(&g, tot * 8)); CK(cudaMalloc(&px0, tot * 8)); CK(cudaMalloc(&py0, tot * 8));
    CK(cudaMalloc(&px1, tot * 8)); CK(cudaMalloc(&py1, tot * 8));
    std::vector<double> h(npix);
    for (size_t i = 0; i < npix; ++i) {
        const int x = (int)(i % n), y = (int)(i / n);
        h[i] = 120.0 + 80.0 * sin(x / 37.0) * cos(y / 23.0) + 40.0 * (((x / 64) + (y / 48)) % 2) + 9.0 * ((double)((x * 7919 + y * 104729) % 1000) / 1000.0 - 0.5);
    }
    for (int b = 0; b < batch; ++b) CK(cudaMemcpy(g + b * npix, h.data(), npix * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(px0, 0, tot * 8)); CK(cudaMemset(py0, 0, tot * 8));
    Control hc; memset(&hc, 0, sizeof hc);
    hc.prox_lambda_theta = 0.3; hc.prox_lambda_run = 0.3; hc.tau = 0.249; hc.tol = 0.0; hc.maxiter = 1 << 30;
    Control* ctl; CK(cudaMalloc(&ctl, sizeof hc)); CK(cudaMemcpy(ctl, &hc, sizeof hc, cudaMemcpyHostToDevice));
    ChambState* st; CK(cudaMalloc(&st, sizeof(ChambState) * batch)); CK(cudaMemset(st, 0, sizeof(ChambState) * batch));
    constexpr int T = HARNESS_T, HL = HARNESS_T, WO = 64 - 2 * HL;
    const int strips = (n + WO - 1) / WO;
    dim3 grid((strips + TV_WARPS - 1) / TV_WARPS, (n + seg - 1) / seg, batch);
    CK(cudaMalloc(&part, sizeof(double) * T * grid.x * grid.y * batch));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](int i) {
        const double* pxi = (i & 1) ? px1 : px0; const double* pyi = (i & 1) ? py1 : py0;
        double* pxo = (i & 1) ? px0 : px1; double* pyo = (i & 1) ? py0 : py1;
        k_chamb_multi<HARNESS_T, false, HARNESS_MINB, false, 0, HARNESS_ERRSUB><<<grid, TV_THREADS>>>(g, pxi, pyi, pxo, pyo, n, n, seg, strips, npix, ctl, st, part, 0, nullptr);
    };
    for (int i = 0; i < 4; ++i) run(i);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) run(i);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<ChambState> hs(batch);
    CK(cudaMemcpy(hs.data(), st, sizeof(ChambState) * batch, cudaMemcpyDeviceToHost));
    double chk = 0; std::vector<double> o(1024);
    CK(cudaMemcpy(o.data(), px0 + npix / 2 + 77, 1024 * 8, cudaMemcpyDeviceToHost));
    for (double v : o) chk += v;
    printf("batch %d seg %d: %.4f ms per T-sweep launch (%.4f ms per sweep of 8 chains-equivalent)  k=%d err=%.10e chk=%.12e\n",
           batch, seg, ms / reps, ms / reps / T * 8 / batch, hs[0].k, hs[0].err, chk);
    return 0;
}
