/* tv_oracle.c - plain-C restatement of the reference's TV prox for LARGE images.
 *
 * ORACLE = TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and the CPU
 * legs of bench.py may load this library (through oracle/tv.py); the product
 * (libsbd.so) never links or loads it.
 *
 * Restates, operation by operation and in the same floating-point order as
 * oracle/tv.py (which is pinned to the reference through tests/golden/):
 *   utils/chambolle_prox_TV_stop.m:120-131  sweep loop and stop test
 *   utils/chambolle_prox_TV_stop.m:149      f = g - lambda*DivergenceIm(px,py)
 *   utils/chambolle_prox_TV_stop.m:152-159  DivergenceIm (last row/col = -p(end))
 *   utils/chambolle_prox_TV_stop.m:161-166  GradientIm (forward differences, zero last row/col)
 *   utils/TVnorm.m:1-2 with SALSA/diffh.m, diffv.m (periodic backward differences)
 * Arrays are numpy C-order a[i*N + j] == MATLAB a(i+1, j+1), M rows x N columns.
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (no FMA contraction, so
 * every element is bit-identical to the numpy restatement; only the order of the
 * err / TV sums differs: rows are summed left to right, then the row sums top to
 * bottom, independent of the number of threads).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* number of threads of the loops below (a launcher such as torchrun exports OMP_NUM_THREADS=1);
 * returns the number now in effect, 1 when built without OpenMP */
int oc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static double div_at(const double* px, const double* py, long M, long N, long i, long j) {
    /* :153-154  v = [p2(:,1)  p2(:,2:end-1)-p2(:,1:end-2)  -p2(:,end)]   (p2 = py, along j) */
    const double v = (j == 0) ? py[i * N] : ((j == N - 1) ? -py[i * N + j] : py[i * N + j] - py[i * N + j - 1]);
    /* :156-157  u = [p1(1,:); p1(2:end-1,:)-p1(1:end-2,:); -p1(end,:)]   (p1 = px, along i) */
    const double u = (i == 0) ? px[j] : ((i == M - 1) ? -px[i * N + j] : px[i * N + j] - px[(i - 1) * N + j]);
    return v + u;                                                           /* :159 */
}

/* px, py: in = starting dual pair, out = final dual pair.  Returns the sweep count k. */
int oc_chambolle(const double* g, long M, long N, double lam, int maxiter, double tol, double tau,
                 double* px, double* py, double* f, double* err_out) {
    double* u = (double*)malloc(sizeof(double) * (size_t)M * N);
    double* rowsum = (double*)malloc(sizeof(double) * (size_t)M);
    if (!u || !rowsum) { free(u); free(rowsum); return -1; }
    int k = 0;
    double err = 0.0;
    for (;;) {                                                              /* :120 */
        k += 1;                                                             /* :121 */
#pragma omp parallel for schedule(static)
        for (long i = 0; i < M; ++i)
            for (long j = 0; j < N; ++j)
                u[i * N + j] = div_at(px, py, M, N, i, j) - g[i * N + j] / lam;     /* :123-124 */
#pragma omp parallel for schedule(static)
        for (long i = 0; i < M; ++i) {
            double s = 0.0;
            for (long j = 0; j < N; ++j) {
                const double uc = u[i * N + j];
                const double upx = (i < M - 1) ? u[(i + 1) * N + j] - uc : 0.0;     /* :162-163 */
                const double upy = (j < N - 1) ? u[i * N + j + 1] - uc : 0.0;       /* :165-166 */
                const double tmp = sqrt(upx * upx + upy * upy);                     /* :127 */
                const double a = -upx + tmp * px[i * N + j], b = -upy + tmp * py[i * N + j];
                s += a * a + b * b;                                                 /* :128 */
                px[i * N + j] = (px[i * N + j] + tau * upx) / (1.0 + tau * tmp);    /* :129 */
                py[i * N + j] = (py[i * N + j] + tau * upy) / (1.0 + tau * tmp);    /* :130 */
            }
            rowsum[i] = s;
        }
        double tot = 0.0;
        for (long i = 0; i < M; ++i) tot += rowsum[i];
        err = sqrt(tot);                                                    /* :128 (...)^0.5 */
        if (!((k < maxiter) && (err > tol))) break;                         /* :131 */
    }
    /* (the in-place update of p is safe: a sweep reads p only through u, formed before, and its own element) */
#pragma omp parallel for schedule(static)
    for (long i = 0; i < M; ++i)
        for (long j = 0; j < N; ++j)
            f[i * N + j] = g[i * N + j] - lam * div_at(px, py, M, N, i, j);         /* :149 */
    free(u); free(rowsum);
    if (err_out) *err_out = err;
    return k;
}

/* utils/TVnorm.m:2  sum(sum(sqrt(diffh(x).^2 + diffv(x).^2))) */
double oc_tvnorm(const double* x, long M, long N) {
    double* rowsum = (double*)malloc(sizeof(double) * (size_t)M);
    if (!rowsum) return NAN;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < M; ++i) {
        const long ip = (i == 0) ? M - 1 : i - 1;
        double s = 0.0;
        for (long j = 0; j < N; ++j) {
            const long jp = (j == 0) ? N - 1 : j - 1;
            const double dh = x[i * N + j] - x[i * N + jp];                 /* diffh.m:2-3 */
            const double dv = x[i * N + j] - x[ip * N + j];                 /* diffv.m:2-3 */
            s += sqrt(dh * dh + dv * dv);
        }
        rowsum[i] = s;
    }
    double tot = 0.0;
    for (long i = 0; i < M; ++i) tot += rowsum[i];
    free(rowsum);
    return tot;
}
