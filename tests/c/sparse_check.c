/* sparse_check.c -- plain C consumer of the C ABI: the sparse form of final_dist (ssg_rerank_finish_sparse,
 * ssg_eps_sparse, ssg_dbscan_sparse) against the dense entry points on the same inputs.
 *   1. every CSR entry equals the dense entry (bytes), rows ascending, every dense entry outside the CSR >= threshold
 *   2. ssg_eps_sparse certified and within 1e-13 (relative) of ssg_eps_estimate, same top_num
 *   3. ssg_dbscan_sparse labels == ssg_dbscan labels (bytes)
 *   4. a rho-slice that reaches past the threshold is reported as NOT certified
 * Run on a GPU box: tests/c/_build/sparse_check [n] [d]   (exit code 0 = all checks passed). */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ssg_b200.h"

#define CK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "FAIL %s:%d %s -> %d (%s)\n", __FILE__, __LINE__, #x, rc_, ssg_last_error()); return 2; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 2; } } while (0)

static double frand(unsigned long long* s) { *s = *s * 6364136223846793005ull + 1442695040888963407ull; return (double)(*s >> 11) / 9007199254740992.0; }

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 2000, d = argc > 2 ? atoi(argv[2]) : 64;
    const int ns = n / 2 + 3, k1 = 20, k2 = 6;
    const double lam = 0.1, rho = 1.6e-3 * 5;
    int fails = 0;
    unsigned long long seed = 777;
    const int nc = n / 20 > 0 ? n / 20 : 1;
    float* cent = (float*)malloc(sizeof(float) * nc * d);
    float* h_t = (float*)malloc(sizeof(float) * (size_t)n * d);
    float* h_s = (float*)malloc(sizeof(float) * (size_t)ns * d);
    for (int i = 0; i < nc * d; ++i) cent[i] = (float)(frand(&seed) * 2 - 1);
    for (int set = 0; set < 2; ++set) {
        float* f = set ? h_s : h_t;
        const int m = set ? ns : n;
        for (int i = 0; i < m; ++i) {
            const int c = (int)(frand(&seed) * nc) % nc;
            double nrm = 0;
            for (int k = 0; k < d; ++k) { f[(size_t)i * d + k] = cent[c * d + k] + 0.25f * (float)(frand(&seed) * 2 - 1); nrm += (double)f[(size_t)i * d + k] * f[(size_t)i * d + k]; }
            for (int k = 0; k < d; ++k) f[(size_t)i * d + k] = (float)(f[(size_t)i * d + k] / sqrt(nrm));
        }
    }
    const size_t nn = (size_t)n * n;
    float *d_t, *d_s; double* d_final;
    CU(cudaMalloc((void**)&d_t, sizeof(float) * (size_t)n * d)); CU(cudaMalloc((void**)&d_s, sizeof(float) * (size_t)ns * d));
    CU(cudaMalloc((void**)&d_final, 8 * nn));
    CU(cudaMemcpy(d_t, h_t, sizeof(float) * (size_t)n * d, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_s, h_s, sizeof(float) * (size_t)ns * d, cudaMemcpyHostToDevice));

    ssg_rerank_plan* rp = NULL; ssg_cluster_plan* cp = NULL;
    CK(ssg_rerank_plan_create(&rp, 0, n, ns, d));
    CK(ssg_cluster_plan_create(&cp, 0, n, 0));
    /* dense reference */
    CK(ssg_rerank_run(rp, d_s, ns, d_t, n, d, k1, k2, lam, SSG_DIST_EXACT, d_final, NULL, NULL));
    double eps0 = 0; long long top0 = 0; int ncl0 = 0;
    CK(ssg_eps_estimate(cp, d_final, SSG_F64, n, rho, &eps0, &top0, NULL));
    int64_t *d_l0, *d_l1;
    CU(cudaMalloc((void**)&d_l0, 8 * (size_t)n)); CU(cudaMalloc((void**)&d_l1, 8 * (size_t)n));
    CK(ssg_dbscan(cp, d_final, SSG_F64, n, eps0, 4, d_l0, &ncl0, NULL));
    double* h_f = (double*)malloc(8 * nn);
    CU(cudaMemcpy(h_f, d_final, 8 * nn, cudaMemcpyDeviceToHost));

    /* sparse form */
    long long nnz = 0, nnz2 = 0; double thr = 0;
    int *d_rp, *d_col; double* d_val;
    CK(ssg_rerank_distance_rows(rp, d_s, ns, d_t, n, d, k1, SSG_DIST_EXACT, 0, n, NULL, NULL));
    CK(ssg_rerank_finish_sparse(rp, d_t, n, d, k1, k2, lam, &nnz, NULL));
    CK(ssg_rerank_sparse_view(rp, &d_rp, &d_col, &d_val, &nnz2, &thr));
    int* h_rp = (int*)malloc(4 * ((size_t)n + 1)); int* h_col = (int*)malloc(4 * (size_t)(nnz + 1)); double* h_val = (double*)malloc(8 * (size_t)(nnz + 1));
    CU(cudaMemcpy(h_rp, d_rp, 4 * ((size_t)n + 1), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_col, d_col, 4 * (size_t)nnz, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(h_val, d_val, 8 * (size_t)nnz, cudaMemcpyDeviceToHost));
    {
        size_t bad_val = 0, bad_order = 0, bad_out = 0, diag = 0;
        unsigned char* in = (unsigned char*)calloc(n, 1);
        for (int i = 0; i < n; ++i) {
            memset(in, 0, n);
            for (int e = h_rp[i]; e < h_rp[i + 1]; ++e) {
                const int m = h_col[e];
                if (m < 0 || m >= n || (e > h_rp[i] && h_col[e - 1] >= m)) { ++bad_order; continue; }
                in[m] = 1;
                diag += m == i;
                bad_val += memcmp(&h_val[e], &h_f[(size_t)i * n + m], 8) != 0;
            }
            for (int m = 0; m < n; ++m) bad_out += !in[m] && !(h_f[(size_t)i * n + m] >= thr);
        }
        const int bad = bad_val || bad_order || bad_out || diag != (size_t)n || nnz != nnz2 || h_rp[n] != nnz;
        fails += bad;
        printf("CSR vs dense: nnz %lld (%.2f %% of n^2), threshold %.17g; value mismatches %zu, order %zu, outside-below-threshold %zu, diagonals %zu: %s\n",
               nnz, 100.0 * (double)nnz / (double)nn, thr, bad_val, bad_order, bad_out, diag, bad ? "FAIL" : "ok");
    }
    {
        double e = 0; long long top = 0; int ok = 0;
        CK(ssg_eps_sparse(cp, n, d_rp, d_col, d_val, thr, rho, &e, &top, &ok, NULL));
        /* the slice can be certified iff the dense matrix holds at least top_num non-zero upper-triangle entries below
         * the bound (all of those are in the CSR); otherwise the refusal is the correct answer */
        long long below = 0;
        for (int i = 0; i < n; ++i)
            for (int m = i + 1; m < n; ++m) { const double v = h_f[(size_t)i * n + m]; below += (v != 0.0 && v < thr); }
        const int expect = top0 <= below;
        const int bad = ok != expect || (ok && (top != top0 || !(fabs(e - eps0) <= 1e-13 * fabs(eps0))));
        fails += bad;
        printf("eps sparse %.17g vs dense %.17g (top %lld vs %lld, %lld entries below the bound, certified %d, expected %d): %s\n",
               e, eps0, top, top0, below, ok, expect, bad ? "FAIL" : "ok");
        CK(ssg_eps_sparse(cp, n, d_rp, d_col, d_val, thr, 0.5, &e, &top, &ok, NULL));
        fails += ok != 0;
        printf("rho = 0.5 must not be certified: certified %d: %s\n", ok, ok ? "FAIL" : "ok");
    }
    {
        int ncl = -1;
        int64_t* h_l0 = (int64_t*)malloc(8 * (size_t)n); int64_t* h_l1 = (int64_t*)malloc(8 * (size_t)n);
        CK(ssg_dbscan_sparse(cp, n, d_rp, d_col, d_val, eps0, 4, d_l1, &ncl, NULL));
        CU(cudaMemcpy(h_l0, d_l0, 8 * (size_t)n, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(h_l1, d_l1, 8 * (size_t)n, cudaMemcpyDeviceToHost));
        const int bad = memcmp(h_l0, h_l1, 8 * (size_t)n) != 0 || ncl != ncl0 || !(eps0 < thr);
        fails += bad;
        printf("dbscan sparse: %d clusters vs %d (eps %.6f < threshold %.6f): %s\n", ncl, ncl0, eps0, thr, bad ? "FAIL" : "ok");
    }
    printf("%s (n=%d, d=%d)\n", fails ? "SPARSE_CHECK FAILED" : "SPARSE_CHECK PASSED", n, d);
    return fails ? 1 : 0;
}
