/* shard_check.c -- plain C consumer of the C ABI (include/ssg_b200.h), no Python, no torch:
 * the row-sharded entry points against the single-GPU ones, with `world` ranks simulated one after another on ONE
 * device (the collectives between them are done by hand through host memory).
 *   1. ssg_rerank_finish_rows blocks  ==  ssg_rerank_run's final_dist            (bytes)
 *   2. sharded eps (3-pass and 6-pass)  ~  ssg_eps_estimate (<= 1e-13 relative), identical on every rank
 *   3. sharded DBSCAN labels           ==  ssg_dbscan labels                      (bytes)
 * Build: tests/c/Makefile.  Run on a GPU box: tests/c/_build/shard_check [n] [world]   (exit code 0 = all equal). */
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ssg_b200.h"

#define CK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "FAIL %s:%d %s -> %d (%s)\n", __FILE__, __LINE__, #x, rc_, ssg_last_error()); return 2; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define MAXW 8

static int lo_of(int n, int w, int r) { int b = n / w, m = n % w; return r * b + (r < m ? r : m); }
static double frand(unsigned long long* s) { *s = *s * 6364136223846793005ull + 1442695040888963407ull; return (double)(*s >> 11) / 9007199254740992.0; }

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1500, world = argc > 2 ? atoi(argv[2]) : 3;
    const int ns = n / 2 + 3, d = 64, k1 = 20, k2 = 6;
    const double lam = 0.1, rho = 0.02;
    int fails = 0;
    if (world < 1 || world > MAXW || n < 64) { fprintf(stderr, "usage: shard_check [n>=64] [world<=8]\n"); return 2; }

    /* clustered unit-norm features (n/20 centres), as bench.py's feature-level generator */
    unsigned long long seed = 12345;
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
    float *d_t, *d_s; double *d_final, *d_blocks;
    const size_t nn = (size_t)n * n;
    CU(cudaMalloc((void**)&d_t, sizeof(float) * (size_t)n * d)); CU(cudaMalloc((void**)&d_s, sizeof(float) * (size_t)ns * d));
    CU(cudaMalloc((void**)&d_final, 8 * nn)); CU(cudaMalloc((void**)&d_blocks, 8 * nn));
    CU(cudaMemcpy(d_t, h_t, sizeof(float) * (size_t)n * d, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_s, h_s, sizeof(float) * (size_t)ns * d, cudaMemcpyHostToDevice));

    /* 1. rows of final_dist */
    ssg_rerank_plan* rp = NULL;
    CK(ssg_rerank_plan_create(&rp, 0, n, ns, d));
    CK(ssg_rerank_run(rp, d_s, ns, d_t, n, d, k1, k2, lam, SSG_DIST_EXACT, d_final, NULL, NULL));
    CU(cudaMemset(d_blocks, 0xff, 8 * nn));
    for (int r = 0; r < world; ++r) {
        const int lo = lo_of(n, world, r), hi = lo_of(n, world, r + 1);
        CK(ssg_rerank_finish_rows(rp, d_t, n, d, k1, k2, lam, lo, hi - lo, d_blocks + (size_t)lo * n, NULL));
    }
    double* h_f = (double*)malloc(8 * nn); double* h_b = (double*)malloc(8 * nn);
    CU(cudaMemcpy(h_f, d_final, 8 * nn, cudaMemcpyDeviceToHost)); CU(cudaMemcpy(h_b, d_blocks, 8 * nn, cudaMemcpyDeviceToHost));
    { int bad = memcmp(h_f, h_b, 8 * nn) != 0; fails += bad; printf("finish_rows == finish (bytes): %s\n", bad ? "FAIL" : "ok"); }
    { size_t asym = 0; for (int i = 0; i < n; ++i) for (int j = 0; j < i; ++j) asym += h_f[(size_t)i * n + j] != h_f[(size_t)j * n + i];
      fails += asym != 0; printf("final_dist symmetric: %s (%zu)\n", asym ? "FAIL" : "ok", asym); }

    /* 2. eps */
    ssg_cluster_plan* cp0 = NULL; ssg_cluster_plan* cp[MAXW];
    CK(ssg_cluster_plan_create(&cp0, 0, n, 0));
    double eps0 = 0; long long top0 = 0;
    CK(ssg_eps_estimate(cp0, d_final, SSG_F64, n, rho, &eps0, &top0, NULL));
    void *hist[MAXW], *state[MAXW], *partial[MAXW], *list[MAXW], *cnt[MAXW], *nbr[MAXW];
    for (int r = 0; r < world; ++r) { CK(ssg_cluster_plan_create(&cp[r], 0, n, 0)); CK(ssg_cluster_buffers(cp[r], &hist[r], &state[r], &partial[r], &list[r], &cnt[r], &nbr[r])); }
    unsigned long long* hh = (unsigned long long*)malloc(8 * 4096); unsigned long long* hsum = (unsigned long long*)malloc(8 * 4096);
    double* hpart = (double*)malloc(8 * (size_t)n); double* hlist = (double*)malloc(8 * (size_t)(1 << 20));
    for (int exact = 0; exact < 2; ++exact) {
        for (int r = 0; r < world; ++r) CK(ssg_eps_shard_begin(cp[r], NULL));
        for (int pass = 0; pass < (exact ? 6 : 2); ++pass) {
            memset(hsum, 0, 8 * 4096);
            for (int r = 0; r < world; ++r) {
                CK(ssg_eps_shard_hist(cp[r], d_final + (size_t)lo_of(n, world, r) * n, SSG_F64, n, world, r, pass, NULL));
                CU(cudaMemcpy(hh, hist[r], 8 * 4096, cudaMemcpyDeviceToHost));
                for (int b = 0; b < 4096; ++b) hsum[b] += hh[b];
            }
            for (int r = 0; r < world; ++r) { CU(cudaMemcpy(hist[r], hsum, 8 * 4096, cudaMemcpyHostToDevice)); CK(ssg_eps_shard_pick(cp[r], pass, rho, NULL)); }
        }
        long long total = 0;
        for (int r = 0; r < world; ++r) {
            const int lo = lo_of(n, world, r), hi = lo_of(n, world, r + 1);
            long long c = 0;
            CK(ssg_eps_shard_gather(cp[r], d_final + (size_t)lo * n, SSG_F64, n, world, r, exact, exact ? NULL : &c, NULL));
            CU(cudaMemcpy(hpart + lo, (double*)partial[r] + lo, 8 * (size_t)(hi - lo), cudaMemcpyDeviceToHost));
            if (!exact) { if (c < 0) { printf("list overflow on rank %d\n", r); return 3; } CU(cudaMemcpy(hlist + total, list[r], 8 * (size_t)c, cudaMemcpyDeviceToHost)); total += c; }
        }
        double e_first = 0;
        for (int r = 0; r < world; ++r) {
            CU(cudaMemcpy(partial[r], hpart, 8 * (size_t)n, cudaMemcpyHostToDevice));
            if (!exact) { CU(cudaMemcpy(list[r], hlist, 8 * (size_t)total, cudaMemcpyHostToDevice)); unsigned long long t = (unsigned long long)total; CU(cudaMemcpy((unsigned long long*)state[r] + 5, &t, 8, cudaMemcpyHostToDevice)); }
            double e = 0; long long top = 0;
            CK(ssg_eps_shard_finish(cp[r], n, exact, &e, &top, NULL));
            if (r == 0) e_first = e;
            const int bad = !(fabs(e - eps0) <= 1e-13 * fabs(eps0)) || top != top0 || memcmp(&e, &e_first, 8) != 0;
            fails += bad;
            if (bad || r == 0) printf("eps %s rank %d: %.17g vs %.17g (top %lld vs %lld, list %lld): %s\n", exact ? "6-pass" : "3-pass", r, e, eps0, top, top0, total, bad ? "FAIL" : "ok");
        }
    }

    /* 3. DBSCAN */
    int64_t *d_lab0, *d_lab; int ncl0 = 0;
    CU(cudaMalloc((void**)&d_lab0, 8 * (size_t)n)); CU(cudaMalloc((void**)&d_lab, 8 * (size_t)n));
    CK(ssg_dbscan(cp0, d_final, SSG_F64, n, eps0, 4, d_lab0, &ncl0, NULL));
    int64_t* h_l0 = (int64_t*)malloc(8 * (size_t)n); int64_t* h_l = (int64_t*)malloc(8 * (size_t)n);
    CU(cudaMemcpy(h_l0, d_lab0, 8 * (size_t)n, cudaMemcpyDeviceToHost));
    int* hcnt = (int*)malloc(4 * (size_t)n);
    for (int r = 0; r < world; ++r) {
        const int lo = lo_of(n, world, r), hi = lo_of(n, world, r + 1);
        CK(ssg_dbscan_shard_count(cp[r], d_final + (size_t)lo * n, SSG_F64, n, lo, hi - lo, eps0, NULL));
        CU(cudaMemcpy(hcnt + lo, (int*)cnt[r] + lo, 4 * (size_t)(hi - lo), cudaMemcpyDeviceToHost));
    }
    long long total = 0;
    int* hn = NULL; int* hns = NULL;
    for (int r = 0; r < world; ++r) {
        const int lo = lo_of(n, world, r), hi = lo_of(n, world, r + 1);
        CU(cudaMemcpy(cnt[r], hcnt, 4 * (size_t)n, cudaMemcpyHostToDevice));
        CK(ssg_dbscan_shard_fill(cp[r], d_final + (size_t)lo * n, SSG_F64, n, lo, hi - lo, eps0, &total, NULL));
        if (!hn) { hn = (int*)malloc(4 * (size_t)(total + 1)); hns = (int*)calloc((size_t)(total + 1), 4); }
        CU(cudaMemcpy(hn, nbr[r], 4 * (size_t)total, cudaMemcpyDeviceToHost));
        for (long long e = 0; e < total; ++e) hns[e] += hn[e];
    }
    for (int r = 0; r < world; ++r) {
        int ncl = -1;
        CU(cudaMemcpy(nbr[r], hns, 4 * (size_t)total, cudaMemcpyHostToDevice));
        CK(ssg_dbscan_shard_label(cp[r], n, 4, d_lab, &ncl, NULL));
        CU(cudaMemcpy(h_l, d_lab, 8 * (size_t)n, cudaMemcpyDeviceToHost));
        const int bad = memcmp(h_l, h_l0, 8 * (size_t)n) != 0 || ncl != ncl0;
        fails += bad;
        if (bad || r == 0) printf("dbscan rank %d: %d clusters vs %d, %lld neighbour pairs: %s\n", r, ncl, ncl0, total, bad ? "FAIL" : "ok");
    }
    printf("%s (n=%d, world=%d simulated on one device, eps=%.6f, clusters=%d)\n", fails ? "SHARD_CHECK FAILED" : "SHARD_CHECK PASSED", n, world, eps0, ncl0);
    return fails ? 1 : 0;
}
