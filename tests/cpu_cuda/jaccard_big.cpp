// jaccard_big.cpp -- emulated-library check of jaccard_sparse_kernel against jaccard_final_kernel on synthetic expanded
// rows with n > 8192, i.e. more than 256 bitmap words: the word-rank scan of the sparse kernel then takes more than
// one pass of its block-wide loop, as it does at the production size (N = 16 702 -> 522 words).  Calls the library's
// internal launch wrappers directly (C++ symbols of libssg_emu.so).   usage: jaccard_big [n=9000]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <algorithm>

#include "cuda_runtime.h"
#include "ssg_b200.h"
#include "kernels.h"

static unsigned long long s_seed = 99;
static double frand() { s_seed = s_seed * 6364136223846793005ull + 1442695040888963407ull; return (double)(s_seed >> 11) / 9007199254740992.0; }

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 9000;
    const double lam = 0.1;
    std::vector<int> q_idx((size_t)n * SSG_VQ_STRIDE, 0), q_cnt(n), colcnt(n), colptr(n + 1), cursor(n), csc_row((size_t)n * SSG_VQ_STRIDE);
    std::vector<float> q_val((size_t)n * SSG_VQ_STRIDE, 0.f), vec(n);
    for (int i = 0; i < n; ++i) {
        vec[i] = (float)frand();
        std::vector<int> cols;
        const int c = 8 + (int)(frand() * 40);
        cols.push_back(i);
        for (int k = 0; k < c; ++k) {
            // mostly near i (clusters), sometimes far away (so that touched columns spread over many bitmap words)
            int m = frand() < 0.8 ? i + (int)(frand() * 60) - 30 : (int)(frand() * n);
            cols.push_back(std::min(std::max(m, 0), n - 1));
        }
        std::sort(cols.begin(), cols.end());
        cols.erase(std::unique(cols.begin(), cols.end()), cols.end());
        q_cnt[i] = (int)cols.size();
        float sum = 0.f;
        for (size_t k = 0; k < cols.size(); ++k) { q_idx[(size_t)i * SSG_VQ_STRIDE + k] = cols[k]; q_val[(size_t)i * SSG_VQ_STRIDE + k] = (float)(0.05 + frand()); sum += q_val[(size_t)i * SSG_VQ_STRIDE + k]; }
        for (size_t k = 0; k < cols.size(); ++k) q_val[(size_t)i * SSG_VQ_STRIDE + k] /= sum;
    }
    if (ssg::launch_csc_build(n, q_idx.data(), q_cnt.data(), colcnt.data(), colptr.data(), cursor.data(), csc_row.data(), nullptr)) { printf("csc_build: %s\n", ssg_last_error()); return 2; }
    std::vector<double> dense((size_t)n * n);
    if (ssg::launch_jaccard_final(n, 0, n, q_idx.data(), q_val.data(), q_cnt.data(), colptr.data(), csc_row.data(), vec.data(), lam, dense.data(), nullptr)) { printf("dense: %s\n", ssg_last_error()); return 2; }
    std::vector<int> sp_cnt(n), sp_rowptr(n + 1);
    if (ssg::launch_jaccard_sparse(n, q_idx.data(), q_val.data(), q_cnt.data(), colptr.data(), csc_row.data(), vec.data(), lam, nullptr, sp_cnt.data(), nullptr, nullptr, nullptr)) { printf("count: %s\n", ssg_last_error()); return 2; }
    sp_rowptr[0] = 0;
    for (int i = 0; i < n; ++i) sp_rowptr[i + 1] = sp_rowptr[i] + sp_cnt[i];
    const size_t nnz = (size_t)sp_rowptr[n];
    std::vector<int> sp_col(nnz + 1);
    std::vector<double> sp_val(nnz + 1);
    if (ssg::launch_jaccard_sparse(n, q_idx.data(), q_val.data(), q_cnt.data(), colptr.data(), csc_row.data(), vec.data(), lam, sp_rowptr.data(), sp_cnt.data(), sp_col.data(), sp_val.data(), nullptr)) { printf("fill: %s\n", ssg_last_error()); return 2; }
    const double thr = (double)(float)(1.0 - lam);
    size_t bad_val = 0, bad_order = 0, bad_out = 0, below = 0;
    std::vector<char> in(n);
    for (int i = 0; i < n; ++i) {
        std::fill(in.begin(), in.end(), 0);
        for (int e = sp_rowptr[i]; e < sp_rowptr[i + 1]; ++e) {
            const int m = sp_col[e];
            if (m < 0 || m >= n || (e > sp_rowptr[i] && sp_col[e - 1] >= m)) { ++bad_order; continue; }
            in[m] = 1;
            bad_val += memcmp(&sp_val[e], &dense[(size_t)i * n + m], 8) != 0;
            below += sp_val[e] < thr;
        }
        for (int m = 0; m < n; ++m) bad_out += !in[m] && !(dense[(size_t)i * n + m] >= thr);
    }
    const int bad = bad_val || bad_order || bad_out || !below;
    printf("n=%d (%d bitmap words), nnz %zu, below threshold %zu; value mismatches %zu, order %zu, outside-below-threshold %zu: %s\n",
           n, (n + 31) / 32, nnz, below, bad_val, bad_order, bad_out, bad ? "JACCARD_BIG FAILED" : "JACCARD_BIG PASSED");
    return bad ? 1 : 0;
}
