// C ABI of libssg_b200: error plumbing, device info, ssg_sqdist and the re-ranking plan
// (reid/rerank.py:27-127).  The eps / DBSCAN entry points live in cluster.cu, the embedding ones in
// embed.cu.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#include "common.cuh"
#include "kernels.h"

using namespace ssg;

static thread_local char g_err[1024] = "";

int ssg_set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char* ssg_last_error(void) { return g_err; }
extern "C" int ssg_version(void) { return 100; }

extern "C" int ssg_device_info(int device, int* n_devices, int* sm) {
    int n = 0;
    SSG_CUDA_TRY(cudaGetDeviceCount(&n));
    if (n_devices) *n_devices = n;
    if (sm) {
        cudaDeviceProp prop;
        SSG_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        *sm = prop.major * 10 + prop.minor;
    }
    return SSG_OK;
}

extern "C" int ssg_sqdist(const float* d_x, int nx, const float* d_y, int ny, int d, int mode, float* d_out,
                          size_t ldo, void* stream) {
    if (!d_x || !d_y || !d_out || nx < 0 || ny < 0 || d <= 0 || ldo < (size_t)ny)
        return ssg_set_error(SSG_ERR_INVALID, "sqdist: bad arguments");
    if (mode == SSG_DIST_EXACT) return launch_sqdist_exact(d_x, nx, d_y, ny, d, d_out, ldo, (cudaStream_t)stream);
    if (mode == SSG_DIST_TENSOR) return launch_sqdist_tensor(d_x, nx, d_y, ny, d, d_out, ldo, (cudaStream_t)stream);
    return ssg_set_error(SSG_ERR_INVALID, "sqdist: unknown mode %d", mode);
}

extern "C" int ssg_dot(const float* d_x, int nx, const float* d_y, int ny, int d, float* d_out, size_t ldo,
                       void* stream) {
    if (!d_x || !d_y || !d_out || nx < 0 || ny < 0 || d <= 0 || ldo < (size_t)ny)
        return ssg_set_error(SSG_ERR_INVALID, "dot: bad arguments");
    return launch_dot_exact(d_x, nx, d_y, ny, d, d_out, ldo, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------- rerank plan
struct ssg_rerank_plan {
    int device, n_max, ns_max, d;
    size_t bytes;
    size_t dmat_elems;           // capacity of the distance block (floats)
    float* dmat;
    float *rowmin, *rowmax, *vec, *scratch;
    int* rank; float* rank_val;
    int* v_idx; float* v_val; int* v_cnt;
    int* q_idx; float* q_val; int* q_cnt;
    int *colcnt, *colptr, *cursor, *csc_row;
    int* flagged;                // [4]: 0 rows that took the exact fallback (total), 1 src flags, 2 tgt flags
    // tensor distance mode (allocated on first use)
    bool tensor_ready;
    void *split_ta, *split_tb, *split_sb;        // bf16 [n,3d] / [ns,3d]
    float *norm_t, *norm_s, *norm_max;           // [n], [ns], [2]
    double* mean_partial; float* mean;           // [64*d], [d]  (centre of the target features)
    int* cand_idx; float *cand_val, *cand_exact; // [n,64]
    int *flag_src, *flag_tgt;                    // [n], [2n]
    float* fb_rows;                              // gathered feature rows for the fallback [FB_ROWS, d]
    float* fb_f32; int* fb_i32;                  // [FB_ROWS, 32]
    // lazily grown device I/O buffers for the host entry point
    float *io_src, *io_tgt; size_t io_src_bytes, io_tgt_bytes;
    double* io_final; size_t io_final_bytes;
    float* io_euclid; size_t io_euclid_bytes;
    // sparse form of final_dist (ssg_rerank_finish_sparse): CSR over the touched columns, grown on demand
    int *sp_cnt, *sp_rowptr, *sp_col; double* sp_val; size_t sp_cap; long long sp_nnz; double sp_threshold;
    // plain kNN-set re-ranking (ssg_rerank_plain): rows with a tied k-th neighbour and their exact fallback
    int *pl_flag_rows, *pl_flags, *pl_sel_idx; float *pl_sel_val, *pl_rows;
    double *lh_minsum, *lh_vec;  // re_ranking_lh: float64 source term
    int rank_cols;               // leading rank columns the last distance stage produced (k1 + 1 of that call)
    int last_n;
};

static int dalloc(void** p, size_t bytes, size_t* total) {
    SSG_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 16));
    if (total) *total += bytes;
    return SSG_OK;
}

extern "C" int ssg_rerank_plan_create(ssg_rerank_plan** out, int device, int n_max, int ns_max, int d) {
    if (!out || n_max <= 0 || ns_max <= 0 || d <= 0)
        return ssg_set_error(SSG_ERR_INVALID, "rerank_plan_create: bad arguments");
    SSG_ON_DEVICE(device);
    ssg_rerank_plan* p = new ssg_rerank_plan();
    memset(p, 0, sizeof(*p));
    p->device = device; p->n_max = n_max; p->ns_max = ns_max; p->d = d;
    const size_t n = (size_t)n_max;
    const size_t cols = (size_t)(n_max > ns_max ? n_max : ns_max);
    // distance block: whole matrix up to 4 GiB, else row blocks (at least 256 rows)
    size_t elems = n * cols;
    const size_t cap = (size_t)1 << 30;   // 2^30 floats = 4 GiB
    if (elems > cap) {
        size_t rows = cap / cols;
        if (rows < 256) rows = 256;
        elems = rows * cols;
    }
    p->dmat_elems = elems;
    int rc = SSG_OK;
#define A(ptr, nbytes) if (rc == SSG_OK) rc = dalloc((void**)&(ptr), (nbytes), &p->bytes)
    A(p->dmat, sizeof(float) * elems);
    A(p->rowmin, sizeof(float) * n);
    A(p->rowmax, sizeof(float) * n);
    A(p->vec, sizeof(float) * n);
    A(p->scratch, sizeof(float) * 4);
    A(p->rank, sizeof(int) * n * SSG_RANK_STRIDE);
    A(p->rank_val, sizeof(float) * n * SSG_RANK_STRIDE);
    A(p->v_idx, sizeof(int) * n * SSG_V_STRIDE);
    A(p->v_val, sizeof(float) * n * SSG_V_STRIDE);
    A(p->v_cnt, sizeof(int) * n);
    A(p->q_idx, sizeof(int) * n * SSG_VQ_STRIDE);
    A(p->q_val, sizeof(float) * n * SSG_VQ_STRIDE);
    A(p->q_cnt, sizeof(int) * n);
    A(p->colcnt, sizeof(int) * n);
    A(p->colptr, sizeof(int) * (n + 1));
    A(p->cursor, sizeof(int) * n);
    A(p->csc_row, sizeof(int) * n * SSG_VQ_STRIDE);
    A(p->flagged, sizeof(int) * 4);
#undef A
    if (rc != SSG_OK) { ssg_rerank_plan_destroy(p); return rc; }
    *out = p;
    return SSG_OK;
}

extern "C" int ssg_rerank_plan_destroy(ssg_rerank_plan* p) {
    if (!p) return SSG_OK;
    SsgDeviceGuard device_guard__(p->device);
    void* ptrs[] = {p->dmat, p->rowmin, p->rowmax, p->vec, p->scratch, p->rank, p->rank_val, p->v_idx,
                    p->v_val, p->v_cnt, p->q_idx, p->q_val, p->q_cnt, p->colcnt, p->colptr, p->cursor,
                    p->csc_row, p->flagged, p->io_src, p->io_tgt, p->io_final, p->io_euclid, p->split_ta,
                    p->split_tb, p->split_sb, p->norm_t, p->norm_s, p->norm_max, p->cand_idx, p->cand_val,
                    p->cand_exact, p->flag_src, p->flag_tgt, p->fb_rows, p->fb_f32, p->fb_i32, p->mean_partial,
                    p->mean, p->sp_cnt, p->sp_rowptr, p->sp_col, p->sp_val, p->pl_flag_rows, p->pl_flags, p->pl_sel_idx,
                    p->pl_sel_val, p->pl_rows, p->lh_minsum, p->lh_vec};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete p;
    return SSG_OK;
}

extern "C" size_t ssg_rerank_plan_bytes(const ssg_rerank_plan* p) {
    return p ? p->bytes + p->io_src_bytes + p->io_tgt_bytes + p->io_final_bytes + p->io_euclid_bytes +
                   p->sp_cap * (sizeof(int) + sizeof(double)) : 0;
}

// capacity of an expanded row (rerank.py:94-98): the union of k2 k-reciprocal rows, each at most
// (k1+1) + (k1+1) * (round(k1/2)+1) columns (rerank.py:76-90) -- it must fit the SSG_VQ_STRIDE slot that query_expand
// writes and jaccard_row stages in shared memory (k2 <= 6 at k1 = 20; larger k2 only with smaller k1)
static int expanded_row_fits(int k1, int k2) {
    const int k1p = k1 + 1, khp = (int)rint(k1 / 2.0) + 1;
    const long long bound = (long long)k2 * (k1p + (long long)k1p * khp);
    if (bound > SSG_VQ_STRIDE)
        return ssg_set_error(SSG_ERR_INVALID, "rerank: k1=%d k2=%d may expand a row to %lld columns, the plan holds %d",
                             k1, k2, bound, SSG_VQ_STRIDE);
    return SSG_OK;
}

static int check_run_args(ssg_rerank_plan* p, const void* src, int ns, const void* tgt, int n, int d,
                          int k1, int k2) {
    if (!p || !src || !tgt) return ssg_set_error(SSG_ERR_INVALID, "rerank: null argument");
    if (n <= 0 || ns <= 0 || n > p->n_max || ns > p->ns_max || d != p->d)
        return ssg_set_error(SSG_ERR_INVALID, "rerank: shape (n=%d, ns=%d, d=%d) exceeds plan (%d, %d, %d)",
                             n, ns, d, p->n_max, p->ns_max, p->d);
    if (k1 < 1 || k1 > 31 || k2 < 1 || k2 > 8)
        return ssg_set_error(SSG_ERR_INVALID, "rerank: k1=%d k2=%d out of range", k1, k2);
    SSG_TRY(expanded_row_fits(k1, k2));
    return SSG_OK;
}

// squared distances of a block of target rows against all of Y, then what the caller needs from it
static int dist_block(ssg_rerank_plan* p, const float* X, int rows, const float* Y, int ny, int d,
                      int dist_mode, cudaStream_t st) {
    if (dist_mode == SSG_DIST_EXACT) return launch_sqdist_exact(X, rows, Y, ny, d, p->dmat, (size_t)ny, st);
    return ssg_set_error(SSG_ERR_UNSUPPORTED, "rerank: dist_mode %d not available in this build", dist_mode);
}

constexpr int CAND_STRIDE = 64;   // candidate table row stride
constexpr int CAND_K = 32;        // nearest-neighbour candidates re-scored per row (k1+1 <= 32 needed)
constexpr int CAND_EXT = 8;       // candidates for the row minimum / maximum
constexpr int FB_ROWS = 1024;     // rows per exact-fallback batch
// Certified error of the tensor-core squared distance, relative to (|x|^2 + max|y|^2):
//   * the bf16 hi/lo split drops products below 3 * 2^-18 |x||y|                      -> 1.15e-5 (x2 for -2*dot, /2 for
//     |x||y| <= (|x|^2+|y|^2)/2);
//   * the tensor core adds each 16-wide partial product to the fp32 accumulator with truncation: 3d/16 updates of at
//     most 2^-23 |x||y| each                                                          -> 2.24e-8 * d.
//   * inside ssg_rerank_run both sets are first centred on the target mean in fp32 (x' = fl(x - mu)): distances are
//     unchanged, |x'| is the spread of the data rather than its norm, and the rounding of the subtraction moves d^2 by
//     at most 2^-22 (|x'|^2+|y'|^2)                                                    -> 2.4e-7.
// The bound is taken relative to the CENTRED norms.  Measured maximum at d = 2048, unit norms, no centring:
// 2.05e-5 of (|x|^2+|y|^2) (tests/test_gpu_tensor.py), bound 5.7e-5.
static inline float tensor_eps_rel(int d) { return 1.18e-5f + 2.24e-8f * (float)d; }

static int ensure_tensor_buffers(ssg_rerank_plan* p) {
    if (p->tensor_ready) return SSG_OK;
    const size_t n = (size_t)p->n_max, ns = (size_t)p->ns_max, d = (size_t)p->d;
    int rc = SSG_OK;
#define A(ptr, nbytes) if (rc == SSG_OK) rc = dalloc((void**)&(ptr), (nbytes), &p->bytes)
    A(p->split_ta, n * 3 * d * 2);
    A(p->split_tb, n * 3 * d * 2);
    A(p->split_sb, ns * 3 * d * 2);
    A(p->norm_t, sizeof(float) * n);
    A(p->norm_s, sizeof(float) * ns);
    A(p->norm_max, sizeof(float) * 2);
    A(p->mean_partial, sizeof(double) * 64 * d);
    A(p->mean, sizeof(float) * d);
    A(p->cand_idx, sizeof(int) * n * CAND_STRIDE);
    A(p->cand_val, sizeof(float) * n * CAND_STRIDE);
    A(p->cand_exact, sizeof(float) * n * CAND_STRIDE);
    A(p->flag_src, sizeof(int) * n);
    A(p->flag_tgt, sizeof(int) * 2 * n);
    A(p->fb_rows, sizeof(float) * FB_ROWS * d);
    A(p->fb_f32, sizeof(float) * FB_ROWS * SSG_RANK_STRIDE);
    A(p->fb_i32, sizeof(int) * FB_ROWS * SSG_RANK_STRIDE);
#undef A
    if (rc == SSG_OK) p->tensor_ready = true;
    return rc;
}

// Tensor distance mode: bf16x3 tcgen05 GEMM for the bulk, exact float64 re-scoring of the candidates, exact
// fallback for the rows whose result the error bound cannot certify.  Synchronises the stream once (flag counts).
static int distance_stages_tensor(ssg_rerank_plan* p, const float* d_src, int ns, const float* d_tgt, int n, int d,
                                  int k1, float* d_euclid, bool want_rank, int row_begin, int row_end,
                                  cudaStream_t st) {
    if (d % 8) return ssg_set_error(SSG_ERR_INVALID, "tensor distance mode needs d %% 8 == 0 (d=%d)", d);
    SSG_TRY(ensure_tensor_buffers(p));
    const int k1p = k1 + 1;
    const int k3 = 3 * d;
    const float TENSOR_EPS_REL = tensor_eps_rel(d);
    int* cnt_src = p->flagged + 1;
    int* cnt_tgt = p->flagged + 2;
    SSG_CUDA_TRY(cudaMemsetAsync(p->flagged, 0, sizeof(int) * 4, st));
    // centre both sets on the target mean (distances are translation invariant; see tensor_eps_rel)
    { SSG_PROF("split_bf16x3", st); SSG_TRY(launch_col_mean(d_tgt, n, d, p->mean_partial, p->mean, st)); }
    { SSG_PROF("split_bf16x3", st); SSG_TRY(launch_split_bf16x3(d_tgt, n, d, 0, p->mean, p->split_ta, p->norm_t, st)); }
    { SSG_PROF("split_bf16x3", st); SSG_TRY(launch_split_bf16x3(d_tgt, n, d, 1, p->mean, p->split_tb, nullptr, st)); }
    { SSG_PROF("split_bf16x3", st); SSG_TRY(launch_split_bf16x3(d_src, ns, d, 1, p->mean, p->split_sb, p->norm_s, st)); }
    SSG_TRY(launch_vec_max(p->norm_t, n, p->norm_max, st));
    SSG_TRY(launch_vec_max(p->norm_s, ns, p->norm_max + 1, st));
    const char* ta = (const char*)p->split_ta;
    // (i) source term: row minimum over the sources
    {
        const int rows_blk = (int)(p->dmat_elems / (size_t)ns < (size_t)n ? p->dmat_elems / (size_t)ns : (size_t)n);
        const int ke = ns < CAND_EXT ? ns : CAND_EXT;
        for (int r0 = row_begin; r0 < row_end; r0 += rows_blk) {
            const int rows = row_end - r0 < rows_blk ? row_end - r0 : rows_blk;
            { SSG_PROF("gemm_dist_tc", st); SSG_TRY(launch_gemm_dist(ta + (size_t)r0 * k3 * 2, p->norm_t + r0, rows, p->split_sb, p->norm_s, ns, k3,
                                     p->dmat, (size_t)ns, st)); }
            { SSG_PROF("row_select", st); SSG_TRY(launch_row_select(p->dmat, (size_t)ns, rows, ns, nullptr, ke, false, p->cand_idx, p->cand_val,
                                      CAND_STRIDE, p->cursor, st)); }
            { SSG_PROF("pair_exact", st); SSG_TRY(launch_pair_exact(d_tgt + (size_t)r0 * d, rows, d_src, d, p->cand_idx, CAND_STRIDE, nullptr, ke,
                                      p->cand_exact, CAND_STRIDE, st)); }
            { SSG_PROF("cand_reduce", st); SSG_TRY(launch_cand_reduce(rows, ns, ke, false, p->cand_exact, p->cand_val, CAND_STRIDE, p->norm_t + r0,
                                       p->norm_max + 1, TENSOR_EPS_REL, p->rowmin + r0, cnt_src, p->flag_src, r0, st)); }
        }
    }
    // (ii)-(iv) target-target distances: row maximum and the leading rank columns
    {
        const int rows_blk = (int)(p->dmat_elems / (size_t)n < (size_t)n ? p->dmat_elems / (size_t)n : (size_t)n);
        const int ke = n < CAND_EXT ? n : CAND_EXT;
        const int kc = n < CAND_K ? n : CAND_K;
        for (int r0 = row_begin; r0 < row_end; r0 += rows_blk) {
            const int rows = row_end - r0 < rows_blk ? row_end - r0 : rows_blk;
            // when one launch covers the whole target x target matrix, compute the tiles of the upper triangle only and
            // mirror them (SSG_DIST_SYM=0 disables; round 2, B200: 18.2 -> 15.0 ms per cycle for the distance GEMMs,
            // outputs bit-equal to the exact mode: tests/test_gpu_next_variants.py)
            static int dist_sym = -1;
            if (dist_sym < 0) { const char* e = getenv("SSG_DIST_SYM"); dist_sym = e ? atoi(e) : 1; }
            const int sym = (dist_sym && r0 == 0 && rows == n) ? 1 : 0;
            { SSG_PROF("gemm_dist_tc", st); SSG_TRY(launch_gemm_dist(ta + (size_t)r0 * k3 * 2, p->norm_t + r0, rows, p->split_tb, p->norm_t, n, k3,
                                     p->dmat, (size_t)n, st, sym)); }
            { SSG_PROF("row_select", st); SSG_TRY(launch_row_select(p->dmat, (size_t)n, rows, n, nullptr, ke, true, p->cand_idx, p->cand_val,
                                      CAND_STRIDE, p->cursor, st)); }
            { SSG_PROF("pair_exact", st); SSG_TRY(launch_pair_exact(d_tgt + (size_t)r0 * d, rows, d_tgt, d, p->cand_idx, CAND_STRIDE, nullptr, ke,
                                      p->cand_exact, CAND_STRIDE, st)); }
            { SSG_PROF("cand_reduce", st); SSG_TRY(launch_cand_reduce(rows, n, ke, true, p->cand_exact, p->cand_val, CAND_STRIDE, p->norm_t + r0,
                                       p->norm_max, TENSOR_EPS_REL, p->rowmax + r0, cnt_tgt, p->flag_tgt, r0, st)); }
            if (want_rank) {
                { SSG_PROF("row_select", st); SSG_TRY(launch_row_select(p->dmat, (size_t)n, rows, n, nullptr, kc, false, p->cand_idx, p->cand_val,
                                          CAND_STRIDE, p->cursor, st)); }
                { SSG_PROF("pair_exact", st); SSG_TRY(launch_pair_exact(d_tgt + (size_t)r0 * d, rows, d_tgt, d, p->cand_idx, CAND_STRIDE, nullptr,
                                          kc, p->cand_exact, CAND_STRIDE, st)); }
                { SSG_PROF("rank_finalize", st); SSG_TRY(launch_rank_finalize(rows, n, kc, k1p < kc ? k1p : kc, p->cand_idx, p->cand_val,
                                             p->cand_exact, CAND_STRIDE, p->rowmax + r0, p->norm_t + r0, p->norm_max,
                                             TENSOR_EPS_REL, p->rank + (size_t)r0 * SSG_RANK_STRIDE,
                                             p->rank_val + (size_t)r0 * SSG_RANK_STRIDE, cnt_tgt, p->flag_tgt, r0, st)); }
            }
            if (d_euclid)
                SSG_CUDA_TRY(cudaMemcpyAsync(d_euclid + (size_t)r0 * n, p->dmat, sizeof(float) * (size_t)rows * n,
                                             cudaMemcpyDeviceToDevice, st));
        }
    }
    // exact fallback for the flagged rows
    int h_flags[4];
    SSG_CUDA_TRY(cudaMemcpyAsync(h_flags, p->flagged, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaStreamSynchronize(st));
    const int nsrc = h_flags[1], ntgt = h_flags[2];
    for (int f0 = 0; f0 < nsrc; f0 += FB_ROWS) {
        int cnt = nsrc - f0 < FB_ROWS ? nsrc - f0 : FB_ROWS;
        const int cap = (int)(p->dmat_elems / (size_t)ns);
        if (cnt > cap) cnt = cap;
        SSG_TRY(launch_gather_rows(d_tgt, d, p->flag_src + f0, cnt, p->fb_rows, st));
        { SSG_PROF("sqdist_exact", st); SSG_TRY(launch_sqdist_exact(p->fb_rows, cnt, d_src, ns, d, p->dmat, (size_t)ns, st)); }
        { SSG_PROF("row_minmax", st); SSG_TRY(launch_row_minmax(p->dmat, (size_t)ns, cnt, ns, p->fb_f32, nullptr, st)); }
        SSG_TRY(launch_scatter_f32(p->fb_f32, cnt, 1, 1, p->flag_src + f0, p->rowmin, 1, st));
        if (cnt < FB_ROWS && f0 + cnt < nsrc) f0 -= (FB_ROWS - cnt);   // capacity-limited batch: continue after it
    }
    for (int f0 = 0; f0 < ntgt; f0 += FB_ROWS) {
        int cnt = ntgt - f0 < FB_ROWS ? ntgt - f0 : FB_ROWS;
        const int cap = (int)(p->dmat_elems / (size_t)n);
        if (cnt > cap) cnt = cap;
        SSG_TRY(launch_gather_rows(d_tgt, d, p->flag_tgt + f0, cnt, p->fb_rows, st));
        { SSG_PROF("sqdist_exact", st); SSG_TRY(launch_sqdist_exact(p->fb_rows, cnt, d_tgt, n, d, p->dmat, (size_t)n, st)); }
        float* fmax = p->cand_val;   // reuse: [cnt] exact row maxima
        { SSG_PROF("row_minmax", st); SSG_TRY(launch_row_minmax(p->dmat, (size_t)n, cnt, n, nullptr, fmax, st)); }
        SSG_TRY(launch_scatter_f32(fmax, cnt, 1, 1, p->flag_tgt + f0, p->rowmax, 1, st));
        if (want_rank) {
            { SSG_PROF("row_select", st); SSG_TRY(launch_row_select(p->dmat, (size_t)n, cnt, n, fmax, k1p, false, p->fb_i32, p->fb_f32,
                                      SSG_RANK_STRIDE, p->cursor, st)); }
            SSG_TRY(launch_scatter_f32((const float*)p->fb_i32, cnt, k1p, SSG_RANK_STRIDE, p->flag_tgt + f0,
                                       (float*)p->rank, SSG_RANK_STRIDE, st));
            SSG_TRY(launch_scatter_f32(p->fb_f32, cnt, k1p, SSG_RANK_STRIDE, p->flag_tgt + f0, p->rank_val,
                                       SSG_RANK_STRIDE, st));
        }
        if (d_euclid)   // exact rows replace the approximate ones
            SSG_TRY(launch_scatter_f32(p->dmat, cnt, n, n, p->flag_tgt + f0, d_euclid, n, st));
        if (cnt < FB_ROWS && f0 + cnt < ntgt) f0 -= (FB_ROWS - cnt);
    }
    h_flags[0] = nsrc + ntgt;
    SSG_CUDA_TRY(cudaMemcpyAsync(p->flagged, h_flags, sizeof(int), cudaMemcpyHostToDevice, st));
    return SSG_OK;
}

// stages (i)-(iv): source vector, squared distance, row normaliser, leading k1+1 rank columns
static int distance_stages_exact(ssg_rerank_plan* p, const float* d_src, int ns, const float* d_tgt, int n, int d,
                           int k1, int dist_mode, float* d_euclid, bool want_rank, int row_begin, int row_end,
                                 cudaStream_t st) {
    const int k1p = k1 + 1;
    // (i) rerank.py:36-40
    {
        const int rows_blk = (int)(p->dmat_elems / (size_t)ns < (size_t)n ? p->dmat_elems / (size_t)ns : (size_t)n);
        for (int r0 = row_begin; r0 < row_end; r0 += rows_blk) {
            const int rows = row_end - r0 < rows_blk ? row_end - r0 : rows_blk;
            { SSG_PROF("sqdist_exact", st); SSG_TRY(dist_block(p, d_tgt + (size_t)r0 * d, rows, d_src, ns, d, dist_mode, st)); }
            { SSG_PROF("row_minmax", st); SSG_TRY(launch_row_minmax(p->dmat, (size_t)ns, rows, ns, p->rowmin + r0, nullptr, st)); }
        }
    }
    // (ii)-(iv) rerank.py:61-70
    {
        const int rows_blk = (int)(p->dmat_elems / (size_t)n < (size_t)n ? p->dmat_elems / (size_t)n : (size_t)n);
        for (int r0 = row_begin; r0 < row_end; r0 += rows_blk) {
            const int rows = row_end - r0 < rows_blk ? row_end - r0 : rows_blk;
            { SSG_PROF("sqdist_exact", st); SSG_TRY(dist_block(p, d_tgt + (size_t)r0 * d, rows, d_tgt, n, d, dist_mode, st)); }
            { SSG_PROF("row_minmax", st); SSG_TRY(launch_row_minmax(p->dmat, (size_t)n, rows, n, nullptr, p->rowmax + r0, st)); }
            if (want_rank)
                { SSG_PROF("row_select", st); SSG_TRY(launch_row_select(p->dmat, (size_t)n, rows, n, p->rowmax + r0, k1p, false,
                                          p->rank + (size_t)r0 * SSG_RANK_STRIDE,
                                          p->rank_val + (size_t)r0 * SSG_RANK_STRIDE, SSG_RANK_STRIDE, p->cursor, st)); }
            if (d_euclid)
                SSG_CUDA_TRY(cudaMemcpyAsync(d_euclid + (size_t)r0 * n, p->dmat, sizeof(float) * (size_t)rows * n,
                                             cudaMemcpyDeviceToDevice, st));
        }
    }
    return SSG_OK;
}

// rows [row_begin, row_end) of: rowmin (source term, before exp/normalisation), rowmax, rank, rank_val
// (d_euclid, when given, is indexed by the absolute row)
static int distance_stages(ssg_rerank_plan* p, const float* d_src, int ns, const float* d_tgt, int n, int d,
                           int k1, int dist_mode, float* d_euclid, bool want_rank, int row_begin, int row_end,
                           cudaStream_t st) {
    if (want_rank) p->rank_cols = k1 + 1;
    if (dist_mode == SSG_DIST_EXACT)
        return distance_stages_exact(p, d_src, ns, d_tgt, n, d, k1, dist_mode, d_euclid, want_rank, row_begin, row_end, st);
    if (dist_mode == SSG_DIST_TENSOR)
        return distance_stages_tensor(p, d_src, ns, d_tgt, n, d, k1, d_euclid, want_rank, row_begin, row_end, st);
    return ssg_set_error(SSG_ERR_INVALID, "rerank: unknown dist_mode %d", dist_mode);
}

extern "C" int ssg_rerank_distance_rows(ssg_rerank_plan* p, const float* d_src, int ns, const float* d_tgt, int n,
                                        int d, int k1, int dist_mode, int row0, int rows, float* d_euclid,
                                        void* stream) {
    SSG_TRY(check_run_args(p, d_src, ns, d_tgt, n, d, k1, 1));
    if (row0 < 0 || rows < 0 || row0 + rows > n)
        return ssg_set_error(SSG_ERR_INVALID, "rerank_distance_rows: rows [%d,%d) outside [0,%d)", row0, row0 + rows, n);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    SSG_CUDA_TRY(cudaMemsetAsync(p->flagged, 0, sizeof(int) * 4, st));
    if (rows == 0) return SSG_OK;
    return distance_stages(p, d_src, ns, d_tgt, n, d, k1, dist_mode, d_euclid, true, row0, row0 + rows, st);
}

extern "C" int ssg_rerank_tables(ssg_rerank_plan* p, float** d_rowmin, float** d_rowmax, int** d_rank,
                                 float** d_rank_val) {
    if (!p) return ssg_set_error(SSG_ERR_INVALID, "rerank_tables: null plan");
    if (d_rowmin) *d_rowmin = p->rowmin;
    if (d_rowmax) *d_rowmax = p->rowmax;
    if (d_rank) *d_rank = p->rank;
    if (d_rank_val) *d_rank_val = p->rank_val;
    return SSG_OK;
}

// Everything after the tables are complete: the sparse stages for all n rows (they read other rows' rank lists and V
// rows, rerank.py:77,97,102), then rows [row0, row0+rows) of final_dist into d_final (which points at row row0).
static int finish_sparse_stages(ssg_rerank_plan* p, const float* d_tgt, int n, int d, int k1, int k2, cudaStream_t st);

static int finish_rows(ssg_rerank_plan* p, const float* d_tgt, int n, int d, int k1, int k2, double lambda_value,
                       int row0, int rows, double* d_final, void* stream) {
    SSG_TRY(check_run_args(p, d_tgt, 1, d_tgt, n, d, k1, k2));
    if (!d_final) return ssg_set_error(SSG_ERR_INVALID, "rerank: d_final is null");
    if (row0 < 0 || rows < 0 || row0 + rows > n)
        return ssg_set_error(SSG_ERR_INVALID, "rerank_finish: rows [%d,%d) outside [0,%d)", row0, row0 + rows, n);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    SSG_TRY(finish_sparse_stages(p, d_tgt, n, d, k1, k2, st));
    { SSG_PROF("jaccard_final", st); SSG_TRY(launch_jaccard_final(n, row0, rows, p->q_idx, p->q_val, p->q_cnt, p->colptr, p->csc_row, p->vec,
                                 lambda_value, d_final, st)); }
    p->last_n = n;
    return SSG_OK;
}

// stages (i, tail) and (v)-(vii, inverted index): source vector, k-reciprocal rows, query expansion, CSC
static int finish_sparse_stages(ssg_rerank_plan* p, const float* d_tgt, int n, int d, int k1, int k2, cudaStream_t st) {
    // rerank.py:77,97 slice initial_rank[:, :k1+1] and [:, :k2]; the table holds the columns of the last distance stage
    if (k1 + 1 > p->rank_cols || k2 > p->rank_cols)
        return ssg_set_error(SSG_ERR_INVALID, "rerank_finish: k1=%d / k2=%d need %d rank columns, the distance stage produced "
                             "%d (run it with k1 >= max(k1, k2 - 1))", k1, k2, (k1 + 1 > k2 ? k1 + 1 : k2), p->rank_cols);
    const int k1p = k1 + 1;
    const int khp = (int)rint(k1 / 2.0) + 1;   // int(np.around(k1/2)) + 1, rerank.py:83
    // (i, tail) rerank.py:38-40: v = 1 - exp(-rowmin); v /= max(v)
    { SSG_PROF("source_vector", st); SSG_TRY(launch_source_vector(p->rowmin, n, p->vec, p->scratch, st)); }
    // (v) rerank.py:74-92
    { SSG_PROF("krecip_build", st); SSG_TRY(launch_krecip_build(p->rank, n, k1p, khp, p->v_idx, p->v_cnt, st)); }
    {
        // entries of R*(i) inside row i's own rank columns already have their exact normalised distance; only the
        // expansion entries are re-scored (scratch: the not-yet-used expanded-row buffers, each [n, VQ_STRIDE])
        int* todo_idx = p->q_idx;
        int* todo_slot = p->csc_row;
        float* todo_od = p->q_val;
        int* todo_cnt = p->q_cnt;
        { SSG_PROF("krecip_build", st); SSG_TRY(launch_krecip_classify(p->rank, p->rank_val, n, k1p, p->v_idx, p->v_cnt, p->v_val, todo_idx, todo_slot, todo_cnt, st)); }
        { SSG_PROF("pair_exact", st); SSG_TRY(launch_pair_exact(d_tgt, n, d_tgt, d, todo_idx, SSG_V_STRIDE, todo_cnt, 0, todo_od, SSG_V_STRIDE, st)); }
        { SSG_PROF("krecip_build", st); SSG_TRY(launch_krecip_scatter(p->rowmax, n, todo_od, todo_slot, todo_cnt, p->v_val, st)); }
    }
    { SSG_PROF("krecip_weights", st); SSG_TRY(launch_krecip_weights(p->rowmax, n, p->v_cnt, p->v_val, 1, st)); }
    // (vi) rerank.py:94-98  (k2 == 1: V is used as it is)
    if (k2 != 1) {
        { SSG_PROF("query_expand", st); SSG_TRY(launch_query_expand(p->rank, n, k2, p->v_idx, p->v_val, p->v_cnt, p->q_idx, p->q_val, p->q_cnt, st)); }
    } else {
        SSG_CUDA_TRY(cudaMemcpy2DAsync(p->q_idx, sizeof(int) * SSG_VQ_STRIDE, p->v_idx, sizeof(int) * SSG_V_STRIDE,
                                       sizeof(int) * SSG_V_STRIDE, n, cudaMemcpyDeviceToDevice, st));
        SSG_CUDA_TRY(cudaMemcpy2DAsync(p->q_val, sizeof(float) * SSG_VQ_STRIDE, p->v_val, sizeof(float) * SSG_V_STRIDE,
                                       sizeof(float) * SSG_V_STRIDE, n, cudaMemcpyDeviceToDevice, st));
        SSG_CUDA_TRY(cudaMemcpyAsync(p->q_cnt, p->v_cnt, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    }
    // (vii)-(viii) rerank.py:101-122
    { SSG_PROF("csc_build", st); SSG_TRY(launch_csc_build(n, p->q_idx, p->q_cnt, p->colcnt, p->colptr, p->cursor, p->csc_row, st)); }
    return SSG_OK;
}

extern "C" int ssg_rerank_finish(ssg_rerank_plan* p, const float* d_tgt, int n, int d, int k1, int k2,
                                 double lambda_value, double* d_final, void* stream) {
    return finish_rows(p, d_tgt, n, d, k1, k2, lambda_value, 0, n, d_final, stream);
}

extern "C" int ssg_rerank_finish_rows(ssg_rerank_plan* p, const float* d_tgt, int n, int d, int k1, int k2,
                                      double lambda_value, int row0, int rows, double* d_final_rows, void* stream) {
    return finish_rows(p, d_tgt, n, d, k1, k2, lambda_value, row0, rows, d_final_rows, stream);
}

extern "C" int ssg_rerank_run(ssg_rerank_plan* p, const float* d_src, int ns, const float* d_tgt, int n, int d,
                              int k1, int k2, double lambda_value, int dist_mode, double* d_final,
                              float* d_euclid, void* stream) {
    SSG_TRY(check_run_args(p, d_src, ns, d_tgt, n, d, k1, k2));
    if (!d_final) return ssg_set_error(SSG_ERR_INVALID, "rerank: d_final is null");
    // k2 > k1 + 1 (legal in the reference, whose argsort keeps every column): make the table wide enough for both
    const int k1d = k1 > k2 - 1 ? k1 : k2 - 1;
    SSG_TRY(ssg_rerank_distance_rows(p, d_src, ns, d_tgt, n, d, k1d, dist_mode, 0, n, d_euclid, stream));
    return ssg_rerank_finish(p, d_tgt, n, d, k1, k2, lambda_value, d_final, stream);
}

// final_dist in sparse form (see rerank.cu, jaccard_sparse_kernel).  Synchronises the stream once (the row lengths are
// data dependent; the CSR buffers grow on demand).
extern "C" int ssg_rerank_finish_sparse(ssg_rerank_plan* p, const float* d_tgt, int n, int d, int k1, int k2,
                                        double lambda_value, long long* h_nnz, void* stream) {
    SSG_TRY(check_run_args(p, d_tgt, 1, d_tgt, n, d, k1, k2));
    if (!(lambda_value >= 0.0 && lambda_value < 1.0))
        return ssg_set_error(SSG_ERR_INVALID, "rerank_finish_sparse: lambda_value %g outside [0, 1)", lambda_value);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!p->sp_cnt) {
        SSG_TRY(dalloc((void**)&p->sp_cnt, sizeof(int) * (size_t)p->n_max, &p->bytes));
        SSG_TRY(dalloc((void**)&p->sp_rowptr, sizeof(int) * ((size_t)p->n_max + 1), &p->bytes));
    }
    SSG_TRY(finish_sparse_stages(p, d_tgt, n, d, k1, k2, st));
    { SSG_PROF("jaccard_sparse", st); SSG_TRY(launch_jaccard_sparse(n, p->q_idx, p->q_val, p->q_cnt, p->colptr, p->csc_row, p->vec, lambda_value,
                                   nullptr, p->sp_cnt, nullptr, nullptr, st)); }
    SSG_TRY(launch_exclusive_scan_i32(p->sp_cnt, p->sp_rowptr, n, st));
    int total = 0;
    SSG_CUDA_TRY(cudaMemcpyAsync(&total, p->sp_rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaStreamSynchronize(st));
    if (total < 0) return ssg_set_error(SSG_ERR_CAPACITY, "rerank_finish_sparse: more than 2^31 touched pairs; use the dense form");
    if ((size_t)total > p->sp_cap) {
        if (p->sp_col) cudaFree(p->sp_col);
        if (p->sp_val) cudaFree(p->sp_val);
        p->sp_col = nullptr; p->sp_val = nullptr; p->sp_cap = 0;
        const size_t cap = (size_t)total + (size_t)total / 4 + 1024;
        SSG_CUDA_TRY(cudaMalloc((void**)&p->sp_col, sizeof(int) * cap));
        SSG_CUDA_TRY(cudaMalloc((void**)&p->sp_val, sizeof(double) * cap));
        p->sp_cap = cap;
    }
    { SSG_PROF("jaccard_sparse", st); SSG_TRY(launch_jaccard_sparse(n, p->q_idx, p->q_val, p->q_cnt, p->colptr, p->csc_row, p->vec, lambda_value,
                                   p->sp_rowptr, p->sp_cnt, p->sp_col, p->sp_val, st)); }
    p->sp_nnz = total;
    p->sp_threshold = (double)(float)(1.0 - lambda_value);      // every entry NOT in the CSR is >= this
    p->last_n = n;
    if (h_nnz) *h_nnz = total;
    return SSG_OK;
}

extern "C" int ssg_rerank_sparse_view(ssg_rerank_plan* p, int** d_rowptr, int** d_col, double** d_val, long long* nnz,
                                      double* threshold) {
    if (!p || !p->sp_rowptr) return ssg_set_error(SSG_ERR_INVALID, "rerank_sparse_view: no sparse result in this plan");
    if (d_rowptr) *d_rowptr = p->sp_rowptr;
    if (d_col) *d_col = p->sp_col;
    if (d_val) *d_val = p->sp_val;
    if (nnz) *nnz = p->sp_nnz;
    if (threshold) *threshold = p->sp_threshold;
    return SSG_OK;
}

// reid/rerank_plain.py:125-178 re_ranking(input_feature_source, input_feature, k=20, lambda_value=0.1): kNN-set Jaccard
// distance + the source term of reid/rerank.py.  d_final: [n,n] float64.  Synchronises the stream (flag counts).
extern "C" int ssg_rerank_plain(ssg_rerank_plan* p, const float* d_src, int ns, const float* d_tgt, int n, int d, int k,
                                double lambda_value, int dist_mode, double* d_final, void* stream) {
    SSG_TRY(check_run_args(p, d_src, ns, d_tgt, n, d, k, 1));
    if (!d_final) return ssg_set_error(SSG_ERR_INVALID, "rerank_plain: d_final is null");
    if (n < k) return ssg_set_error(SSG_ERR_INVALID, "rerank_plain: k=%d exceeds the %d targets (np.partition raises, "
                                    "rerank_plain.py:167)", k, n);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int PL_ROWS = 256;                 // rows per exact-fallback batch
    if (!p->pl_flags) {
        SSG_TRY(dalloc((void**)&p->pl_flag_rows, sizeof(int) * (size_t)p->n_max, &p->bytes));
        SSG_TRY(dalloc((void**)&p->pl_flags, sizeof(int) * 4, &p->bytes));
        SSG_TRY(dalloc((void**)&p->pl_sel_idx, sizeof(int) * PL_ROWS * SSG_RANK_STRIDE, &p->bytes));
        SSG_TRY(dalloc((void**)&p->pl_sel_val, sizeof(float) * PL_ROWS * SSG_RANK_STRIDE, &p->bytes));
        SSG_TRY(dalloc((void**)&p->pl_rows, sizeof(float) * (size_t)PL_ROWS * p->d, &p->bytes));
    }
    // distance stages with k1 = k: source row minimum, row maximum, k+1 leading rank columns
    SSG_TRY(distance_stages(p, d_src, ns, d_tgt, n, d, k, dist_mode, nullptr, true, 0, n, st));
    { SSG_PROF("source_vector", st); SSG_TRY(launch_source_vector(p->rowmin, n, p->vec, p->scratch, st)); }
    SSG_CUDA_TRY(cudaMemsetAsync(p->pl_flags, 0, sizeof(int) * 4, st));
    { SSG_PROF("knn_sets", st); SSG_TRY(launch_knn_sets(p->rank, p->rank_val, n, k, p->q_idx, p->q_cnt, p->pl_flag_rows, p->pl_flags, st)); }
    int h_flags[4];
    SSG_CUDA_TRY(cudaMemcpyAsync(h_flags, p->pl_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaStreamSynchronize(st));
    const int nflag = h_flags[0];
    const int cap_rows = (int)(p->dmat_elems / (size_t)n < (size_t)PL_ROWS ? p->dmat_elems / (size_t)n : (size_t)PL_ROWS);
    for (int f0 = 0; f0 < nflag; f0 += cap_rows) {
        const int cnt = nflag - f0 < cap_rows ? nflag - f0 : cap_rows;
        SSG_TRY(launch_gather_rows(d_tgt, d, p->pl_flag_rows + f0, cnt, p->pl_rows, st));
        { SSG_PROF("sqdist_exact", st); SSG_TRY(launch_sqdist_exact(p->pl_rows, cnt, d_tgt, n, d, p->dmat, (size_t)n, st)); }
        { SSG_PROF("row_select", st); SSG_TRY(launch_row_select(p->dmat, (size_t)n, cnt, n, nullptr, k, false, p->pl_sel_idx, p->pl_sel_val,
                                  SSG_RANK_STRIDE, p->cursor, st)); }
        { SSG_PROF("knn_sets", st); SSG_TRY(launch_knn_scan(p->dmat, (size_t)n, cnt, n, p->pl_sel_val, SSG_RANK_STRIDE, k, p->pl_flag_rows + f0,
                                p->q_idx, p->q_cnt, p->pl_flags + 1, st)); }
    }
    if (nflag > 0) {
        SSG_CUDA_TRY(cudaMemcpyAsync(h_flags, p->pl_flags, sizeof(h_flags), cudaMemcpyDeviceToHost, st));
        SSG_CUDA_TRY(cudaStreamSynchronize(st));
        if (h_flags[1])
            return ssg_set_error(SSG_ERR_CAPACITY, "rerank_plain: a neighbour set exceeds %d entries (massive ties at the "
                                 "k-th distance)", SSG_VQ_STRIDE);
    }
    { SSG_PROF("csc_build", st); SSG_TRY(launch_csc_build(n, p->q_idx, p->q_cnt, p->colcnt, p->colptr, p->cursor, p->csc_row, st)); }
    { SSG_PROF("jaccard_plain", st); SSG_TRY(launch_jaccard_plain(n, p->q_idx, p->q_cnt, p->colptr, p->csc_row, p->vec, lambda_value, d_final, st)); }
    p->last_n = n;
    return SSG_OK;
}

// reid/rerank_plain.py:27-123 re_ranking_lh(input_feature_source, input_feature, k1=20, k2=6, lambda_value=0.2): the
// k-reciprocal / Jaccard part of ssg_rerank_run with the float64 source term of that function.
extern "C" int ssg_rerank_lh(ssg_rerank_plan* p, const float* d_src, int ns, const float* d_tgt, int n, int d, int k1,
                             int k2, double lambda_value, int dist_mode, double* d_final, void* stream) {
    SSG_TRY(check_run_args(p, d_src, ns, d_tgt, n, d, k1, k2));
    if (!d_final) return ssg_set_error(SSG_ERR_INVALID, "rerank_lh: d_final is null");
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (!p->lh_vec) {
        SSG_TRY(dalloc((void**)&p->lh_minsum, sizeof(double) * (size_t)p->n_max, &p->bytes));
        SSG_TRY(dalloc((void**)&p->lh_vec, sizeof(double) * (size_t)p->n_max, &p->bytes));
    }
    const int k1d = k1 > k2 - 1 ? k1 : k2 - 1;
    SSG_TRY(ssg_rerank_distance_rows(p, d_src, ns, d_tgt, n, d, k1d, dist_mode, 0, n, nullptr, stream));
    SSG_TRY(finish_sparse_stages(p, d_tgt, n, d, k1, k2, st));
    { SSG_PROF("source_vector", st); SSG_TRY(launch_source_vec_f64(d_tgt, n, d_src, ns, d, p->lh_minsum, p->lh_vec, st)); }
    { SSG_PROF("jaccard_final", st); SSG_TRY(launch_jaccard_final_lh(n, p->q_idx, p->q_val, p->q_cnt, p->colptr, p->csc_row, p->lh_vec, lambda_value,
                                    d_final, st)); }
    p->last_n = n;
    return SSG_OK;
}

static int grow(void** ptr, size_t* have, size_t need) {
    if (need <= *have) return SSG_OK;
    if (*ptr) cudaFree(*ptr);
    *ptr = nullptr; *have = 0;
    SSG_CUDA_TRY(cudaMalloc(ptr, need));
    *have = need;
    return SSG_OK;
}

extern "C" int ssg_rerank_host(ssg_rerank_plan* p, const float* h_src, int ns, const float* h_tgt, int n, int d,
                               int k1, int k2, double lambda_value, int dist_mode, int no_rerank,
                               double* h_final, float* h_euclid) {
    SSG_TRY(check_run_args(p, h_src, ns, h_tgt, n, d, k1, k2));
    if (!no_rerank && !h_final) return ssg_set_error(SSG_ERR_INVALID, "rerank_host: h_final is null");
    SSG_ON_DEVICE(p->device);
    const size_t nn = (size_t)n * n;
    SSG_TRY(grow((void**)&p->io_src, &p->io_src_bytes, sizeof(float) * (size_t)ns * d));
    SSG_TRY(grow((void**)&p->io_tgt, &p->io_tgt_bytes, sizeof(float) * (size_t)n * d));
    if (!no_rerank) SSG_TRY(grow((void**)&p->io_final, &p->io_final_bytes, sizeof(double) * nn));
    if (h_euclid) SSG_TRY(grow((void**)&p->io_euclid, &p->io_euclid_bytes, sizeof(float) * nn));
    cudaStream_t st = nullptr;
    SSG_CUDA_TRY(cudaMemcpyAsync(p->io_src, h_src, sizeof(float) * (size_t)ns * d, cudaMemcpyHostToDevice, st));
    SSG_CUDA_TRY(cudaMemcpyAsync(p->io_tgt, h_tgt, sizeof(float) * (size_t)n * d, cudaMemcpyHostToDevice, st));
    if (no_rerank) {
        // rerank.py:65-66: the source term is still computed before the early return; only
        // euclidean_dist is handed back.
        SSG_TRY(distance_stages(p, p->io_src, ns, p->io_tgt, n, d, k1, dist_mode, h_euclid ? p->io_euclid : nullptr,
                                false, 0, n, st));
    } else {
        SSG_TRY(ssg_rerank_run(p, p->io_src, ns, p->io_tgt, n, d, k1, k2, lambda_value, dist_mode, p->io_final,
                               h_euclid ? p->io_euclid : nullptr, st));
        SSG_CUDA_TRY(cudaMemcpyAsync(h_final, p->io_final, sizeof(double) * nn, cudaMemcpyDeviceToHost, st));
    }
    if (h_euclid)
        SSG_CUDA_TRY(cudaMemcpyAsync(h_euclid, p->io_euclid, sizeof(float) * nn, cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaStreamSynchronize(st));
    return SSG_OK;
}

// reid/rerank_initial.py:40-99 on precomputed similarity blocks (device float32, row-major).
extern "C" int ssg_rerank_init(ssg_rerank_plan* p, const float* d_qg, const float* d_qq, const float* d_gg, int q,
                               int g, int k1, int k2, double lambda_value, float* d_out, void* stream) {
    if (!p || !d_qg || !d_qq || !d_gg || !d_out || q <= 0 || g <= 0)
        return ssg_set_error(SSG_ERR_INVALID, "rerank_init: bad arguments");
    const int n = q + g;
    if (n > p->n_max || (size_t)n * n > p->dmat_elems)
        return ssg_set_error(SSG_ERR_INVALID, "rerank_init: q+g=%d exceeds the plan (n_max=%d)", n, p->n_max);
    if (k1 < 1 || k1 > 31 || k2 < 1 || k2 > 8) return ssg_set_error(SSG_ERR_INVALID, "rerank_init: k1/k2 out of range");
    SSG_TRY(expanded_row_fits(k1, k2));
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const int k1p = k1 + 1, khp = (int)rint(k1 / 2.0) + 1;
    { SSG_PROF("init_assemble", st); SSG_TRY(launch_init_assemble(d_qg, d_qq, d_gg, q, g, p->dmat, st)); }
    { SSG_PROF("row_minmax", st); SSG_TRY(launch_row_minmax(p->dmat, (size_t)n, n, n, nullptr, p->rowmax, st)); }
    // the query expansion reads rank[:, :k2]: the table must hold max(k1 + 1, k2) sorted columns (k2 > k1 + 1 lies
    // outside the prefix np.argpartition sorts in the reference, rerank_initial.py:52 -- sorted order is what we define)
    const int kcols = k1p > k2 ? k1p : k2;
    { SSG_PROF("row_select", st); SSG_TRY(launch_row_select(p->dmat, (size_t)n, n, n, p->rowmax, kcols, false, p->rank, p->rank_val,
                                      SSG_RANK_STRIDE, p->cursor, st)); }
    p->rank_cols = kcols;
    { SSG_PROF("krecip_build", st); SSG_TRY(launch_krecip_build(p->rank, n, k1p, khp, p->v_idx, p->v_cnt, st)); }
    SSG_TRY(launch_gather_row_vals(p->dmat, (size_t)n, n, p->v_idx, p->v_cnt, SSG_V_STRIDE, p->v_val, st));
    { SSG_PROF("krecip_weights", st); SSG_TRY(launch_krecip_weights(p->rowmax, n, p->v_cnt, p->v_val, 0, st)); }
    if (k2 != 1) {
        { SSG_PROF("query_expand", st); SSG_TRY(launch_query_expand(p->rank, n, k2, p->v_idx, p->v_val, p->v_cnt, p->q_idx, p->q_val, p->q_cnt, st)); }
    } else {
        SSG_CUDA_TRY(cudaMemcpy2DAsync(p->q_idx, sizeof(int) * SSG_VQ_STRIDE, p->v_idx, sizeof(int) * SSG_V_STRIDE,
                                       sizeof(int) * SSG_V_STRIDE, n, cudaMemcpyDeviceToDevice, st));
        SSG_CUDA_TRY(cudaMemcpy2DAsync(p->q_val, sizeof(float) * SSG_VQ_STRIDE, p->v_val, sizeof(float) * SSG_V_STRIDE,
                                       sizeof(float) * SSG_V_STRIDE, n, cudaMemcpyDeviceToDevice, st));
        SSG_CUDA_TRY(cudaMemcpyAsync(p->q_cnt, p->v_cnt, sizeof(int) * (size_t)n, cudaMemcpyDeviceToDevice, st));
    }
    { SSG_PROF("csc_build", st); SSG_TRY(launch_csc_build(n, p->q_idx, p->q_cnt, p->colcnt, p->colptr, p->cursor, p->csc_row, st)); }
    { SSG_PROF("jaccard_init", st); SSG_TRY(launch_jaccard_init(n, q, p->q_idx, p->q_val, p->q_cnt, p->colptr, p->csc_row, p->dmat, p->rowmax,
                                        lambda_value, d_out, st)); }
    p->last_n = n;
    return SSG_OK;
}

extern "C" int ssg_rerank_get_stage(ssg_rerank_plan* p, int stage, void* h_dst, size_t bytes) {
    if (!p || !h_dst) return ssg_set_error(SSG_ERR_INVALID, "get_stage: null argument");
    const size_t n = (size_t)p->last_n;
    const void* src = nullptr;
    size_t need = 0;
    switch (stage) {
        case SSG_STAGE_VEC: src = p->vec; need = 4 * n; break;
        case SSG_STAGE_ROWMAX: src = p->rowmax; need = 4 * n; break;
        case SSG_STAGE_RANK: src = p->rank; need = 4 * n * SSG_RANK_STRIDE; break;
        case SSG_STAGE_RANK_VAL: src = p->rank_val; need = 4 * n * SSG_RANK_STRIDE; break;
        case SSG_STAGE_V_CNT: src = p->v_cnt; need = 4 * n; break;
        case SSG_STAGE_V_IDX: src = p->v_idx; need = 4 * n * SSG_V_STRIDE; break;
        case SSG_STAGE_V_VAL: src = p->v_val; need = 4 * n * SSG_V_STRIDE; break;
        case SSG_STAGE_VQ_CNT: src = p->q_cnt; need = 4 * n; break;
        case SSG_STAGE_VQ_IDX: src = p->q_idx; need = 4 * n * SSG_VQ_STRIDE; break;
        case SSG_STAGE_VQ_VAL: src = p->q_val; need = 4 * n * SSG_VQ_STRIDE; break;
        case SSG_STAGE_FLAGGED: src = p->flagged; need = 4; break;
        default: return ssg_set_error(SSG_ERR_INVALID, "get_stage: unknown stage %d", stage);
    }
    if (bytes < need) return ssg_set_error(SSG_ERR_INVALID, "get_stage: buffer too small (%zu < %zu)", bytes, need);
    SSG_ON_DEVICE(p->device);
    SSG_CUDA_TRY(cudaDeviceSynchronize());
    SSG_CUDA_TRY(cudaMemcpy(h_dst, src, need, cudaMemcpyDeviceToHost));
    return SSG_OK;
}
