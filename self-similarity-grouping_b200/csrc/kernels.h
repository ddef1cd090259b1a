// Internal launch wrappers shared by the translation units of libssg_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace ssg {

// dist.cu
int launch_sqdist_exact(const float* X, int nx, const float* Y, int ny, int d, float* out, size_t ldo,
                        cudaStream_t st);
int launch_dot_exact(const float* X, int nx, const float* Y, int ny, int d, float* out, size_t ldo,
                     cudaStream_t st);
int launch_row_minmax(const float* M, size_t ld, int rows, int cols, float* rmin, float* rmax,
                      cudaStream_t st);
int launch_row_select(const float* M, size_t ld, int rows, int cols, const float* scale, int K, bool largest,
                      int* out_idx, float* out_val, int out_stride, int* row_done, cudaStream_t st);
int launch_cand_reduce(int rows, int cols, int K, bool is_max, const float* exact, const float* cand_approx,
                       int stride, const float* norm_row, const float* norm_other_max, float eps_rel, float* out,
                       int* flag_cnt, int* flag_rows, int row_offset, cudaStream_t st);
int launch_rank_finalize(int rows, int cols, int K, int k1p, const int* cand_idx, const float* cand_approx,
                         const float* exact, int stride, const float* rowmax, const float* norm_row,
                         const float* norm_other_max, float eps_rel, int* rank, float* rank_val, int* flag_cnt,
                         int* flag_rows, int row_offset, cudaStream_t st);
int launch_gather_rows(const float* X, int d, const int* rows, int cnt, float* out, cudaStream_t st);
int launch_scatter_f32(const float* in, int cnt, int width, int in_stride, const int* rows, float* out,
                       int out_stride, cudaStream_t st);
int launch_vec_max(const float* vec, int n, float* out_max, cudaStream_t st);
int launch_pair_exact(const float* A, int rows, const float* B, int d, const int* idx, int idx_stride,
                      const int* cnt, int fixed_cnt, float* out, int out_stride, cudaStream_t st);

// gemm_tc.cu
int launch_sqdist_tensor(const float* X, int nx, const float* Y, int ny, int d, float* out, size_t ldo,
                         cudaStream_t st);
int launch_split_bf16x3(const float* x, int n, int d, int which, const float* centre, void* out_bf16, float* norm2,
                        cudaStream_t st);
int launch_col_mean(const float* x, int n, int d, double* partial, float* mean, cudaStream_t st);
// sym != 0: A and B are the two splits of the same rows and `out` is the whole square matrix -- only the tiles that
// touch the upper triangle are computed, the rest is mirrored (half the tensor work)
int launch_gemm_dist(const void* a_split, const float* na, int m, const void* b_split, const float* nb, int n,
                     int k, float* out, size_t ldc, cudaStream_t st, int sym = 0);

// rerank.cu
int launch_source_vector(const float* rowmin, int n, float* vec, float* scratch, cudaStream_t st);
int launch_krecip_build(const int* rank, int n, int k1p, int khp, int* v_idx, int* v_cnt, cudaStream_t st);
int launch_krecip_weights(const float* rowmax, int n, const int* v_cnt, float* v_val, int normalised,
                          cudaStream_t st);
int launch_krecip_classify(const int* rank, const float* rank_val, int n, int k1p, const int* v_idx, const int* v_cnt,
                           float* v_val, int* todo_idx, int* todo_slot, int* todo_cnt, cudaStream_t st);
int launch_krecip_scatter(const float* rowmax, int n, const float* todo_od, const int* todo_slot, const int* todo_cnt,
                          float* v_val, cudaStream_t st);
int launch_query_expand(const int* rank, int n, int k2, const int* v_idx, const float* v_val,
                        const int* v_cnt, int* q_idx, float* q_val, int* q_cnt, cudaStream_t st);
int launch_csc_build(int n, const int* q_idx, const int* q_cnt, int* colcnt, int* colptr, int* cursor,
                     int* csc_row, cudaStream_t st);
// rows [row0, row0+rows) of final_dist; `final_dist` points at the first of those rows (leading dimension n)
int launch_jaccard_final(int n, int row0, int rows, const int* q_idx, const float* q_val, const int* q_cnt,
                         const int* colptr, const int* csc_row, const float* vec, double lambda_value,
                         double* final_dist, cudaStream_t st);
// sparse form of final_dist: sp_rowptr == NULL counts the touched columns per row into sp_cnt, otherwise the rows are
// written (ascending columns) at sp_rowptr[i]
int launch_jaccard_sparse(int n, const int* q_idx, const float* q_val, const int* q_cnt, const int* colptr,
                          const int* csc_row, const float* vec, double lambda_value, const int* sp_rowptr, int* sp_cnt,
                          int* sp_col, double* sp_val, cudaStream_t st);
// re_ranking_lh (reid/rerank_plain.py:27-123): float64 source term on the un-squared distances
int launch_source_vec_f64(const float* tgt, int n, const float* src, int ns, int d, double* minsum, double* vec,
                          cudaStream_t st);
int launch_jaccard_final_lh(int n, const int* q_idx, const float* q_val, const int* q_cnt, const int* colptr,
                            const int* csc_row, const double* vec, double lambda_value, double* final_dist,
                            cudaStream_t st);
// plain kNN-set re-ranking (reid/rerank_plain.py:125-178)
int launch_knn_sets(const int* rank, const float* rank_val, int n, int k, int* set_idx, int* set_cnt, int* flag_rows,
                    int* flag_cnt, cudaStream_t st);
int launch_knn_scan(const float* M, size_t ld, int rows, int cols, const float* sel_val, int sel_stride, int k,
                    const int* rows_list, int* set_idx, int* set_cnt, int* overflow, cudaStream_t st);
int launch_jaccard_plain(int n, const int* s_idx, const int* s_cnt, const int* colptr, const int* csc_row,
                         const float* vec, double lambda_value, double* final_dist, cudaStream_t st);
int launch_jaccard_init(int n, int q, const int* q_idx, const float* q_val, const int* q_cnt, const int* colptr,
                        const int* csc_row, const float* dmat, const float* rowmax, double lambda_value, float* out,
                        cudaStream_t st);
int launch_init_assemble(const float* qg, const float* qq, const float* gg, int q, int g, float* dt, cudaStream_t st);
int launch_gather_row_vals(const float* M, size_t ld, int rows, const int* idx, const int* cnt, int stride,
                           float* out, cudaStream_t st);
int launch_exclusive_scan_i32(const int* in, int* out, int n, cudaStream_t st);  // out has n+1 entries

// cluster.cu (plan-level entry points live there as well)

}  // namespace ssg
