/* ssg_b200.h — C ABI of the B200-native pseudo-label hot path of Self-Similarity Grouping.
 *
 * The reference (SHI-Labs/Self-Similarity-Grouping) is pure Python and has no FFI of its own
 * (SURVEY.md §8b); this header is the boundary a maintainer would bind with ctypes/cffi (see
 * INTEGRATION.md).  Each group of entry points cites the reference interface it replaces.
 *
 * Conventions
 *   - every function returns SSG_OK (0) or a negative SSG_ERR_* code; ssg_last_error() returns a
 *     thread-local message for the last failure.  No exception crosses this boundary.
 *   - `d_*` pointers are DEVICE pointers on the plan's device, `h_*` pointers are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All device entry
 *     points are stream-ordered and asynchronous unless they return a host scalar.
 *   - plans own their workspace, allocated at creation; nothing is allocated per call, except that the opt-in
 *     entry points (sparse form, rerank_plain, rerank_lh) add their buffers on first use and the CSR of the sparse
 *     form grows when a result needs more room.
 *   - plans are bound to one device and are not thread-safe.
 */
#ifndef SSG_B200_H
#define SSG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSG_OK 0
#define SSG_ERR_INVALID (-1)     /* bad argument (shape, alignment, null pointer)           */
#define SSG_ERR_CUDA (-2)        /* a CUDA runtime call failed; message has the CUDA string  */
#define SSG_ERR_CAPACITY (-3)    /* a bounded workspace overflowed; re-create with more room */
#define SSG_ERR_UNSUPPORTED (-4) /* feature not built / device is not sm_100                 */

int ssg_version(void);
const char* ssg_last_error(void);
/* number of visible CUDA devices and compute capability (major*10+minor) of `device`. */
int ssg_device_info(int device, int* n_devices, int* sm);

/* ------------------------------------------------------------------------------------------------
 * Pairwise squared Euclidean distance  (scipy cdist as used by reid/rerank.py:37,61-62; and
 * reid/evaluators.py:63-85 pairwise_distance).
 *   out[i*ldo + j] = fl32( fl32( sqrt( sum_k (x_ik - y_jk)^2 ) )^2 ), the sum taken sequentially in
 *   float64 exactly as cdist does.
 * mode: SSG_DIST_EXACT (float64 direct difference, bit-identical to cdist) or
 *       SSG_DIST_TENSOR (bf16x3 split on tcgen05 tensor cores, |err| <~ 2e-5; candidates are
 *       re-scored exactly inside ssg_rerank_run).
 * ------------------------------------------------------------------------------------------------ */
#define SSG_DIST_EXACT 0
#define SSG_DIST_TENSOR 1
int ssg_sqdist(const float* d_x, int nx, const float* d_y, int ny, int d, int mode, float* d_out,
               size_t ldo, void* stream);

/* Dot-product block  out[i*ldo + j] = fl32( sum_k x_ik * y_jk )  (products exact, summed sequentially in float64):
 * the similarity blocks of the cosine re-ranking -- np.dot at reid/rerank.py:174-176 (re_ranking_init from features)
 * and reid/eug.py:223-225 (EUG label estimation) -- which then go to ssg_rerank_init. */
int ssg_dot(const float* d_x, int nx, const float* d_y, int ny, int d, float* d_out, size_t ldo, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Source-aware k-reciprocal / Jaccard re-ranking: reid/rerank.py:27-127 re_ranking(
 *   input_feature_source, input_feature, k1=20, k2=6, lambda_value, ...), O-f32 arithmetic
 *   (float16 -> float32, stable argsort).
 * ------------------------------------------------------------------------------------------------ */
typedef struct ssg_rerank_plan ssg_rerank_plan;

#define SSG_RANK_STRIDE 32   /* rank table row stride; k1+1 <= 32          */
#define SSG_V_STRIDE 256     /* k-reciprocal row capacity (bound 21+21*11) */
#define SSG_VQ_STRIDE 1536   /* expanded row capacity (bound 6*252)        */

int ssg_rerank_plan_create(ssg_rerank_plan** plan, int device, int n_max, int ns_max, int d);
int ssg_rerank_plan_destroy(ssg_rerank_plan* plan);
size_t ssg_rerank_plan_bytes(const ssg_rerank_plan* plan);

/* Device in / device out.  d_final: [n,n] float64 (row-major) = final_dist of the reference;
 * d_euclid: optional [n,n] float32 = euclidean_dist of the reference (pre-normalisation squared
 * distances), may be NULL.  k1 <= 20 (the k-reciprocal row capacity: (k1+1) * (round(k1/2) + 2) <= SSG_V_STRIDE;
 * larger values return SSG_ERR_INVALID), k2 <= 8; k2 > k1 + 1 is allowed, as in the reference. */
int ssg_rerank_run(ssg_rerank_plan* plan, const float* d_src, int ns, const float* d_tgt, int n, int d,
                   int k1, int k2, double lambda_value, int dist_mode, double* d_final,
                   float* d_euclid, void* stream);
/* The same computation in two halves, for row-block sharding across GPUs:
 *   ssg_rerank_distance_rows : stages (i)-(iv) for target rows [row0, row0+rows): fills those rows of the plan's
 *                              tables (row minimum over the sources, row maximum, k1+1 leading rank columns and
 *                              their normalised distances).  d_euclid (optional) is the full [n,n] matrix; only the
 *                              rows of the block are written.
 *   ssg_rerank_tables        : device pointers of the tables (rowmin float[n], rowmax float[n], rank int32[n,32],
 *                              rank_val float[n,32]) so that the caller can all-gather the row blocks in place.
 *   ssg_rerank_finish        : everything after the tables are complete (source vector, stages (v)-(viii)). */
int ssg_rerank_distance_rows(ssg_rerank_plan* plan, const float* d_src, int ns, const float* d_tgt, int n, int d,
                             int k1, int dist_mode, int row0, int rows, float* d_euclid, void* stream);
int ssg_rerank_tables(ssg_rerank_plan* plan, float** d_rowmin, float** d_rowmax, int** d_rank, float** d_rank_val);
int ssg_rerank_finish(ssg_rerank_plan* plan, const float* d_tgt, int n, int d, int k1, int k2, double lambda_value,
                      double* d_final, void* stream);

/* Host in / host out (the literal drop-in of re_ranking): copies features up, results down. */
int ssg_rerank_host(ssg_rerank_plan* plan, const float* h_src, int ns, const float* h_tgt, int n, int d,
                    int k1, int k2, double lambda_value, int dist_mode, int no_rerank, double* h_final,
                    float* h_euclid);

/* reid/rerank_initial.py:40-99 re_ranking_init(q_g_dist, q_q_dist, g_g_dist, k1=20, k2=6, lambda_value=0.3) (also
 * reid/rerank.py:171-234): float32 k-reciprocal re-ranking on precomputed SIMILARITY blocks (device, row-major
 * [q,g], [q,q], [g,g]); d_out is [q,g] float32.  Needs (q+g) <= n_max of the plan.  Ties are ordered by index
 * (the reference's argpartition leaves them unspecified). */
int ssg_rerank_init(ssg_rerank_plan* plan, const float* d_qg, const float* d_qq, const float* d_gg, int q, int g,
                    int k1, int k2, double lambda_value, float* d_out, void* stream);

/* reid/rerank_plain.py:125-178 re_ranking(input_feature_source, input_feature, k=20, lambda_value=0.1) -- the plain
 * kNN-set variant both drivers carry as a commented-out import (selftraining.py:29):
 *   S_i = { j != i : d2[i,j] <= k-th smallest entry of row i (diagonal included) },
 *   J = scipy cdist(S, S, 'jaccard'),  final = J*(1-lambda) + (v_i+v_j)*lambda with v as in reid/rerank.py:36-40.
 * d_final: [n,n] float64 (the reference returns it twice).  1 <= k <= 31, n >= k.  float16 -> float32 as for
 * ssg_rerank_run.  Synchronises the stream. */
int ssg_rerank_plain(ssg_rerank_plan* plan, const float* d_src, int ns, const float* d_tgt, int n, int d, int k,
                     double lambda_value, int dist_mode, double* d_final, void* stream);

/* reid/rerank_plain.py:27-123 re_ranking_lh(input_feature_source, input_feature, k1=20, k2=6, lambda_value=0.2): the
 * k-reciprocal / Jaccard re-ranking of ssg_rerank_run with that function's source term -- v_i = min_j cdist(t_i, s_j) on
 * the un-squared float64 distances, v /= max(v), source_dist = v_i + v_j, all in float64.  d_final: [n,n] float64. */
int ssg_rerank_lh(ssg_rerank_plan* plan, const float* d_src, int ns, const float* d_tgt, int n, int d, int k1, int k2,
                  double lambda_value, int dist_mode, double* d_final, void* stream);

/* Intermediate results of the last ssg_rerank_run, copied to the host (stage-isolated parity tests). */
#define SSG_STAGE_VEC 0       /* float  [n]        normalised source vector v (rerank.py:36-40)     */
#define SSG_STAGE_ROWMAX 1    /* float  [n]        row maximum of the squared distance (rerank.py:68) */
#define SSG_STAGE_RANK 2      /* int32  [n,32]     initial_rank[:, :k1+1] (rerank.py:70)            */
#define SSG_STAGE_RANK_VAL 3  /* float  [n,32]     normalised distance of those entries             */
#define SSG_STAGE_V_CNT 4     /* int32  [n]        nnz of V rows (rerank.py:74-92)                  */
#define SSG_STAGE_V_IDX 5     /* int32  [n,256]                                                     */
#define SSG_STAGE_V_VAL 6     /* float  [n,256]                                                     */
#define SSG_STAGE_VQ_CNT 7    /* int32  [n]        nnz of expanded rows (rerank.py:94-98)           */
#define SSG_STAGE_VQ_IDX 8    /* int32  [n,1536]                                                    */
#define SSG_STAGE_VQ_VAL 9    /* float  [n,1536]                                                    */
#define SSG_STAGE_FLAGGED 10  /* int32  [1]        rows that took the exact fallback (tensor mode)  */
int ssg_rerank_get_stage(ssg_rerank_plan* plan, int stage, void* h_dst, size_t bytes);

/* ------------------------------------------------------------------------------------------------
 * eps estimate and DBSCAN: selftraining.py:289-306 (np.triu/nonzero/sort/mean; sklearn
 * DBSCAN(eps, min_samples=4, metric='precomputed').fit_predict on the dense matrix).
 * dtype: SSG_F64 or SSG_F32 matrix elements (comparisons are made in the matrix dtype, as numpy does).
 * ------------------------------------------------------------------------------------------------ */
#define SSG_F32 0
#define SSG_F64 1
typedef struct ssg_cluster_plan ssg_cluster_plan;

/* max_neighbors bounds the total number of (i,j) pairs with dist<=eps the plan can hold. */
int ssg_cluster_plan_create(ssg_cluster_plan** plan, int device, int n_max, long long max_neighbors);
int ssg_cluster_plan_destroy(ssg_cluster_plan* plan);
size_t ssg_cluster_plan_bytes(const ssg_cluster_plan* plan);

/* eps = mean of the round-half-even(rho * M) smallest non-zero entries above the diagonal, M = their
 * count.  Synchronises the stream (returns host scalars).  eps is NaN when the count rounds to 0. */
int ssg_eps_estimate(ssg_cluster_plan* plan, const void* d_dist, int dtype, int n, double rho,
                     double* h_eps, long long* h_top_num, void* stream);
/* labels: int64 [n] on the device, -1 = noise, cluster ids as sklearn assigns them.
 * The call synchronises the stream once to read the neighbour-capacity flag (SSG_ERR_CAPACITY on overflow: the labels
 * are then undefined); h_n_clusters is optional. */
int ssg_dbscan(ssg_cluster_plan* plan, const void* d_dist, int dtype, int n, double eps, int min_samples,
               int64_t* d_labels, int* h_n_clusters, void* stream);
/* core-sample mask of the last ssg_dbscan (uint8 [n], host). */
int ssg_dbscan_core_mask(ssg_cluster_plan* plan, uint8_t* h_core, int n);

/* Host-matrix conveniences (sklearn drop-in: matrix is copied to the device first). */
int ssg_eps_estimate_host(ssg_cluster_plan* plan, const void* h_dist, int dtype, int n, double rho,
                          double* h_eps, long long* h_top_num);
int ssg_dbscan_host(ssg_cluster_plan* plan, const void* h_dist, int dtype, int n, double eps,
                    int min_samples, int64_t* h_labels, int* h_n_clusters);

/* ------------------------------------------------------------------------------------------------
 * Row-sharded variants for one process per GPU.  Rank r of `world` holds the rows
 * [lo_r, hi_r) of final_dist (lo_r = r*(n/world) + min(r, n%world): the first n%world ranks own one row more),
 * as a [rows, n] block with leading dimension n.  The library launches the kernels; the CALLER runs the collectives
 * (torch.distributed / NCCL) on the plan's buffers, whose device pointers ssg_cluster_buffers returns:
 *   hist uint64[4096], state uint64[8], partial double[n_max], list double[2^20], cnt int32[n_max],
 *   nbr int32[max_neighbors].
 *
 * ssg_rerank_finish_rows: ssg_rerank_finish, but only rows [row0, row0+rows) of final_dist are produced, into
 *   d_final_rows (which points at row row0).  The sparse stages run for all n rows on every rank (they are cheap and
 *   read other rows' tables); the tables must be complete (all-gathered) as for ssg_rerank_finish.
 *
 * eps (selftraining.py:289-293) on a SYMMETRIC matrix; every unordered pair is visited by exactly one rank
 * (diagonal block: strict upper triangle; off-diagonal block (r,c): taken by rank r iff r<c and r+c odd, or r>c and
 * r+c even).  Sequence, identical on every rank:
 *   ssg_eps_shard_begin
 *   for pass in 0, 1:  ssg_eps_shard_hist(pass); all-reduce(sum) hist; ssg_eps_shard_pick(pass, rho)
 *   ssg_eps_shard_gather(exact_threshold=0, &count)   -- per-row sums below the threshold's 24-bit bin into
 *       partial[global row], the bin's entries into list[0, count); count == -1: more than 2^20 such entries
 *   if every rank's count >= 0 and their sum <= 2^20:
 *       all-gather the rows of partial in place; concatenate the lists of all ranks (rank order) into list and store
 *       the total into state[5]; ssg_eps_shard_finish(exact_threshold=0)
 *   else (massive ties):
 *       for pass in 2..5: hist, all-reduce, pick;  ssg_eps_shard_gather(exact_threshold=1, NULL);
 *       all-gather partial; ssg_eps_shard_finish(exact_threshold=1)
 * The list's entries are summed exactly on their integer mantissas, the per-row partial sums in a fixed order: eps
 * is bit-reproducible for a given (n, world) and agrees with ssg_eps_estimate to the last few ulps (the order of
 * the float64 additions differs).
 *
 * DBSCAN: ssg_dbscan_shard_count (cnt of the local rows); all-gather cnt rows in place;
 *   ssg_dbscan_shard_fill (global CSR offsets from cnt, zero nbr[0,total), fill the local rows; SSG_ERR_CAPACITY
 *   when total exceeds the plan's max_neighbors -- total and the error are the same on every rank);
 *   all-reduce(sum) nbr[0,total); ssg_dbscan_shard_label (every rank labels all n rows: identical results).
 * ------------------------------------------------------------------------------------------------ */
int ssg_rerank_finish_rows(ssg_rerank_plan* plan, const float* d_tgt, int n, int d, int k1, int k2,
                           double lambda_value, int row0, int rows, double* d_final_rows, void* stream);
int ssg_cluster_buffers(ssg_cluster_plan* plan, void** d_hist, void** d_state, void** d_partial, void** d_list,
                        void** d_cnt, void** d_nbr);
int ssg_eps_shard_begin(ssg_cluster_plan* plan, void* stream);
int ssg_eps_shard_hist(ssg_cluster_plan* plan, const void* d_rows, int dtype, int n, int world, int rank, int pass,
                       void* stream);
int ssg_eps_shard_pick(ssg_cluster_plan* plan, int pass, double rho, void* stream);
int ssg_eps_shard_gather(ssg_cluster_plan* plan, const void* d_rows, int dtype, int n, int world, int rank,
                         int exact_threshold, long long* h_list_count, void* stream);
int ssg_eps_shard_finish(ssg_cluster_plan* plan, int n, int exact_threshold, double* h_eps, long long* h_top_num,
                         void* stream);
int ssg_dbscan_shard_count(ssg_cluster_plan* plan, const void* d_rows, int dtype, int n, int row0, int rows,
                           double eps, void* stream);
int ssg_dbscan_shard_fill(ssg_cluster_plan* plan, const void* d_rows, int dtype, int n, int row0, int rows,
                          double eps, long long* h_total, void* stream);
int ssg_dbscan_shard_label(ssg_cluster_plan* plan, int n, int min_samples, int64_t* d_labels, int* h_n_clusters,
                           void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sparse form of final_dist, for callers that only want eps and the labels (selftraining.py:289-306 is all the driver
 * uses the matrix for).  A column m that shares no k-reciprocal neighbour with row i has Jaccard distance 1, hence
 * final_dist[i,m] = fl32(1-lambda) + fl32(v_i+v_m)*lambda >= fl32(1-lambda) =: threshold (rerank.py:115-122, v >= 0).
 * Only the other ("touched", ~1 %) entries can be smaller:
 *   ssg_rerank_finish_sparse : ssg_rerank_finish, but the result is a CSR over the touched columns (ascending inside a
 *                              row, values bit-identical to the dense matrix) kept in the plan; 0 <= lambda < 1.
 *                              Synchronises the stream (row lengths are data dependent; buffers grow on demand).
 *   ssg_rerank_sparse_view   : device pointers of that CSR (rowptr int32[n+1], col int32[nnz], val double[nnz]) and the
 *                              threshold below which an entry is guaranteed to be in it.
 *   ssg_eps_sparse           : ssg_eps_estimate on the CSR.  *h_certified = 1 iff the rho-slice lies entirely below
 *                              the threshold (then eps is the value the dense matrix gives, up to the order of the
 *                              float64 additions); 0: the caller must materialise the matrix (ssg_rerank_finish).
 *   ssg_dbscan_sparse        : ssg_dbscan on the CSR; the caller guarantees eps < threshold (no entry outside the CSR
 *                              is a neighbour then).  Labels are identical to the dense ones.
 * ------------------------------------------------------------------------------------------------ */
int ssg_rerank_finish_sparse(ssg_rerank_plan* plan, const float* d_tgt, int n, int d, int k1, int k2,
                             double lambda_value, long long* h_nnz, void* stream);
int ssg_rerank_sparse_view(ssg_rerank_plan* plan, int** d_rowptr, int** d_col, double** d_val, long long* nnz,
                           double* threshold);
int ssg_eps_sparse(ssg_cluster_plan* plan, int n, const int* d_rowptr, const int* d_col, const double* d_val,
                   double threshold, double rho, double* h_eps, long long* h_top_num, int* h_certified, void* stream);
int ssg_dbscan_sparse(ssg_cluster_plan* plan, int n, const int* d_rowptr, const int* d_col, const double* d_val,
                      double eps, int min_samples, int64_t* d_labels, int* h_n_clusters, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Embedding: reid/evaluators.py:18-60 extract_features + reid/feature_extraction/cnn.py:10-23 +
 * reid/models/resnet.py:86-134 (ResNet-50 trunk, num_classes=0, cluster=False) for 256x128 inputs.
 *   forward: images fp32 NCHW [n,3,256,128] (already mean/std normalised, as the reference's loaders
 *   deliver them) -> per image the (num_split>1 ? num_split+1 : 1) pooled 2048-d banks of
 *   model(x) + model(fliplr(x)) (flip != 0), L2-normalised:
 *     eval_mode == 0 : d_feat[bank * bank_stride + (row0+i) * 2048 + c]      (each bank normalised alone)
 *     eval_mode & 1  : d_feat[((row0+i) * banks + bank) * 2048 + c]           (one norm over the concatenation)
 *     eval_mode & 2  : skip the L2 normalisation (raw pooled banks, what one model(x) call returns)
 *   Convolutions run in bf16 with fp32 accumulation on the tcgen05 tensor cores; eval-mode BatchNorm is
 *   folded into the weights when a layer is loaded.
 * Layers are indexed in a fixed order (stem, then conv1, conv2, conv3[, downsample] per bottleneck);
 * ssg_embed_layer_info gives the torchvision state_dict key prefixes of each index.
 * ------------------------------------------------------------------------------------------------ */
typedef struct ssg_embed_plan ssg_embed_plan;
int ssg_embed_num_layers(void);
int ssg_embed_layer_info(int idx, int* cin, int* cout, int* ksize, int* stride, char* conv_key, char* bn_key,
                         size_t cap);
int ssg_embed_plan_create(ssg_embed_plan** plan, int device, int batch_max, int height, int width);
int ssg_embed_plan_destroy(ssg_embed_plan* plan);
size_t ssg_embed_plan_bytes(const ssg_embed_plan* plan);
/* d_w: conv weight fp32 [cout,cin,k,k]; BatchNorm weight/bias/running_mean/running_var fp32 [cout]. */
int ssg_embed_load_layer(ssg_embed_plan* plan, int idx, const float* d_w, const float* d_gamma,
                         const float* d_beta, const float* d_mean, const float* d_var, float eps, void* stream);
int ssg_embed_forward(ssg_embed_plan* plan, const float* d_images, int n, int num_split, int eval_mode, int flip,
                      float* d_feat, size_t bank_stride, int row0, void* stream);
/* The same from raw pixels: d_images_u8 is uint8 HWC [n,256,128,3] (a decoded, resized RGB
 * image as PIL hands it over); the loader's ToTensor + Normalize of selftraining.py:36-45 -- x/255, then
 * (x - mean[c]) / std[c], IEEE float32 in that order -- run on the device on the way to bf16, so the features are
 * bit-identical to ssg_embed_forward on the normalised float32 tensor while 4x fewer bytes cross PCIe.
 * h_mean / h_std: HOST float[3]. */
int ssg_embed_forward_u8(ssg_embed_plan* plan, const uint8_t* d_images_u8, const float* h_mean, const float* h_std,
                         int n, int num_split, int eval_mode, int flip, float* d_feat, size_t bank_stride, int row0,
                         void* stream);

/* Building blocks of the trunk (NHWC bf16 activations, weights bf16 [cout][k][k][cin] with BatchNorm folded):
 * exported for stage-isolated parity tests and for callers with their own graph.
 *   ssg_op_conv   : k x k convolution (k in {1,3}, padding k/2, stride in {1,2}) + bias (+ residual, 1x1 only)
 *                   (+ ReLU).  H, W are the INPUT map size; stride 2 needs a scratch buffer of the input's size.
 *                   3x3 needs W (output) to divide 128 and H*W (output) to be a multiple or a divisor of 128.
 *   ssg_op_fold_bn: fp32 [cout,cin,k,k] conv weight + BatchNorm statistics -> bf16 [cout,kpad] + fp32 bias.
 *   ssg_op_stem   : 7x7/2 conv (K padded 147->192: d_w is bf16 [64,192], d_col scratch bf16 [images*8192,192]) + ReLU on fp32 NCHW images (and their mirror images when
 *                   flip != 0), optionally followed by the 3x3/2 max-pool.
 *   ssg_op_pooled_tail : global + stripe average pools, flip sum, L2 normalisation (see ssg_embed_forward). */
int ssg_op_conv(const void* d_x, int B, int H, int W, int cin, int ksize, int stride, const void* d_w,
                const float* d_bias, int cout, const void* d_res, int relu, void* d_y, void* d_scratch, void* stream);
int ssg_op_fold_bn(const float* d_w, int cout, int cin, int ksize, const float* d_gamma, const float* d_beta,
                   const float* d_mean, const float* d_var, float eps, int kpad, void* d_wout, float* d_bout,
                   void* stream);
int ssg_op_stem(const float* d_images, int n, int flip, const void* d_w, const float* d_bias, void* d_col,
                void* d_conv_out, void* d_pool_out, void* stream);
int ssg_op_pooled_tail(const void* d_x, int n, int num_split, int eval_mode, int flip, float* d_feat,
                       size_t bank_stride, int row0, void* stream);

/* Training-side convolution operators: what loss.backward() of reid/trainers.py:204-271
 * (FinedTrainer2.train -> _forward) asks of the ResNet-50 convolutions of reid/models/resnet.py:52-70, on the same
 * tcgen05 GEMM kernels as the forward path.  NHWC bf16 activations and gradients; fp32 master weights and weight
 * gradients [cout,cin,k,k] as torch holds them; k in {1,3}, padding k/2, stride in {1,2}; H, W = INPUT map size.
 *   ssg_op_conv_pack_weight: fp32 weights -> bf16 GEMM operand; transposed == 0: [cout][k][k][cin], the forward operand
 *                            of ssg_op_conv (with a zero bias: training-mode BatchNorm cannot be folded); transposed != 0:
 *                            [cin][k][k][cout] with mirrored taps, the operand of the data gradient.
 *   ssg_op_conv_dgrad      : d_dx bf16 [B,H,W,cin] from d_dy bf16 [B,H/stride,W/stride,cout] (cin, cout multiples of 64;
 *                            the map sizes ssg_op_conv accepts for the INPUT map).
 *   ssg_op_conv_wgrad      : d_dw fp32 [cout,cin,k,k] from d_x bf16 [B,H,W,cin] and d_dy: split-K GEMM over the B*Ho*Wo
 *                            output pixels; the partial products are summed in a fixed order (deterministic).
 *   ssg_op_stem_im2col     : the stem's im2col (7x7/2 windows, K padded 147 -> 192, (kh,kw,ci) order) on its own, so that
 *                            the stem convolution and its weight gradient are 1x1 operators on d_col [n*8192,192].
 * Scratch memory comes from the stream-ordered allocator (cudaMallocAsync). */
int ssg_op_conv_pack_weight(const float* d_w, int cout, int cin, int ksize, int transposed, void* d_out, void* stream);
int ssg_op_conv_dgrad(const void* d_dy, int B, int H, int W, int cout, int ksize, int stride, const float* d_w, int cin,
                      void* d_dx, void* stream);
int ssg_op_conv_wgrad(const void* d_x, int B, int H, int W, int cin, const void* d_dy, int cout, int ksize, int stride,
                      float* d_dw, void* stream);
int ssg_op_stem_im2col(const float* d_images, int n, int flip, void* d_col, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fine-tune loss: reid/loss/triplet.py:11-77 TripletLoss(margin, num_instances, use_semi)
 * .forward(inputs, targets, epoch) with w = None, as called per feature bank by reid/trainers.py:257-271
 * (FinedTrainer2._forward).  The batch is P = n / num_instances identities x num_instances consecutive rows.
 *   forward : d_x fp32 [n,d] (16-byte aligned), d_targets int64 [n]  ->  d_loss_prec fp32 [2] = {loss, prec};
 *             d_dist fp32 [n,n] (the clamped pairwise distances, triplet.py:27-31) and d_coef fp32 [n,n]
 *             ((d loss / d dist) / dist, consumed by backward) are caller-allocated.  use_semi != 0: every
 *             (anchor, later row of its group) pair with the anchor's closest other-label row (triplet.py:49-56);
 *             use_semi == 0: per row the farthest same-label and closest other-label rows (triplet.py:57-60).
 *             d_status int32 [2]: {1 if some anchor has no other-label row in the batch (the reference raises
 *             there; the loss is then NaN), that anchor's index}.  n <= 4096.
 *   backward: d_grad_x fp32 [n,d] = *d_grad_loss (device scalar, NULL = 1) * d loss / d x.
 * Deterministic (fixed reduction order); ties between candidate negatives take the lowest index.
 * ------------------------------------------------------------------------------------------------ */
int ssg_triplet_forward(const float* d_x, const int64_t* d_targets, int n, int d, int num_instances, float margin,
                        int use_semi, float* d_dist, float* d_coef, float* d_loss_prec, int* d_status, void* stream);
int ssg_triplet_backward(const float* d_x, int n, int d, const float* d_coef, const float* d_grad_loss,
                         float* d_grad_x, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Retrieval metrics of the evaluation step: reid/evaluation_metrics/ranking.py:18-79 cmc(...)
 * and 82-115 mean_ap(...), as called by reid/evaluators.py:88-133 evaluate_all, from the q x g distance matrix
 * WITHOUT sorting it: per query i and per match s (gallery entries with the query's id that are not filtered out --
 * same id AND same camera; with separate_camera_set also same camera -- in ascending gallery index)
 *   d_slots[i*SSG_RANK_MAX_MATCHES + s] = number of valid non-matching entries ranked before the match under the order
 *                                         (distance, gallery index) = the reference's  k - j  (ranking.py:67-75),
 *   d_nmatch[i] = number of matches (0: the reference skips the query; -1: more than SSG_RANK_MAX_MATCHES, and
 *                 d_flags[0] = 1),
 *   d_ap[i]     = sklearn.metrics.average_precision_score(matches, -dist) over the valid entries (ranking.py:105-111:
 *                 tied distances form one threshold).
 * dtype: SSG_F32 or SSG_F64 matrix elements; ids and cameras int64.  The host sums d_ap and histograms d_slots.
 * ------------------------------------------------------------------------------------------------ */
#define SSG_RANK_MAX_MATCHES 1024
int ssg_rank_metrics(const void* d_dist, int dtype, int m, int n, const long long* d_query_ids,
                     const long long* d_gallery_ids, const long long* d_query_cams, const long long* d_gallery_cams,
                     int separate_camera_set, double* d_ap, int* d_nmatch, int* d_slots, int* d_flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Per-kernel CUDA-event timers (the reference only has wall-clock AverageMeters, reid/utils/meters.py:4-23,
 * printed from reid/evaluators.py:48-57).  Disabled by default; when enabled every kernel group launched by
 * this library is bracketed by events on its own stream.  collect() synchronises the device, accumulates
 * and returns the number of distinct names; entry(i) reads one accumulated row.
 * ------------------------------------------------------------------------------------------------ */
int ssg_profile_enable(int on);
int ssg_profile_reset(void);
int ssg_profile_collect(void);
int ssg_profile_entry(int i, char* name, size_t cap, double* ms, long long* launches);

#ifdef __cplusplus
}
#endif
#endif /* SSG_B200_H */
