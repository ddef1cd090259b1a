// k-reciprocal encoding, query expansion, inverted index and Jaccard/final distance of
// reid/rerank.py:74-122 (O-f32 arithmetic).  All rows are kept sparse (index-sorted) on the device;
// only the final N x N float64 matrix the reference API returns is dense.
#include <limits.h>

#include "common.cuh"
#include "kernels.h"

namespace ssg {

// ---------------------------------------------------------------------------------------------------
// generic exclusive scan of int32 (single CTA; inputs here are at most a few 100k entries)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) exclusive_scan_i32_kernel(const int* __restrict__ in,
                                                                   int* __restrict__ out, int n) {
    __shared__ int wsum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int j = base + threadIdx.x;
        const int v = j < n ? in[j] : 0;
        int total;
        const int ex = block_exclusive_scan<1024>(v, wsum, total);
        const int c = carry;
        if (j < n) out[j] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

int launch_exclusive_scan_i32(const int* in, int* out, int n, cudaStream_t st) {
    exclusive_scan_i32_kernel<<<1, 1024, 0, st>>>(in, out, n);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// source vector: v_i = 1 - exp(-min_j d2(t_i, s_j)); v /= max(v)        (rerank.py:38-40)
// exp/subtract are monotone, so the row minimum is taken on the distances.
// ---------------------------------------------------------------------------------------------------
__global__ void source_vec_kernel(const float* __restrict__ rowmin, int n, float* __restrict__ vec) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vec[i] = __fsub_rn(1.0f, expf(-rowmin[i]));
}
__global__ void __launch_bounds__(1024) vec_max_kernel(const float* __restrict__ vec, int n,
                                                       float* __restrict__ out_max) {
    float m = -INFINITY;
    bool nan = false;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { m = fmaxf(m, vec[i]); nan |= isnan(vec[i]); }
    __shared__ float sm[32];
    __shared__ int snan;
    if (threadIdx.x == 0) snan = 0;
    __syncthreads();
    if (nan) snan = 1;
    m = warp_max_f(m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) m = fmaxf(m, sm[w]);
        *out_max = snan ? NAN : m;   // np.max propagates NaN
    }
}
__global__ void vec_div_kernel(float* __restrict__ vec, int n, const float* __restrict__ mx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vec[i] = __fdiv_rn(vec[i], *mx);
}

int launch_vec_max(const float* vec, int n, float* out_max, cudaStream_t st) {
    vec_max_kernel<<<1, 1024, 0, st>>>(vec, n, out_max);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int launch_source_vector(const float* rowmin, int n, float* vec, float* scratch, cudaStream_t st) {
    source_vec_kernel<<<ssg_cdiv(n, 256), 256, 0, st>>>(rowmin, n, vec);
    SSG_CHECK_LAUNCH();
    vec_max_kernel<<<1, 1024, 0, st>>>(vec, n, scratch);
    SSG_CHECK_LAUNCH();
    vec_div_kernel<<<ssg_cdiv(n, 256), 256, 0, st>>>(vec, n, scratch);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// k-reciprocal sets with half-k expansion (rerank.py:74-90): one CTA per row, output = sorted unique
// index list R* (at most SSG_V_STRIDE entries).
// ---------------------------------------------------------------------------------------------------
constexpr int KR_NT = 256;

__global__ void __launch_bounds__(KR_NT)
krecip_build_kernel(const int* __restrict__ rank, int n, int k1p, int khp, int* __restrict__ v_idx,
                    int* __restrict__ v_cnt) {
    __shared__ int R[32];
    __shared__ int list[SSG_V_STRIDE];
    __shared__ int s_nR, s_nlist;
    __shared__ int wsum[KR_NT / 32];
    const int i = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int RS = SSG_RANK_STRIDE;

    for (int t = tid; t < SSG_V_STRIDE; t += KR_NT) list[t] = INT_MAX;
    if (wid == 0) {
        // R(i,k1): forward neighbours whose own top-(k1+1) list contains i, in rank order
        int c = -1;
        bool found = false;
        if (lane < k1p) {
            c = rank[(size_t)i * RS + lane];
            if (c >= 0)
                for (int b = 0; b < k1p; ++b) found |= (rank[(size_t)c * RS + b] == i);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, found);
        if (found) R[__popc(mask & ((1u << lane) - 1u))] = c;
        if (lane == 0) { s_nR = __popc(mask); s_nlist = __popc(mask); }
    }
    __syncthreads();
    const int nR = s_nR;
    if (tid < nR) list[tid] = R[tid];
    __syncthreads();
    // expansion: candidates are distributed over the warps
    for (int ci = wid; ci < nR; ci += KR_NT / 32) {
        const int c = R[ci];
        int e = -1;
        bool found = false;
        if (lane < khp) {
            e = rank[(size_t)c * RS + lane];
            if (e >= 0)
                for (int b = 0; b < khp; ++b) found |= (rank[(size_t)e * RS + b] == c);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, found);
        bool inR = false;
        if (found)
            for (int q = 0; q < nR; ++q) inR |= (R[q] == e);
        const int ninter = __popc(__ballot_sync(0xffffffffu, inR));
        const int nrc = __popc(mask);
        // len(intersect1d(Rc, R)) > 2/3*len(Rc)   (rerank.py:86, double arithmetic)
        if ((double)ninter > (2.0 / 3.0) * (double)nrc) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&s_nlist, nrc);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (found) {
                const int pos = base + __popc(mask & ((1u << lane) - 1u));
                if (pos < SSG_V_STRIDE) list[pos] = e;
            }
        }
    }
    __syncthreads();
    // bitonic sort of the 256-entry list (INT_MAX padding), then unique (np.unique, rerank.py:90)
    for (int k = 2; k <= SSG_V_STRIDE; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int t = tid, p = t ^ j;
            if (p > t) {
                const int a = list[t], b = list[p];
                const bool up = (t & k) == 0;
                if ((a > b) == up) { list[t] = b; list[p] = a; }
            }
            __syncthreads();
        }
    }
    const int val = list[tid];
    const bool head = (val != INT_MAX) && (tid == 0 || list[tid - 1] != val);
    int total;
    const int pos = block_exclusive_scan<KR_NT>(head ? 1 : 0, wsum, total);
    if (head) v_idx[(size_t)i * SSG_V_STRIDE + pos] = val;
    if (tid == 0) v_cnt[i] = total;
}

int launch_krecip_build(const int* rank, int n, int k1p, int khp, int* v_idx, int* v_cnt,
                        cudaStream_t st) {
    static_assert(KR_NT == SSG_V_STRIDE, "one thread per list slot");
    if (k1p > 32 || khp > 32 || k1p + k1p * khp > SSG_V_STRIDE)
        return ssg_set_error(SSG_ERR_INVALID, "k1=%d exceeds the k-reciprocal row capacity", k1p - 1);
    krecip_build_kernel<<<n, KR_NT, 0, st>>>(rank, n, k1p, khp, v_idx, v_cnt);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// weights (rerank.py:91-92): v_val holds the exact squared distances d2(i, R*) on entry;
// w = exp(-d2/rowmax_i); V = w / np.sum(w).  np.sum's pairwise summation is replicated exactly
// (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum: 8 partial sums, blocks of 128).
// ---------------------------------------------------------------------------------------------------
__device__ float np_pairwise_sum_f32(const float* a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    if (n <= 128) {
        float r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = a[q];
        int i;
        for (i = 8; i < n - (n % 8); i += 8)
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = __fadd_rn(r[q], a[i + q]);
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __fadd_rn(np_pairwise_sum_f32(a, n2), np_pairwise_sum_f32(a + n2, n - n2));
}

// Most entries of R*(i) are members of row i's own leading rank columns, whose exact normalised distance is already in
// the rank table: copy those, and list only the remaining (expansion) entries for exact re-scoring.
__global__ void __launch_bounds__(256)
krecip_classify_kernel(const int* __restrict__ rank, const float* __restrict__ rank_val, int k1p,
                       const int* __restrict__ v_idx, const int* __restrict__ v_cnt, float* __restrict__ v_val,
                       int* __restrict__ todo_idx, int* __restrict__ todo_slot, int* __restrict__ todo_cnt) {
    __shared__ int s_rank[32];
    __shared__ float s_rv[32];
    __shared__ int s_n;
    const int i = blockIdx.x, s = threadIdx.x;
    if (s < 32) { s_rank[s] = s < k1p ? rank[(size_t)i * SSG_RANK_STRIDE + s] : -2; s_rv[s] = rank_val[(size_t)i * SSG_RANK_STRIDE + s]; }
    if (s == 0) s_n = 0;
    __syncthreads();
    const int c = v_cnt[i];
    if (s < c) {
        const int m = v_idx[(size_t)i * SSG_V_STRIDE + s];
        int hit = -1;
        for (int q = 0; q < k1p; ++q) if (s_rank[q] == m) hit = q;
        if (hit >= 0) {
            v_val[(size_t)i * SSG_V_STRIDE + s] = s_rv[hit];
        } else {
            const int t = atomicAdd(&s_n, 1);            // order inside the todo list is irrelevant
            todo_idx[(size_t)i * SSG_V_STRIDE + t] = m;
            todo_slot[(size_t)i * SSG_V_STRIDE + t] = s;
        }
    }
    __syncthreads();
    if (s == 0) todo_cnt[i] = s_n;
}
__global__ void __launch_bounds__(256)
krecip_scatter_kernel(const float* __restrict__ rowmax, const float* __restrict__ todo_od,
                      const int* __restrict__ todo_slot, const int* __restrict__ todo_cnt, float* __restrict__ v_val) {
    const int i = blockIdx.x, t = threadIdx.x;
    if (t < todo_cnt[i])
        v_val[(size_t)i * SSG_V_STRIDE + todo_slot[(size_t)i * SSG_V_STRIDE + t]] =
            __fdiv_rn(todo_od[(size_t)i * SSG_V_STRIDE + t], rowmax[i]);
}
int launch_krecip_classify(const int* rank, const float* rank_val, int n, int k1p, const int* v_idx, const int* v_cnt,
                           float* v_val, int* todo_idx, int* todo_slot, int* todo_cnt, cudaStream_t st) {
    krecip_classify_kernel<<<n, 256, 0, st>>>(rank, rank_val, k1p, v_idx, v_cnt, v_val, todo_idx, todo_slot, todo_cnt);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int launch_krecip_scatter(const float* rowmax, int n, const float* todo_od, const int* todo_slot, const int* todo_cnt,
                          float* v_val, cudaStream_t st) {
    krecip_scatter_kernel<<<n, 256, 0, st>>>(rowmax, todo_od, todo_slot, todo_cnt, v_val);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// `normalised` != 0: v_val already holds d2/rowmax on entry (else the squared distance d2).
__global__ void __launch_bounds__(128)
krecip_weights_kernel(const float* __restrict__ rowmax, int n, const int* __restrict__ v_cnt,
                      float* __restrict__ v_val, int normalised) {
    __shared__ float w[4][SSG_V_STRIDE];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + wid;
    if (i >= n) return;
    const int c = v_cnt[i];
    const float mx = rowmax[i];
    float* row = v_val + (size_t)i * SSG_V_STRIDE;
    for (int s = lane; s < c; s += 32) w[wid][s] = expf(-(normalised ? row[s] : __fdiv_rn(row[s], mx)));
    __syncwarp();
    float sum = 0.f;
    if (lane == 0) sum = np_pairwise_sum_f32(w[wid], c);
    sum = __shfl_sync(0xffffffffu, sum, 0);
    for (int s = lane; s < c; s += 32) row[s] = __fdiv_rn(w[wid][s], sum);
}

int launch_krecip_weights(const float* rowmax, int n, const int* v_cnt, float* v_val, int normalised,
                          cudaStream_t st) {
    krecip_weights_kernel<<<ssg_cdiv(n, 4), 128, 0, st>>>(rowmax, n, v_cnt, v_val, normalised);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// query expansion (rerank.py:94-98): V_qe[i] = mean(V[rank[i,:k2]], axis=0); np.mean reduces the k2
// rows sequentially in float32 and divides by k2.  One CTA per row: gather the k2 sparse rows,
// bitonic-sort by (column, j), sum each column's run in j order, compact.
// ---------------------------------------------------------------------------------------------------
constexpr int QE_NT = 256;
constexpr int QE_CAP = 2048;

__global__ void __launch_bounds__(QE_NT)
query_expand_kernel(const int* __restrict__ rank, int n, int k2, const int* __restrict__ v_idx,
                    const float* __restrict__ v_val, const int* __restrict__ v_cnt,
                    int* __restrict__ q_idx, float* __restrict__ q_val, int* __restrict__ q_cnt) {
    __shared__ uint32_t key[QE_CAP];
    __shared__ float val[QE_CAP];
    __shared__ int wsum[QE_NT / 32];
    __shared__ int s_base;
    const int i = blockIdx.x, tid = threadIdx.x;
    int off[9];
    int rows[8];
    off[0] = 0;
    for (int j = 0; j < k2; ++j) {
        rows[j] = rank[(size_t)i * SSG_RANK_STRIDE + j];
        off[j + 1] = off[j] + (rows[j] >= 0 ? v_cnt[rows[j]] : 0);
    }
    const int total = off[k2];
    int P = 32;
    while (P < total) P <<= 1;
    for (int t = tid; t < P; t += QE_NT) key[t] = 0xffffffffu;
    __syncthreads();
    for (int j = 0; j < k2; ++j) {
        const int r = rows[j];
        const int c = off[j + 1] - off[j];
        for (int s = tid; s < c; s += QE_NT) {
            key[off[j] + s] = ((uint32_t)v_idx[(size_t)r * SSG_V_STRIDE + s] << 3) | (uint32_t)j;
            val[off[j] + s] = v_val[(size_t)r * SSG_V_STRIDE + s];
        }
    }
    __syncthreads();
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < P; t += QE_NT) {
                const int p = t ^ j;
                if (p > t) {
                    const uint32_t a = key[t], b = key[p];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) {
                        key[t] = b; key[p] = a;
                        const float fa = val[t]; val[t] = val[p]; val[p] = fa;
                    }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) s_base = 0;
    __syncthreads();
    // np.mean over the rows rank[i, :k2] that exist (rerank.py:97): k2 of them unless the whole set is smaller than k2
    int nvalid = 0;
    for (int j = 0; j < k2; ++j) nvalid += rows[j] >= 0;
    const float kf = (float)nvalid;
    for (int base = 0; base < total; base += QE_NT) {
        const int e = base + tid;
        bool head = false;
        uint32_t col = 0;
        if (e < total) {
            col = key[e] >> 3;
            head = (e == 0) || ((key[e - 1] >> 3) != col);
        }
        int tot;
        const int pos = block_exclusive_scan<QE_NT>(head ? 1 : 0, wsum, tot);
        const int b0 = s_base;
        if (head) {
            float s = val[e];
            for (int q = e + 1; q < total && (key[q] >> 3) == col; ++q) s = __fadd_rn(s, val[q]);
            q_idx[(size_t)i * SSG_VQ_STRIDE + b0 + pos] = (int)col;
            q_val[(size_t)i * SSG_VQ_STRIDE + b0 + pos] = __fdiv_rn(s, kf);
        }
        __syncthreads();
        if (tid == 0) s_base = b0 + tot;
        __syncthreads();
    }
    if (tid == 0) q_cnt[i] = s_base;
}

int launch_query_expand(const int* rank, int n, int k2, const int* v_idx, const float* v_val,
                        const int* v_cnt, int* q_idx, float* q_val, int* q_cnt, cudaStream_t st) {
    if (k2 < 1 || k2 > 8 || k2 * 252 > QE_CAP)
        return ssg_set_error(SSG_ERR_INVALID, "k2=%d out of range (1..8)", k2);
    // (the callers check that k2 rows of the k1-dependent V bound fit the SSG_VQ_STRIDE slot: api.cu expanded_row_fits)
    query_expand_kernel<<<n, QE_NT, 0, st>>>(rank, n, k2, v_idx, v_val, v_cnt, q_idx, q_val, q_cnt);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// inverted index (rerank.py:101-103): CSC row lists of the expanded matrix.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
csc_count_kernel(int n, const int* __restrict__ q_idx, const int* __restrict__ q_cnt,
                 int* __restrict__ colcnt) {
    const int i = blockIdx.x;
    const int c = q_cnt[i];
    for (int s = threadIdx.x; s < c; s += blockDim.x) atomicAdd(&colcnt[q_idx[(size_t)i * SSG_VQ_STRIDE + s]], 1);
}
__global__ void __launch_bounds__(128)
csc_fill_kernel(int n, const int* __restrict__ q_idx, const int* __restrict__ q_cnt,
                const int* __restrict__ colptr, int* __restrict__ cursor, int* __restrict__ csc_row) {
    const int i = blockIdx.x;
    const int c = q_cnt[i];
    for (int s = threadIdx.x; s < c; s += blockDim.x) {
        const int k = q_idx[(size_t)i * SSG_VQ_STRIDE + s];
        csc_row[colptr[k] + atomicAdd(&cursor[k], 1)] = i;
    }
}

int launch_csc_build(int n, const int* q_idx, const int* q_cnt, int* colcnt, int* colptr, int* cursor,
                     int* csc_row, cudaStream_t st) {
    SSG_CUDA_TRY(cudaMemsetAsync(colcnt, 0, sizeof(int) * (size_t)n, st));
    SSG_CUDA_TRY(cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)n, st));
    csc_count_kernel<<<n, 128, 0, st>>>(n, q_idx, q_cnt, colcnt);
    SSG_CHECK_LAUNCH();
    SSG_TRY(launch_exclusive_scan_i32(colcnt, colptr, n, st));
    csc_fill_kernel<<<n, 128, 0, st>>>(n, q_idx, q_cnt, colptr, cursor, csc_row);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// Jaccard + final distance (rerank.py:105-122): one CTA per row i.
//   final[i,m] = fl32(J*fl32(1-lambda)) + fl32(v_i+v_m)*lambda   (float64),  J = 1 - S/(2-S),
//   S = sum over common columns k (ascending) of min(Vq[i,k], Vq[m,k]) in float32.
// Every column m that shares no k with row i has S = 0 (J = 1): the row is first filled with that
// base value; the touched columns (found through the inverted index, de-duplicated with a
// shared-memory bitmap) are then overwritten by the thread that claimed them, which merges the two
// index-sorted sparse rows.  Same terms in the same order for (i,m) and (m,i): exactly symmetric.
// ---------------------------------------------------------------------------------------------------
constexpr int JF_NT = 256;

// output policies: how a Jaccard value becomes the stored final distance
struct OutSourceF64 {        // rerank.py:122  final = J*(1-lambda) + (v_i+v_m)*lambda, float64 [n,n]
    const float* vec; float vi; double lambda_value; float oml; double* out;
    __device__ __forceinline__ bool wants(int m) const { return true; }
    __device__ __forceinline__ void store(int m, float J) const {
        const float Jm = __fmul_rn(J, oml);
        out[m] = __dadd_rn((double)Jm, __dmul_rn((double)__fadd_rn(vec[m], vi), lambda_value));
    }
};
struct OutInitF32 {          // rerank_initial.py:94,98  final = J*(1-lambda) + od_norm[i,m]*lambda, float32 [q,g]
    const float* drow; float rowmax; float lam; float oml; float* out; int q;
    __device__ __forceinline__ bool wants(int m) const { return m >= q; }
    __device__ __forceinline__ void store(int m, float J) const {
        const float od = __fdiv_rn(drow[m], rowmax);
        out[m - q] = __fadd_rn(__fmul_rn(J, oml), __fmul_rn(od, lam));
    }
};

template <class Out>
__device__ __forceinline__ void jaccard_row(int n, int i, const int* __restrict__ q_idx,
                                            const float* __restrict__ q_val, const int* __restrict__ q_cnt,
                                            const int* __restrict__ colptr, const int* __restrict__ csc_row,
                                            const Out& o, unsigned char* jf_smem) {
    int* si = reinterpret_cast<int*>(jf_smem);                       // [VQ_STRIDE]
    float* sv = reinterpret_cast<float*>(si + SSG_VQ_STRIDE);        // [VQ_STRIDE]
    int* pref = reinterpret_cast<int*>(sv + SSG_VQ_STRIDE);          // [VQ_STRIDE + 1]
    unsigned* bitmap = reinterpret_cast<unsigned*>(pref + SSG_VQ_STRIDE + 1);  // [(n+31)/32]
    __shared__ int wsum[JF_NT / 32];
    __shared__ int s_carry;

    const int tid = threadIdx.x;
    const int ci = q_cnt[i];
    const int nwords = (n + 31) >> 5;
    for (int s = tid; s < ci; s += JF_NT) {
        si[s] = q_idx[(size_t)i * SSG_VQ_STRIDE + s];
        sv[s] = q_val[(size_t)i * SSG_VQ_STRIDE + s];
    }
    for (int w = tid; w < nwords; w += JF_NT) bitmap[w] = 0u;
    // base fill (S = 0  ->  J = 1)
    for (int m = tid; m < n; m += JF_NT)
        if (o.wants(m)) o.store(m, 1.0f);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    // prefix sums of the inverted-list lengths of this row's columns
    for (int base = 0; base < ci; base += JF_NT) {
        const int s = base + tid;
        int len = 0;
        if (s < ci) len = colptr[si[s] + 1] - colptr[si[s]];
        int tot;
        const int ex = block_exclusive_scan<JF_NT>(len, wsum, tot);
        const int c = s_carry;
        if (s < ci) pref[s] = c + ex;
        __syncthreads();
        if (tid == 0) s_carry = c + tot;
        __syncthreads();
    }
    const int T = s_carry;
    if (tid == 0) pref[ci] = T;
    __syncthreads();

    for (int f = tid; f < T; f += JF_NT) {
        int lo = 0, hi = ci - 1;          // largest s with pref[s] <= f
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (pref[mid] <= f) lo = mid; else hi = mid - 1;
        }
        const int m = csc_row[colptr[si[lo]] + (f - pref[lo])];
        if (!o.wants(m)) continue;
        const unsigned bit = 1u << (m & 31);
        const unsigned old = atomicOr(&bitmap[m >> 5], bit);
        if (old & bit) continue;          // another thread owns column m
        const int cm = q_cnt[m];
        const int* mi = q_idx + (size_t)m * SSG_VQ_STRIDE;
        const float* mv = q_val + (size_t)m * SSG_VQ_STRIDE;
        float S = 0.f;
        int a = 0, b = 0;
        int ib = cm > 0 ? mi[0] : INT_MAX;
        while (a < ci && b < cm) {
            const int ia = si[a];
            if (ia == ib) {
                S = __fadd_rn(S, fminf(sv[a], mv[b]));
                ++a; ++b;
                ib = b < cm ? mi[b] : INT_MAX;
            } else if (ia < ib) {
                ++a;
            } else {
                ++b;
                ib = b < cm ? mi[b] : INT_MAX;
            }
        }
        float J = __fsub_rn(1.0f, __fdiv_rn(S, __fsub_rn(2.0f, S)));
        o.store(m, J);
    }
}

__global__ void __launch_bounds__(JF_NT)
jaccard_final_kernel(int n, int row0, const int* __restrict__ q_idx, const float* __restrict__ q_val,
                     const int* __restrict__ q_cnt, const int* __restrict__ colptr,
                     const int* __restrict__ csc_row, const float* __restrict__ vec, double lambda_value,
                     float one_minus_lambda, double* __restrict__ final_dist) {
    extern __shared__ unsigned char jf_smem[];
    const int i = row0 + (int)blockIdx.x;          // final_dist holds rows [row0, row0 + gridDim.x) of the matrix
    // rerank.py:117-118 clamps J < 0 to 0; J = 1 - S/(2-S) with 0 <= S <= 1(+rounding) cannot go below -1e-7,
    // and the clamp is applied inside store through fmaxf
    struct Clamp : OutSourceF64 {
        __device__ __forceinline__ void store(int m, float J) const { OutSourceF64::store(m, fmaxf(J, 0.f)); }
    };
    Clamp o;
    o.vec = vec; o.vi = vec[i]; o.lambda_value = lambda_value; o.oml = one_minus_lambda;
    o.out = final_dist + (size_t)blockIdx.x * n;
    jaccard_row(n, i, q_idx, q_val, q_cnt, colptr, csc_row, o, jf_smem);
}

// rerank_initial.py:85-98: only the first `q` rows, no clamp, float32, columns [q, n)
__global__ void __launch_bounds__(JF_NT)
jaccard_init_kernel(int n, int q, const int* __restrict__ q_idx, const float* __restrict__ q_val,
                    const int* __restrict__ q_cnt, const int* __restrict__ colptr, const int* __restrict__ csc_row,
                    const float* __restrict__ dmat, const float* __restrict__ rowmax, float lam, float oml,
                    float* __restrict__ out) {
    extern __shared__ unsigned char jf_smem[];
    const int i = blockIdx.x;
    OutInitF32 o{dmat + (size_t)i * n, rowmax[i], lam, oml, out + (size_t)i * (n - q), q};
    jaccard_row(n, i, q_idx, q_val, q_cnt, colptr, csc_row, o, jf_smem);
}

static size_t jaccard_smem(int n) {
    return sizeof(int) * SSG_VQ_STRIDE + sizeof(float) * SSG_VQ_STRIDE + sizeof(int) * (SSG_VQ_STRIDE + 1) +
           sizeof(unsigned) * (size_t)((n + 31) / 32 + 1);
}

int launch_jaccard_final(int n, int row0, int rows, const int* q_idx, const float* q_val, const int* q_cnt,
                         const int* colptr, const int* csc_row, const float* vec, double lambda_value,
                         double* final_dist, cudaStream_t st) {
    if (rows <= 0) return SSG_OK;
    const size_t smem = jaccard_smem(n);
    if (smem > 220 * 1024) return ssg_set_error(SSG_ERR_INVALID, "jaccard: n=%d too large for the bitmap", n);
    SSG_CUDA_TRY(cudaFuncSetAttribute(jaccard_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    // (1 - lambda_value) is a Python float cast to float32 by numpy when it multiplies the fp32 array
    const float oml = (float)(1.0 - lambda_value);
    jaccard_final_kernel<<<rows, JF_NT, smem, st>>>(n, row0, q_idx, q_val, q_cnt, colptr, csc_row, vec, lambda_value,
                                                    oml, final_dist);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int launch_jaccard_init(int n, int q, const int* q_idx, const float* q_val, const int* q_cnt, const int* colptr,
                        const int* csc_row, const float* dmat, const float* rowmax, double lambda_value, float* out,
                        cudaStream_t st) {
    const size_t smem = jaccard_smem(n);
    if (smem > 220 * 1024) return ssg_set_error(SSG_ERR_INVALID, "jaccard: n=%d too large for the bitmap", n);
    SSG_CUDA_TRY(cudaFuncSetAttribute(jaccard_init_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jaccard_init_kernel<<<q, JF_NT, smem, st>>>(n, q, q_idx, q_val, q_cnt, colptr, csc_row, dmat, rowmax,
                                                (float)lambda_value, (float)(1.0 - lambda_value), out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// Sparse form of final_dist (same arithmetic as jaccard_final_kernel).  A column m that shares no non-zero column with
// row i has J = 1 and final[i,m] = fl32(1 - lambda) + fl32(v_i + v_m) * lambda >= fl32(1 - lambda) (v >= 0): only the
// TOUCHED columns (~1 % of a row) can be smaller, and those are all eps and DBSCAN ever look at while eps stays below
// that bound.  Row i of the CSR = its touched columns in ascending order with their final values.
// One CTA per row, two launches: FILL = false counts the touched columns (bitmap popcount), FILL = true writes them.
// The position of a column inside its row is the rank of its bit in the bitmap, so the layout is deterministic.
// ---------------------------------------------------------------------------------------------------
template <bool FILL>
__global__ void __launch_bounds__(JF_NT)
jaccard_sparse_kernel(int n, const int* __restrict__ q_idx, const float* __restrict__ q_val,
                      const int* __restrict__ q_cnt, const int* __restrict__ colptr, const int* __restrict__ csc_row,
                      const float* __restrict__ vec, double lambda_value, float oml, const int* __restrict__ sp_rowptr,
                      int* __restrict__ sp_cnt, int* __restrict__ sp_col, double* __restrict__ sp_val) {
    extern __shared__ unsigned char jf_smem[];
    int* si = reinterpret_cast<int*>(jf_smem);                       // [VQ_STRIDE]
    float* sv = reinterpret_cast<float*>(si + SSG_VQ_STRIDE);        // [VQ_STRIDE]
    int* pref = reinterpret_cast<int*>(sv + SSG_VQ_STRIDE);          // [VQ_STRIDE + 1]
    const int nwords = (n + 31) >> 5;
    unsigned* bitmap = reinterpret_cast<unsigned*>(pref + SSG_VQ_STRIDE + 1);  // [nwords + 1]
    int* wpre = reinterpret_cast<int*>(bitmap + nwords + 1);         // [nwords + 1] exclusive popcount prefix
    __shared__ int wsum[JF_NT / 32];
    __shared__ int s_carry;

    const int i = blockIdx.x, tid = threadIdx.x;
    const int ci = q_cnt[i];
    for (int s = tid; s < ci; s += JF_NT) {
        si[s] = q_idx[(size_t)i * SSG_VQ_STRIDE + s];
        sv[s] = q_val[(size_t)i * SSG_VQ_STRIDE + s];
    }
    for (int w = tid; w < nwords; w += JF_NT) bitmap[w] = 0u;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ci; base += JF_NT) {                   // prefix sums of the inverted-list lengths
        const int s = base + tid;
        int len = 0;
        if (s < ci) len = colptr[si[s] + 1] - colptr[si[s]];
        int tot;
        const int ex = block_exclusive_scan<JF_NT>(len, wsum, tot);
        const int c = s_carry;
        if (s < ci) pref[s] = c + ex;
        __syncthreads();
        if (tid == 0) s_carry = c + tot;
        __syncthreads();
    }
    const int T = s_carry;
    if (tid == 0) pref[ci] = T;
    __syncthreads();
    for (int f = tid; f < T; f += JF_NT) {
        int lo = 0, hi = ci - 1;                                     // largest s with pref[s] <= f
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (pref[mid] <= f) lo = mid; else hi = mid - 1;
        }
        const int m = csc_row[colptr[si[lo]] + (f - pref[lo])];
        atomicOr(&bitmap[m >> 5], 1u << (m & 31));
    }
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < nwords; base += JF_NT) {               // rank of every word's first bit
        const int w = base + tid;
        const int pc = w < nwords ? __popc(bitmap[w]) : 0;
        int tot;
        const int ex = block_exclusive_scan<JF_NT>(pc, wsum, tot);
        const int c = s_carry;
        if (w < nwords) wpre[w] = c + ex;
        __syncthreads();
        if (tid == 0) s_carry = c + tot;
        __syncthreads();
    }
    const int cnt = s_carry;
    if (!FILL) {
        if (tid == 0) sp_cnt[i] = cnt;
        return;
    }
    const size_t out0 = (size_t)sp_rowptr[i];
    for (int w = tid; w < nwords; w += JF_NT) {
        unsigned bits = bitmap[w];
        int pos = wpre[w];
        while (bits) {
            const int b = __ffs(bits) - 1;
            sp_col[out0 + pos++] = (w << 5) + b;
            bits &= bits - 1u;
        }
    }
    __syncthreads();                                                 // the column list of this row is complete
    const float vi = vec[i];
    for (int e = tid; e < cnt; e += JF_NT) {
        const int m = sp_col[out0 + e];
        const int cm = q_cnt[m];
        const int* mi = q_idx + (size_t)m * SSG_VQ_STRIDE;
        const float* mv = q_val + (size_t)m * SSG_VQ_STRIDE;
        float S = 0.f;
        int a = 0, b = 0;
        int ib = cm > 0 ? mi[0] : INT_MAX;
        while (a < ci && b < cm) {                                   // same merge, same order as jaccard_row
            const int ia = si[a];
            if (ia == ib) {
                S = __fadd_rn(S, fminf(sv[a], mv[b]));
                ++a; ++b;
                ib = b < cm ? mi[b] : INT_MAX;
            } else if (ia < ib) {
                ++a;
            } else {
                ++b;
                ib = b < cm ? mi[b] : INT_MAX;
            }
        }
        const float J = fmaxf(__fsub_rn(1.0f, __fdiv_rn(S, __fsub_rn(2.0f, S))), 0.f);
        const float Jm = __fmul_rn(J, oml);
        sp_val[out0 + e] = __dadd_rn((double)Jm, __dmul_rn((double)__fadd_rn(vec[m], vi), lambda_value));
    }
}

static size_t jaccard_sparse_smem(int n) {
    return sizeof(int) * SSG_VQ_STRIDE + sizeof(float) * SSG_VQ_STRIDE + sizeof(int) * (SSG_VQ_STRIDE + 1) +
           2 * sizeof(unsigned) * (size_t)((n + 31) / 32 + 1);
}

// sp_rowptr == NULL: count pass (sp_cnt[i] = touched columns of row i); else fill pass.
int launch_jaccard_sparse(int n, const int* q_idx, const float* q_val, const int* q_cnt, const int* colptr,
                          const int* csc_row, const float* vec, double lambda_value, const int* sp_rowptr, int* sp_cnt,
                          int* sp_col, double* sp_val, cudaStream_t st) {
    const size_t smem = jaccard_sparse_smem(n);
    if (smem > 220 * 1024) return ssg_set_error(SSG_ERR_INVALID, "jaccard: n=%d too large for the bitmap", n);
    const float oml = (float)(1.0 - lambda_value);
    if (!sp_rowptr) {
        SSG_CUDA_TRY(cudaFuncSetAttribute(jaccard_sparse_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jaccard_sparse_kernel<false><<<n, JF_NT, smem, st>>>(n, q_idx, q_val, q_cnt, colptr, csc_row, vec, lambda_value, oml,
                                                             nullptr, sp_cnt, nullptr, nullptr);
    } else {
        SSG_CUDA_TRY(cudaFuncSetAttribute(jaccard_sparse_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        jaccard_sparse_kernel<true><<<n, JF_NT, smem, st>>>(n, q_idx, q_val, q_cnt, colptr, csc_row, vec, lambda_value, oml,
                                                            sp_rowptr, sp_cnt, sp_col, sp_val);
    }
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// re_ranking_lh (reid/rerank_plain.py:27-123): reid/rerank.py's re_ranking with a float64 source term built on the
// UN-squared distances: v_i = min_j cdist(t_i, s_j) (float64, sequential sum as cdist), v /= max(v), source_dist =
// v_i + v_j in float64; final = fl32(J * fl32(1 - lambda)) + (v_m + v_i) * lambda.  sqrt is monotone and correctly
// rounded, so min_j sqrt(sum_j) = sqrt(min_j sum_j): the kernel keeps the smallest float64 sum per target row.
// ---------------------------------------------------------------------------------------------------
constexpr int LH_NT = 128;
__global__ void __launch_bounds__(LH_NT)
source_min_f64_kernel(const float* __restrict__ tgt, const float* __restrict__ src, int ns, int d,
                      double* __restrict__ out_minsum) {
    extern __shared__ unsigned char jf_smem[];
    double* a = reinterpret_cast<double*>(jf_smem);          // [d] target row as float64
    __shared__ double red[LH_NT];
    const int i = blockIdx.x, tid = threadIdx.x;
    for (int k = tid; k < d; k += LH_NT) a[k] = (double)tgt[(size_t)i * d + k];
    __syncthreads();
    double best = __longlong_as_double(0x7ff0000000000000ll);        // +inf
    for (int j = tid; j < ns; j += LH_NT) {
        const float* b = src + (size_t)j * d;
        double acc = 0.0;
        for (int k = 0; k < d; ++k) {
            const double df = __dsub_rn(a[k], (double)b[k]);
            acc = __dadd_rn(acc, __dmul_rn(df, df));
        }
        best = acc < best ? acc : best;
    }
    red[tid] = best;
    __syncthreads();
    for (int o = LH_NT / 2; o > 0; o >>= 1) {
        if (tid < o) red[tid] = red[tid + o] < red[tid] ? red[tid + o] : red[tid];
        __syncthreads();
    }
    if (tid == 0) out_minsum[i] = red[0];
}

// v = sqrt(minsum); v /= max(v)   (one CTA; float64 throughout, rerank_plain.py:37-38)
__global__ void __launch_bounds__(1024)
source_vec_f64_kernel(const double* __restrict__ minsum, int n, double* __restrict__ vec) {
    __shared__ double red[1024];
    const int tid = threadIdx.x;
    double mx = 0.0;
    for (int i = tid; i < n; i += 1024) {
        const double v = sqrt(minsum[i]);
        vec[i] = v;
        mx = v > mx ? v : mx;
    }
    red[tid] = mx;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) red[tid] = red[tid + o] > red[tid] ? red[tid + o] : red[tid];
        __syncthreads();
    }
    const double m = red[0];
    for (int i = tid; i < n; i += 1024) vec[i] = vec[i] / m;
}

struct OutSourceLH {         // rerank_plain.py:41-43,120: final = J*(1-lambda) + (v_m + v_i)*lambda with float64 v
    const double* vec; double vi; double lambda_value; float oml; double* out;
    __device__ __forceinline__ bool wants(int m) const { return true; }
    __device__ __forceinline__ void store(int m, float J) const {
        const float Jm = __fmul_rn(fmaxf(J, 0.f), oml);
        out[m] = __dadd_rn((double)Jm, __dmul_rn(__dadd_rn(vec[m], vi), lambda_value));
    }
};

__global__ void __launch_bounds__(JF_NT)
jaccard_final_lh_kernel(int n, const int* __restrict__ q_idx, const float* __restrict__ q_val,
                        const int* __restrict__ q_cnt, const int* __restrict__ colptr, const int* __restrict__ csc_row,
                        const double* __restrict__ vec, double lambda_value, float oml, double* __restrict__ final_dist) {
    extern __shared__ unsigned char jf_smem[];
    const int i = blockIdx.x;
    OutSourceLH o{vec, vec[i], lambda_value, oml, final_dist + (size_t)i * n};
    jaccard_row(n, i, q_idx, q_val, q_cnt, colptr, csc_row, o, jf_smem);
}

int launch_source_vec_f64(const float* tgt, int n, const float* src, int ns, int d, double* minsum, double* vec,
                          cudaStream_t st) {
    const size_t smem = sizeof(double) * (size_t)d;
    if (smem > 200 * 1024) return ssg_set_error(SSG_ERR_INVALID, "rerank_lh: d=%d too large", d);
    SSG_CUDA_TRY(cudaFuncSetAttribute(source_min_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    source_min_f64_kernel<<<n, LH_NT, smem, st>>>(tgt, src, ns, d, minsum);
    SSG_CHECK_LAUNCH();
    source_vec_f64_kernel<<<1, 1024, 0, st>>>(minsum, n, vec);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int launch_jaccard_final_lh(int n, const int* q_idx, const float* q_val, const int* q_cnt, const int* colptr,
                            const int* csc_row, const double* vec, double lambda_value, double* final_dist,
                            cudaStream_t st) {
    const size_t smem = jaccard_smem(n);
    if (smem > 220 * 1024) return ssg_set_error(SSG_ERR_INVALID, "jaccard: n=%d too large for the bitmap", n);
    SSG_CUDA_TRY(cudaFuncSetAttribute(jaccard_final_lh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jaccard_final_lh_kernel<<<n, JF_NT, smem, st>>>(n, q_idx, q_val, q_cnt, colptr, csc_row, vec, lambda_value,
                                                    (float)(1.0 - lambda_value), final_dist);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// Plain kNN-set re-ranking: reid/rerank_plain.py:125-178 re_ranking(source, target, k, lambda).
//   S_i    = { j != i : od[i,j] <= k-th smallest entry of row i of od }   (od = squared distances, diagonal included)
//   J[i,j] = scipy cdist(S, S, 'jaccard') = (|S_i u S_j| - |S_i n S_j|) / |S_i u S_j|   (0 when both are empty)
//   final  = fl(J * fl(1 - lambda)) + fl(v_i + v_j) * lambda, v as in reid/rerank.py:36-40.
// The sets come from the rank table of the distance stages (sorted by (od / rowmax, index); the division is
// monotone, so while the k-th and (k+1)-th normalised values differ strictly the first k columns ARE the k smallest
// raw distances); rows whose boundary is tied are handed to the exact fallback (knn_scan_kernel on recomputed rows).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
knn_sets_kernel(const int* __restrict__ rank, const float* __restrict__ rank_val, int n, int k,
                int* __restrict__ set_idx, int* __restrict__ set_cnt, int* __restrict__ flag_rows,
                int* __restrict__ flag_cnt) {
    const int i = blockIdx.x, lane = threadIdx.x;
    constexpr int RS = SSG_RANK_STRIDE;
    const bool tied = n > k && !(rank_val[(size_t)i * RS + k - 1] < rank_val[(size_t)i * RS + k]);
    if (tied) {
        if (lane == 0) { flag_rows[atomicAdd(flag_cnt, 1)] = i; set_cnt[i] = 0; }
        return;
    }
    int c = lane < k ? rank[(size_t)i * RS + lane] : INT_MAX;
    if (c == i || c < 0) c = INT_MAX;                    // knn_bool[i, i] = False
    int pos = 0;                                          // ascending index order (the Jaccard kernel merges sorted lists)
    for (int l = 0; l < 32; ++l) {
        const int o = __shfl_sync(0xffffffffu, c, l);
        pos += (o < c) || (o == c && l < lane);
    }
    if (c != INT_MAX) set_idx[(size_t)i * SSG_VQ_STRIDE + pos] = c;
    const unsigned have = __ballot_sync(0xffffffffu, c != INT_MAX);
    if (lane == 0) set_cnt[i] = __popc(have);
}

// exact fallback: block b = flagged row rows_list[b], whose raw squared distances are row b of M and whose k smallest
// raw values (sorted) are sel_val[b, 0..k): S = { j != i : M[b, j] <= sel_val[b, k-1] } in ascending j.
__global__ void __launch_bounds__(256)
knn_scan_kernel(const float* __restrict__ M, size_t ld, int cols, const float* __restrict__ sel_val, int sel_stride,
                int k, const int* __restrict__ rows_list, int* __restrict__ set_idx, int* __restrict__ set_cnt,
                int* __restrict__ overflow) {
    __shared__ int wsum[8];
    __shared__ int s_carry;
    const int b = blockIdx.x, tid = threadIdx.x;
    const int i = rows_list[b];
    const float thr = sel_val[(size_t)b * sel_stride + k - 1];
    const float* row = M + (size_t)b * ld;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int j0 = 0; j0 < cols; j0 += 256) {
        const int j = j0 + tid;
        const int hit = (j < cols && j != i && row[j] <= thr) ? 1 : 0;
        int tot;
        const int ex = block_exclusive_scan<256>(hit, wsum, tot);
        const int base = s_carry;
        if (hit) {
            if (base + ex < SSG_VQ_STRIDE) set_idx[(size_t)i * SSG_VQ_STRIDE + base + ex] = j;
            else *overflow = 1;
        }
        __syncthreads();
        if (tid == 0) s_carry = base + tot;
        __syncthreads();
    }
    if (tid == 0) set_cnt[i] = s_carry < SSG_VQ_STRIDE ? s_carry : SSG_VQ_STRIDE;
}

__global__ void __launch_bounds__(JF_NT)
jaccard_plain_kernel(int n, const int* __restrict__ s_idx, const int* __restrict__ s_cnt,
                     const int* __restrict__ colptr, const int* __restrict__ csc_row, const float* __restrict__ vec,
                     double lambda_value, float oml, double* __restrict__ final_dist) {
    extern __shared__ unsigned char jf_smem[];
    int* si = reinterpret_cast<int*>(jf_smem);                       // [VQ_STRIDE]
    int* pref = si + SSG_VQ_STRIDE;                                  // [VQ_STRIDE + 1]
    unsigned* bitmap = reinterpret_cast<unsigned*>(pref + SSG_VQ_STRIDE + 1);
    __shared__ int wsum[JF_NT / 32];
    __shared__ int s_carry;
    const int i = blockIdx.x, tid = threadIdx.x;
    const int ci = s_cnt[i];
    const int nwords = (n + 31) >> 5;
    const float vi = vec[i];
    double* out = final_dist + (size_t)i * n;
    auto store = [&](int m, float J) {
        const float Jm = __fmul_rn(J, oml);
        out[m] = __dadd_rn((double)Jm, __dmul_rn((double)__fadd_rn(vec[m], vi), lambda_value));
    };
    for (int s = tid; s < ci; s += JF_NT) si[s] = s_idx[(size_t)i * SSG_VQ_STRIDE + s];
    for (int w = tid; w < nwords; w += JF_NT) bitmap[w] = 0u;
    // no common neighbour: J = 1 (0 when both sets are empty, as scipy defines the distance of two all-False rows)
    for (int m = tid; m < n; m += JF_NT) store(m, (ci + s_cnt[m]) > 0 ? 1.0f : 0.0f);
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < ci; base += JF_NT) {
        const int s = base + tid;
        int len = 0;
        if (s < ci) len = colptr[si[s] + 1] - colptr[si[s]];
        int tot;
        const int ex = block_exclusive_scan<JF_NT>(len, wsum, tot);
        const int c = s_carry;
        if (s < ci) pref[s] = c + ex;
        __syncthreads();
        if (tid == 0) s_carry = c + tot;
        __syncthreads();
    }
    const int T = s_carry;
    if (tid == 0) pref[ci] = T;
    __syncthreads();
    for (int f = tid; f < T; f += JF_NT) {
        int lo = 0, hi = ci - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (pref[mid] <= f) lo = mid; else hi = mid - 1;
        }
        const int m = csc_row[colptr[si[lo]] + (f - pref[lo])];
        const unsigned bit = 1u << (m & 31);
        if (atomicOr(&bitmap[m >> 5], bit) & bit) continue;          // another thread owns column m
        const int cm = s_cnt[m];
        const int* mi = s_idx + (size_t)m * SSG_VQ_STRIDE;
        int inter = 0, a = 0, b = 0;
        while (a < ci && b < cm) {
            const int ia = si[a], ib = mi[b];
            if (ia == ib) { ++inter; ++a; ++b; }
            else if (ia < ib) ++a;
            else ++b;
        }
        const int uni = ci + cm - inter;
        // scipy: (number of positions where exactly one is set) / (number where at least one is set), in double
        store(m, uni > 0 ? (float)((double)(uni - inter) / (double)uni) : 0.0f);
    }
}

int launch_knn_sets(const int* rank, const float* rank_val, int n, int k, int* set_idx, int* set_cnt, int* flag_rows,
                    int* flag_cnt, cudaStream_t st) {
    if (k < 1 || k > 31) return ssg_set_error(SSG_ERR_INVALID, "rerank_plain: k=%d out of range (1..31)", k);
    knn_sets_kernel<<<n, 32, 0, st>>>(rank, rank_val, n, k, set_idx, set_cnt, flag_rows, flag_cnt);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int launch_knn_scan(const float* M, size_t ld, int rows, int cols, const float* sel_val, int sel_stride, int k,
                    const int* rows_list, int* set_idx, int* set_cnt, int* overflow, cudaStream_t st) {
    if (rows <= 0) return SSG_OK;
    knn_scan_kernel<<<rows, 256, 0, st>>>(M, ld, cols, sel_val, sel_stride, k, rows_list, set_idx, set_cnt, overflow);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int launch_jaccard_plain(int n, const int* s_idx, const int* s_cnt, const int* colptr, const int* csc_row,
                         const float* vec, double lambda_value, double* final_dist, cudaStream_t st) {
    const size_t smem = sizeof(int) * (2 * SSG_VQ_STRIDE + 1) + sizeof(unsigned) * (size_t)((n + 31) / 32 + 1);
    if (smem > 220 * 1024) return ssg_set_error(SSG_ERR_INVALID, "jaccard: n=%d too large for the bitmap", n);
    SSG_CUDA_TRY(cudaFuncSetAttribute(jaccard_plain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    jaccard_plain_kernel<<<n, JF_NT, smem, st>>>(n, s_idx, s_cnt, colptr, csc_row, vec, lambda_value,
                                                 (float)(1.0 - lambda_value), final_dist);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// re_ranking_init front end (rerank_initial.py:43-48): D'[r,c] = 2 - 2*S[c,r] with S = [[qq, qg],[qg^T, gg]]
// (the reference normalises by the column max and transposes; D' is that transpose before the division).
// ---------------------------------------------------------------------------------------------------
__global__ void init_assemble_kernel(const float* __restrict__ qg, const float* __restrict__ qq,
                                     const float* __restrict__ gg, int q, int g, float* __restrict__ dt) {
    const int n = q + g;
    const size_t total = (size_t)n * n;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(e / n), c = (int)(e % n);     // D'[r,c] = od[c,r]
        float s;
        if (c < q) s = r < q ? qq[(size_t)c * q + r] : qg[(size_t)c * g + (r - q)];
        else       s = r < q ? qg[(size_t)r * g + (c - q)] : gg[(size_t)(c - q) * g + (r - q)];
        dt[e] = __fsub_rn(2.0f, __fmul_rn(2.0f, s));
    }
}
int launch_init_assemble(const float* qg, const float* qq, const float* gg, int q, int g, float* dt, cudaStream_t st) {
    const size_t total = (size_t)(q + g) * (q + g);
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    init_assemble_kernel<<<grid, 256, 0, st>>>(qg, qq, gg, q, g, dt);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// v_val[i, s] = M[i, v_idx[i, s]]
__global__ void gather_row_vals_kernel(const float* __restrict__ M, size_t ld, const int* __restrict__ idx,
                                       const int* __restrict__ cnt, int stride, float* __restrict__ out) {
    const int i = blockIdx.x;
    for (int s = threadIdx.x; s < cnt[i]; s += blockDim.x)
        out[(size_t)i * stride + s] = M[(size_t)i * ld + idx[(size_t)i * stride + s]];
}
int launch_gather_row_vals(const float* M, size_t ld, int rows, const int* idx, const int* cnt, int stride,
                           float* out, cudaStream_t st) {
    gather_row_vals_kernel<<<rows, 64, 0, st>>>(M, ld, idx, cnt, stride, out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // namespace ssg
