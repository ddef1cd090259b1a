// Distance-side kernels of the re-ranking path.
//   * sqdist_exact      : bit-identical restatement of scipy cdist + fp32 squaring (rerank.py:37,61-62)
//   * row_minmax        : row minimum / maximum of an fp32 matrix block       (rerank.py:39,68)
//   * row_select        : K smallest entries of every row ordered by (value, index), i.e. the
//                         leading columns of a stable argsort                 (rerank.py:70)
//   * pair_exact        : exact distance for sparse (row, column) pairs      (rerank.py:91 gathers)
#include <limits.h>

#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace ssg {

// out = fl32(fl32(sqrt(s))^2)
__device__ __forceinline__ float finish_sqdist(double s) {
    float r = (float)sqrt(s);
    return __fmul_rn(r, r);
}

// ---------------------------------------------------------------------------------------------------
// sqdist_exact: 64x64 output tile per CTA, 4x4 outputs per thread, float64 accumulation in the exact
// order cdist uses (sequential over k, separate multiply and add: no FMA contraction).
// FP64-pipe bound by design: this is the reference-exact path (fallback + verification); the fast path
// is the tensor-core GEMM in gemm_tc.cu.
// ---------------------------------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 16;

// DOT = true: out = fl32( sum_k x_ik * y_jk ) instead, the products exact in float64 and summed sequentially in
// float64 (the dot-product blocks of the cosine re-ranking, np.dot at reid/rerank.py:174-176, reid/eug.py:223-225; at
// least as accurate as the reference's float32 GEMM).
template <bool DOT>
__global__ void __launch_bounds__(256)
sqdist_exact_kernel(const float* __restrict__ X, int nx, const float* __restrict__ Y, int ny, int d,
                    float* __restrict__ out, size_t ldo) {
    __shared__ double sx[TK][TM];
    __shared__ double sy[TK][TN];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row0 = blockIdx.y * TM, col0 = blockIdx.x * TN;
    const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;
    const bool vec_ok = (d & 3) == 0;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;

    for (int k0 = 0; k0 < d; k0 += TK) {
        float vx[4] = {0.f, 0.f, 0.f, 0.f}, vy[4] = {0.f, 0.f, 0.f, 0.f};
        const int gx = row0 + lr, gy = col0 + lr, kk = k0 + lk;
        if (gx < nx) {
            const float* p = X + (size_t)gx * d + kk;
            if (vec_ok && kk + 3 < d) {
                float4 t = *reinterpret_cast<const float4*>(p);
                vx[0] = t.x; vx[1] = t.y; vx[2] = t.z; vx[3] = t.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (kk + q < d) vx[q] = p[q];
            }
        }
        if (gy < ny) {
            const float* p = Y + (size_t)gy * d + kk;
            if (vec_ok && kk + 3 < d) {
                float4 t = *reinterpret_cast<const float4*>(p);
                vy[0] = t.x; vy[1] = t.y; vy[2] = t.z; vy[3] = t.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) if (kk + q < d) vy[q] = p[q];
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            sx[lk + q][lr] = (double)vx[q];
            sy[lk + q][lr] = (double)vy[q];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TK; ++k) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sx[k][ty + 16 * i]; b[i] = sy[k][tx + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (DOT) {
                        acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(a[i], b[j]));
                    } else {
                        const double df = __dsub_rn(a[i], b[j]);
                        acc[i][j] = __dadd_rn(acc[i][j], __dmul_rn(df, df));
                    }
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = row0 + ty + 16 * i;
        if (r >= nx) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = col0 + tx + 16 * j;
            if (c < ny) out[(size_t)r * ldo + c] = DOT ? (float)acc[i][j] : finish_sqdist(acc[i][j]);
        }
    }
}

int launch_sqdist_exact(const float* X, int nx, const float* Y, int ny, int d, float* out, size_t ldo,
                        cudaStream_t st) {
    if (nx <= 0 || ny <= 0) return SSG_OK;
    dim3 grid(ssg_cdiv(ny, TN), ssg_cdiv(nx, TM));
    sqdist_exact_kernel<false><<<grid, 256, 0, st>>>(X, nx, Y, ny, d, out, ldo);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int launch_dot_exact(const float* X, int nx, const float* Y, int ny, int d, float* out, size_t ldo,
                     cudaStream_t st) {
    if (nx <= 0 || ny <= 0) return SSG_OK;
    dim3 grid(ssg_cdiv(ny, TN), ssg_cdiv(nx, TM));
    sqdist_exact_kernel<true><<<grid, 256, 0, st>>>(X, nx, Y, ny, d, out, ldo);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// row_minmax: one CTA per row.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
row_minmax_kernel(const float* __restrict__ M, size_t ld, int cols, float* __restrict__ rmin,
                  float* __restrict__ rmax) {
    const float* row = M + (size_t)blockIdx.x * ld;
    float lo = INFINITY, hi = -INFINITY;
    for (int j = threadIdx.x; j < cols; j += blockDim.x) {
        const float v = row[j];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    __shared__ float slo[8], shi[8];
    lo = warp_min_f(lo);
    hi = warp_max_f(hi);
    if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) { lo = fminf(lo, slo[w]); hi = fmaxf(hi, shi[w]); }
        if (rmin) rmin[blockIdx.x] = lo;
        if (rmax) rmax[blockIdx.x] = hi;
    }
}

int launch_row_minmax(const float* M, size_t ld, int rows, int cols, float* rmin, float* rmax,
                      cudaStream_t st) {
    if (rows <= 0) return SSG_OK;
    row_minmax_kernel<<<rows, 256, 0, st>>>(M, ld, cols, rmin, rmax);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// row_select: K smallest (value, index) pairs of each row, ascending — the first K columns of
// np.argsort(row, kind='stable').  value = M[j] / scale[row] (IEEE division) when `scale` is given.
// Three-pass radix select on the order-preserving 32-bit key in shared memory, then an
// index-ordered collection of threshold ties, then a rank sort of the K survivors.
// ---------------------------------------------------------------------------------------------------
constexpr int SEL_NT = 256;
constexpr int SEL_BINS = 2048;
constexpr int SEL_KMAX = 64;

__device__ __forceinline__ float sel_value(const float* row, int j, bool scaled, float s) {
    const float v = row[j];
    return scaled ? __fdiv_rn(v, s) : v;
}
// `largest` selects by descending value (ties still by ascending index): complement the key.
__device__ __forceinline__ uint32_t sel_key(float v, bool largest) {
    const uint32_t k = f32_key(v);
    return largest ? ~k : k;
}

__global__ void __launch_bounds__(SEL_NT)
row_select_kernel(const float* __restrict__ M, size_t ld, int cols, const float* __restrict__ scale,
                  int K, bool largest, int* __restrict__ out_idx, float* __restrict__ out_val, int out_stride,
                  const int* __restrict__ row_done) {
    if (row_done && row_done[blockIdx.x]) return;     // already served by the sampled fast path
    __shared__ int hist[SEL_BINS];
    __shared__ int wsum[SEL_NT / 32];
    __shared__ uint32_t s_prefix;
    __shared__ int s_remaining;
    __shared__ uint32_t c_key[SEL_KMAX];
    __shared__ int c_idx[SEL_KMAX];
    __shared__ int s_nless, s_ties_seen;

    const int row_id = blockIdx.x;
    const float* row = M + (size_t)row_id * ld;
    const bool scaled = scale != nullptr;
    const float s = scaled ? scale[row_id] : 1.0f;
    const int Keff = min(K, cols);
    const int tid = threadIdx.x;

    if (tid == 0) { s_prefix = 0u; s_remaining = Keff; s_nless = 0; s_ties_seen = 0; }
    // pass p examines bits [shift, shift+width)
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int p = 0; p < 3; ++p) {
        for (int b = tid; b < SEL_BINS; b += SEL_NT) hist[b] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        const int shift = shifts[p], width = widths[p];
        const uint32_t mask = (1u << width) - 1u;
        for (int j = tid; j < cols; j += SEL_NT) {
            const uint32_t k = sel_key(sel_value(row, j, scaled, s), largest);
            const bool in = (p == 0) || ((k >> (shift + width)) == (prefix >> (shift + width)));
            if (in) atomicAdd(&hist[(k >> shift) & mask], 1);
        }
        __syncthreads();
        // find the bin holding the `remaining`-th smallest element
        constexpr int PER = SEL_BINS / SEL_NT;  // 8 bins per thread
        int local[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) { local[q] = hist[tid * PER + q]; sum += local[q]; }
        int total;
        int excl = block_exclusive_scan<SEL_NT>(sum, wsum, total);
        const int rem = s_remaining;
        __syncthreads();
        if (excl < rem && rem <= excl + sum) {
            int c = excl;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                if (c < rem && rem <= c + local[q]) {
                    s_prefix = prefix | ((uint32_t)(tid * PER + q) << shift);
                    s_remaining = rem - c;
                }
                c += local[q];
            }
        }
        __syncthreads();
    }
    const uint32_t thr = s_prefix;     // exact key of the Keff-th smallest element
    const int take_ties = s_remaining; // how many elements equal to thr belong to the result

    // collect: everything below thr, plus the first `take_ties` ties in index order
    for (int base = 0; base < cols; base += SEL_NT) {
        const int j = base + tid;
        uint32_t k = 0xffffffffu;
        bool valid = j < cols;
        if (valid) k = sel_key(sel_value(row, j, scaled, s), largest);
        const bool less = valid && k < thr;
        const bool tie = valid && k == thr;
        if (less) {
            const int pos = atomicAdd(&s_nless, 1);
            c_key[pos] = k;
            c_idx[pos] = j;
        }
        if (__syncthreads_or(tie)) {
            int total;
            const int pos = block_exclusive_scan<SEL_NT>(tie ? 1 : 0, wsum, total);
            const int seen = s_ties_seen;
            __syncthreads();
            if (tie && seen + pos < take_ties) {
                const int slot = (Keff - take_ties) + seen + pos;
                c_key[slot] = k;
                c_idx[slot] = j;
            }
            if (tid == 0) s_ties_seen = seen + total;
            __syncthreads();
        }
    }
    __syncthreads();
    // rank sort of the Keff survivors by (key, idx)
    if (tid < Keff) {
        const uint32_t k = c_key[tid];
        const int ix = c_idx[tid];
        int r = 0;
        for (int q = 0; q < Keff; ++q) {
            const uint32_t kq = c_key[q];
            r += (kq < k) || (kq == k && c_idx[q] < ix);
        }
        out_idx[(size_t)row_id * out_stride + r] = ix;
        out_val[(size_t)row_id * out_stride + r] = sel_value(row, ix, scaled, s);
    }
    for (int q = Keff + tid; q < K; q += SEL_NT) {   // short rows: pad
        out_idx[(size_t)row_id * out_stride + q] = -1;
        out_val[(size_t)row_id * out_stride + q] = INFINITY;
    }
}


// ---------------------------------------------------------------------------------------------------
// row_select, sampled fast path: a sorted 512-element sample of the row gives a pivot that is, with overwhelming
// probability, above the K-th smallest element and below ~2 % of the row; ONE pass over the row then collects
// everything <= pivot into shared memory, and the K smallest by (value, index) are ranked there.  Rows for which
// the pivot turns out too low (fewer than K collected) or too high (list overflow) are left to the radix-select
// kernel above (row_done[row] = 0), so the result is always exact.
// ---------------------------------------------------------------------------------------------------
constexpr int SMP_N = 256;
constexpr int SMP_CAP = 1024;

__global__ void __launch_bounds__(SEL_NT)
row_select_sampled_kernel(const float* __restrict__ M, size_t ld, int cols, const float* __restrict__ scale,
                          int K, bool largest, int* __restrict__ out_idx, float* __restrict__ out_val,
                          int out_stride, int* __restrict__ row_done) {
    __shared__ uint32_t samp[SMP_N];
    __shared__ uint32_t l_key[SMP_CAP];
    __shared__ int l_idx[SMP_CAP];
    __shared__ int s_cnt;
    const int row_id = blockIdx.x, tid = threadIdx.x;
    const float* row = M + (size_t)row_id * ld;
    const bool scaled = scale != nullptr;
    const float s = scaled ? scale[row_id] : 1.0f;
    const int Keff = min(K, cols);
    if (cols <= SMP_CAP / 2) {                   // short rows: not worth sampling
        if (tid == 0) row_done[row_id] = 0;
        return;
    }
    for (int t = tid; t < SMP_N; t += SEL_NT) {
        const int j = (int)(((long long)t * cols) / SMP_N);
        samp[t] = sel_key(sel_value(row, j, scaled, s), largest);
    }
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    for (int k = 2; k <= SMP_N; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < SMP_N; t += SEL_NT) {
                const int p = t ^ j;
                if (p > t) {
                    const uint32_t a = samp[t], b = samp[p];
                    const bool up = (t & k) == 0;
                    if ((a > b) == up) { samp[t] = b; samp[p] = a; }
                }
            }
            __syncthreads();
        }
    }
    // expected rank of the K-th smallest in the sample is K*SMP_N/cols; take 3x that plus a safety margin
    int pi = (int)(((long long)Keff * SMP_N * 2 + cols - 1) / cols) + 5;
    if (pi > SMP_N - 1) pi = SMP_N - 1;
    const uint32_t pivot = samp[pi];
    // the row is streamed with 8 independent loads in flight per thread (the loop is latency bound otherwise)
    for (int j0 = tid; j0 < cols; j0 += SEL_NT * 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u * SEL_NT;
            v[u] = j < cols ? row[j] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int j = j0 + u * SEL_NT;
            if (j < cols) {
                const uint32_t k = sel_key(scaled ? __fdiv_rn(v[u], s) : v[u], largest);
                if (k <= pivot) {
                    const int pos = atomicAdd(&s_cnt, 1);
                    if (pos < SMP_CAP) { l_key[pos] = k; l_idx[pos] = j; }
                }
            }
        }
    }
    __syncthreads();
    const int c = s_cnt;
    if (c > SMP_CAP || c < Keff) {
        if (tid == 0) row_done[row_id] = 0;
        return;
    }
    // K smallest of the c collected (key, index) pairs: 3-pass radix select on the 32-bit keys inside shared memory,
    // ties at the threshold key resolved by index, then the K survivors are rank-sorted (K^2 comparisons only).
    __shared__ int hist[SEL_BINS];
    __shared__ int wsum2[SEL_NT / 32];
    __shared__ uint32_t s_prefix2;
    __shared__ int s_rem2, s_nsel;
    __shared__ uint32_t c_key[SEL_KMAX];
    __shared__ int c_idx[SEL_KMAX];
    if (tid == 0) { s_prefix2 = 0u; s_rem2 = Keff; s_nsel = 0; }
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int p = 0; p < 3; ++p) {
        for (int bq = tid; bq < SEL_BINS; bq += SEL_NT) hist[bq] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix2;
        const int shift = shifts[p], width = widths[p];
        const uint32_t mask = (1u << width) - 1u;
        for (int e = tid; e < c; e += SEL_NT) {
            const uint32_t k = l_key[e];
            if (p == 0 || ((k >> (shift + width)) == (prefix >> (shift + width)))) atomicAdd(&hist[(k >> shift) & mask], 1);
        }
        __syncthreads();
        constexpr int PER = SEL_BINS / SEL_NT;
        int local[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) { local[q] = hist[tid * PER + q]; sum += local[q]; }
        int total;
        const int excl = block_exclusive_scan<SEL_NT>(sum, wsum2, total);
        const int rem = s_rem2;
        __syncthreads();
        if (excl < rem && rem <= excl + sum) {
            int cc = excl;
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                if (cc < rem && rem <= cc + local[q]) {
                    s_prefix2 = prefix | ((uint32_t)(tid * PER + q) << shift);
                    s_rem2 = rem - cc;
                }
                cc += local[q];
            }
        }
        __syncthreads();
    }
    const uint32_t thr = s_prefix2;        // key of the Keff-th smallest
    const int take_ties = s_rem2;          // how many entries with key == thr belong to the answer (lowest indices)
    for (int e = tid; e < c; e += SEL_NT) {
        const uint32_t k = l_key[e];
        bool take = k < thr;
        if (k == thr) {
            const int ix = l_idx[e];
            int r = 0;
            for (int q = 0; q < c; ++q) r += (l_key[q] == thr && l_idx[q] < ix);   // ties are rare: short loop in practice
            take = r < take_ties;
        }
        if (take) {
            const int pos = atomicAdd(&s_nsel, 1);
            c_key[pos] = k;
            c_idx[pos] = l_idx[e];
        }
    }
    __syncthreads();
    if (tid < Keff) {
        const uint32_t k = c_key[tid];
        const int ix = c_idx[tid];
        int r = 0;
        for (int q = 0; q < Keff; ++q) {
            const uint32_t kq = c_key[q];
            r += (kq < k) || (kq == k && c_idx[q] < ix);
        }
        out_idx[(size_t)row_id * out_stride + r] = ix;
        out_val[(size_t)row_id * out_stride + r] = sel_value(row, ix, scaled, s);
    }
    for (int q = Keff + tid; q < K; q += SEL_NT) {
        out_idx[(size_t)row_id * out_stride + q] = -1;
        out_val[(size_t)row_id * out_stride + q] = INFINITY;
    }
    if (tid == 0) row_done[row_id] = 1;
}

int launch_row_select(const float* M, size_t ld, int rows, int cols, const float* scale, int K, bool largest,
                      int* out_idx, float* out_val, int out_stride, int* row_done, cudaStream_t st) {
    if (K < 1 || K > SEL_KMAX || K > out_stride)
        return ssg_set_error(SSG_ERR_INVALID, "row_select: K=%d out of range", K);
    if (rows <= 0) return SSG_OK;
    if (row_done)
        row_select_sampled_kernel<<<rows, SEL_NT, 0, st>>>(M, ld, cols, scale, K, largest, out_idx, out_val, out_stride,
                                                          row_done);
    row_select_kernel<<<rows, SEL_NT, 0, st>>>(M, ld, cols, scale, K, largest, out_idx, out_val, out_stride, row_done);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// pair_exact: out[i, s] = exact squared distance between A[row0+i] and B[idx[i, s]] for s < cnt[i]
// (cnt == nullptr -> fixed_cnt).  One CTA per row, the row staged in shared memory as float64, one
// thread per pair walking its partner row sequentially (cdist order, see sqdist_exact).
// ---------------------------------------------------------------------------------------------------
// v3 layout: a CTA owns ROWS rows; thread = (row, slot) with KP slots per row, and every thread carries PPT
// independent accumulator chains (pair slots slot, slot+KP, ...): the float64 add chain of one pair is strictly
// sequential (k = 0..d-1, cdist order), so instruction-level parallelism has to come from different pairs.  The row
// block is staged through shared memory in CHUNK-wide float64 chunks shared by all the pairs of a row.
// VEC8 (opt-in, SSG_PAIR_VEC8=1): every chain loads 32 contiguous bytes (two float4 = one whole 32-byte sector) of its
// partner row per step instead of 16.  With 16-byte steps each sector is requested twice, a step apart, and the
// shared-memory carve-out leaves almost no L1 to catch the second request.  Working hypothesis for why the kernel sits
// at ~3 TB/s of useful bytes with DRAM and the FP64 pipe both far from busy (profiles/r01g_rerank_full.md: 0.55 GB of
// DRAM reads, 12 % SM throughput): to be confirmed by the A/B run.  Same subtraction / product / sum per element in
// the same k order, hence the same bits.
template <int KP, int PPT, int ROWS, int CHUNK, bool VEC8 = false>
__global__ void __launch_bounds__(KP * ROWS)
pair_exact_kernel(const float* __restrict__ A, int rows, const float* __restrict__ B, int d,
                  const int* __restrict__ idx, int idx_stride, const int* __restrict__ cnt,
                  int fixed_cnt, float* __restrict__ out, int out_stride) {
    constexpr int NT = KP * ROWS;
    __shared__ double sa[ROWS][CHUNK];
    __shared__ int s_maxcnt;
    const int r_in = threadIdx.x / KP, slot = threadIdx.x % KP;
    const int row = blockIdx.x * ROWS + r_in;
    const int c = row < rows ? (cnt ? cnt[row] : fixed_cnt) : 0;
    if (threadIdx.x == 0) s_maxcnt = 0;
    __syncthreads();
    if (slot == 0) atomicMax(&s_maxcnt, c);
    __syncthreads();
    const int maxcnt = s_maxcnt;
    const bool vec_ok = (d & 3) == 0;
    for (int sbase = 0; sbase < maxcnt; sbase += KP * PPT) {
        int m[PPT];
        const float* b[PPT];
        double acc[PPT];
        bool act[PPT];
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            const int sidx = sbase + slot + u * KP;
            act[u] = sidx < c;
            m[u] = act[u] ? idx[(size_t)row * idx_stride + sidx] : -1;
            b[u] = B + (size_t)(m[u] < 0 ? 0 : m[u]) * d;
            acc[u] = 0.0;
        }
        for (int k0 = 0; k0 < d; k0 += CHUNK) {
            __syncthreads();
            for (int e = threadIdx.x; e < ROWS * CHUNK; e += NT) {
                const int rr = e / CHUNK, kk = e % CHUNK;
                const int gr = blockIdx.x * ROWS + rr;
                sa[rr][kk] = (gr < rows && k0 + kk < d) ? (double)A[(size_t)gr * d + k0 + kk] : 0.0;
            }
            __syncthreads();
            const int kend = min(CHUNK, d - k0);
            const double* ar = sa[r_in];
            if (VEC8 && (d & 7) == 0) {
                for (int k = 0; k < kend; k += 8) {
                    float4 t[PPT], t2[PPT];
#pragma unroll
                    for (int u = 0; u < PPT; ++u) {
                        const bool on = act[u] && m[u] >= 0;
                        t[u] = on ? *reinterpret_cast<const float4*>(b[u] + k0 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                        t2[u] = on ? *reinterpret_cast<const float4*>(b[u] + k0 + k + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    const double a0 = ar[k], a1 = ar[k + 1], a2 = ar[k + 2], a3 = ar[k + 3];
                    const double a4 = ar[k + 4], a5 = ar[k + 5], a6 = ar[k + 6], a7 = ar[k + 7];
#pragma unroll
                    for (int u = 0; u < PPT; ++u) {
                        double df;
                        df = __dsub_rn(a0, (double)t[u].x); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a1, (double)t[u].y); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a2, (double)t[u].z); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a3, (double)t[u].w); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a4, (double)t2[u].x); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a5, (double)t2[u].y); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a6, (double)t2[u].z); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a7, (double)t2[u].w); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                    }
                }
            } else if (vec_ok) {
                for (int k = 0; k < kend; k += 4) {
                    float4 t[PPT];
#pragma unroll
                    for (int u = 0; u < PPT; ++u)
                        t[u] = (act[u] && m[u] >= 0) ? *reinterpret_cast<const float4*>(b[u] + k0 + k) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const double a0 = ar[k], a1 = ar[k + 1], a2 = ar[k + 2], a3 = ar[k + 3];
#pragma unroll
                    for (int u = 0; u < PPT; ++u) {
                        double df;
                        df = __dsub_rn(a0, (double)t[u].x); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a1, (double)t[u].y); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a2, (double)t[u].z); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        df = __dsub_rn(a3, (double)t[u].w); acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                    }
                }
            } else {
                for (int k = 0; k < kend; ++k) {
#pragma unroll
                    for (int u = 0; u < PPT; ++u) {
                        if (act[u] && m[u] >= 0) {
                            const double df = __dsub_rn(ar[k], (double)b[u][k0 + k]);
                            acc[u] = __dadd_rn(acc[u], __dmul_rn(df, df));
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < PPT; ++u)
            if (act[u]) out[(size_t)row * out_stride + sbase + slot + u * KP] = m[u] < 0 ? INFINITY : finish_sqdist(acc[u]);
    }
}

// v4 (opt-in, SSG_PAIR_STAGED=1; MEASURED SLOWER on the B200: 15.0 vs 9.9 ms per cycle, profiles/r02h_ab_*.json -- one
// dependent float64 add chain per lane leaves the FP64 pipe latency bound, v3's four chains per thread matter more than
// its half-used sectors): one WARP per (row, group of 32 pairs), lane = pair.  The partner rows are
// staged through shared memory in 64-float chunks with coalesced loads (half a warp reads the 256 contiguous bytes of
// one partner row), then every lane walks its own pair's chunk out of shared memory.  In v3 a lane read its partner row
// straight from global memory, 16 bytes at a time: 32 lanes x 32 different rows = 32 half-used sectors per instruction,
// i.e. 16 KB of L2 -> SM traffic per 8 KB row and the kernel ran at the L2 rate (1.65 ms per launch for 0.67 M pairs at
// N = 16 702, round-2 profile) with the FP64 pipe idle.  Same subtraction / product / sum per element in the same k
// order as v3 and as cdist: the same bits.
constexpr int PX_CHUNK = 64, PX_PAD = 4, PX_WARPS = 4;
__global__ void __launch_bounds__(32 * PX_WARPS)
pair_exact_staged_kernel(const float* __restrict__ A, int rows, const float* __restrict__ B, int d,
                         const int* __restrict__ idx, int idx_stride, const int* __restrict__ cnt, int fixed_cnt,
                         int groups_per_row, float* __restrict__ out, int out_stride) {
    __shared__ float4 sp4[PX_WARPS][32][(PX_CHUNK + PX_PAD) / 4];         // partner rows of this warp's 32 pairs
    float (*sp)[32][PX_CHUNK + PX_PAD] = reinterpret_cast<float (*)[32][PX_CHUNK + PX_PAD]>(&sp4[0][0][0]);
    __shared__ double sa[PX_WARPS][PX_CHUNK];                             // the row itself, as float64
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long units = (long long)rows * groups_per_row;
    for (long long u = (long long)blockIdx.x * PX_WARPS + warp; u < units; u += (long long)gridDim.x * PX_WARPS) {
        const int row = (int)(u / groups_per_row), base = (int)(u % groups_per_row) * 32;
        const int c = cnt ? cnt[row] : fixed_cnt;
        if (base >= c) continue;                                          // warp-uniform
        const int npairs = min(32, c - base);
        const int m = lane < npairs ? idx[(size_t)row * idx_stride + base + lane] : -1;
        double acc = 0.0;
        const float* arow = A + (size_t)row * d;
        for (int k0 = 0; k0 < d; k0 += PX_CHUNK) {
            const int kend = min(PX_CHUNK, d - k0);
            __syncwarp();
            // the row chunk (float64) and the partner chunks: 16 lanes x float4 = one 256-byte run of one partner row
            for (int k = lane; k < PX_CHUNK; k += 32) sa[warp][k] = k < kend ? (double)arow[k0 + k] : 0.0;
            const int half = lane >> 4, l16 = lane & 15;
            if ((d & 3) == 0) {
                for (int p0 = 0; p0 < npairs; p0 += 2) {
                    const int p = p0 + half;
                    const int mp = __shfl_sync(0xffffffffu, m, p < 32 ? p : 31);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p < npairs && mp >= 0 && 4 * l16 < kend) v = *reinterpret_cast<const float4*>(B + (size_t)mp * d + k0 + 4 * l16);
                    if (p < 32) *reinterpret_cast<float4*>(&sp[warp][p][4 * l16]) = v;
                }
            } else {
                for (int p = 0; p < npairs; ++p) {
                    const int mp = __shfl_sync(0xffffffffu, m, p);
                    for (int k = lane; k < PX_CHUNK; k += 32)
                        sp[warp][p][k] = (mp >= 0 && k < kend) ? B[(size_t)mp * d + k0 + k] : 0.f;
                }
            }
            __syncwarp();
            if (m >= 0) {
                const float* br = sp[warp][lane];
                const double* ar = sa[warp];
                for (int k = 0; k < kend; k += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(br + k);
                    double df;
                    df = __dsub_rn(ar[k], (double)t.x); acc = __dadd_rn(acc, __dmul_rn(df, df));
                    if (k + 1 < kend) { df = __dsub_rn(ar[k + 1], (double)t.y); acc = __dadd_rn(acc, __dmul_rn(df, df)); }
                    if (k + 2 < kend) { df = __dsub_rn(ar[k + 2], (double)t.z); acc = __dadd_rn(acc, __dmul_rn(df, df)); }
                    if (k + 3 < kend) { df = __dsub_rn(ar[k + 3], (double)t.w); acc = __dadd_rn(acc, __dmul_rn(df, df)); }
                }
            }
        }
        if (lane < npairs) out[(size_t)row * out_stride + base + lane] = m < 0 ? INFINITY : finish_sqdist(acc);
    }
}

int launch_pair_exact(const float* A, int rows, const float* B, int d, const int* idx, int idx_stride,
                      const int* cnt, int fixed_cnt, float* out, int out_stride, cudaStream_t st) {
    if (rows <= 0) return SSG_OK;
    static int staged = -1;
    if (staged < 0) { const char* e = getenv("SSG_PAIR_STAGED"); staged = e ? atoi(e) : 0; }
    if (staged && (reinterpret_cast<uintptr_t>(B) & 15) == 0) {
        // pair groups per row: the widest row decides (cnt[] holds at most idx_stride entries per row)
        const int maxc = cnt ? idx_stride : fixed_cnt;
        const int groups = (maxc + 31) / 32;
        if (groups > 0) {
            const long long units = (long long)rows * groups;
            long long grid = (units + PX_WARPS - 1) / PX_WARPS;
            if (grid > 148 * 16) grid = 148 * 16;
            pair_exact_staged_kernel<<<(int)grid, 32 * PX_WARPS, 0, st>>>(A, rows, B, d, idx, idx_stride, cnt, fixed_cnt, groups,
                                                                         out, out_stride);
            SSG_CHECK_LAUNCH();
        }
        return SSG_OK;
    }
    static int vec8 = -1;
    if (vec8 < 0) { const char* e = getenv("SSG_PAIR_VEC8"); vec8 = e ? atoi(e) : 0; }
    if (vec8 && (d & 7) == 0 && (reinterpret_cast<uintptr_t>(B) & 31) == 0) {
        if (!cnt && fixed_cnt <= 8)
            pair_exact_kernel<2, 4, 128, 32, true><<<ssg_cdiv(rows, 128), 256, 0, st>>>(A, rows, B, d, idx, idx_stride, cnt,
                                                                                     fixed_cnt, out, out_stride);
        else
            pair_exact_kernel<8, 4, 32, 128, true><<<ssg_cdiv(rows, 32), 256, 0, st>>>(A, rows, B, d, idx, idx_stride, cnt,
                                                                                    fixed_cnt, out, out_stride);
        SSG_CHECK_LAUNCH();
        return SSG_OK;
    }
    if (!cnt && fixed_cnt <= 8) {
        // few pairs per row (row min / max candidates): 2 threads x 4 chains per row, 128 rows per CTA
        pair_exact_kernel<2, 4, 128, 32><<<ssg_cdiv(rows, 128), 256, 0, st>>>(A, rows, B, d, idx, idx_stride, cnt, fixed_cnt,
                                                                           out, out_stride);
    } else {
        pair_exact_kernel<8, 4, 32, 128><<<ssg_cdiv(rows, 32), 256, 0, st>>>(A, rows, B, d, idx, idx_stride, cnt, fixed_cnt,
                                                                          out, out_stride);
    }
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// Tensor distance mode: candidates picked on the approximate matrix are re-scored exactly; each result
// is accepted only if the approximation error bound E proves that no non-candidate could change it,
// otherwise the row is flagged for the exact fallback.
// ---------------------------------------------------------------------------------------------------
// row extreme from K exact candidate values.  cand_approx is sorted by selection order, so its last
// entry is the weakest candidate: every non-candidate is at least (min) / at most (max) that good.
__global__ void cand_reduce_kernel(int rows, int cols, int K, bool is_max, const float* __restrict__ exact,
                                   const float* __restrict__ cand_approx, int stride,
                                   const float* __restrict__ norm_row, const float* __restrict__ norm_other_max,
                                   float eps_rel, float* __restrict__ out, int* __restrict__ flag_cnt,
                                   int* __restrict__ flag_rows, int row_offset) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int kk = min(K, cols);
    float best = is_max ? -INFINITY : INFINITY;
    for (int q = 0; q < kk; ++q) {
        const float v = exact[(size_t)r * stride + q];
        best = is_max ? fmaxf(best, v) : fminf(best, v);
    }
    out[r] = best;
    if (cols > K) {
        const float E = eps_rel * (norm_row[r] + *norm_other_max);
        const float weakest = cand_approx[(size_t)r * stride + kk - 1];
        const bool ok = is_max ? (best >= weakest + E) : (best <= weakest - E);
        if (!ok) flag_rows[atomicAdd(flag_cnt, 1)] = row_offset + r;
    }
}

int launch_cand_reduce(int rows, int cols, int K, bool is_max, const float* exact, const float* cand_approx,
                       int stride, const float* norm_row, const float* norm_other_max, float eps_rel, float* out,
                       int* flag_cnt, int* flag_rows, int row_offset, cudaStream_t st) {
    if (rows <= 0) return SSG_OK;
    cand_reduce_kernel<<<ssg_cdiv(rows, 128), 128, 0, st>>>(rows, cols, K, is_max, exact, cand_approx, stride,
                                                            norm_row, norm_other_max, eps_rel, out, flag_cnt,
                                                            flag_rows, row_offset);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// leading k1p rank columns from K exactly re-scored candidates: order by (od/rowmax, index).
__global__ void __launch_bounds__(64)
rank_finalize_kernel(int cols, int K, int k1p, const int* __restrict__ cand_idx,
                     const float* __restrict__ cand_approx, const float* __restrict__ exact, int stride,
                     const float* __restrict__ rowmax, const float* __restrict__ norm_row,
                     const float* __restrict__ norm_other_max, float eps_rel, int* __restrict__ rank,
                     float* __restrict__ rank_val, int* __restrict__ flag_cnt, int* __restrict__ flag_rows,
                     int row_offset) {
    __shared__ float s_v[64];
    __shared__ int s_i[64];
    const int r = blockIdx.x, t = threadIdx.x;
    const float mx = rowmax[r];
    float v = INFINITY;
    int ix = INT_MAX;
    if (t < K) {
        const int ci = cand_idx[(size_t)r * stride + t];
        if (ci >= 0) { ix = ci; v = __fdiv_rn(exact[(size_t)r * stride + t], mx); }
    }
    s_v[t] = v;
    s_i[t] = ix;
    __syncthreads();
    const uint32_t kv = f32_key(v);
    int pos = 0;
    for (int q = 0; q < 64; ++q) {
        const uint32_t kq = f32_key(s_v[q]);
        pos += (kq < kv) || (kq == kv && s_i[q] < ix);
    }
    if (t < K && pos < k1p) {
        rank[(size_t)r * SSG_RANK_STRIDE + pos] = ix == INT_MAX ? -1 : ix;
        rank_val[(size_t)r * SSG_RANK_STRIDE + pos] = v;
        if (pos == k1p - 1 && cols > K) {
            // every non-candidate has approx >= weakest candidate, hence exact >= weakest - E
            const float E = eps_rel * (norm_row[r] + *norm_other_max);
            const float weakest = cand_approx[(size_t)r * stride + K - 1];
            const float bound = __fdiv_rn(fmaxf(weakest - E, 0.f), mx);
            if (!(v < bound)) flag_rows[atomicAdd(flag_cnt, 1)] = row_offset + r;
        }
    }
}

int launch_rank_finalize(int rows, int cols, int K, int k1p, const int* cand_idx, const float* cand_approx,
                         const float* exact, int stride, const float* rowmax, const float* norm_row,
                         const float* norm_other_max, float eps_rel, int* rank, float* rank_val, int* flag_cnt,
                         int* flag_rows, int row_offset, cudaStream_t st) {
    if (rows <= 0) return SSG_OK;
    if (K > 64 || k1p > K) return ssg_set_error(SSG_ERR_INVALID, "rank_finalize: K=%d k1p=%d", K, k1p);
    rank_finalize_kernel<<<rows, 64, 0, st>>>(cols, K, k1p, cand_idx, cand_approx, exact, stride, rowmax, norm_row,
                                              norm_other_max, eps_rel, rank, rank_val, flag_cnt, flag_rows,
                                              row_offset);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// fallback plumbing: gather the flagged feature rows, scatter the exact results back
__global__ void gather_rows_kernel(const float* __restrict__ X, int d, const int* __restrict__ rows,
                                   float* __restrict__ out) {
    const float* src = X + (size_t)rows[blockIdx.x] * d;
    float* dst = out + (size_t)blockIdx.x * d;
    for (int k = threadIdx.x; k < d; k += blockDim.x) dst[k] = src[k];
}
__global__ void scatter_f32_kernel(const float* __restrict__ in, int cnt, int width, int in_stride,
                                   const int* __restrict__ rows, float* __restrict__ out, int out_stride) {
    const int f = blockIdx.x;
    for (int k = threadIdx.x; k < width; k += blockDim.x)
        out[(size_t)rows[f] * out_stride + k] = in[(size_t)f * in_stride + k];
}
int launch_gather_rows(const float* X, int d, const int* rows, int cnt, float* out, cudaStream_t st) {
    if (cnt <= 0) return SSG_OK;
    gather_rows_kernel<<<cnt, 256, 0, st>>>(X, d, rows, out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
int launch_scatter_f32(const float* in, int cnt, int width, int in_stride, const int* rows, float* out,
                       int out_stride, cudaStream_t st) {
    if (cnt <= 0) return SSG_OK;
    scatter_f32_kernel<<<cnt, 32, 0, st>>>(in, cnt, width, in_stride, rows, out, out_stride);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // namespace ssg
