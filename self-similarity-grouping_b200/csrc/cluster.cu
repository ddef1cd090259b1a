// eps estimate (selftraining.py:289-293) and DBSCAN on a dense precomputed matrix
// (sklearn DBSCAN(eps, min_samples, metric='precomputed').fit_predict, selftraining.py:295-306).
//
// eps   : exact order statistic by 6-pass radix select over the 64-bit order-preserving keys of the
//         non-zero strict-upper-triangle entries (HBM-bound row scans), then one masked-sum pass.
// DBSCAN: order-free restatement validated against sklearn: region query = row scan
//         (count, then neighbour lists), union-find over core-core edges hooked towards the smaller
//         index (root = minimum core index of the component), cluster id = rank of the root, border
//         point = minimum id over its core neighbours, noise = -1.
#include <math.h>
#include <limits.h>

#include "common.cuh"
#include "kernels.h"

struct ssg_cluster_plan {
    int device;
    int n_max;
    long long max_nbr;
    size_t bytes;
    // eps
    unsigned long long* hist;      // [4096]
    unsigned long long* state;     // [8]: 0 prefix key, 1 remaining, 2 top_num, 3 M, 4 done-flag
    double* partial;               // [n_max]
    double* eps_out;               // [2]: eps, (unused)
    double* list;                  // [EPS_LIST_CAP] values of the threshold bin (3-pass eps)
    // dbscan
    int* cnt;                      // [n_max]
    int* rowptr;                   // [n_max+1]
    int* nbr;                      // [max_nbr]
    int* parent;                   // [n_max]
    int* isroot;                   // [n_max]
    int* cid;                      // [n_max+1]
    unsigned char* core;           // [n_max]
    int* flags;                    // [4]: 0 overflow
    void* staging;                 // device copy of a host matrix (host entry points), grown lazily
    size_t staging_bytes;
    int last_n;
};

namespace ssg {

template <typename T> __device__ __forceinline__ double ld_as_double(const T* p, size_t i);
template <> __device__ __forceinline__ double ld_as_double<double>(const double* p, size_t i) { return p[i]; }
template <> __device__ __forceinline__ double ld_as_double<float>(const float* p, size_t i) { return (double)p[i]; }

// ----------------------------------------------------------------------------------------------- eps
constexpr int EPS_NT = 256;
constexpr int EPS_BINS = 4096;
constexpr int EPS_LIST_CAP = 1 << 20;
// pass p looks at key bits [shift, shift+width)
__constant__ int c_eps_shift[6] = {52, 40, 28, 16, 4, 0};
__constant__ int c_eps_width[6] = {12, 12, 12, 12, 12, 4};

template <typename T>
__global__ void __launch_bounds__(EPS_NT)
eps_hist_kernel(const T* __restrict__ D, int n, int pass, const unsigned long long* __restrict__ state,
                unsigned long long* __restrict__ ghist) {
    __shared__ unsigned int hist[EPS_BINS];
    for (int b = threadIdx.x; b < EPS_BINS; b += EPS_NT) hist[b] = 0u;
    __syncthreads();
    const int shift = c_eps_shift[pass], width = c_eps_width[pass];
    const unsigned long long prefix = state[0];
    const unsigned mask = (1u << width) - 1u;
    const int hs = shift + width;                     // bits above are the established prefix
    // rows are paired (i, n-1-i) so that every CTA sees ~n elements of the upper triangle
    for (int half = 0; half < 2; ++half) {
        const int i = half == 0 ? (int)blockIdx.x : n - 1 - (int)blockIdx.x;
        if (half == 1 && i <= (int)blockIdx.x) break;
        const T* row = D + (size_t)i * n;
        for (int jb = i + 1; jb < n; jb += EPS_NT * 4) {
            // 4 independent loads in flight per thread (the scan is latency bound otherwise)
            double vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = jb + u * EPS_NT + threadIdx.x;
                vv[u] = j < n ? ld_as_double<T>(row, j) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double v = vv[u];
                bool in = false;
                unsigned bin = 0;
                if (v != 0.0) {
                    const unsigned long long k = f64_key(v);
                    in = (pass == 0) || ((k >> hs) == (prefix >> hs));
                    bin = (unsigned)(k >> shift) & mask;
                }
                // warp-aggregated shared-memory histogram (the values are heavily clustered)
                const unsigned act = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const unsigned peers = __match_any_sync(act, bin);
                    if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], __popc(peers));
                }
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < EPS_BINS; b += EPS_NT) {
        const unsigned c = hist[b];
        if (c) atomicAdd(&ghist[b], (unsigned long long)c);
    }
}

// single CTA: pick the bin holding the `remaining`-th smallest key, extend the prefix.
__global__ void __launch_bounds__(1024)
eps_pick_kernel(unsigned long long* __restrict__ ghist, unsigned long long* __restrict__ state, int pass,
                double rho, long long n_pairs = -1, const unsigned long long* __restrict__ zeros = nullptr) {
    __shared__ unsigned long long cum[EPS_BINS];
    const int tid = threadIdx.x;
    // serial-ish scan: 4096 bins, 1024 threads x 4 bins, then thread 0 stitches 1024 partials
    unsigned long long loc[4], s = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { loc[q] = ghist[tid * 4 + q]; s += loc[q]; }
    cum[tid] = s;
    __syncthreads();
    if (tid == 0) {
        unsigned long long run = 0;
        for (int t = 0; t < 1024; ++t) { const unsigned long long v = cum[t]; cum[t] = run; run += v; }
        cum[1024] = run;
    }
    __syncthreads();
    const unsigned long long total = cum[1024];
    if (pass == 0 && tid == 0) {
        // M = number of non-zero upper-triangle entries; top_num = np.round(rho*M) (half to even)
        // sparse form (n_pairs >= 0): the histogram holds only the entries that lie below every entry outside the CSR;
        // M = all pairs minus the exact zeros, and the selection is valid iff it stays inside the histogram
        // (state[7] = 1 otherwise: the caller falls back to the dense matrix).
        const unsigned long long M = n_pairs >= 0 ? (unsigned long long)n_pairs - zeros[0] : total;
        double t = rint(rho * (double)M);
        if (t < 0.0) t = 0.0;
        unsigned long long top = (unsigned long long)t;
        if (top > M) top = M;
        state[3] = M;
        state[2] = top;
        state[1] = top;      // remaining
        state[0] = 0ull;     // prefix
        if (n_pairs >= 0 && top > total) { state[7] = 1ull; state[1] = 0ull; }
    }
    __syncthreads();
    const unsigned long long rem = state[1];
    __syncthreads();
    if (rem > 0) {
        unsigned long long c = cum[tid];
        const int shift = c_eps_shift[pass];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (c < rem && rem <= c + loc[q]) {
                state[0] = state[0] | ((unsigned long long)(tid * 4 + q) << shift);
                state[1] = rem - c;
            }
            c += loc[q];
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) ghist[tid * 4 + q] = 0ull;   // ready for the next pass
}

template <typename T>
__global__ void __launch_bounds__(EPS_NT)
eps_sum_kernel(const T* __restrict__ D, int n, const unsigned long long* __restrict__ state,
               double* __restrict__ partial) {
    const unsigned long long thr = state[0];
    double acc = 0.0;
    if (state[2] > 0) {
        for (int half = 0; half < 2; ++half) {
            const int i = half == 0 ? (int)blockIdx.x : n - 1 - (int)blockIdx.x;
            if (half == 1 && i <= (int)blockIdx.x) break;
            const T* row = D + (size_t)i * n;
            for (int j = i + 1 + threadIdx.x; j < n; j += EPS_NT) {
                const double v = ld_as_double<T>(row, j);
                if (v != 0.0 && f64_key(v) < thr) acc += v;
            }
        }
    }
    // deterministic tree: fixed thread->element map and fixed combine order
    __shared__ double sh[EPS_NT];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = EPS_NT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(1024)
eps_final_kernel(const double* __restrict__ partial, int nparts, const unsigned long long* __restrict__ state,
                 double* __restrict__ eps_out) {
    __shared__ double sh[1024];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 1024) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const unsigned long long top = state[2];
        if (top == 0) {
            eps_out[0] = __longlong_as_double(0x7ff8000000000000ll);   // mean of an empty slice
        } else {
            const double thr = f64_from_key(state[0]);
            eps_out[0] = (sh[0] + (double)state[1] * thr) / (double)top;
        }
    }
}

// ---- 3-pass variant: two histogram passes fix the leading 24 key bits of the threshold; the third pass sums every
// entry below that 24-bit bin and gathers the (few) entries inside it; the order statistic is finished on that list.
template <typename T>
__global__ void __launch_bounds__(EPS_NT)
eps_gather_kernel(const T* __restrict__ D, int n, unsigned long long* __restrict__ state, double* __restrict__ list,
                  double* __restrict__ partial) {
    const unsigned long long pre = state[0] >> 40;
    double acc = 0.0;
    if (state[2] > 0) {
        for (int half = 0; half < 2; ++half) {
            const int i = half == 0 ? (int)blockIdx.x : n - 1 - (int)blockIdx.x;
            if (half == 1 && i <= (int)blockIdx.x) break;
            const T* row = D + (size_t)i * n;
            for (int jb = i + 1 + threadIdx.x; jb < n; jb += EPS_NT * 4) {
                double vv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = jb + u * EPS_NT;
                    vv[u] = j < n ? ld_as_double<T>(row, j) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {      // same element order per thread as a plain strided loop
                    const double v = vv[u];
                    if (v == 0.0) continue;
                    const unsigned long long h = f64_key(v) >> 40;
                    if (h < pre) acc += v;
                    else if (h == pre) {
                        const unsigned long long pos = atomicAdd(&state[5], 1ull);
                        if (pos < (unsigned long long)EPS_LIST_CAP) list[pos] = v; else state[6] = 1ull;
                    }
                }
            }
        }
    }
    __shared__ double sh[EPS_NT];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = EPS_NT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// single CTA: finish the radix select (40 low key bits, 4 passes) on the gathered list, then the mean.
__global__ void __launch_bounds__(1024)
eps_list_finish_kernel(const double* __restrict__ list, unsigned long long* __restrict__ state,
                       const double* __restrict__ partial, int nparts, double* __restrict__ eps_out) {
    __shared__ unsigned int hist[EPS_BINS];
    __shared__ unsigned long long s_prefix, s_rem;
    __shared__ double sh[1024];
    const int tid = threadIdx.x;
    const unsigned long long top = state[2];
    if (top == 0) {
        if (tid == 0) eps_out[0] = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    if (state[6]) return;                         // list overflow: the host falls back to the 6-pass path
    const int L = (int)state[5];
    if (tid == 0) { s_prefix = state[0]; s_rem = state[1]; }
    __syncthreads();
    const int shifts[4] = {28, 16, 4, 0};
    const int widths[4] = {12, 12, 12, 4};
    for (int p = 0; p < 4; ++p) {
        for (int b = tid; b < EPS_BINS; b += 1024) hist[b] = 0u;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        const int shift = shifts[p], hs = shifts[p] + widths[p];
        const unsigned mask = (1u << widths[p]) - 1u;
        for (int e = tid; e < L; e += 1024) {
            const unsigned long long k = f64_key(list[e]);
            if ((k >> hs) == (prefix >> hs)) atomicAdd(&hist[(unsigned)(k >> shift) & mask], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long c = 0, rem = s_rem;
            for (int b = 0; b < EPS_BINS; ++b) {
                const unsigned long long hb = hist[b];
                if (c < rem && rem <= c + hb) { s_prefix = prefix | ((unsigned long long)b << shift); s_rem = rem - c; break; }
                c += hb;
            }
        }
        __syncthreads();
    }
    const unsigned long long thr = s_prefix;
    // The list is filled through an atomic cursor, so its order differs from run to run; all its entries share sign
    // and exponent (equal leading 24 key bits), so their sum is taken exactly on the integer mantissas and is
    // therefore independent of the order: eps is bit-reproducible.
    unsigned __int128 isum = 0;
    for (int e = tid; e < L; e += 1024) {
        const double v = list[e];
        if (f64_key(v) < thr) {
            const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
            const unsigned long long ex = (bits >> 52) & 0x7ffull;
            isum += (unsigned __int128)((bits & 0xfffffffffffffull) | (ex ? (1ull << 52) : 0ull));
        }
    }
    __shared__ unsigned long long sh_lo[1024], sh_hi[1024];
    sh_lo[tid] = (unsigned long long)isum;
    sh_hi[tid] = (unsigned long long)(isum >> 64);
    double acc = 0.0;
    for (int i = tid; i < nparts; i += 1024) acc += partial[i];
    sh[tid] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) {
            sh[tid] += sh[tid + o];
            const unsigned __int128 a = ((unsigned __int128)sh_hi[tid] << 64) | sh_lo[tid];
            const unsigned __int128 b = ((unsigned __int128)sh_hi[tid + o] << 64) | sh_lo[tid + o];
            const unsigned __int128 c = a + b;
            sh_lo[tid] = (unsigned long long)c;
            sh_hi[tid] = (unsigned long long)(c >> 64);
        }
        __syncthreads();
    }
    if (tid == 0) {
        const double tv = f64_from_key(thr);
        const unsigned long long tb = (unsigned long long)__double_as_longlong(tv);
        const int ex = (int)((tb >> 52) & 0x7ffull);
        const int e2 = (ex ? ex : 1) - 1075;                         // value = mantissa * 2^e2
        double lsum = ldexp((double)sh_hi[0], 64 + e2) + ldexp((double)sh_lo[0], e2);
        if (tb >> 63) lsum = -lsum;
        state[0] = thr; state[1] = s_rem;
        eps_out[0] = (sh[0] + lsum + (double)s_rem * tv) / (double)top;
    }
}

template <typename T>
static int eps_run3(ssg_cluster_plan* p, const T* D, int n, double rho, cudaStream_t st) {
    SSG_CUDA_TRY(cudaMemsetAsync(p->hist, 0, sizeof(unsigned long long) * EPS_BINS, st));
    SSG_CUDA_TRY(cudaMemsetAsync(p->state, 0, sizeof(unsigned long long) * 8, st));
    const int grid = (n + 1) / 2;
    for (int pass = 0; pass < 2; ++pass) {
        { SSG_PROF("eps_hist", st); eps_hist_kernel<T><<<grid, EPS_NT, 0, st>>>(D, n, pass, p->state, p->hist); }
        SSG_CHECK_LAUNCH();
        eps_pick_kernel<<<1, 1024, 0, st>>>(p->hist, p->state, pass, rho);
        SSG_CHECK_LAUNCH();
    }
    { SSG_PROF("eps_gather", st); eps_gather_kernel<T><<<grid, EPS_NT, 0, st>>>(D, n, p->state, p->list, p->partial); }
    SSG_CHECK_LAUNCH();
    eps_list_finish_kernel<<<1, 1024, 0, st>>>(p->list, p->state, p->partial, grid, p->eps_out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

template <typename T>
static int eps_run(ssg_cluster_plan* p, const T* D, int n, double rho, cudaStream_t st) {
    SSG_CUDA_TRY(cudaMemsetAsync(p->hist, 0, sizeof(unsigned long long) * EPS_BINS, st));
    SSG_CUDA_TRY(cudaMemsetAsync(p->state, 0, sizeof(unsigned long long) * 8, st));
    const int grid = (n + 1) / 2;
    for (int pass = 0; pass < 6; ++pass) {
        { SSG_PROF("eps_hist", st); eps_hist_kernel<T><<<grid, EPS_NT, 0, st>>>(D, n, pass, p->state, p->hist); }
        SSG_CHECK_LAUNCH();
        eps_pick_kernel<<<1, 1024, 0, st>>>(p->hist, p->state, pass, rho);
        SSG_CHECK_LAUNCH();
    }
    { SSG_PROF("eps_sum", st); eps_sum_kernel<T><<<grid, EPS_NT, 0, st>>>(D, n, p->state, p->partial); }
    SSG_CHECK_LAUNCH();
    eps_final_kernel<<<1, 1024, 0, st>>>(p->partial, grid, p->state, p->eps_out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---- row-sharded eps (multi-GPU): rank r of `world` holds the rows [lo_r, hi_r) of a SYMMETRIC
// matrix (ssg_b200.dist.shard_bounds: the first n % world ranks own one row more).  Every unordered pair {a, b},
// a != b, is visited by exactly one rank, and the rows of a rank are read in contiguous column runs:
//   * inside the diagonal block of the rank: the strict upper triangle, i.e. columns (i, hi_r) of row i;
//   * an off-diagonal block (r, c) is taken by rank r when r < c and r + c is odd, or r > c and r + c is even
//     (the mirror block (c, r) holds the same values and is skipped by rank c) -- every rank reads about half of its
//     row block, where the plain upper triangle would leave rank 0 with twice the work of the average.
// The multiset of visited values equals that of np.triu(D, 1) because D is exactly symmetric (rerank.cu).
struct ShardGeom {
    int n, world, rank, base, rem;
    __host__ __device__ int lo(int r) const { return r * base + (r < rem ? r : rem); }
    __host__ __device__ bool takes(int c) const {
        if (c == rank) return true;
        const bool odd = ((rank + c) & 1) != 0;
        return rank < c ? odd : !odd;
    }
};
static ShardGeom make_geom(int n, int world, int rank) {
    ShardGeom g;
    g.n = n; g.world = world; g.rank = rank; g.base = n / world; g.rem = n % world;
    return g;
}

// Calls body(value, valid) for every assigned column of global row i, EPS_NT * 4 columns per step with four
// independent loads in flight per thread; all threads of the CTA make the same number of calls (body may use
// warp-synchronous primitives), `valid` is false for the padding slots.
template <typename T, class Body>
__device__ __forceinline__ void shard_row_scan(const T* __restrict__ row, int i, const ShardGeom& g, Body body) {
    for (int c = 0; c < g.world; ++c) {
        if (!g.takes(c)) continue;
        const int jb0 = c == g.rank ? i + 1 : g.lo(c);
        const int je = g.lo(c + 1);
        for (int jb = jb0; jb < je; jb += EPS_NT * 4) {
            double vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = jb + u * EPS_NT + (int)threadIdx.x;
                vv[u] = j < je ? ld_as_double<T>(row, j) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) body(vv[u], jb + u * EPS_NT + (int)threadIdx.x < je);
        }
    }
}

// one CTA per local row (block row blockIdx.x = global row lo_rank + blockIdx.x)
template <typename T>
__global__ void __launch_bounds__(EPS_NT)
eps_hist_rows_kernel(const T* __restrict__ D, ShardGeom g, int pass, const unsigned long long* __restrict__ state,
                     unsigned long long* __restrict__ ghist) {
    __shared__ unsigned int hist[EPS_BINS];
    for (int b = threadIdx.x; b < EPS_BINS; b += EPS_NT) hist[b] = 0u;
    __syncthreads();
    const int shift = c_eps_shift[pass], width = c_eps_width[pass];
    const unsigned long long prefix = state[0];
    const unsigned mask = (1u << width) - 1u;
    const int hs = shift + width;
    const int i = g.lo(g.rank) + (int)blockIdx.x;
    shard_row_scan<T>(D + (size_t)blockIdx.x * g.n, i, g, [&](double v, bool valid) {
        bool in = false;
        unsigned bin = 0;
        if (valid && v != 0.0) {
            const unsigned long long k = f64_key(v);
            in = (pass == 0) || ((k >> hs) == (prefix >> hs));
            bin = (unsigned)(k >> shift) & mask;
        }
        const unsigned act = __ballot_sync(0xffffffffu, in);
        if (in) {
            const unsigned peers = __match_any_sync(act, bin);
            if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(&hist[bin], __popc(peers));
        }
    });
    __syncthreads();
    for (int b = threadIdx.x; b < EPS_BINS; b += EPS_NT) {
        const unsigned c = hist[b];
        if (c) atomicAdd(&ghist[b], (unsigned long long)c);
    }
}

// mode 0 (3-pass path): sum of the entries below the threshold's 24-bit bin -> partial[global row]; the entries
//                       inside the bin are appended to `list` (cursor state[5], overflow flag state[6]);
// mode 1 (6-pass path): sum of the entries whose key is below the exact threshold state[0] -> partial[global row].
template <typename T>
__global__ void __launch_bounds__(EPS_NT)
eps_gather_rows_kernel(const T* __restrict__ D, ShardGeom g, int mode, unsigned long long* __restrict__ state,
                       double* __restrict__ list, double* __restrict__ partial) {
    const unsigned long long pre = state[0] >> 40, thr = state[0];
    const int i = g.lo(g.rank) + (int)blockIdx.x;
    double acc = 0.0;
    if (state[2] > 0) {
        shard_row_scan<T>(D + (size_t)blockIdx.x * g.n, i, g, [&](double v, bool valid) {
            if (!valid || v == 0.0) return;
            const unsigned long long k = f64_key(v);
            if (mode == 1) {
                if (k < thr) acc += v;
                return;
            }
            const unsigned long long h = k >> 40;
            if (h < pre) acc += v;
            else if (h == pre) {
                const unsigned long long pos = atomicAdd(&state[5], 1ull);
                if (pos < (unsigned long long)EPS_LIST_CAP) list[pos] = v; else state[6] = 1ull;
            }
        });
    }
    __shared__ double sh[EPS_NT];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = EPS_NT / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[i] = sh[0];
}

// ---- eps over the sparse form of final_dist (CSR of the touched columns, rerank.cu): one warp per row, only the
// strict upper triangle (col > row), zeros dropped (and counted: they are missing from M), entries >= thr ignored --
// they cannot be told apart from the entries outside the CSR, so the selection is certified by eps_pick_kernel.
constexpr int SP_NT = 256;

__global__ void __launch_bounds__(SP_NT)
eps_sp_hist_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int n,
                   double thr, int pass, const unsigned long long* __restrict__ state,
                   unsigned long long* __restrict__ ghist, unsigned long long* __restrict__ zeros) {
    __shared__ unsigned int hist[EPS_BINS];
    for (int b = threadIdx.x; b < EPS_BINS; b += SP_NT) hist[b] = 0u;
    __syncthreads();
    const int shift = c_eps_shift[pass], width = c_eps_width[pass];
    const unsigned long long prefix = state[0];
    const unsigned mask = (1u << width) - 1u;
    const int hs = shift + width;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (SP_NT / 32) + (threadIdx.x >> 5);
    if (i < n && !state[7]) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        int nz = 0;
        for (int base = beg; base < end; base += 32) {
            const int e = base + lane;
            const bool up = e < end && col[e] > i;
            const double v = up ? val[e] : 1.0;
            nz += (up && v == 0.0);
            bool in = false;
            unsigned bin = 0;
            if (up && v != 0.0 && v < thr) {
                const unsigned long long k = f64_key(v);
                in = (pass == 0) || ((k >> hs) == (prefix >> hs));
                bin = (unsigned)(k >> shift) & mask;
            }
            const unsigned act = __ballot_sync(0xffffffffu, in);
            if (in) {
                const unsigned peers = __match_any_sync(act, bin);
                if ((int)(__ffs(peers) - 1) == lane) atomicAdd(&hist[bin], __popc(peers));
            }
        }
        if (pass == 0) {
            nz = warp_sum_i(nz);
            if (lane == 0 && nz) atomicAdd(zeros, (unsigned long long)nz);
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < EPS_BINS; b += SP_NT) {
        const unsigned c = hist[b];
        if (c) atomicAdd(&ghist[b], (unsigned long long)c);
    }
}

__global__ void __launch_bounds__(SP_NT)
eps_sp_gather_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int n,
                     double thr, unsigned long long* __restrict__ state, double* __restrict__ list,
                     double* __restrict__ partial) {
    const unsigned long long pre = state[0] >> 40;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (SP_NT / 32) + (threadIdx.x >> 5);
    if (i >= n) return;
    double acc = 0.0;
    if (state[2] > 0 && !state[7]) {
        const int beg = rowptr[i], end = rowptr[i + 1];
        for (int e = beg + lane; e < end; e += 32) {
            if (col[e] <= i) continue;
            const double v = val[e];
            if (v == 0.0 || !(v < thr)) continue;
            const unsigned long long h = f64_key(v) >> 40;
            if (h < pre) acc += v;
            else if (h == pre) {
                const unsigned long long pos = atomicAdd(&state[5], 1ull);
                if (pos < (unsigned long long)EPS_LIST_CAP) list[pos] = v; else state[6] = 1ull;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);   // fixed butterfly: deterministic
    if (lane == 0) partial[i] = acc;
}

// region query on the sparse form: valid while eps is below every entry outside the CSR (checked by the caller)
__global__ void __launch_bounds__(SP_NT)
db_sp_count_kernel(const int* __restrict__ rowptr, const double* __restrict__ val, int n, double eps,
                   int* __restrict__ cnt) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (SP_NT / 32) + (threadIdx.x >> 5);
    if (i >= n) return;
    int c = 0;
    for (int e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) c += (val[e] <= eps);
    c = warp_sum_i(c);
    if (lane == 0) cnt[i] = c;
}

__global__ void __launch_bounds__(SP_NT)
db_sp_fill_kernel(const int* __restrict__ sp_rowptr, const int* __restrict__ col, const double* __restrict__ val, int n,
                  double eps, const int* __restrict__ rowptr, long long max_nbr, int* __restrict__ nbr,
                  int* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (SP_NT / 32) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int obeg = rowptr[i], oend = rowptr[i + 1];
    if (oend == obeg) return;
    if ((long long)oend > max_nbr) { if (lane == 0) flags[0] = 1; return; }
    int out = obeg;
    const int beg = sp_rowptr[i], end = sp_rowptr[i + 1];
    for (int base = beg; base < end; base += 32) {
        const int e = base + lane;
        const bool hit = e < end && val[e] <= eps;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) nbr[out + __popc(m & ((1u << lane) - 1u))] = col[e];
        out += __popc(m);
    }
}

// -------------------------------------------------------------------------------------------- DBSCAN
constexpr int DB_NT = 256;

// region query, pass 1: neighbour count per row (all j, the diagonal included: sklearn does not force
// self-inclusion for a dense precomputed matrix).
template <typename T>
__global__ void __launch_bounds__(DB_NT)
db_count_kernel(const T* __restrict__ D, int n, double eps, int* __restrict__ cnt) {
    const T* row = D + (size_t)blockIdx.x * n;
    int c = 0;
    for (int j = threadIdx.x; j < n; j += DB_NT) c += (ld_as_double<T>(row, j) <= eps);
    c = warp_sum_i(c);
    __shared__ int sh[DB_NT / 32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < DB_NT / 32; ++w) t += sh[w];
        cnt[blockIdx.x] = t;
    }
}

// pass 2: neighbour lists (CSR).  Order inside a row is irrelevant to the order-free formulation.
template <typename T>
__global__ void __launch_bounds__(DB_NT)
db_fill_kernel(const T* __restrict__ D, int n, double eps, const int* __restrict__ rowptr,
               long long max_nbr, int* __restrict__ nbr, int* __restrict__ flags) {
    const int i = blockIdx.x;
    if (flags[0]) return;                         // the 64-bit total already exceeds the plan (db_total_check_kernel)
    const int beg = rowptr[i], end = rowptr[i + 1];
    if (end == beg) return;
    if ((long long)end > max_nbr) { if (threadIdx.x == 0) flags[0] = 1; return; }
    __shared__ int cursor;
    if (threadIdx.x == 0) cursor = 0;
    __syncthreads();
    const T* row = D + (size_t)i * n;
    for (int j0 = 0; j0 < n; j0 += DB_NT) {
        const int j = j0 + threadIdx.x;
        const bool hit = j < n && ld_as_double<T>(row, j) <= eps;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) {
            int base = 0;
            if ((threadIdx.x & 31) == 0) base = atomicAdd(&cursor, __popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (hit) nbr[beg + base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = j;
        }
    }
}

// 64-bit total of the per-row counts: the int32 CSR offsets wrap once the neighbour pairs exceed INT_MAX (possible for
// n > 46 340), so the capacity test cannot be left to the offsets alone.  One CTA; sets flags[0] on overflow.
__global__ void __launch_bounds__(1024)
db_total_check_kernel(int n, const int* __restrict__ cnt, long long max_nbr, int* __restrict__ flags) {
    __shared__ unsigned long long sh[1024];
    unsigned long long acc = 0;
    for (int i = threadIdx.x; i < n; i += 1024) acc += (unsigned long long)cnt[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0 && sh[0] > (unsigned long long)max_nbr) flags[0] = 1;
}

__global__ void db_init_kernel(int n, const int* __restrict__ cnt, int min_samples, int* __restrict__ parent,
                               unsigned char* __restrict__ core) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { parent[i] = i; core[i] = cnt[i] >= min_samples; }
}

__device__ __forceinline__ int uf_find(int* parent, int x) {
    int p = ((volatile int*)parent)[x];
    while (p != x) { x = p; p = ((volatile int*)parent)[x]; }
    return x;
}

// one warp per row: union every core-core edge (j < i) of a core row.
__global__ void __launch_bounds__(DB_NT)
db_union_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ nbr,
                const unsigned char* __restrict__ core, int* parent) {
    const int i = blockIdx.x * (DB_NT / 32) + (threadIdx.x >> 5);
    if (i >= n || !core[i]) return;
    const int beg = rowptr[i], end = rowptr[i + 1];
    for (int e = beg + (threadIdx.x & 31); e < end; e += 32) {
        const int j = nbr[e];
        if (j >= i || !core[j]) continue;
        int a = i, b = j;
        while (true) {
            a = uf_find(parent, a);
            b = uf_find(parent, b);
            if (a == b) break;
            if (a < b) { const int t = a; a = b; b = t; }     // hook the larger root under the smaller
            const int old = atomicCAS(&parent[a], a, b);
            if (old == a) break;
        }
    }
}

__global__ void db_flatten_kernel(int n, const unsigned char* __restrict__ core, int* parent,
                                  int* __restrict__ isroot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = 0;
    if (core[i]) r = (uf_find(parent, i) == i);
    isroot[i] = r;
}

__global__ void __launch_bounds__(DB_NT)
db_label_kernel(int n, const int* __restrict__ rowptr, const int* __restrict__ nbr,
                const unsigned char* __restrict__ core, int* parent, const int* __restrict__ cid,
                int64_t* __restrict__ labels) {
    const int i = blockIdx.x * (DB_NT / 32) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int lane = threadIdx.x & 31;
    if (core[i]) {
        if (lane == 0) labels[i] = (int64_t)cid[uf_find(parent, i)];
        return;
    }
    int best = INT_MAX;
    for (int e = rowptr[i] + lane; e < rowptr[i + 1]; e += 32) {
        const int j = nbr[e];
        if (core[j]) best = min(best, cid[uf_find(parent, j)]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
    if (lane == 0) labels[i] = best == INT_MAX ? (int64_t)-1 : (int64_t)best;
}

static int dbscan_label_stage(ssg_cluster_plan* p, int n, int min_samples, int64_t* labels, cudaStream_t st);

template <typename T>
static int dbscan_run(ssg_cluster_plan* p, const T* D, int n, double eps, int min_samples,
                      int64_t* labels, cudaStream_t st) {
    SSG_CUDA_TRY(cudaMemsetAsync(p->flags, 0, sizeof(int) * 4, st));
    { SSG_PROF("dbscan_count", st); db_count_kernel<T><<<n, DB_NT, 0, st>>>(D, n, eps, p->cnt); }
    SSG_CHECK_LAUNCH();
    db_total_check_kernel<<<1, 1024, 0, st>>>(n, p->cnt, p->max_nbr, p->flags);
    SSG_CHECK_LAUNCH();
    SSG_TRY(launch_exclusive_scan_i32(p->cnt, p->rowptr, n, st));
    { SSG_PROF("dbscan_fill", st); db_fill_kernel<T><<<n, DB_NT, 0, st>>>(D, n, eps, p->rowptr, p->max_nbr, p->nbr, p->flags); }
    SSG_CHECK_LAUNCH();
    return dbscan_label_stage(p, n, min_samples, labels, st);
}

// union-find over the core-core edges of the complete neighbour CSR, cluster ids, border points (needs cnt, rowptr
// and nbr of all n rows)
static int dbscan_label_stage(ssg_cluster_plan* p, int n, int min_samples, int64_t* labels, cudaStream_t st) {
    db_init_kernel<<<ssg_cdiv(n, 256), 256, 0, st>>>(n, p->cnt, min_samples, p->parent, p->core);
    SSG_CHECK_LAUNCH();
    const int wgrid = ssg_cdiv(n, DB_NT / 32);
    SSG_PROF("dbscan_label", st);
    db_union_kernel<<<wgrid, DB_NT, 0, st>>>(n, p->rowptr, p->nbr, p->core, p->parent);
    SSG_CHECK_LAUNCH();
    db_flatten_kernel<<<ssg_cdiv(n, 256), 256, 0, st>>>(n, p->core, p->parent, p->isroot);
    SSG_CHECK_LAUNCH();
    SSG_TRY(launch_exclusive_scan_i32(p->isroot, p->cid, n, st));
    db_label_kernel<<<wgrid, DB_NT, 0, st>>>(n, p->rowptr, p->nbr, p->core, p->parent, p->cid, labels);
    SSG_CHECK_LAUNCH();
    p->last_n = n;
    return SSG_OK;
}

}  // namespace ssg

// ------------------------------------------------------------------------------------------ C ABI
using namespace ssg;

static int dev_alloc(void** p, size_t bytes, size_t* total) {
    SSG_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 16));
    *total += bytes;
    return SSG_OK;
}

extern "C" int ssg_cluster_plan_create(ssg_cluster_plan** out, int device, int n_max, long long max_neighbors) {
    if (!out || n_max <= 0) return ssg_set_error(SSG_ERR_INVALID, "cluster_plan_create: bad arguments");
    if (max_neighbors <= 0) max_neighbors = 64ll * n_max + (1 << 20);
    if (max_neighbors > 0x7fffffffll) max_neighbors = 0x7fffffffll;
    SSG_ON_DEVICE(device);
    ssg_cluster_plan* p = new ssg_cluster_plan();
    p->device = device; p->n_max = n_max; p->max_nbr = max_neighbors; p->bytes = 0;
    p->staging = nullptr; p->staging_bytes = 0; p->last_n = 0;
    size_t n = (size_t)n_max;
    int rc = SSG_OK;
#define A(ptr, nbytes) if (rc == SSG_OK) rc = dev_alloc((void**)&(ptr), (nbytes), &p->bytes)
    A(p->hist, sizeof(unsigned long long) * 4096);
    A(p->state, sizeof(unsigned long long) * 8);
    A(p->partial, sizeof(double) * n);
    A(p->eps_out, sizeof(double) * 2);
    A(p->list, sizeof(double) * EPS_LIST_CAP);
    A(p->cnt, sizeof(int) * n);
    A(p->rowptr, sizeof(int) * (n + 1));
    A(p->nbr, sizeof(int) * (size_t)max_neighbors);
    A(p->parent, sizeof(int) * n);
    A(p->isroot, sizeof(int) * n);
    A(p->cid, sizeof(int) * (n + 1));
    A(p->core, n);
    A(p->flags, sizeof(int) * 4);
#undef A
    if (rc != SSG_OK) { ssg_cluster_plan_destroy(p); return rc; }
    *out = p;
    return SSG_OK;
}

extern "C" int ssg_cluster_plan_destroy(ssg_cluster_plan* p) {
    if (!p) return SSG_OK;
    SsgDeviceGuard device_guard__(p->device);
    void* ptrs[] = {p->hist, p->state, p->partial, p->eps_out, p->list, p->cnt, p->rowptr, p->nbr, p->parent,
                    p->isroot, p->cid, p->core, p->flags, p->staging};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete p;
    return SSG_OK;
}

extern "C" size_t ssg_cluster_plan_bytes(const ssg_cluster_plan* p) { return p ? p->bytes + p->staging_bytes : 0; }

extern "C" int ssg_eps_estimate(ssg_cluster_plan* p, const void* d_dist, int dtype, int n, double rho,
                                double* h_eps, long long* h_top_num, void* stream) {
    if (!p || !d_dist || n <= 0 || n > p->n_max || !h_eps)
        return ssg_set_error(SSG_ERR_INVALID, "eps_estimate: bad arguments (n=%d, n_max=%d)", n, p ? p->n_max : -1);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype != SSG_F64 && dtype != SSG_F32) return ssg_set_error(SSG_ERR_INVALID, "eps_estimate: dtype %d", dtype);
    if (dtype == SSG_F64) SSG_TRY(eps_run3<double>(p, (const double*)d_dist, n, rho, st));
    else SSG_TRY(eps_run3<float>(p, (const float*)d_dist, n, rho, st));
    unsigned long long hs[8];
    SSG_CUDA_TRY(cudaMemcpyAsync(h_eps, p->eps_out, sizeof(double), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaMemcpyAsync(hs, p->state, sizeof(hs), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaStreamSynchronize(st));
    if (hs[6]) {
        // more than EPS_LIST_CAP entries share the threshold's leading 24 key bits (massive ties): full radix select
        if (dtype == SSG_F64) SSG_TRY(eps_run<double>(p, (const double*)d_dist, n, rho, st));
        else SSG_TRY(eps_run<float>(p, (const float*)d_dist, n, rho, st));
        SSG_CUDA_TRY(cudaMemcpyAsync(h_eps, p->eps_out, sizeof(double), cudaMemcpyDeviceToHost, st));
        SSG_CUDA_TRY(cudaMemcpyAsync(hs, p->state, sizeof(hs), cudaMemcpyDeviceToHost, st));
        SSG_CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (h_top_num) *h_top_num = (long long)hs[2];
    return SSG_OK;
}

extern "C" int ssg_dbscan(ssg_cluster_plan* p, const void* d_dist, int dtype, int n, double eps,
                          int min_samples, int64_t* d_labels, int* h_n_clusters, void* stream) {
    if (!p || !d_dist || !d_labels || n <= 0 || n > p->n_max)
        return ssg_set_error(SSG_ERR_INVALID, "dbscan: bad arguments (n=%d, n_max=%d)", n, p ? p->n_max : -1);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == SSG_F64) SSG_TRY(dbscan_run<double>(p, (const double*)d_dist, n, eps, min_samples, d_labels, st));
    else if (dtype == SSG_F32) SSG_TRY(dbscan_run<float>(p, (const float*)d_dist, n, eps, min_samples, d_labels, st));
    else return ssg_set_error(SSG_ERR_INVALID, "dbscan: dtype %d", dtype);
    {
        // the overflow flag is ALWAYS read back (one stream synchronisation): on overflow the neighbour lists are
        // incomplete and the labels meaningless, which must never pass silently
        int flags[4], ncl = 0;
        SSG_CUDA_TRY(cudaMemcpyAsync(flags, p->flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        SSG_CUDA_TRY(cudaMemcpyAsync(&ncl, p->cid + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        SSG_CUDA_TRY(cudaStreamSynchronize(st));
        if (flags[0])
            return ssg_set_error(SSG_ERR_CAPACITY, "dbscan: more than %lld neighbour pairs within eps; "
                                 "re-create the plan with a larger max_neighbors", p->max_nbr);
        if (h_n_clusters) *h_n_clusters = ncl;
    }
    return SSG_OK;
}

extern "C" int ssg_dbscan_core_mask(ssg_cluster_plan* p, uint8_t* h_core, int n) {
    if (!p || !h_core || n <= 0 || n > p->n_max) return ssg_set_error(SSG_ERR_INVALID, "core_mask: bad arguments");
    SSG_ON_DEVICE(p->device);
    SSG_CUDA_TRY(cudaMemcpy(h_core, p->core, (size_t)n, cudaMemcpyDeviceToHost));
    return SSG_OK;
}

static int stage_host_matrix(ssg_cluster_plan* p, const void* h, int dtype, int n) {
    const size_t bytes = (size_t)n * n * (dtype == SSG_F64 ? 8 : 4);
    if (bytes > p->staging_bytes) {
        if (p->staging) cudaFree(p->staging);
        p->staging = nullptr; p->staging_bytes = 0;
        SSG_CUDA_TRY(cudaMalloc(&p->staging, bytes));
        p->staging_bytes = bytes;
    }
    SSG_CUDA_TRY(cudaMemcpy(p->staging, h, bytes, cudaMemcpyHostToDevice));
    return SSG_OK;
}

extern "C" int ssg_eps_estimate_host(ssg_cluster_plan* p, const void* h_dist, int dtype, int n, double rho,
                                     double* h_eps, long long* h_top_num) {
    if (!p || !h_dist || n <= 0 || n > p->n_max || (dtype != SSG_F32 && dtype != SSG_F64))
        return ssg_set_error(SSG_ERR_INVALID, "eps_estimate_host: bad arguments");
    SSG_ON_DEVICE(p->device);
    SSG_TRY(stage_host_matrix(p, h_dist, dtype, n));
    return ssg_eps_estimate(p, p->staging, dtype, n, rho, h_eps, h_top_num, nullptr);
}

extern "C" int ssg_dbscan_host(ssg_cluster_plan* p, const void* h_dist, int dtype, int n, double eps,
                               int min_samples, int64_t* h_labels, int* h_n_clusters) {
    if (!p || !h_dist || !h_labels || n <= 0 || n > p->n_max || (dtype != SSG_F32 && dtype != SSG_F64))
        return ssg_set_error(SSG_ERR_INVALID, "dbscan_host: bad arguments");
    SSG_ON_DEVICE(p->device);
    SSG_TRY(stage_host_matrix(p, h_dist, dtype, n));
    int64_t* d_labels = nullptr;
    SSG_CUDA_TRY(cudaMalloc(&d_labels, sizeof(int64_t) * (size_t)n));
    int ncl = 0;
    int rc = ssg_dbscan(p, p->staging, dtype, n, eps, min_samples, d_labels, &ncl, nullptr);
    if (rc == SSG_OK) {
        cudaError_t e = cudaMemcpy(h_labels, d_labels, sizeof(int64_t) * (size_t)n, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = ssg_set_error(SSG_ERR_CUDA, "dbscan_host: %s", cudaGetErrorString(e));
    }
    cudaFree(d_labels);
    if (rc == SSG_OK && h_n_clusters) *h_n_clusters = ncl;
    return rc;
}

// ------------------------------------------------------------------------- row-sharded eps / DBSCAN
// The caller owns the collectives (torch.distributed / NCCL on the buffers ssg_cluster_buffers exposes); see
// include/ssg_b200.h for the call sequence and ssg_b200/dist.py for the one this repo uses.
extern "C" int ssg_cluster_buffers(ssg_cluster_plan* p, void** d_hist, void** d_state, void** d_partial,
                                   void** d_list, void** d_cnt, void** d_nbr) {
    if (!p) return ssg_set_error(SSG_ERR_INVALID, "cluster_buffers: null plan");
    if (d_hist) *d_hist = p->hist;
    if (d_state) *d_state = p->state;
    if (d_partial) *d_partial = p->partial;
    if (d_list) *d_list = p->list;
    if (d_cnt) *d_cnt = p->cnt;
    if (d_nbr) *d_nbr = p->nbr;
    return SSG_OK;
}

static int check_shard(const ssg_cluster_plan* p, const void* d_rows, int dtype, int n, int world, int rank,
                       const char* who) {
    if (!p || n <= 0 || n > p->n_max || world <= 0 || rank < 0 || rank >= world)
        return ssg_set_error(SSG_ERR_INVALID, "%s: bad arguments (n=%d, world=%d, rank=%d)", who, n, world, rank);
    if (dtype != SSG_F64 && dtype != SSG_F32) return ssg_set_error(SSG_ERR_INVALID, "%s: dtype %d", who, dtype);
    const ShardGeom g = make_geom(n, world, rank);
    if (!d_rows && g.lo(rank + 1) > g.lo(rank)) return ssg_set_error(SSG_ERR_INVALID, "%s: null row block", who);
    return SSG_OK;
}

extern "C" int ssg_eps_shard_begin(ssg_cluster_plan* p, void* stream) {
    if (!p) return ssg_set_error(SSG_ERR_INVALID, "eps_shard_begin: null plan");
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    SSG_CUDA_TRY(cudaMemsetAsync(p->hist, 0, sizeof(unsigned long long) * EPS_BINS, st));
    SSG_CUDA_TRY(cudaMemsetAsync(p->state, 0, sizeof(unsigned long long) * 8, st));
    return SSG_OK;
}

extern "C" int ssg_eps_shard_hist(ssg_cluster_plan* p, const void* d_rows, int dtype, int n, int world, int rank,
                                  int pass, void* stream) {
    SSG_TRY(check_shard(p, d_rows, dtype, n, world, rank, "eps_shard_hist"));
    if (pass < 0 || pass > 5) return ssg_set_error(SSG_ERR_INVALID, "eps_shard_hist: pass %d", pass);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const ShardGeom g = make_geom(n, world, rank);
    const int rows = g.lo(rank + 1) - g.lo(rank);
    if (rows == 0) return SSG_OK;
    SSG_PROF("eps_hist", st);
    if (dtype == SSG_F64)
        eps_hist_rows_kernel<double><<<rows, EPS_NT, 0, st>>>((const double*)d_rows, g, pass, p->state, p->hist);
    else
        eps_hist_rows_kernel<float><<<rows, EPS_NT, 0, st>>>((const float*)d_rows, g, pass, p->state, p->hist);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

extern "C" int ssg_eps_shard_pick(ssg_cluster_plan* p, int pass, double rho, void* stream) {
    if (!p || pass < 0 || pass > 5) return ssg_set_error(SSG_ERR_INVALID, "eps_shard_pick: bad arguments");
    SSG_ON_DEVICE(p->device);
    eps_pick_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(p->hist, p->state, pass, rho);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

extern "C" int ssg_eps_shard_gather(ssg_cluster_plan* p, const void* d_rows, int dtype, int n, int world, int rank,
                                    int exact_threshold, long long* h_list_count, void* stream) {
    SSG_TRY(check_shard(p, d_rows, dtype, n, world, rank, "eps_shard_gather"));
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const ShardGeom g = make_geom(n, world, rank);
    const int rows = g.lo(rank + 1) - g.lo(rank);
    const int mode = exact_threshold ? 1 : 0;
    if (rows > 0) {
        SSG_PROF(mode ? "eps_sum" : "eps_gather", st);
        if (dtype == SSG_F64)
            eps_gather_rows_kernel<double><<<rows, EPS_NT, 0, st>>>((const double*)d_rows, g, mode, p->state, p->list,
                                                                    p->partial);
        else
            eps_gather_rows_kernel<float><<<rows, EPS_NT, 0, st>>>((const float*)d_rows, g, mode, p->state, p->list,
                                                                   p->partial);
        SSG_CHECK_LAUNCH();
    }
    if (h_list_count) {
        unsigned long long hs[8];
        SSG_CUDA_TRY(cudaMemcpyAsync(hs, p->state, sizeof(hs), cudaMemcpyDeviceToHost, st));
        SSG_CUDA_TRY(cudaStreamSynchronize(st));
        // more entries than the list holds share the threshold's leading 24 key bits: report the overflow as -1
        *h_list_count = hs[6] ? -1ll : (long long)hs[5];
    }
    return SSG_OK;
}

extern "C" int ssg_eps_shard_finish(ssg_cluster_plan* p, int n, int exact_threshold, double* h_eps,
                                    long long* h_top_num, void* stream) {
    if (!p || n <= 0 || n > p->n_max || !h_eps) return ssg_set_error(SSG_ERR_INVALID, "eps_shard_finish: bad arguments");
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (exact_threshold) eps_final_kernel<<<1, 1024, 0, st>>>(p->partial, n, p->state, p->eps_out);
    else eps_list_finish_kernel<<<1, 1024, 0, st>>>(p->list, p->state, p->partial, n, p->eps_out);
    SSG_CHECK_LAUNCH();
    unsigned long long hs[8];
    SSG_CUDA_TRY(cudaMemcpyAsync(h_eps, p->eps_out, sizeof(double), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaMemcpyAsync(hs, p->state, sizeof(hs), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaStreamSynchronize(st));
    if (h_top_num) *h_top_num = (long long)hs[2];
    return SSG_OK;
}

extern "C" int ssg_dbscan_shard_count(ssg_cluster_plan* p, const void* d_rows, int dtype, int n, int row0, int rows,
                                      double eps, void* stream) {
    if (!p || n <= 0 || n > p->n_max || row0 < 0 || rows < 0 || row0 + rows > n || (!d_rows && rows > 0) ||
        (dtype != SSG_F64 && dtype != SSG_F32))
        return ssg_set_error(SSG_ERR_INVALID, "dbscan_shard_count: bad arguments (n=%d, rows [%d,%d))", n, row0,
                             row0 + rows);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    SSG_CUDA_TRY(cudaMemsetAsync(p->flags, 0, sizeof(int) * 4, st));
    if (rows == 0) return SSG_OK;
    SSG_PROF("dbscan_count", st);
    if (dtype == SSG_F64) db_count_kernel<double><<<rows, DB_NT, 0, st>>>((const double*)d_rows, n, eps, p->cnt + row0);
    else db_count_kernel<float><<<rows, DB_NT, 0, st>>>((const float*)d_rows, n, eps, p->cnt + row0);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

extern "C" int ssg_dbscan_shard_fill(ssg_cluster_plan* p, const void* d_rows, int dtype, int n, int row0, int rows,
                                     double eps, long long* h_total, void* stream) {
    if (!p || n <= 0 || n > p->n_max || row0 < 0 || rows < 0 || row0 + rows > n || (!d_rows && rows > 0) ||
        (dtype != SSG_F64 && dtype != SSG_F32) || !h_total)
        return ssg_set_error(SSG_ERR_INVALID, "dbscan_shard_fill: bad arguments (n=%d, rows [%d,%d))", n, row0,
                             row0 + rows);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    // cnt holds the counts of ALL rows by now (the caller gathered them): global CSR offsets, the same on every rank
    SSG_TRY(launch_exclusive_scan_i32(p->cnt, p->rowptr, n, st));
    int total = 0;
    SSG_CUDA_TRY(cudaMemcpyAsync(&total, p->rowptr + n, sizeof(int), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaStreamSynchronize(st));
    *h_total = (long long)total;
    if (total < 0 || (long long)total > p->max_nbr)
        return ssg_set_error(SSG_ERR_CAPACITY, "dbscan: %d neighbour pairs within eps exceed the plan's %lld; "
                             "re-create the plan with a larger max_neighbors", total, p->max_nbr);
    // rows of the other ranks stay zero: the caller completes the list with a sum all-reduce over nbr[0, total)
    SSG_CUDA_TRY(cudaMemsetAsync(p->nbr, 0, sizeof(int) * (size_t)total, st));
    if (rows == 0) return SSG_OK;
    SSG_PROF("dbscan_fill", st);
    if (dtype == SSG_F64)
        db_fill_kernel<double><<<rows, DB_NT, 0, st>>>((const double*)d_rows, n, eps, p->rowptr + row0, p->max_nbr, p->nbr, p->flags);
    else
        db_fill_kernel<float><<<rows, DB_NT, 0, st>>>((const float*)d_rows, n, eps, p->rowptr + row0, p->max_nbr, p->nbr, p->flags);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

extern "C" int ssg_dbscan_shard_label(ssg_cluster_plan* p, int n, int min_samples, int64_t* d_labels,
                                      int* h_n_clusters, void* stream) {
    if (!p || !d_labels || n <= 0 || n > p->n_max) return ssg_set_error(SSG_ERR_INVALID, "dbscan_shard_label: bad arguments");
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    SSG_TRY(dbscan_label_stage(p, n, min_samples, d_labels, st));
    if (h_n_clusters) {
        SSG_CUDA_TRY(cudaMemcpyAsync(h_n_clusters, p->cid + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        SSG_CUDA_TRY(cudaStreamSynchronize(st));
    }
    return SSG_OK;
}

// ------------------------------------------------------------- eps / DBSCAN on the sparse form of final_dist
// (ssg_rerank_finish_sparse: CSR rows = touched columns ascending; every entry outside the CSR is >= threshold)
extern "C" int ssg_eps_sparse(ssg_cluster_plan* p, int n, const int* d_rowptr, const int* d_col, const double* d_val,
                              double threshold, double rho, double* h_eps, long long* h_top_num, int* h_certified,
                              void* stream) {
    if (!p || !d_rowptr || !d_col || !d_val || n <= 0 || n > p->n_max || !h_eps || !h_certified)
        return ssg_set_error(SSG_ERR_INVALID, "eps_sparse: bad arguments (n=%d, n_max=%d)", n, p ? p->n_max : -1);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    SSG_CUDA_TRY(cudaMemsetAsync(p->hist, 0, sizeof(unsigned long long) * EPS_BINS, st));
    SSG_CUDA_TRY(cudaMemsetAsync(p->state, 0, sizeof(unsigned long long) * 8, st));
    SSG_CUDA_TRY(cudaMemsetAsync(p->eps_out, 0, sizeof(double) * 2, st));
    unsigned long long* zeros = reinterpret_cast<unsigned long long*>(p->eps_out + 1);   // spare 8-byte slot
    const int grid = ssg_cdiv(n, SP_NT / 32);
    const long long n_pairs = (long long)n * (n - 1) / 2;
    for (int pass = 0; pass < 2; ++pass) {
        { SSG_PROF("eps_hist", st); eps_sp_hist_kernel<<<grid, SP_NT, 0, st>>>(d_rowptr, d_col, d_val, n, threshold, pass, p->state, p->hist, zeros); }
        SSG_CHECK_LAUNCH();
        eps_pick_kernel<<<1, 1024, 0, st>>>(p->hist, p->state, pass, rho, n_pairs, zeros);
        SSG_CHECK_LAUNCH();
    }
    { SSG_PROF("eps_gather", st); eps_sp_gather_kernel<<<grid, SP_NT, 0, st>>>(d_rowptr, d_col, d_val, n, threshold, p->state, p->list, p->partial); }
    SSG_CHECK_LAUNCH();
    eps_list_finish_kernel<<<1, 1024, 0, st>>>(p->list, p->state, p->partial, n, p->eps_out);
    SSG_CHECK_LAUNCH();
    unsigned long long hs[8];
    SSG_CUDA_TRY(cudaMemcpyAsync(h_eps, p->eps_out, sizeof(double), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaMemcpyAsync(hs, p->state, sizeof(hs), cudaMemcpyDeviceToHost, st));
    SSG_CUDA_TRY(cudaStreamSynchronize(st));
    // not certified: the rho-slice reaches past the entries the CSR can vouch for, or > 2^20 ties around the threshold
    *h_certified = (hs[7] || hs[6]) ? 0 : 1;
    if (h_top_num) *h_top_num = (long long)hs[2];
    return SSG_OK;
}

extern "C" int ssg_dbscan_sparse(ssg_cluster_plan* p, int n, const int* d_rowptr, const int* d_col, const double* d_val,
                                 double eps, int min_samples, int64_t* d_labels, int* h_n_clusters, void* stream) {
    if (!p || !d_rowptr || !d_col || !d_val || !d_labels || n <= 0 || n > p->n_max)
        return ssg_set_error(SSG_ERR_INVALID, "dbscan_sparse: bad arguments (n=%d, n_max=%d)", n, p ? p->n_max : -1);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    SSG_CUDA_TRY(cudaMemsetAsync(p->flags, 0, sizeof(int) * 4, st));
    const int grid = ssg_cdiv(n, SP_NT / 32);
    { SSG_PROF("dbscan_count", st); db_sp_count_kernel<<<grid, SP_NT, 0, st>>>(d_rowptr, d_val, n, eps, p->cnt); }
    SSG_CHECK_LAUNCH();
    SSG_TRY(launch_exclusive_scan_i32(p->cnt, p->rowptr, n, st));
    { SSG_PROF("dbscan_fill", st); db_sp_fill_kernel<<<grid, SP_NT, 0, st>>>(d_rowptr, d_col, d_val, n, eps, p->rowptr, p->max_nbr, p->nbr, p->flags); }
    SSG_CHECK_LAUNCH();
    SSG_TRY(dbscan_label_stage(p, n, min_samples, d_labels, st));
    if (h_n_clusters) {
        int flags[4], ncl = 0;
        SSG_CUDA_TRY(cudaMemcpyAsync(flags, p->flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
        SSG_CUDA_TRY(cudaMemcpyAsync(&ncl, p->cid + n, sizeof(int), cudaMemcpyDeviceToHost, st));
        SSG_CUDA_TRY(cudaStreamSynchronize(st));
        if (flags[0])
            return ssg_set_error(SSG_ERR_CAPACITY, "dbscan: more than %lld neighbour pairs within eps; "
                                 "re-create the plan with a larger max_neighbors", p->max_nbr);
        *h_n_clusters = ncl;
    }
    return SSG_OK;
}
