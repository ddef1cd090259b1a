// Fine-tune loss of the SSG iteration: reid/loss/triplet.py:11-77 TripletLoss.forward
// (w = None) and its gradient, as three small kernels.  A batch is n = P*K rows (n <= 4096) of d features; the
// whole op is ~n*n*d flops (33 MFLOP at n = 128, d = 2048) and launch-latency bound, so the design goal is
// "no Python loops, no host round trips, deterministic": the reference spends O(P*K^2) Python iterations and
// tensor cats per call (triplet.py:50-56).
//
//   triplet_dist_kernel  : dist[a,b] = sqrt(max(sum_k (x_ak - x_bk)^2, 1e-12))   (direct difference, fp32 FMA: no
//                          cancellation, diagonal exactly clamped; the reference's |x|^2+|y|^2-2xy is the same value
//                          up to fp32 GEMM rounding)
//   triplet_mine_kernel  : one CTA.  Phase A, warp per anchor: closest row with a different label (first minimum)
//                          [OHEM mode: and the farthest row with the same label].  Phase B, thread per anchor: hinge
//                          max(0, ap - an + margin) over the anchor's pairs, precision count, and row a of
//                          coef[a,b] = (d loss / d dist[a,b]) / dist[a,b]; sums are reduced in a fixed order.
//   triplet_grad_kernel  : grad_x[i] = g * sum_b (coef[i,b] + coef[b,i]) (x_i - x_b), CTA per row, the non-zero
//                          partners compacted in ascending order (deterministic accumulation).
#include "common.cuh"
#include "kernels.h"

namespace ssg {

#define TRIPLET_MAX_N 4096
#define TRIPLET_CLAMP 1e-12f

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) triplet_dist_kernel(const float* __restrict__ X, int n, int d,
                                                           float* __restrict__ dist, int* __restrict__ status) {
    constexpr int T = 32, KC = 32;
    __shared__ float As[T][KC + 1], Bs[T][KC + 1];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 2) status[threadIdx.x] = threadIdx.x == 0 ? 0 : -1;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int row0 = blockIdx.y * T, col0 = blockIdx.x * T;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};   // fp32 inside a 32-wide chunk, fp64 across chunks
    for (int k0 = 0; k0 < d; k0 += KC) {
        for (int e = threadIdx.x; e < T * KC; e += 256) {
            const int r = e / KC, k = e % KC;
            const int ga = row0 + r, gb = col0 + r, gk = k0 + k;
            As[r][k] = (ga < n && gk < d) ? X[(size_t)ga * d + gk] : 0.f;
            Bs[r][k] = (gb < n && gk < d) ? X[(size_t)gb * d + gk] : 0.f;
        }
        __syncthreads();
        float c00 = 0.f, c01 = 0.f, c10 = 0.f, c11 = 0.f;
#pragma unroll 8
        for (int k = 0; k < KC; ++k) {
            const float a0 = As[ty * 2][k], a1 = As[ty * 2 + 1][k];
            const float b0 = Bs[tx * 2][k], b1 = Bs[tx * 2 + 1][k];
            float t;
            t = a0 - b0; c00 = fmaf(t, t, c00);
            t = a0 - b1; c01 = fmaf(t, t, c01);
            t = a1 - b0; c10 = fmaf(t, t, c10);
            t = a1 - b1; c11 = fmaf(t, t, c11);
        }
        acc[0][0] += (double)c00; acc[0][1] += (double)c01;
        acc[1][0] += (double)c10; acc[1][1] += (double)c11;
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int r = row0 + ty * 2 + i, c = col0 + tx * 2 + j;
            if (r < n && c < n) dist[(size_t)r * n + c] = sqrtf(fmaxf((float)acc[i][j], TRIPLET_CLAMP));
        }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) triplet_mine_kernel(const float* __restrict__ dist,
                                                            const int64_t* __restrict__ tg, int n, int K,
                                                            float margin, int semi, long long T,
                                                            float* __restrict__ coef, float* __restrict__ out,
                                                            int* __restrict__ status) {
    __shared__ int s_an[TRIPLET_MAX_N];
    __shared__ int s_ap[TRIPLET_MAX_N];
    __shared__ double s_sum[1024];
    __shared__ int s_cnt[1024];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n_anchor = semi ? (n / K) * K : n;
    for (size_t e = tid; e < (size_t)n * n; e += 1024) coef[e] = 0.f;
    // phase A: warp per anchor
    for (int a = wid; a < n_anchor; a += 32) {
        const int64_t ta = tg[a];
        const float* row = dist + (size_t)a * n;
        float nv = __int_as_float(0x7f800000);   // +inf
        int ni = 0x7fffffff;
        float pv = -1.f;
        int pi = 0x7fffffff;
        for (int b = lane; b < n; b += 32) {
            const float v = row[b];
            if (tg[b] != ta) {
                if (v < nv) { nv = v; ni = b; }     // ascending b per lane: keeps the first minimum
            } else if (v > pv) { pv = v; pi = b; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, nv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, ni, o);
            if (ov < nv || (ov == nv && oi < ni)) { nv = ov; ni = oi; }
            const float qv = __shfl_xor_sync(0xffffffffu, pv, o);
            const int qi = __shfl_xor_sync(0xffffffffu, pi, o);
            if (qv > pv || (qv == pv && qi < pi)) { pv = qv; pi = qi; }
        }
        if (lane == 0) {
            s_an[a] = ni;
            s_ap[a] = pi;
            if (ni == 0x7fffffff) { status[0] = 1; atomicMax(&status[1], a); }
        }
    }
    __syncthreads();
    // phase B: thread per anchor (row a of coef is written by this thread only)
    const float invT = 1.0f / (float)T;
    const float clamp_s = sqrtf(TRIPLET_CLAMP);
    double sum = 0.0;
    int cnt = 0;
    for (int a = tid; a < n_anchor; a += 1024) {
        const int m = s_an[a];
        if (m == 0x7fffffff) continue;
        const float* row = dist + (size_t)a * n;
        float* crow = coef + (size_t)a * n;
        const float an = row[m];
        int active = 0;
        const int p_lo = semi ? a + 1 : s_ap[a];
        const int p_hi = semi ? (a / K + 1) * K : s_ap[a] + 1;
        for (int p = p_lo; p < p_hi; ++p) {
            const float ap = row[p];
            const float h = -(an - ap) + margin;        // MarginRankingLoss(margin)(an, ap, y = 1), triplet.py:74
            cnt += an > ap;
            if (h > 0.f) {
                sum += (double)h;
                ++active;
                if (ap > clamp_s) crow[p] += invT / ap;
            }
        }
        if (active && an > clamp_s) crow[m] -= (float)active * invT / an;
    }
    s_sum[tid] = sum;
    s_cnt[tid] = cnt;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if (tid < o) { s_sum[tid] += s_sum[tid + o]; s_cnt[tid] += s_cnt[tid + o]; }
        __syncthreads();
    }
    if (tid == 0) {
        const bool bad = status[0] != 0;
        out[0] = bad ? __int_as_float(0x7fc00000) : (float)(s_sum[0] / (double)T);
        out[1] = (float)((double)s_cnt[0] / (double)T);
    }
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) triplet_grad_kernel(const float* __restrict__ X, int n, int d,
                                                           const float* __restrict__ coef,
                                                           const float* __restrict__ gout, float* __restrict__ gx) {
    __shared__ float s_w[TRIPLET_MAX_N];
    __shared__ int s_b[TRIPLET_MAX_N];
    __shared__ int s_cnt;
    const int i = blockIdx.x, tid = threadIdx.x;
    for (int b = tid; b < n; b += 256) s_w[b] = coef[(size_t)i * n + b] + coef[(size_t)b * n + i];
    __syncthreads();
    if (tid < 32) {   // order-preserving compaction of the non-zero partners
        int pos = 0;
        for (int b0 = 0; b0 < n; b0 += 32) {
            const int b = b0 + tid;
            const float w = b < n ? s_w[b] : 0.f;
            const unsigned m = __ballot_sync(0xffffffffu, w != 0.f && b != i);
            __syncwarp();
            if (w != 0.f && b != i) {
                const int r = pos + __popc(m & ((1u << tid) - 1u));
                s_w[r] = w;
                s_b[r] = b;
            }
            pos += __popc(m);
            __syncwarp();
        }
        if (tid == 0) s_cnt = pos;
    }
    __syncthreads();
    const int cnt = s_cnt;
    const float g = gout ? *gout : 1.f;
    const float* xi = X + (size_t)i * d;
    float* gi = gx + (size_t)i * d;
    if ((d & 3) == 0) {
        for (int c = tid; c < d / 4; c += 256) {
            const float4 a = reinterpret_cast<const float4*>(xi)[c];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int t = 0; t < cnt; ++t) {
                const float w = s_w[t];
                const float4 v = reinterpret_cast<const float4*>(X + (size_t)s_b[t] * d)[c];
                acc.x = fmaf(w, a.x - v.x, acc.x);
                acc.y = fmaf(w, a.y - v.y, acc.y);
                acc.z = fmaf(w, a.z - v.z, acc.z);
                acc.w = fmaf(w, a.w - v.w, acc.w);
            }
            reinterpret_cast<float4*>(gi)[c] = make_float4(g * acc.x, g * acc.y, g * acc.z, g * acc.w);
        }
    } else {
        for (int c = tid; c < d; c += 256) {
            const float a = xi[c];
            float acc = 0.f;
            for (int t = 0; t < cnt; ++t) acc = fmaf(s_w[t], a - X[(size_t)s_b[t] * d + c], acc);
            gi[c] = g * acc;
        }
    }
}

}  // namespace ssg

using namespace ssg;

static long long triplet_count(int n, int K, int semi) {
    if (!semi) return n;
    return (long long)(n / K) * K * (K - 1) / 2;
}

extern "C" int ssg_triplet_forward(const float* d_x, const int64_t* d_targets, int n, int d, int num_instances,
                                   float margin, int use_semi, float* d_dist, float* d_coef, float* d_loss_prec,
                                   int* d_status, void* stream) {
    if (!d_x || !d_targets || !d_dist || !d_coef || !d_loss_prec || !d_status)
        return ssg_set_error(SSG_ERR_INVALID, "triplet_forward: null pointer");
    if (n <= 0 || n > TRIPLET_MAX_N || d <= 0)
        return ssg_set_error(SSG_ERR_INVALID, "triplet_forward: n=%d (1..%d), d=%d", n, TRIPLET_MAX_N, d);
    if (num_instances <= 0 || (use_semi && triplet_count(n, num_instances, 1) == 0))
        return ssg_set_error(SSG_ERR_INVALID,
                             "triplet_forward: no triplets for n=%d, num_instances=%d (the reference fails in "
                             "torch.cat of an empty list, triplet.py:61)", n, num_instances);
    if ((reinterpret_cast<uintptr_t>(d_x) & 15) != 0)
        return ssg_set_error(SSG_ERR_INVALID, "triplet_forward: d_x must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    {
        SSG_PROF("triplet_dist", st);
        dim3 grid(ssg_cdiv(n, 32), ssg_cdiv(n, 32));
        triplet_dist_kernel<<<grid, 256, 0, st>>>(d_x, n, d, d_dist, d_status);
        SSG_CHECK_LAUNCH();
    }
    {
        SSG_PROF("triplet_mine", st);
        triplet_mine_kernel<<<1, 1024, 0, st>>>(d_dist, d_targets, n, num_instances, margin, use_semi ? 1 : 0,
                                                triplet_count(n, num_instances, use_semi), d_coef, d_loss_prec,
                                                d_status);
        SSG_CHECK_LAUNCH();
    }
    return SSG_OK;
}

extern "C" int ssg_triplet_backward(const float* d_x, int n, int d, const float* d_coef, const float* d_grad_loss,
                                    float* d_grad_x, void* stream) {
    if (!d_x || !d_coef || !d_grad_x) return ssg_set_error(SSG_ERR_INVALID, "triplet_backward: null pointer");
    if (n <= 0 || n > TRIPLET_MAX_N || d <= 0)
        return ssg_set_error(SSG_ERR_INVALID, "triplet_backward: n=%d (1..%d), d=%d", n, TRIPLET_MAX_N, d);
    if (((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_grad_x)) & 15) != 0)
        return ssg_set_error(SSG_ERR_INVALID, "triplet_backward: d_x / d_grad_x must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    SSG_PROF("triplet_grad", st);
    triplet_grad_kernel<<<n, 256, 0, st>>>(d_x, n, d, d_coef, d_grad_loss, d_grad_x);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
