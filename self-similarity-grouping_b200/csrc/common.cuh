// Shared helpers for the ssg_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include "../../include/ssg_b200.h"

// ---- host-side error plumbing (api.cu owns the storage) -------------------------------------------
int ssg_set_error(int code, const char* fmt, ...);

#define SSG_CUDA_TRY(expr)                                                                     \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return ssg_set_error(SSG_ERR_CUDA, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,   \
                                 cudaGetErrorString(e__));                                     \
    } while (0)

#define SSG_TRY(expr)                  \
    do {                               \
        int rc__ = (expr);             \
        if (rc__ != SSG_OK) return rc__; \
    } while (0)

#define SSG_CHECK_LAUNCH() SSG_CUDA_TRY(cudaGetLastError())

// Entry points run on the plan's device and hand the caller's current device back on every return path.
struct SsgDeviceGuard {
    int prev_;
    cudaError_t err;
    explicit SsgDeviceGuard(int dev) : prev_(-1), err(cudaSuccess) {
        err = cudaGetDevice(&prev_);
        if (err == cudaSuccess && prev_ != dev) err = cudaSetDevice(dev); else if (err == cudaSuccess) prev_ = -1;
    }
    ~SsgDeviceGuard() { if (prev_ >= 0) cudaSetDevice(prev_); }
};
#define SSG_ON_DEVICE(dev)            \
    SsgDeviceGuard device_guard__(dev); \
    SSG_CUDA_TRY(device_guard__.err)

namespace ssg {
// RAII CUDA-event timer around a group of launches on one stream (no-op unless ssg_profile_enable(1)).
struct ProfScope {
    ProfScope(const char* name, cudaStream_t st);
    ~ProfScope();
    const char* name_;
    cudaStream_t st_;
    void* e0_;
};
void prof_suspend(bool on);   // no timers while a stream capture records launches
}  // namespace ssg
#define SSG_PROF(name, st) ssg::ProfScope prof_scope__(name, st)

static inline int ssg_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// ---- order-preserving integer keys ------------------------------------------------------------------
__device__ __forceinline__ uint32_t f32_key(float f) {
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ uint64_t f64_key(double f) {
    uint64_t b = (uint64_t)__double_as_longlong(f);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_key(uint64_t k) {
    uint64_t b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

// ---- warp / block reductions -------------------------------------------------------------------------
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Exclusive scan of one int per thread over a CTA of NT threads (NT multiple of 32, <= 1024).
// `total` receives the CTA sum.  Two __syncthreads.
template <int NT>
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums /* [NT/32] smem */, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        int s = warp_sums[w];
        if (w < wid) woff += s;
        tot += s;
    }
    total = tot;
    __syncthreads();
    return woff + inc - v;
}
#endif  // __CUDACC__
