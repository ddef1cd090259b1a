// cuda_runtime.h STAND-IN for the CPU execution of the library's plain-CUDA kernels (tests/cpu_cuda/README.md).
// TEST INFRASTRUCTURE: never on the product path.  It gives g++ what the .cu sources need -- launch geometry, the
// __shared__ / __global__ qualifiers, __syncthreads, warp collectives, atomics, the rounding intrinsics -- with
// CUDA's execution model emulated by cooperative fibers on ONE host thread (emu.cpp): the threads of a block run
// round-robin and switch only at barriers / warp collectives, blocks run one after another, "device memory" is host
// memory.  Data races are therefore not modelled; barrier / collective protocol errors show up as a detected deadlock.
#pragma once
#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
#include <functional>

struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r = {x, y}; return r; }
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
static inline float2 make_float2(float x, float y) { float2 r = {x, y}; return r; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;
#endif

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static          /* blocks run one after another: one instance per kernel is one per block */
#define __constant__ static

typedef int cudaError_t;
#define cudaSuccess 0
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp { int major, minor; char name[256]; };

#ifdef __cplusplus
extern "C" {
#endif
cudaError_t cudaMalloc(void** p, size_t bytes);
cudaError_t cudaFree(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, enum cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, enum cudaMemcpyKind kind, cudaStream_t st);
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                              enum cudaMemcpyKind kind, cudaStream_t st);
cudaError_t cudaMemset(void* p, int v, size_t bytes);
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t st);
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaGetDeviceProperties(struct cudaDeviceProp* prop, int d);
cudaError_t cudaStreamSynchronize(cudaStream_t st);
cudaError_t cudaDeviceSynchronize(void);
cudaError_t cudaGetLastError(void);
const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaEventCreate(cudaEvent_t* e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t st);
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b);
#ifdef __cplusplus
}

template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
// driver entry point (tensor-map encoder), device attributes, stream-ordered allocation
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
enum { cudaEnableDefault = 0, cudaDevAttrMultiProcessorCount = 16 };
cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long flags, cudaDriverEntryPointQueryResult* res);
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 3; return cudaSuccess; }     // 3 "SMs": persistent CTAs walk several tiles
static inline cudaError_t cudaMallocAsync(void** p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
// streams / events / graphs as embed.cu uses them: a "capture" simply executes, a graph launch does nothing more
typedef void* cudaGraph_t;
typedef void* cudaGraphExec_t;
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaStreamCaptureModeThreadLocal = 1 };
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)0x10; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = (cudaEvent_t)0x20; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaSuccess; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) { *g = (cudaGraph_t)0x30; return cudaSuccess; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) { *e = (cudaGraphExec_t)0x40; return cudaSuccess; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t bytes) { return cudaMalloc((void**)p, bytes); }

namespace emu {
// runs fn once per thread of every block of the grid (blocks sequentially, threads of a block as fibers)
void launch(const std::function<void()>& fn, dim3 grid, dim3 block, size_t smem = 0, cudaStream_t st = nullptr);
extern unsigned char* dyn_smem;                       // what `extern __shared__` arrays point at
void syncthreads();
int syncthreads_or(int pred);
const uint64_t* exchange(unsigned mask, uint64_t v);  // warp collective: every lane in `mask` deposits v, all get all
int lane_id();
}  // namespace emu

static inline void __syncthreads() { emu::syncthreads(); }
static inline int __syncthreads_or(int p) { return emu::syncthreads_or(p); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { emu::exchange(mask, 0); }
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    const uint64_t* s = emu::exchange(mask, pred ? 1 : 0);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if (((mask >> l) & 1u) && s[l]) r |= 1u << l;
    return r;
}
template <class T> static inline uint64_t emu_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> static inline T emu_from(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
template <class T> static inline T __shfl_sync(unsigned mask, T v, int src) { return emu_from<T>(emu::exchange(mask, emu_bits(v))[src & 31]); }
template <class T> static inline T __shfl_xor_sync(unsigned mask, T v, int o) { return emu_from<T>(emu::exchange(mask, emu_bits(v))[(emu::lane_id() ^ o) & 31]); }
template <class T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d) {
    const uint64_t* s = emu::exchange(mask, emu_bits(v));
    const int l = emu::lane_id();
    return l >= (int)d ? emu_from<T>(s[l - (int)d]) : v;
}
template <class T> static inline unsigned __match_any_sync(unsigned mask, T v) {
    const uint64_t mine = emu_bits(v);
    const uint64_t* s = emu::exchange(mask, mine);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if (((mask >> l) & 1u) && s[l] == mine) r |= 1u << l;
    return r;
}
// one host thread: atomics are plain read-modify-writes
template <class T, class U> static inline T atomicAdd(T* p, U v) { T o = *p; *p = (T)(o + (T)v); return o; }
template <class T, class U> static inline T atomicOr(T* p, U v) { T o = *p; *p = (T)(o | (T)v); return o; }
template <class T, class U> static inline T atomicMax(T* p, U v) { T o = *p; if ((T)v > o) *p = (T)v; return o; }
template <class T> static inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }

static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
// compiled with -ffp-contract=off, no fast-math: every operation below is one IEEE-754 rounding, as the _rn intrinsics
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline double __longlong_as_double(long long l) { double r; memcpy(&r, &l, 8); return r; }
static inline unsigned __float_as_uint(float f) { unsigned r; memcpy(&r, &f, 4); return r; }
static inline float __uint_as_float(unsigned u) { float r; memcpy(&r, &u, 4); return r; }
static inline float __int_as_float(int i) { float r; memcpy(&r, &i, 4); return r; }
static inline int __float_as_int(float f) { int r; memcpy(&r, &f, 4); return r; }
template <class A, class B> static inline auto min(A a, B b) -> decltype(a + b) { return a < b ? a : b; }
template <class A, class B> static inline auto max(A a, B b) -> decltype(a + b) { return a > b ? a : b; }
#endif
