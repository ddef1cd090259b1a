// cuda_bf16.h STAND-IN for the CPU emulation (tests/cpu_cuda): bfloat16 storage type with round-to-nearest-even
// conversion, which is all the plain-CUDA operand-split kernel needs.  TEST INFRASTRUCTURE.
#pragma once
#include <stdint.h>
#include <string.h>
struct __nv_bfloat16 { uint16_t x; };
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    __nv_bfloat16 r;
    if ((u & 0x7fffffffu) > 0x7f800000u) { r.x = 0x7fff; return r; }           // NaN
    u += 0x7fffu + ((u >> 16) & 1u);                                              // round to nearest, ties to even
    r.x = (uint16_t)(u >> 16);
    return r;
}
static inline float __bfloat162float(__nv_bfloat16 h) {
    uint32_t u = (uint32_t)h.x << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

struct __nv_bfloat162 { __nv_bfloat16 x, y; };
struct emu_float2 { float x, y; };
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { __nv_bfloat162 r; r.x = __float2bfloat16_rn(a); r.y = __float2bfloat16_rn(b); return r; }
#ifdef __cplusplus
#include <cuda_runtime.h>
static inline float2 __bfloat1622float2(__nv_bfloat162 v) { return make_float2(__bfloat162float(v.x), __bfloat162float(v.y)); }
static inline __nv_bfloat162 __hmax2(__nv_bfloat162 a, __nv_bfloat162 b) {
    __nv_bfloat162 r;
    r.x = __bfloat162float(a.x) >= __bfloat162float(b.x) ? a.x : b.x;
    r.y = __bfloat162float(a.y) >= __bfloat162float(b.y) ? a.y : b.y;
    return r;
}
#endif
