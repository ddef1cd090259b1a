// tc_common.cuh STAND-IN: FUNCTIONAL emulation of the Blackwell primitives the tcgen05 GEMM kernel uses -- mbarriers,
// TMA tensor loads / stores (boxes, traversal strides, out-of-bounds zero fill, 64- and 128-byte shared-memory swizzle),
// tcgen05.alloc / mma (kind::f16, shared-memory matrix descriptors) / commit / ld, named barriers -- same names and
// signatures as csrc/tc_common.cuh, so that gemm_tc.cuh compiles unchanged against it (tests/cpu_cuda, TEST
// INFRASTRUCTURE).  Everything asynchronous completes at issue; threads are the cooperative fibers of emu.cpp, and a
// failed mbarrier wait yields.  The accumulation is plain float (the tensor core's internal order and truncation are
// not modelled): results are compared with float references to a tolerance and BETWEEN kernel variants bit for bit.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <map>

#define __grid_constant__

namespace emu {
void yield_now();
void named_barrier(int id, int count);
void note_event();
extern float tmem[128][512];
struct MBar { int count = 0, pending = 0; long long tx = 0; unsigned phase = 0; };
std::map<const void*, MBar>& mbars();
}  // namespace emu

namespace ssg {
namespace tc {

// shared-memory "addresses" are offsets into the emulated dynamic shared memory (1024-byte aligned, as on the device)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - emu::dyn_smem); }
static inline unsigned char* smem_ptr(uint32_t a) { return emu::dyn_smem + a; }
__device__ __forceinline__ bool elect_one() { return emu::lane_id() == 0; }

// ---- mbarrier
static inline void mbar_maybe_flip(emu::MBar& b) {
    emu::note_event();
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; }
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    emu::MBar& b = emu::mbars()[bar];
    b.count = b.pending = (int)count; b.tx = 0; b.phase = 0;
}
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    emu::MBar& b = emu::mbars()[bar];
    b.tx += bytes; --b.pending;
    mbar_maybe_flip(b);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    emu::MBar& b = emu::mbars()[bar];
    --b.pending;
    mbar_maybe_flip(b);
}
static inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
    emu::MBar& b = emu::mbars()[bar];
    b.tx -= bytes;
    mbar_maybe_flip(b);
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    if (emu::mbars()[bar].phase != parity) return true;      // the phase with this parity has completed
    emu::yield_now();
    return false;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- swizzle: XOR of the 16-byte-chunk bits [4, 4+B) with the row bits [7, 7+B) of the shared-memory offset
static inline uint32_t swz(uint32_t off, uint32_t mode) {        // mode: CUtensorMapSwizzle
    const uint32_t bits = mode == CU_TENSOR_MAP_SWIZZLE_128B ? 3 : mode == CU_TENSOR_MAP_SWIZZLE_64B ? 2 : mode == CU_TENSOR_MAP_SWIZZLE_32B ? 1 : 0;
    const uint32_t m = ((off >> 7) & ((1u << bits) - 1u)) << 4;
    return off ^ m;
}

// ---- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap*) {}
static inline void tma_box(const CUtensorMap* m, unsigned char* smem_dst, const int* c, bool store) {
    const uint32_t dst0 = (uint32_t)(smem_dst - emu::dyn_smem);
    uint32_t nb[5] = {1, 1, 1, 1, 1};
    for (uint32_t i = 0; i < m->rank; ++i) nb[i] = (m->box[i] + m->estr[i] - 1) / m->estr[i];
    uint32_t lin = 0;
    for (uint32_t b4 = 0; b4 < nb[4]; ++b4)
    for (uint32_t b3 = 0; b3 < nb[3]; ++b3)
    for (uint32_t b2 = 0; b2 < nb[2]; ++b2)
    for (uint32_t b1 = 0; b1 < nb[1]; ++b1)
    for (uint32_t b0 = 0; b0 < nb[0]; ++b0, ++lin) {
        const uint32_t bi[5] = {b0, b1, b2, b3, b4};
        bool inb = true;
        uint64_t goff = 0;
        for (uint32_t i = 0; i < m->rank; ++i) {
            const long long g = (long long)c[i] + (long long)bi[i] * m->estr[i];
            if (g < 0 || g >= (long long)m->dims[i]) { inb = false; break; }
            goff += i == 0 ? (uint64_t)g * 2 : (uint64_t)g * m->strides[i - 1];
        }
        unsigned char* s = emu::dyn_smem + swz(dst0 + lin * 2, m->swizzle);
        uint16_t* g16 = (uint16_t*)(uintptr_t)(m->base + goff);
        if (store) { if (inb) *g16 = *(uint16_t*)s; }
        else *(uint16_t*)s = inb ? *g16 : (uint16_t)0;
    }
}
static inline uint32_t tma_box_bytes(const CUtensorMap* m) {
    uint32_t n = 2;
    for (uint32_t i = 0; i < m->rank; ++i) n *= (m->box[i] + m->estr[i] - 1) / m->estr[i];
    return n;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    const int c[5] = {c0, c1, 0, 0, 0};
    tma_box(m, (unsigned char*)smem_dst, c, false);
    mbar_complete_tx(bar, tma_box_bytes(m));
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    const int c[5] = {c0, c1, c2, c3, 0};
    tma_box(m, (unsigned char*)smem_dst, c, false);
    mbar_complete_tx(bar, tma_box_bytes(m));
}

// ---- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t) { *smem_result = 0u; }
__device__ __forceinline__ void tmem_dealloc(uint32_t, uint32_t) {}
__device__ __forceinline__ void tc_fence_before() {}
__device__ __forceinline__ void tc_fence_after() {}

// element (row r, k-column e) of a K-major operand described by `desc`: start + (r / 8) * SBO + (r % 8) * row_bytes + 2e,
// then the swizzle XOR on the resulting shared-memory offset
static inline float desc_elem(uint64_t desc, int r, int e) {
    const uint32_t start = (uint32_t)(desc & 0x3fffu) << 4;
    const uint32_t sbo = (uint32_t)((desc >> 32) & 0x3fffu) << 4;
    const uint32_t layout = (uint32_t)(desc >> 61) & 7u;                     // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    const uint32_t row_bytes = layout == 2 ? 128 : layout == 4 ? 64 : 32;
    const uint32_t mode = layout == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : layout == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const uint32_t off = start + (uint32_t)(r / 8) * sbo + (uint32_t)(r % 8) * row_bytes + (uint32_t)e * 2;
    __nv_bfloat16 h;
    h.x = *(const uint16_t*)(emu::dyn_smem + swz(off, mode));
    return __bfloat162float(h);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    const int N = (int)((idesc >> 17) & 0x3fu) << 3, M = (int)((idesc >> 24) & 0x1fu) << 4;
    const int col0 = (int)(tmem_d & 0xffffu);
    float a[128][16], b[256][16];
    for (int r = 0; r < M; ++r) for (int e = 0; e < 16; ++e) a[r][e] = desc_elem(desc_a, r, e);
    for (int n = 0; n < N; ++n) for (int e = 0; e < 16; ++e) b[n][e] = desc_elem(desc_b, n, e);
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
            float acc = accumulate ? emu::tmem[r][col0 + n] : 0.f;
            for (int e = 0; e < 16; ++e) acc += a[r][e] * b[n][e];
            emu::tmem[r][col0 + n] = acc;
        }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { mbar_arrive(bar); }     // the MMAs above have completed
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    const int row = (int)(taddr >> 16) + emu::lane_id(), col = (int)(taddr & 0xffffu);
    for (int i = 0; i < 32; ++i) memcpy(&v[i], &emu::tmem[row][col + i], 4);
}
__device__ __forceinline__ void tmem_ld_wait() {}

__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)1u << 16;
    d |= (uint64_t)(1024u >> 4) << 32;
    d |= (uint64_t)1u << 46;
    d |= (uint64_t)2u << 61;
    return d;
}
__device__ __forceinline__ uint64_t make_desc_k_sw64(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)1u << 16;
    d |= (uint64_t)(512u >> 4) << 32;
    d |= (uint64_t)1u << 46;
    d |= (uint64_t)4u << 61;
    return d;
}
__host__ __device__ constexpr uint32_t make_idesc_bf16_f32(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- what gemm_tc.cuh defines with inline PTX (cut from its text by the build script)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    const int c[5] = {c0, c1, 0, 0, 0};
    tma_box(m, (unsigned char*)const_cast<void*>(smem_src), c, true);
}
__device__ __forceinline__ void tma_store_commit() {}
template <int N> __device__ __forceinline__ void tma_store_wait_read() {}
__device__ __forceinline__ void tma_store_wait_all() {}

}  // namespace tc
}  // namespace ssg
