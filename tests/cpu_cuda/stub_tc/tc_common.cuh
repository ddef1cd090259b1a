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
extern float tmem[128][512];
// asynchronous-unit model (emu_tc.cpp): eager or, with SSG_EMU_ASYNC=late, latest-legal completion
void tc_mbar_init(const void* bar, uint32_t count);
void tc_mbar_arrive(const void* bar);
void tc_mbar_arrive_expect_tx(const void* bar, uint32_t bytes);
bool tc_mbar_poll(const void* bar, uint32_t parity);
void tc_tma_load(void* smem_dst, const CUtensorMap* m, const void* bar, const int* c);
void tc_tma_store(const CUtensorMap* m, const void* smem_src, const int* c);
void tc_store_commit();
void tc_store_wait(int allowed_pending);
void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate);
void tc_mma_commit(const void* bar);
}  // namespace emu

namespace ssg {
namespace tc {

// shared-memory "addresses" are offsets into the emulated dynamic shared memory (1024-byte aligned, as on the device)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)((const unsigned char*)p - emu::dyn_smem); }
__device__ __forceinline__ bool elect_one() { return emu::lane_id() == 0; }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { emu::tc_mbar_init(bar, count); }
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { emu::tc_mbar_arrive_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { emu::tc_mbar_arrive(bar); }
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    if (emu::tc_mbar_poll(bar, parity)) return true;         // the phase with this parity has completed
    emu::yield_now();
    return false;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ---- TMA
#define SSG_NO_LAUNCH_EX 1          // gemm_tc.cuh: plain <<<>>> launches (rewritten to emu::launch) instead of cudaLaunchKernelEx
__device__ __forceinline__ void pdl_wait() {}                      // launches run one after another under emulation
__device__ __forceinline__ void pdl_launch_dependents() {}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap*) {}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    const int c[5] = {c0, c1, 0, 0, 0};
    emu::tc_tma_load(smem_dst, m, bar, c);
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
    const int c[5] = {c0, c1, c2, c3, 0};
    emu::tc_tma_load(smem_dst, m, bar, c);
}

// ---- tcgen05
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t) { *smem_result = 0u; }
__device__ __forceinline__ void tmem_dealloc(uint32_t, uint32_t) {}
__device__ __forceinline__ void tc_fence_before() {}
__device__ __forceinline__ void tc_fence_after() {}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    emu::tc_mma(tmem_d, desc_a, desc_b, idesc, accumulate);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) { emu::tc_mma_commit(bar); }
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
    emu::tc_tma_store(m, smem_src, c);
}
__device__ __forceinline__ void tma_store_commit() { emu::tc_store_commit(); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() { emu::tc_store_wait(N); }
__device__ __forceinline__ void tma_store_wait_all() { emu::tc_store_wait(0); }

}  // namespace tc
}  // namespace ssg
