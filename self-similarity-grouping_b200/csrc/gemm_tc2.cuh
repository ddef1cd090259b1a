// Two-CTA (cta_group::2) form of the convolution GEMM of gemm_tc.cuh: a cluster of two CTAs on one TPC computes a
// 256 x BN output tile.  Each CTA loads the A tile of ITS 128 rows and HALF of the B tile (BN/2 weight rows); the
// leader CTA issues `tcgen05.mma.cta_group::2` with M = 256, which reads A and the two B halves out of both CTAs'
// shared memory and accumulates each CTA's 128 rows into that CTA's own TMEM.  Per CTA and MMA the shared-memory operand
// read drops from 4 KB (A) + BN*32 B (B) to 4 KB + BN*16 B and the L2 -> shared-memory traffic of the weights halves --
// the two things that bound the single-CTA tiles (profiles/r02d_conv_l1l2_full.md: ~64 B/clk of operand bandwidth per
// UMMA gives 128x128 tiles a ceiling of 50 % and 128x256 tiles 67 % of the tensor peak; pairs lift that to 67 / 100 %).
// Protocol (after the CUTLASS sm100 2-SM kernels): both producers signal the LEADER's `full` barrier (TMA
// `.cta_group::2`, the leader arms it with the bytes of both CTAs); `tcgen05.commit ... multicast::cluster` frees the
// operand stage and publishes the accumulator in BOTH CTAs; the epilogue warps of both CTAs arrive on the leader's
// `tempty` barrier (remote mbarrier arrive).  Epilogue per CTA as in gemm_tc.cuh (bias, residual ring, ReLU, bf16,
// swizzled staging, TMA store).  Same MMA order per output element as the single-CTA kernels, hence the same bits.
#pragma once
#include "gemm_tc.cuh"

namespace ssg {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_addr` (a shared::cta address of this CTA) in CTA `cta` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of the pair: destination in this CTA, completion bytes on the barrier at `bar_cluster` (the leader's)
__device__ __forceinline__ void tma2_load_2d(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                             int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at this shared-memory offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void umma2_commit_both(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

// KHS (kernel-row sharing, the C = 64 3x3 convolutions of layer 1; see gemm_tc.cuh): a stage holds this CTA's haloed
// activation box for one kernel COLUMN (192 pixel rows) and its halves of the three weight tiles of that column.
template <int BN, bool RES, bool KHS = false>
struct Smem2 {
    static constexpr int A_BYTES = (KHS ? 192 : BM) * BK * 2;    // this CTA's 128 rows (KHS: + halo)
    static constexpr int B_TILE = (BN / 2) * BK * 2;             // this CTA's half of one weight tile
    static constexpr int B_BYTES = (KHS ? 3 : 1) * B_TILE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SUB_BYTES = BM * 128;
    static constexpr int NSUB = BN / 64;
    static constexpr int RSLOTS = 3;
    static constexpr int C_BYTES = 2 * SUB_BYTES;
    static constexpr int R_BYTES = RES ? RSLOTS * SUB_BYTES : 0;
    static constexpr int BUDGET = 232448 - 1024 /* static */ - 384 - 1024 - C_BYTES - R_BYTES;
    static constexpr int STAGES = BUDGET / STAGE_BYTES > 6 ? 6 : BUDGET / STAGE_BYTES;
    static constexpr int C_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int BAR_OFFSET = C_OFFSET + C_BYTES + R_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 384 + 1024;
    static_assert(STAGES >= 3, "pair kernel: too few operand stages");
    static_assert(!(KHS && (RES || BN != 64)), "pair kernel: KHS is the 64 -> 64 channel 3x3 convolution");
};

template <int BN, bool RES, bool KHS = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm2_kernel(const __grid_constant__ AOperand A, const __grid_constant__ CUtensorMap mapB, int M, int N, int num_k_blocks,
             const __grid_constant__ StagedEpi epi) {
    using L = Smem2<BN, RES, KHS>;
    constexpr int STAGES = L::STAGES, NSUB = L::NSUB, RS = L::RSLOTS;
    constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256 : 512;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;      // [2]
    uint64_t* tempty_bar = tfull_bar + 2;          // [2] (used in the leader CTA: 2 x EPI_WARPS arrivals)
    uint64_t* res_bar = tempty_bar + 2;            // [3]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(res_bar + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();       // 0 = leader
    const int pair = (int)blockIdx.x >> 1, num_pairs = (int)gridDim.x >> 1;
    const int m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
    const int mp_blocks = (m_blocks + 1) / 2;      // pair tiles along M (the last one may have an empty second half)
    const int num_tiles = mp_blocks * n_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&A.map[0]);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 2 * EPI_WARPS); }
        for (int s = 0; s < 3; ++s) mbar_init(&res_bar[s], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc2(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();                            // barriers initialised and TMEM allocated in BOTH CTAs
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    pdl_wait();                                    // see gemm_tc.cuh: the prologue overlapped the previous kernel's tail
    pdl_launch_dependents();

    auto tile_coords = [&](int t, int& mp, int& n_blk) {
        const int per_group = GROUP_M * n_blocks;
        const int g = t / per_group;
        const int first_m = g * GROUP_M;
        const int gsz = min(GROUP_M, mp_blocks - first_m);
        const int r = t - g * per_group;
        mp = first_m + r % gsz;
        n_blk = r / gsz;
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                int mp, n_blk;
                tile_coords(t, mp, n_blk);
                const int m_blk = 2 * mp + (int)rank;
                int b0 = 0, h0 = 0;
                if (A.mode == 1 || (A.kb_split > 0 && A.mode1 == 1)) {
                    if (A.bb > 1) { b0 = m_blk * A.bb; }
                    else { b0 = m_blk / A.tiles_per_img; h0 = (m_blk % A.tiles_per_img) * A.bh; }
                }
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * L::STAGE_BYTES;
                    unsigned char* sb = sa + L::A_BYTES;
                    const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::STAGE_BYTES);   // both CTAs' bytes
                    if constexpr (KHS) {
                        // stage kb = kernel column kw (one channel block): rows h0-1 .. h0+bh of this CTA's tile, and this
                        // CTA's half (32 output channels) of the weight tiles of taps (kh, kw), kh = 0..2
                        tma2_load_4d(sa, &A.map[0], full_leader, 0, kb - 1, h0 - 1, b0);
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh)
                            tma2_load_2d(sb + kh * L::B_TILE, &mapB, full_leader, (kh * 3 + kb) * BK, (int)rank * (BN / 2));
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (A.kb_split > 0 && kb >= A.kb_split) {
                        const int kb2 = kb - A.kb_split;
                        if (A.mode1 == 0) tma2_load_2d(sa, &A.map[1], full_leader, kb2 * BK, m_blk * BM);
                        else tma2_load_4d(sa, &A.map[1], full_leader, kb2 * BK, 0, h0 * A.hmul, b0);
                    } else if (A.mode == 0) {
                        tma2_load_2d(sa, &A.map[0], full_leader, kb * BK, m_blk * BM);
                    } else {
                        const int tap = kb / A.cblks, cb = kb - tap * A.cblks;
                        tma2_load_4d(sa, &A.map[A.tap_plane[tap]], full_leader, cb * BK, A.tap_dw[tap],
                                     h0 * A.hmul + A.tap_dh[tap], b0);
                    }
                    tma2_load_2d(sb, &mapB, full_leader, kb * BK, n_blk * BN + (int)rank * (BN / 2));
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16_f32(2 * BM, BN);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);       // both CTAs' epilogues have drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                        if constexpr (KHS) {
                            // three kernel rows out of one haloed buffer: A starts kh*bw pixel rows in (both CTAs alike)
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh) {
                                const uint64_t da = make_desc_k_sw128(sa + (uint32_t)(kh * A.khs_row_bytes));
                                const uint64_t db = make_desc_k_sw128(sa + L::A_BYTES + (uint32_t)(kh * L::B_TILE));
#pragma unroll
                                for (int k = 0; k < BK / UMMA_K; ++k)
                                    umma2_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                              (kb > 0 || kh > 0 || k > 0) ? 1u : 0u);
                            }
                        } else {
                        const uint64_t da = make_desc_k_sw128(sa), db = make_desc_k_sw128(sa + L::A_BYTES);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma2_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        }
                        umma2_commit_both(&empty_bar[stage]);
                        if (kb == num_k_blocks - 1) umma2_commit_both(&tfull_bar[acc]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9, both CTAs) =====================
        const int q = warp & 3;
        const int grp = (warp - 2) >> 2;
        int acc = 0;
        uint32_t acc_phase = 0;
        unsigned char* c_s = smem + L::C_OFFSET;
        unsigned char* r_s = c_s + L::C_BYTES;
        const bool has_res = RES && epi.has_res;
        const bool leader = (warp == 2 && lane == 0);
        const int r_in = q * 32 + lane;
        const uint32_t row_off = (uint32_t)r_in * 128u;
        const uint32_t sw = (uint32_t)(r_in & 7);
        __shared__ float s_bias[BN];
        const int epi_tid = threadIdx.x - 64;
        // residual sub-tile s (s counts this CTA's 64-column sub-tiles) -> ring slot s % RS, fetched RS - 1 ahead
        auto load_residual_sub = [&](int s) {
            const int tile = pair + (s / NSUB) * num_pairs;
            if (tile >= num_tiles) return;
            int mp, nb;
            tile_coords(tile, mp, nb);
            const int slot = s % RS;
            mbar_arrive_expect_tx(&res_bar[slot], L::SUB_BYTES);
            tma_load_2d(r_s + slot * L::SUB_BYTES, &epi.mapR, &res_bar[slot], nb * BN + (s % NSUB) * 64, (2 * mp + (int)rank) * BM);
        };
        if (leader && has_res) {
            for (int s0 = 0; s0 < RS - 1; ++s0) load_residual_sub(s0);
        }
        const uint32_t tempty_leader0 = mapa_u32(smem_u32(&tempty_bar[0]), 0), tempty_leader1 = mapa_u32(smem_u32(&tempty_bar[1]), 0);
        int it = 0;
        for (int t = pair; t < num_tiles; t += num_pairs, ++it) {
            int mp, n_blk;
            tile_coords(t, mp, n_blk);
            const int m_blk = 2 * mp + (int)rank;
            if (epi_tid < BN) s_bias[epi_tid] = epi.bias[n_blk * BN + epi_tid];      // visible after the next barrier
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int j = 0; j < NSUB; ++j) {
                const int g = it * NSUB + j, b = g & 1;
                if (leader) tma_store_wait_read<1>();                   // the store issued from buffer b has been read
                epi_bar_sync();
                unsigned char* csub = c_s + b * L::SUB_BYTES + row_off;
                const unsigned char* rsub = nullptr;
                if (has_res) {
                    if (leader) load_residual_sub(g + RS - 1);            // slot of sub-tile g - 1: every thread is past it
                    mbar_wait(&res_bar[g % RS], (uint32_t)((g / RS) & 1));
                    rsub = r_s + (g % RS) * L::SUB_BYTES + row_off;
                }
                const int c = 2 * j + grp;
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                tmem_ld_wait();
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const uint32_t chunk = ((uint32_t)(grp * 4 + gq) ^ sw) << 4;
                    float f[8];
                    const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c * 32 + 8 * gq);
                    const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c * 32 + 8 * gq + 4);
                    f[0] = __uint_as_float(v[8 * gq + 0]) + b0.x; f[1] = __uint_as_float(v[8 * gq + 1]) + b0.y;
                    f[2] = __uint_as_float(v[8 * gq + 2]) + b0.z; f[3] = __uint_as_float(v[8 * gq + 3]) + b0.w;
                    f[4] = __uint_as_float(v[8 * gq + 4]) + b1.x; f[5] = __uint_as_float(v[8 * gq + 5]) + b1.y;
                    f[6] = __uint_as_float(v[8 * gq + 6]) + b1.z; f[7] = __uint_as_float(v[8 * gq + 7]) + b1.w;
                    if (rsub != nullptr) {
                        const uint4 rr = *reinterpret_cast<const uint4*>(rsub + chunk);
                        const __nv_bfloat162* rp = reinterpret_cast<const __nv_bfloat162*>(&rr);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 ff = __bfloat1622float2(rp[e]);
                            f[2 * e] += ff.x;
                            f[2 * e + 1] += ff.y;
                        }
                    }
                    if (epi.relu) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                    }
                    uint4 pk;
                    __nv_bfloat162* pp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                    for (int e = 0; e < 4; ++e) pp[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                    *reinterpret_cast<uint4*>(csub + chunk) = pk;
                }
                if (j == NSUB - 1) {                                     // accumulator drained: tell the leader's MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
                }
                fence_proxy_async();
                epi_bar_sync();
                if (leader) {
                    tma_store_2d(&epi.mapC, c_s + b * L::SUB_BYTES, n_blk * BN + j * 64, m_blk * BM);
                    tma_store_commit();
                }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (leader) tma_store_wait_all();
    }
    tc_fence_before();
    cluster_sync_all();                            // nobody touches the peer's barriers / TMEM after this point
    if (warp == 1) tmem_dealloc2(tmem_base, TMEM_COLS);
}

// Launch: clusters of two CTAs, one cluster per pair of SMs.
template <int BN, bool RES, bool KHS = false>
int launch_gemm2_op(const AOperand& A, int m, const void* b, int n, int k, const StagedEpi& epi, cudaStream_t st) {
    using L = Smem2<BN, RES, KHS>;
    if (KHS && (n != BN || k != 9 * BK || A.cblks != 1))
        return ssg_set_error(SSG_ERR_INVALID, "gemm2: the KHS pair kernel needs C = N = 64 (N=%d, K=%d)", n, k);
    if (k % 8 || n % BN) return ssg_set_error(SSG_ERR_INVALID, "gemm2: N=%d must be a multiple of %d, K=%d of 8", n, BN, k);
    CUtensorMap mapB;
    SSG_TRY(make_tmap_2d_bf16(&mapB, b, (uint64_t)n, (uint64_t)k, (uint64_t)k, BN / 2));
    int sms = 0;
    SSG_TRY(tc_num_sms(&sms));
    const int m_blocks = (m + BM - 1) / BM;
    const int tiles = ((m_blocks + 1) / 2) * (n / BN);
    int pairs = tiles < sms / 2 ? tiles : sms / 2;
    if (pairs < 1) pairs = 1;
    auto kern = gemm2_kernel<BN, RES, KHS>;
    SSG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    const int nkb = KHS ? 3 : (k + BK - 1) / BK;         // KHS: one K block per kernel column
    SSG_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, A, mapB, m, n, nkb, epi));
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // namespace tc
}  // namespace ssg
