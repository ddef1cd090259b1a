// tcgen05 GEMM kernel template (see gemm_tc.cu for the overview).  Included by gemm_tc.cu and conv.cu.
#pragma once
#include <string.h>
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>

namespace ssg {

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows);
int make_tmap_nhwc_bf16(CUtensorMap* map, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C,
                        uint32_t bw, uint32_t bh, uint32_t bb, uint32_t stride = 1);
int make_tmap_stem_windows(CUtensorMap* map, const void* base, uint64_t images);
int make_tmap_stem_windows64(CUtensorMap* map, const void* base, uint64_t images, uint32_t box_rows = 4);
int make_tmap_2d_bf16_sw64(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
int tc_num_sms(int* out);
bool pdl_enabled();            // SSG_PDL (default 0: measured no gain): launch the GEMM kernels with programmatic stream serialization

namespace tc {

constexpr int BM = 128;        // UMMA M (one TMEM lane per output row)
constexpr int BK = 64;         // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;          // two warps per TMEM lane quarter: each takes every other 32-column chunk
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
constexpr int GROUP_M = 16;    // tile rasterisation: 16 m-blocks share each B tile while it is L2-hot

// A operand: plain [M,K] matrix (mode 0) or implicit-GEMM view of NHWC activations (mode 1): the K axis runs over
// (tap, channel block); each tap reads a shifted box of one of up to four (stride-parity) planes; halo
// elements are zero-filled by TMA.
struct AOperand {
    CUtensorMap map[4];
    int mode;
    int cblks;             // channel blocks of 64 per tap
    int taps;
    int bh, bb;            // box rows / images per 128-pixel tile
    int tiles_per_img;     // >= 1
    int hmul;              // input rows per output row (2 for stride-2 boxes read straight from the input map)
    int kb_split;          // > 0: K blocks >= kb_split come from a second source, map[1] (K-concatenated GEMM)
    int mode1;             // second source: 0 = plain [M,K1] matrix, 1 = single-tap implicit view (uses bh/bb/hmul)
    int khs_row_bytes;     // KHS kernels: bytes of one image row of the tile in the haloed A buffer (bw * 128)
    int ksplit_mblks;      // > 0 (plain mode, split-K batches stacked along M: weight-gradient GEMM): m-block mb belongs to
                           // split mb / ksplit_mblks and reads B's K blocks from (mb / ksplit_mblks) * num_k_blocks on
    signed char tap_plane[9], tap_dh[9], tap_dw[9];
};

// Staged epilogue (bf16 output through shared memory + TMA store, residual tile fetched by TMA): parameters.
struct StagedEpi {
    CUtensorMap mapC;      // output  [M, N] bf16, box 64 cols x 128 rows, 128B swizzle
    CUtensorMap mapR;      // residual, same geometry (valid when has_res)
    const float* bias;     // [N]
    int relu;
    int has_res;
    void* pool_out;        // VAR_BRES only: != NULL fuses the 3x3/2 max-pool behind the stem (the conv map is not stored)
    // VAR_CHAIN only: the NEXT 1x1 convolution (the following block's conv1) fused behind this one
    CUtensorMap mapW2;     // its weights [n2, 256] bf16, box 64 cols x n2 rows, 128B swizzle
    CUtensorMap mapC2;     // its output  [M, n2] bf16, box 64 cols x 128 rows
    const float* bias2;    // [n2]
    int n2;                // 64 or 128 output channels
    static constexpr bool kSkippable = false;   // every tile is computed
};

// BN <= 128 (staged): one output sub-buffer per 64-column sub-tile + two residual tile buffers, 3-5 operand stages.
// BN == 256 (staged): K-heavy convolutions without residual: a ring of two output sub-buffers, no residual staging,
//                     4 operand stages (128x256 tiles halve the shared-memory operand traffic per MMA).
// KHS ("kernel-row sharing", 3x3 stride-1 convolutions with few channels): one stage holds the input rows of a tile
// plus a one-row halo above and below ((bh+2)*bw <= 192 pixel rows) for one kernel column, and the three weight
// tiles of that column; the three kernel rows are three MMAs whose A descriptors start kh*bw rows into the same
// buffer — the activation is read 3 times from L2 instead of 9.
// VAR (kernel variant, staged epilogue only):
//   VAR_BRES  (stem, BN = 64, K = 256): the whole B operand (the 32 KB of stem weights) is loaded ONCE per CTA into a
//             resident region and the operand stages carry A only (8 stages of 16 KB): the per-tile weight re-load was a
//             third of the kernel's L2 -> shared-memory traffic.
//   VAR_RRING (BN = 256 WITH a residual): 3 operand stages, and the residual arrives through a ring of three 64-column
//             sub-tile buffers fetched by TMA two sub-tiles ahead (a whole-tile double buffer does not fit next to
//             128x256 operand stages).
//             (A six-slot ring with two operand stages was measured for the short-K launches and changed nothing:
//             those are bound by the epilogue, not by HBM latency -- profiles/r01m_conv_variants.md.)
//   VAR_BRESP (stem, "parity planes"): VAR_BRES whose operand stage is a whole tile: the 10 input rows a two-row output
//             tile touches are loaded ONCE, as two stride-2 TMA boxes of five rows (even / odd offsets: 2 x 20 KB), and
//             kernel row kh reads plane kh & 1 starting (kh >> 1) rows in -- 40 KB per tile instead of the 64 KB of eight
//             per-kernel-row boxes.  Same MMA sequence, hence the same bits.
//   VAR_KHSB  (KHS with resident weights; C = 64 -> 64 3x3 convolutions of layer 1): all nine 64x64 weight tiles
//             (72 KB) are loaded ONCE per CTA and the operand stages carry the haloed activation box only (4 stages of
//             24 KB).  The plain KHS kernel re-fetches the 72 KB of weights for every 128-pixel tile, which is half of
//             its L2 -> shared-memory traffic (144 KB per tile at ~42 B/clk/SM: the launch is L2-bandwidth bound).
//             Same MMA sequence, hence the same bits.
//   VAR_CHAIN (BN = 256 = all output channels of a layer-1 block; K = 64 or 128): conv3 (+ residual or K-concatenated
//             downsample, ReLU) AND the next block's conv1 in one kernel.  The bf16 output sub-tiles that the epilogue
//             stages for the TMA store are valid 128B-swizzled K-major A operands: after sub-tile j (64 output channels
//             = K block j of the next 1x1) is staged, the MMA warp issues D2 += ysub_j * W1'_j^T into a second TMEM
//             region while the TMA store drains the same buffer; when the four sub-tiles are done the epilogue drains
//             D2 (+bias', ReLU) and stores t1' -- the next block never re-reads y for its conv1 (512 of layer 1's
//             2048 HBM bytes per pixel and block) and one launch per block disappears.  Both weight matrices
//             (W3: 32 / 64 KB, W1': 32 / 64 KB) are resident in shared memory, the operand stages carry A only, the
//             accumulator is single-buffered (the kernel is HBM bound: ~6 K cycles per tile against ~0.5 K of MMAs).
//             Same MMA order per output element as the two separate launches, hence the same bits.
//   VAR_NORES (BN <= 128, no residual): the two whole-tile residual buffers of the generic layout (32 / 64 KB) become
//             operand stages -- 6 stages of 32 KB (BN = 128) or 8 of 24 KB (BN = 64) instead of 3 / 5.  ncu (round 2,
//             profiles/r02d_*): in the 128x128 3x3 launches the MMA warp waits for operands 65 % of the time and the
//             producer for a free stage, with L2 and DRAM far from saturated -- the loads are latency bound, so bytes
//             in flight are what counts.  Same MMA sequence, hence the same bits.
constexpr int VAR_NONE = 0, VAR_BRES = 1, VAR_RRING = 2, VAR_BRESP = 3, VAR_KHSB = 4, VAR_CHAIN = 5, VAR_NORES = 6;
constexpr int CHAIN_WRES_BYTES = 96 * 1024;  // resident W3 + W1' (never both 64 KB)
constexpr int CHAIN_D2_COL = 256;            // TMEM column of the second accumulator
constexpr int STEM_ROW_BYTES = 64 * 64;     // one input row of a stem tile in shared memory: 64 windows x 64 B
constexpr int BRES_K = 256;

// EPI2 (opt-in, SSG_CONV_EPI2=1; generic staged epilogue only): ONE named barrier per 64-column sub-tile instead of
// two (the leader waits for the previous TMA store to leave the other staging buffer BEFORE the barrier that publishes
// the current sub-tile, so that barrier also frees the next buffer), software-pipelined tcgen05.ld (the load of
// sub-tile j+1 is in flight while j is converted), the accumulator released as soon as its last columns are in
// registers, bias rows double-buffered per tile, residual ring refilled a full ring ahead.  Always two staging buffers.
template <int BN, bool STAGED, bool KHS = false, int VAR = VAR_NONE, bool EPI2 = false>
struct SmemLayout {
    static constexpr bool PLANES = VAR == VAR_BRESP;
    static constexpr bool BRES = VAR == VAR_BRES || PLANES, RRING = VAR == VAR_RRING;
    static constexpr bool KRES = VAR == VAR_KHSB;                // KHS + resident weights
    static constexpr bool CHAIN = VAR == VAR_CHAIN;              // conv3 + next conv1
    static constexpr bool NORES = VAR == VAR_NORES || KHS;       // no residual staging (KHS kernels never have one)
    static constexpr int RSLOTS = 3;                             // residual ring slots (RRING)
    static constexpr int A_BYTES = PLANES ? 2 * 5 * STEM_ROW_BYTES : (KHS ? 192 : BM) * BK * 2;
    static constexpr int B_TILE = BN * BK * 2;
    static constexpr int B_BYTES = (BRES || KRES || CHAIN) ? 0 : (KHS ? 3 : 1) * B_TILE;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SUB_BYTES = BM * 128;                   // one 128-row x 64-col bf16 sub-tile
    static constexpr int NSUB = BN / 64;
    static constexpr int NBUF = EPI2 ? 2 : (NSUB > 2 ? 2 : NSUB);   // output sub-buffers
    static constexpr bool HAS_R = (BN <= 128 && !NORES) || RRING || CHAIN;   // residual staging available
    static constexpr int C_BYTES = STAGED ? NBUF * SUB_BYTES : 0;
    static constexpr int R_BYTES = (STAGED && BN <= 128 && !NORES) ? NSUB * SUB_BYTES : 0;   // one residual tile
    static constexpr int RSTAGE_BYTES = (RRING || CHAIN) ? RSLOTS * SUB_BYTES : 2 * R_BYTES;  // residual staging in total
    static constexpr int STAGES = PLANES ? 3 : BRES ? 8 : RRING ? 3 : KRES ? 5 : CHAIN ? 2 : KHS ? 4 :
        VAR == VAR_NORES ? (BN <= 64 ? 8 : 6) :
        (STAGED ? (BN <= 64 ? 5 : (BN <= 128 ? 3 : 4)) : ((BN <= 64) ? 8 : (BN <= 128 ? 6 : 4)));
    static constexpr int BRES_OFFSET = STAGES * STAGE_BYTES;     // resident B operand (VAR_BRES)
    static constexpr int BRES_BYTES = BRES ? BN * BRES_K * 2 : KRES ? 9 * B_TILE : CHAIN ? CHAIN_WRES_BYTES : 0;
    static constexpr int C_OFFSET = BRES_OFFSET + BRES_BYTES;    // output staging, then the residual staging buffers
    static constexpr int W_OFFSET = BRES_OFFSET;                 // resident weights (VAR_CHAIN)
    static constexpr int BAR_OFFSET = BRES_OFFSET + BRES_BYTES + C_BYTES + RSTAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 384 + 1024;        // barriers (up to 2*8 + 16 of them) + alignment slack
    static_assert(!(BRES && (KHS || !STAGED || BN != 64)), "VAR_BRES is the stem kernel");
    static_assert(!(RRING && (KHS || !STAGED || BN != 256)), "VAR_RRING is the 128x256 residual kernel");
    static_assert(!(KRES && (!KHS || !STAGED || BN != 64)), "VAR_KHSB is the 64-channel kernel-row-sharing kernel");
    static_assert(!(CHAIN && (KHS || !STAGED || BN != 256 || EPI2)), "VAR_CHAIN is the layer-1 conv3 + conv1 kernel");
    static_assert(!(VAR == VAR_NORES && (KHS || !STAGED || BN > 128 || EPI2)), "VAR_NORES: generic staged kernel, BN <= 128");
    static_assert(!(EPI2 && (BRES || !STAGED)), "EPI2 is a variant of the generic staged epilogue");
    static_assert(TOTAL <= 232448, "shared-memory layout exceeds 227 KB");
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * EPI_WARPS) : "memory"); }

template <int BN, class Epi, bool STAGED, bool KHS = false, int VAR = VAR_NONE, bool EPI2 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ AOperand A, const __grid_constant__ CUtensorMap mapB, int M, int N,
            int num_k_blocks, const __grid_constant__ Epi epi) {
    using L = SmemLayout<BN, STAGED, KHS, VAR, EPI2>;
    constexpr int STAGES = L::STAGES;
    constexpr uint32_t TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;     // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;         // [2] accumulator drained
    uint64_t* res_bar = tempty_bar + 2;           // [6] residual tile / sub-tile landed (staged epilogue)
    uint64_t* bres_bar = res_bar + 6;             // [1] resident B operand landed (VAR_BRES / KHSB / CHAIN)
    uint64_t* ysub_full = bres_bar + 1;           // [2] VAR_CHAIN: staged output sub-tile ready as an A operand
    uint64_t* ysub_free = ysub_full + 2;          // [2] VAR_CHAIN: the MMAs reading that buffer have retired
    uint64_t* d2_full = ysub_free + 2;            // [1] VAR_CHAIN: second accumulator complete
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d2_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
    const int num_tiles = m_blocks * n_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&A.map[0]);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], EPI_WARPS); }
        for (int s = 0; s < 6; ++s) mbar_init(&res_bar[s], 1);
        mbar_init(bres_bar, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&ysub_full[s], 1); mbar_init(&ysub_free[s], 1); }
        mbar_init(d2_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped
    // the tail of the previous kernel of the stream; nothing below may touch global memory before that kernel is done.
    // EVERY kernel of the chain waits, also one that does not read its predecessor's output: completion is transitive
    // only that way (a residual comes from two or three launches back).
    pdl_wait();
    pdl_launch_dependents();

    // tile index -> (m_blk, n_blk), grouped along M so that concurrently resident CTAs share B tiles in L2
    auto tile_coords = [&](int t, int& m_blk, int& n_blk) {
        const int per_group = GROUP_M * n_blocks;
        const int g = t / per_group;
        const int first_m = g * GROUP_M;
        const int gsz = min(GROUP_M, m_blocks - first_m);
        const int r = t - g * per_group;
        m_blk = first_m + r % gsz;
        n_blk = r / gsz;
    };

    // tile sequence of this CTA: round-robin, except for the stem with the fused max-pool, where a CTA owns whole
    // images and walks their 64 two-row tiles in order (the pool needs the previous tile's last row)
    int t_begin = blockIdx.x, t_end = num_tiles, t_step = gridDim.x;
    if constexpr (L::BRES) {
        if (epi.pool_out != nullptr) {
            const long long images = m_blocks / 64;
            t_begin = (int)(images * blockIdx.x / gridDim.x) * 64;
            t_end = (int)(images * (blockIdx.x + 1) / gridDim.x) * 64;
            t_step = 1;
        }
    }

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            if constexpr (L::BRES) {
                // the whole [BN, 256] weight matrix, once: K block kb at kb * B_TILE, two 64-byte-row halves each
                unsigned char* bres = smem + L::BRES_OFFSET;
                mbar_arrive_expect_tx(bres_bar, L::BRES_BYTES);
#pragma unroll
                for (int kb = 0; kb < BRES_K / BK; ++kb) {
                    tma_load_2d(bres + kb * L::B_TILE, &mapB, bres_bar, kb * BK, 0);
                    tma_load_2d(bres + kb * L::B_TILE + L::B_TILE / 2, &mapB, bres_bar, kb * BK + 32, 0);
                }
            }
            if constexpr (L::CHAIN) {
                // W3 [256, K] as K/64 blocks of [256 x 64] (32 KB each), then W1' [n2, 256] as four [n2 x 64] K blocks
                unsigned char* wres = smem + L::W_OFFSET;
                const int n2 = epi.n2;
                mbar_arrive_expect_tx(bres_bar, (uint32_t)(num_k_blocks * L::B_TILE + 4 * n2 * 128));
                for (int kb = 0; kb < num_k_blocks; ++kb) tma_load_2d(wres + kb * L::B_TILE, &mapB, bres_bar, kb * BK, 0);
                for (int j = 0; j < 4; ++j)
                    tma_load_2d(wres + num_k_blocks * L::B_TILE + j * n2 * 128, &epi.mapW2, bres_bar, j * BK, 0);
            }
            if constexpr (L::KRES) {
                // the nine [64, 64] weight tiles (tap t = kh*3 + kw at K offset t*64: one channel block), once
                unsigned char* bres = smem + L::BRES_OFFSET;
                mbar_arrive_expect_tx(bres_bar, L::BRES_BYTES);
#pragma unroll
                for (int tp = 0; tp < 9; ++tp) tma_load_2d(bres + tp * L::B_TILE, &mapB, bres_bar, tp * BK, 0);
            }
            for (int t = t_begin; t < t_end; t += t_step) {
                int m_blk, n_blk;
                tile_coords(t, m_blk, n_blk);
                if constexpr (Epi::kSkippable) { if (epi.skip_tile(m_blk, n_blk, BN)) continue; }
                int b0 = 0, h0 = 0;
                if (A.mode == 1 || (A.kb_split > 0 && A.mode1 == 1)) {
                    if (A.bb > 1) { b0 = m_blk * A.bb; }
                    else { b0 = m_blk / A.tiles_per_img; h0 = (m_blk % A.tiles_per_img) * A.bh; }
                }
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char* sa = smem + stage * L::STAGE_BYTES;
                    unsigned char* sb = sa + L::A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
                    if constexpr (KHS) {
                        // stage kb = (kernel column kw, channel block cb): rows h0-1 .. h0+bh of the tile's image
                        const int kw = kb / A.cblks, cb = kb - kw * A.cblks;
                        tma_load_4d(sa, &A.map[0], &full_bar[stage], cb * BK, kw - 1, h0 - 1, b0);
                        if constexpr (!L::KRES) {
#pragma unroll
                            for (int kh = 0; kh < 3; ++kh)
                                tma_load_2d(sb + kh * L::B_TILE, &mapB, &full_bar[stage],
                                            ((kh * 3 + kw) * A.cblks + cb) * BK, n_blk * BN);
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (A.kb_split > 0 && kb >= A.kb_split) {
                        const int kb2 = kb - A.kb_split;
                        if (A.mode1 == 0) tma_load_2d(sa, &A.map[1], &full_bar[stage], kb2 * BK, m_blk * BM);
                        else tma_load_4d(sa, &A.map[1], &full_bar[stage], kb2 * BK, 0, h0 * A.hmul, b0);
                    } else if (A.mode == 0) {
                        tma_load_2d(sa, &A.map[0], &full_bar[stage], kb * BK, m_blk * BM);
                    } else if (L::PLANES) {
                        // stem, whole tile per stage: input rows r0, r0+2, .., r0+8 (plane 0) and r0+1, .., r0+9 (plane 1)
                        const int r0 = 4 * (m_blk & 63) - 3;
                        tma_load_4d(sa, &A.map[0], &full_bar[stage], 0, 0, r0, m_blk >> 6);
                        tma_load_4d(sa + L::A_BYTES / 2, &A.map[0], &full_bar[stage], 0, 0, r0 + 1, m_blk >> 6);
                    } else if (A.mode == 3) {
                        // stem, 64-byte rows: the stage holds two kernel rows (kh = 2kb, 2kb+1) as two [128 x 64 B] halves
                        const int ih = 4 * (m_blk & 63) + 2 * kb - 3;
                        tma_load_4d(sa, &A.map[0], &full_bar[stage], 0, 0, ih, m_blk >> 6);
                        tma_load_4d(sa + L::A_BYTES / 2, &A.map[0], &full_bar[stage], 0, 0, ih + 1, m_blk >> 6);
                    } else if (A.mode == 2) {
                        // stem: K block kb = kernel row kh; tile = output rows (2t, 2t+1) of image m_blk / 64
                        tma_load_4d(sa, &A.map[0], &full_bar[stage], 0, 0, 4 * (m_blk & 63) + kb - 3, m_blk >> 6);
                    } else {
                        const int tap = kb / A.cblks, cb = kb - tap * A.cblks;
                        tma_load_4d(sa, &A.map[A.tap_plane[tap]], &full_bar[stage], cb * BK, A.tap_dw[tap],
                                    h0 * A.hmul + A.tap_dh[tap], b0);
                    }
                    if constexpr (L::BRES || L::CHAIN) {
                        // weights are resident: the stage carries A only
                    } else if (A.mode == 3) {
                        tma_load_2d(sb, &mapB, &full_bar[stage], kb * BK, n_blk * BN);
                        tma_load_2d(sb + L::B_TILE / 2, &mapB, &full_bar[stage], kb * BK + 32, n_blk * BN);
                    } else {
                        const int kb_b = A.ksplit_mblks > 0 ? (m_blk / A.ksplit_mblks) * num_k_blocks + kb : kb;
                        tma_load_2d(sb, &mapB, &full_bar[stage], kb_b * BK, n_blk * BN);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_bf16_f32(BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        if constexpr (L::BRES || L::KRES || L::CHAIN) {
            mbar_wait(bres_bar, 0);                           // resident weights have landed
            tc_fence_after();
        }
        if constexpr (L::CHAIN) {
            // per tile: D1 = A * W3^T (single accumulator at column 0), then for each staged output sub-tile j
            // D2 (+)= ysub_j * W1'_j^T (column CHAIN_D2_COL).  Staging step g = it * steps + j uses buffer g & 1.
            const uint32_t idesc2 = make_idesc_bf16_f32(BM, epi.n2);
            const uint32_t wres = smem_u32(smem + L::W_OFFSET);
            const uint32_t w2 = wres + (uint32_t)(num_k_blocks * L::B_TILE);
            const uint32_t cs = smem_u32(smem + L::C_OFFSET);
            const uint32_t w2_blk = (uint32_t)(epi.n2 * 128);
            const int steps = 4 + epi.n2 / 64;
            uint32_t yph = 0;                                  // phase bit of ysub_full[b] in bit b
            int it = 0;
            for (int t = t_begin; t < t_end; t += t_step, ++it) {
                mbar_wait(&tempty_bar[0], (uint32_t)((it & 1) ^ 1));    // epilogue has drained D1 of the previous tile
                tc_fence_after();
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t da = make_desc_k_sw128(smem_u32(smem + stage * L::STAGE_BYTES));
                        const uint64_t db = make_desc_k_sw128(wres + (uint32_t)(kb * L::B_TILE));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&empty_bar[stage]);
                        if (kb == num_k_blocks - 1) umma_commit(&tfull_bar[0]);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                for (int j = 0; j < 4; ++j) {
                    const int b = (it * steps + j) & 1;
                    mbar_wait(&ysub_full[b], (yph >> b) & 1u);
                    yph ^= (1u << b);
                    tc_fence_after();
                    if (lane == 0) {
                        const uint64_t da = make_desc_k_sw128(cs + (uint32_t)(b * L::SUB_BYTES));
                        const uint64_t db = make_desc_k_sw128(w2 + (uint32_t)j * w2_blk);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            umma_f16(tmem_base + (uint32_t)CHAIN_D2_COL, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc2,
                                     (j > 0 || k > 0) ? 1u : 0u);
                        umma_commit(&ysub_free[b]);                       // staging buffer b may be overwritten
                        if (j == 3) umma_commit(d2_full);                  // second accumulator complete
                    }
                    __syncwarp();
                }
            }
            t_begin = t_end;                                              // nothing left for the generic loop below
        }
        for (int t = t_begin; t < t_end; t += t_step) {
            if constexpr (Epi::kSkippable) {
                int m_blk, n_blk;
                tile_coords(t, m_blk, n_blk);
                if (epi.skip_tile(m_blk, n_blk, BN)) continue;    // the same tiles in all three roles
            }
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);       // epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + stage * L::STAGE_BYTES);
                    const uint32_t sb = L::BRES ? smem_u32(smem + L::BRES_OFFSET) + (uint32_t)(kb * L::B_TILE)
                                                : sa + L::A_BYTES;
                    if constexpr (KHS) {
                        // three kernel rows out of one haloed buffer: A starts kh*bw pixel rows (kh*bw*128 B) in
#pragma unroll
                        for (int kh = 0; kh < 3; ++kh) {
                            const uint64_t da = make_desc_k_sw128(sa + (uint32_t)(kh * A.khs_row_bytes));
                            // resident weights: tap (kh, kw = kb) sits (kh*3 + kb) tiles into the resident region
                            const uint64_t db = make_desc_k_sw128(
                                L::KRES ? smem_u32(smem + L::BRES_OFFSET) + (uint32_t)((kh * 3 + kb) * L::B_TILE)
                                        : sb + (uint32_t)(kh * L::B_TILE));
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; ++k)
                                umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                         (kb > 0 || kh > 0 || k > 0) ? 1u : 0u);
                        }
                    } else if (L::PLANES) {
                        // kernel row kh: output rows (0, 1) of the tile read plane rows (kh >> 1, (kh >> 1) + 1) of plane
                        // kh & 1 -- 128 contiguous 64-byte rows; weights of kernel row kh sit kh * 4 KB into the
                        // resident B operand.  Same (kh, k-step) order as the per-kernel-row stages.
                        const uint32_t bres = smem_u32(smem + L::BRES_OFFSET);
                        // (kernel row 7 of the K = 8 x 32 layout carries zero weights only: its two MMAs are skipped --
                        // adding exact zeros cannot change the accumulator, and the MMA issue rate bounds this kernel)
#pragma unroll
                        for (int kh = 0; kh < 7; ++kh) {
                            const uint64_t da = make_desc_k_sw64(sa + (uint32_t)((kh & 1) * (L::A_BYTES / 2) +
                                                                              (kh >> 1) * STEM_ROW_BYTES));
                            const uint64_t db = make_desc_k_sw64(bres + (uint32_t)(kh * (L::B_TILE / 2)));
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks)
                                umma_f16(d_tmem, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc,
                                         (kh > 0 || ks > 0) ? 1u : 0u);
                        }
                    } else if (A.mode == 3) {
                        // two half-stages of 64-byte-swizzled rows, two 16-element K steps each
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t da = make_desc_k_sw64(sa + (k >> 1) * (L::A_BYTES / 2));
                            const uint64_t db = make_desc_k_sw64(sb + (k >> 1) * (L::B_TILE / 2));
                            umma_f16(d_tmem, da + (uint64_t)(2 * (k & 1)), db + (uint64_t)(2 * (k & 1)), idesc,
                                     (kb > 0 || k > 0) ? 1u : 0u);
                        }
                    } else {
                        const uint64_t da = make_desc_k_sw128(sa), db = make_desc_k_sw128(sb);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            // advance 32 bytes (16 bf16) inside the swizzle row: +2 in the (addr >> 4) field
                            umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                     (kb > 0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);                        // smem slot free once the MMAs retire
                    if (kb == num_k_blocks - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const int q = warp & 3;               // TMEM lane quarter this warp may access
        const int grp = (warp - 2) >> 2;      // 0/1: which 32-column chunk of every 64 this warp handles
        int acc = 0;
        uint32_t acc_phase = 0;
        if constexpr (!STAGED) {
            for (int t = t_begin; t < t_end; t += t_step) {
                int m_blk, n_blk;
                tile_coords(t, m_blk, n_blk);
                if constexpr (Epi::kSkippable) { if (epi.skip_tile(m_blk, n_blk, BN)) continue; }
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                const int row = m_blk * BM + q * 32 + lane;
#pragma unroll 1
                for (int c = grp; c < BN / 32; c += 2) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                    tmem_ld_wait();
                    const int col0 = n_blk * BN + c * 32;
                    if (row < M && col0 < N) epi(row, col0, min(32, N - col0), v);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        } else {
            // bf16 output staged in 128B-swizzled shared memory and written by TMA (full-line, asynchronous stores), one
            // 64-column sub-tile at a time so that the store of one sub-tile overlaps the math of the next; the residual
            // tile is fetched by TMA one tile ahead (double buffered).
            constexpr int NSUB = L::NSUB, NBUF = L::NBUF;
            unsigned char* c_s = smem + L::C_OFFSET;
            unsigned char* r_s = c_s + L::C_BYTES;                    // 2 buffers of R_BYTES (BN <= 128)
            const bool has_res = L::HAS_R && epi.has_res;
            const bool leader = (warp == 2 && lane == 0);
            const int r_in = q * 32 + lane;                           // row inside the tile
            const uint32_t row_off = (uint32_t)r_in * 128u;
            const uint32_t sw = (uint32_t)(r_in & 7);
            __shared__ float s_bias[EPI2 ? 2 * BN : BN];
            const int epi_tid = threadIdx.x - 64;
            auto load_residual = [&](int tile, int buf) {
                int mb, nb;
                tile_coords(tile, mb, nb);
                mbar_arrive_expect_tx(&res_bar[buf], L::R_BYTES);
#pragma unroll
                for (int j = 0; j < NSUB; ++j)
                    tma_load_2d(r_s + buf * L::R_BYTES + j * L::SUB_BYTES, &epi.mapR, &res_bar[buf], nb * BN + j * 64,
                                mb * BM);
            };
            if constexpr (L::BRES) {
                if (epi.pool_out != nullptr) {
                    // ---- stem with the 3x3/2 max-pool fused behind it.  Tile t = conv rows (2tr, 2tr+1) of image b,
                    // 64 px x 64 ch each, written as bf16 into one of two staging buffers (buffer = tile parity, same
                    // swizzled layout as for a TMA store); pooled row tr = max over conv rows 2tr-1 (second half of the
                    // OTHER buffer: the previous tile of the same image), 2tr, 2tr+1 and columns 2pw-1..2pw+1, clamped
                    // at the borders (a duplicate tap cannot change a maximum).  One thread = one pooled pixel x 8
                    // channels; the 64 x 32 x 64 pooled map is the only thing stored.
                    __nv_bfloat16* pool = reinterpret_cast<__nv_bfloat16*>(epi.pool_out);
                    const int pw = epi_tid >> 3, pch = epi_tid & 7;
                    int it = 0;
                    for (int t = t_begin; t < t_end; t += t_step, ++it) {
                        const int b = t >> 6, tr = t & 63;
                        if (epi_tid < BN) s_bias[epi_tid] = epi.bias[epi_tid];
                        mbar_wait(&tfull_bar[acc], acc_phase);
                        tc_fence_after();
                        epi_bar_sync();                                  // bias visible; buffer (it & 1) free (see below)
                        unsigned char* cur = c_s + (it & 1) * L::SUB_BYTES;
                        const unsigned char* prv = c_s + ((it & 1) ^ 1) * L::SUB_BYTES;
                        {
                            uint32_t v[32];
                            tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + grp * 32), v);
                            tmem_ld_wait();
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                const uint32_t chunk = ((uint32_t)(grp * 4 + g) ^ sw) << 4;
                                uint4 pk;
                                __nv_bfloat162* pp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float f0 = fmaxf(__uint_as_float(v[8 * g + 2 * e]) + s_bias[grp * 32 + 8 * g + 2 * e], 0.f);
                                    const float f1 = fmaxf(__uint_as_float(v[8 * g + 2 * e + 1]) + s_bias[grp * 32 + 8 * g + 2 * e + 1], 0.f);
                                    pp[e] = __floats2bfloat162_rn(f0, f1);
                                }
                                *reinterpret_cast<uint4*>(cur + row_off + chunk) = pk;
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);      // accumulator drained
                        epi_bar_sync();                                  // the whole conv tile is in `cur`
                        {
                            // staging row index: conv row 2tr -> rows 0..63 of cur, 2tr+1 -> rows 64..127 of cur,
                            // 2tr-1 -> rows 64..127 of prv (tr > 0), else clamp to conv row 0
                            uint4 m = make_uint4(0u, 0u, 0u, 0u);         // post-ReLU values are >= 0
                            __nv_bfloat162* pm = reinterpret_cast<__nv_bfloat162*>(&m);
#pragma unroll
                            for (int ry = 0; ry < 3; ++ry) {
                                const unsigned char* rowbuf = (ry == 0) ? (tr > 0 ? prv + 64 * 128 : cur)
                                                                        : cur + (ry - 1) * 64 * 128;
#pragma unroll
                                for (int kx = 0; kx < 3; ++kx) {
                                    const int ow = max(2 * pw - 1 + kx, 0);
                                    const uint4 vv = *reinterpret_cast<const uint4*>(
                                        rowbuf + ow * 128 + (((uint32_t)pch ^ (uint32_t)(ow & 7)) << 4));
                                    const __nv_bfloat162* pv = reinterpret_cast<const __nv_bfloat162*>(&vv);
#pragma unroll
                                    for (int e = 0; e < 4; ++e) pm[e] = __hmax2(pm[e], pv[e]);
                                }
                            }
                            *reinterpret_cast<uint4*>(pool + (((size_t)b * 64 + tr) * 32 + pw) * 64 + pch * 8) = m;
                        }
                        // buffer `prv` is overwritten by tile it+1 only after the barrier at the top of that tile, which
                        // every thread reaches after the reads above
                        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                    }
                    t_begin = t_end;                                     // nothing left for the generic epilogue below
                }
            }
            // VAR_RRING: residual sub-tile s (s counts this CTA's 64-column sub-tiles: tile it, sub-tile j -> it*NSUB + j)
            // lives in ring slot s % RS and is fetched RS - 1 sub-tiles ahead
            constexpr int RS = L::RSLOTS;
            auto load_residual_sub = [&](int s) {
                const int tile = (int)blockIdx.x + (s / NSUB) * (int)gridDim.x;
                if (tile >= num_tiles) return;
                int mb, nb;
                tile_coords(tile, mb, nb);
                const int slot = s % RS;
                mbar_arrive_expect_tx(&res_bar[slot], L::SUB_BYTES);
                tma_load_2d(r_s + slot * L::SUB_BYTES, &epi.mapR, &res_bar[slot], nb * BN + (s % NSUB) * 64, mb * BM);
            };
            if constexpr (EPI2) {
                // ---- one barrier per sub-tile, pipelined TMEM loads (see SmemLayout).  g = it * NSUB + j counts this
                // CTA's sub-tiles: sub-tile g is staged in buffer g % 2, its residual (RRING) sits in ring slot g % RS.
                if constexpr (L::RRING) {
                    if (leader && has_res) {
                        for (int s0 = 0; s0 < RS; ++s0) load_residual_sub(s0);      // the whole ring
                    }
                } else {
                    if (leader && has_res && (int)blockIdx.x < num_tiles) load_residual(blockIdx.x, 0);
                }
                if (t_begin < t_end) {
                    int mb0, nb0;
                    tile_coords(t_begin, mb0, nb0);
                    if (epi_tid < BN) s_bias[epi_tid] = epi.bias[nb0 * BN + epi_tid];
                }
                epi_bar_sync();                                           // bias row of the first tile visible
                int it2 = 0;
                for (int t = t_begin; t < t_end; t += t_step, ++it2) {
                    int m_blk, n_blk;
                    tile_coords(t, m_blk, n_blk);
                    const int rb = it2 & 1;
                    if constexpr (!L::RRING) {
                        // the other residual buffer was last read by tile it2-1, whose last barrier every thread has passed
                        if (leader && has_res && t + (int)gridDim.x < num_tiles) load_residual(t + gridDim.x, rb ^ 1);
                    }
                    if (t + t_step < t_end) {
                        // bias row of the NEXT tile into the other half (last read by tile it2-1); it becomes visible
                        // through the barriers of this tile
                        int mb2, nb2;
                        tile_coords(t + t_step, mb2, nb2);
                        if (epi_tid < BN) s_bias[((it2 + 1) & 1) * BN + epi_tid] = epi.bias[nb2 * BN + epi_tid];
                    }
                    const float* sb = s_bias + (it2 & 1) * BN;
                    mbar_wait(&tfull_bar[acc], acc_phase);
                    tc_fence_after();
                    if constexpr (!L::RRING) {
                        if (has_res) mbar_wait(&res_bar[rb], (uint32_t)((it2 >> 1) & 1));
                    }
                    const unsigned char* rbuf = r_s + rb * L::R_BYTES;
                    const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + grp * 32);
                    uint32_t v[2][32];
                    tmem_ld_32x32(t_addr, v[0]);
#pragma unroll
                    for (int j = 0; j < NSUB; ++j) {
                        tmem_ld_wait();                                   // columns of sub-tile j are in v[j & 1]
                        if (j + 1 < NSUB) tmem_ld_32x32(t_addr + (uint32_t)((j + 1) * 64), v[(j + 1) & 1]);
                        if (j == NSUB - 1) {                              // accumulator drained: MMAs may reuse it
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                        }
                        const int g = it2 * NSUB + j;
                        unsigned char* csub = c_s + (g & 1) * L::SUB_BYTES + row_off;
                        const unsigned char* rsub = rbuf + j * L::SUB_BYTES + row_off;
                        if constexpr (L::RRING) {
                            if (has_res) mbar_wait(&res_bar[g % RS], (uint32_t)((g / RS) & 1));
                            rsub = r_s + (g % RS) * L::SUB_BYTES + row_off;
                        }
                        const int c = 2 * j + grp;
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {                  // 8 columns = one 16-byte chunk
                            const uint32_t chunk = ((uint32_t)(grp * 4 + gq) ^ sw) << 4;
                            float f[8];
                            const float4 b0 = *reinterpret_cast<const float4*>(sb + c * 32 + 8 * gq);
                            const float4 b1 = *reinterpret_cast<const float4*>(sb + c * 32 + 8 * gq + 4);
                            f[0] = __uint_as_float(v[j & 1][8 * gq + 0]) + b0.x; f[1] = __uint_as_float(v[j & 1][8 * gq + 1]) + b0.y;
                            f[2] = __uint_as_float(v[j & 1][8 * gq + 2]) + b0.z; f[3] = __uint_as_float(v[j & 1][8 * gq + 3]) + b0.w;
                            f[4] = __uint_as_float(v[j & 1][8 * gq + 4]) + b1.x; f[5] = __uint_as_float(v[j & 1][8 * gq + 5]) + b1.y;
                            f[6] = __uint_as_float(v[j & 1][8 * gq + 6]) + b1.z; f[7] = __uint_as_float(v[j & 1][8 * gq + 7]) + b1.w;
                            if (has_res) {
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
                        fence_proxy_async();                              // smem writes -> visible to the TMA engine
                        if (leader) tma_store_wait_read<0>();             // store g-1 has left buffer (g+1) & 1
                        epi_bar_sync();                                   // sub-tile g complete; next buffer free
                        if (leader) {
                            tma_store_2d(&epi.mapC, c_s + (g & 1) * L::SUB_BYTES, n_blk * BN + j * 64, m_blk * BM);
                            tma_store_commit();
                            if constexpr (L::RRING) {
                                if (has_res) load_residual_sub(g + RS);   // slot g % RS was read before the barrier
                            }
                        }
                    }
                    if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
                if (leader) tma_store_wait_all();
                t_begin = t_end;                                          // nothing left for the loop below
            }
            if constexpr (L::CHAIN) {
                // ---- conv3 + next conv1.  Per tile: four output sub-tiles (bias, residual, ReLU -> bf16 staging -> TMA
                // store AND A operand of the second GEMM), then n2/64 sub-tiles of the second accumulator (bias', ReLU
                // -> staging -> TMA store through mapC2).  Staging step g = it * steps + j uses buffer g & 1; a buffer
                // is re-written once the TMA store issued from it two steps ago has been read out AND -- if that step
                // was an output sub-tile -- the MMAs reading it have retired (ysub_free).
                const int n2 = epi.n2, steps = 4 + n2 / 64;
                __shared__ float s_bias2[128];
                if (epi_tid < BN) s_bias[epi_tid] = epi.bias[epi_tid];
                if (epi_tid < n2) s_bias2[epi_tid] = epi.bias2[epi_tid];
                // (a deeper residual ring -- 5 slots where the weights leave room -- was measured in round 2: no change,
                // profiles/r02k_ab_ring*.json; the ring stays at the three slots of VAR_RRING)
                if (leader && has_res) {
                    for (int s0 = 0; s0 < RS - 1; ++s0) load_residual_sub(s0);
                }
                uint32_t fph = 0, pend = 0;                               // leader only: phase / pending bits per buffer
                epi_bar_sync();                                           // bias rows visible
                int itc = 0;
                for (int t = t_begin; t < t_end; t += t_step, ++itc) {
                    int m_blk, n_blk;
                    tile_coords(t, m_blk, n_blk);
                    mbar_wait(&tfull_bar[0], (uint32_t)(itc & 1));
                    tc_fence_after();
#pragma unroll 1
                    for (int j = 0; j < steps; ++j) {
                        const int b = (itc * steps + j) & 1;
                        if (leader) {
                            tma_store_wait_read<1>();                     // the store issued from buffer b has been read
                            if (pend & (1u << b)) {                       // ... and the MMAs that read it have retired
                                mbar_wait(&ysub_free[b], (fph >> b) & 1u);
                                fph ^= (1u << b);
                                pend &= ~(1u << b);
                            }
                        }
                        if (j == 4) {                                     // (first sub-tile of the second accumulator)
                            mbar_wait(d2_full, (uint32_t)(itc & 1));
                            tc_fence_after();
                        }
                        epi_bar_sync();
                        unsigned char* csub = c_s + b * L::SUB_BYTES + row_off;
                        const bool first = j < 4;
                        const unsigned char* rsub = nullptr;
                        if (first) {
                            const int s = itc * 4 + j;
                            // every thread is past the barrier above, i.e. done with sub-tile s-1: its slot is free
                            if (leader && has_res) load_residual_sub(s + RS - 1);
                            if (has_res) {
                                mbar_wait(&res_bar[s % RS], (uint32_t)((s / RS) & 1));
                                rsub = r_s + (s % RS) * L::SUB_BYTES + row_off;
                            }
                        }
                        const int c = 2 * (first ? j : j - 4) + grp;      // 32-column chunk of the accumulator
                        const float* sb = first ? s_bias : s_bias2;
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((first ? 0 : CHAIN_D2_COL) + c * 32), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int gq = 0; gq < 4; ++gq) {                  // 8 columns = one 16-byte chunk
                            const uint32_t chunk = ((uint32_t)(grp * 4 + gq) ^ sw) << 4;
                            float f[8];
                            const float4 b0 = *reinterpret_cast<const float4*>(sb + c * 32 + 8 * gq);
                            const float4 b1 = *reinterpret_cast<const float4*>(sb + c * 32 + 8 * gq + 4);
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
                            if (!first || epi.relu) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] = fmaxf(f[e], 0.f);
                            }
                            uint4 pk;
                            __nv_bfloat162* pp = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
                            for (int e = 0; e < 4; ++e) pp[e] = __floats2bfloat162_rn(f[2 * e], f[2 * e + 1]);
                            *reinterpret_cast<uint4*>(csub + chunk) = pk;
                        }
                        // TMEM reads ordered before what follows the barriers below: j == 3 hands D1 back to the MMA warp,
                        // the last step of a tile precedes the next tile's first write to D2
                        tc_fence_before();
                        if (j == 3) {
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&tempty_bar[0]);
                        }
                        fence_proxy_async();                              // smem writes -> visible to TMA and the tensor core
                        epi_bar_sync();
                        if (leader) {
                            if (first) {
                                tma_store_2d(&epi.mapC, c_s + b * L::SUB_BYTES, j * 64, m_blk * BM);
                                tma_store_commit();
                                mbar_arrive(&ysub_full[b]);               // K block j of the second GEMM is staged
                                pend |= (1u << b);
                            } else {
                                tma_store_2d(&epi.mapC2, c_s + b * L::SUB_BYTES, (j - 4) * 64, m_blk * BM);
                                tma_store_commit();
                            }
                        }
                    }
                }
                if (leader) tma_store_wait_all();
                t_begin = t_end;                                          // nothing left for the loop below
            } else if constexpr (L::RRING) {
                if (leader && has_res && !EPI2) {
                    for (int s0 = 0; s0 < RS - 1; ++s0) load_residual_sub(s0);
                }
            } else {
                if (leader && has_res && !EPI2 && (int)blockIdx.x < num_tiles) load_residual(blockIdx.x, 0);
            }
            int it = 0;
            for (int t = t_begin; t < t_end; t += t_step, ++it) {
                int m_blk, n_blk;
                tile_coords(t, m_blk, n_blk);
                const int rb = it & 1;
                if constexpr (!L::RRING) {
                    // the other residual buffer was last read by tile it-1, which every thread has left: prefetch tile it+1
                    if (leader && has_res && t + (int)gridDim.x < num_tiles) load_residual(t + gridDim.x, rb ^ 1);
                }
                if (epi_tid < BN) s_bias[epi_tid] = epi.bias[n_blk * BN + epi_tid];   // visible after the next barrier
                mbar_wait(&tfull_bar[acc], acc_phase);
                tc_fence_after();
                if constexpr (!L::RRING) {
                    if (has_res) mbar_wait(&res_bar[rb], (uint32_t)((it >> 1) & 1));
                }
                const unsigned char* rbuf = r_s + rb * L::R_BYTES;
#pragma unroll 1
                for (int j = 0; j < NSUB; ++j) {
                    // sub-buffer j is free once the store issued for it one tile ago has been read out
                    if (leader) tma_store_wait_read<NBUF - 1>();
                    epi_bar_sync();
                    unsigned char* csub = c_s + (j % NBUF) * L::SUB_BYTES + row_off;
                    const unsigned char* rsub = rbuf + j * L::SUB_BYTES + row_off;
                    if constexpr (L::RRING) {
                        const int s = it * NSUB + j;
                        // every epilogue thread is past the barrier above, i.e. done with sub-tile s-1: its ring slot
                        // (s + RS - 1) % RS == (s - 1) % RS is free for the sub-tile RS - 1 ahead
                        if (leader && has_res) load_residual_sub(s + RS - 1);
                        if (has_res) mbar_wait(&res_bar[s % RS], (uint32_t)((s / RS) & 1));
                        rsub = r_s + (s % RS) * L::SUB_BYTES + row_off;
                    }
                    {
                        const int h = grp;
                        const int c = 2 * j + h;
                        uint32_t v[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int g = 0; g < 4; ++g) {                 // 8 columns = one 16-byte chunk
                            const uint32_t chunk = ((uint32_t)(h * 4 + g) ^ sw) << 4;
                            float f[8];
                            const float4 b0 = *reinterpret_cast<const float4*>(s_bias + c * 32 + 8 * g);
                            const float4 b1 = *reinterpret_cast<const float4*>(s_bias + c * 32 + 8 * g + 4);
                            f[0] = __uint_as_float(v[8 * g + 0]) + b0.x; f[1] = __uint_as_float(v[8 * g + 1]) + b0.y;
                            f[2] = __uint_as_float(v[8 * g + 2]) + b0.z; f[3] = __uint_as_float(v[8 * g + 3]) + b0.w;
                            f[4] = __uint_as_float(v[8 * g + 4]) + b1.x; f[5] = __uint_as_float(v[8 * g + 5]) + b1.y;
                            f[6] = __uint_as_float(v[8 * g + 6]) + b1.z; f[7] = __uint_as_float(v[8 * g + 7]) + b1.w;
                            if (has_res) {
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
                    }
                    if (j == NSUB - 1) {                               // accumulator drained: MMAs may reuse it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    fence_proxy_async();                               // smem writes -> visible to the TMA engine
                    epi_bar_sync();
                    if (leader) {
                        tma_store_2d(&epi.mapC, c_s + (j % NBUF) * L::SUB_BYTES, n_blk * BN + j * 64, m_blk * BM);
                        tma_store_commit();
                    }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (leader) tma_store_wait_all();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// Launch with a fully prepared A operand.
template <int BN, class Epi, bool STAGED = false, bool KHS = false, int VAR = VAR_NONE, bool EPI2 = false>
int launch_gemm_op(const AOperand& A, int m, const void* b, int n, int k, const Epi& epi, cudaStream_t st,
                   int b_cols = 0 /* > 0: B is [n, b_cols] with b_cols > k (split-K: A.ksplit_mblks) */) {
    using L = SmemLayout<BN, STAGED, KHS, VAR, EPI2>;
    if ((VAR == VAR_BRES || VAR == VAR_BRESP) && (A.mode != 3 || n != BN || k != BRES_K))
        return ssg_set_error(SSG_ERR_INVALID, "gemm: the resident-B variant is the stem kernel (N=%d, K=%d)", n, k);
    if constexpr (VAR == VAR_CHAIN) {
        if (n != BN || k % BK || k / BK < 1 || (epi.n2 != 64 && epi.n2 != 128) ||
            (k / BK) * SmemLayout<BN, STAGED, KHS, VAR, EPI2>::B_TILE + 4 * epi.n2 * 128 > CHAIN_WRES_BYTES)
            return ssg_set_error(SSG_ERR_INVALID, "gemm: the chained kernel needs N = 256, K = 64 / 128 and resident weights "
                                 "within %d bytes (N=%d, K=%d, n2=%d)", CHAIN_WRES_BYTES, n, k, epi.n2);
    }
    if (VAR == VAR_KHSB && (n != BN || k != 9 * BK || A.cblks != 1))
        return ssg_set_error(SSG_ERR_INVALID, "gemm: the resident-weight KHS variant needs C = N = 64 (N=%d, K=%d)", n, k);
    // K need not be a multiple of BK: the last K block reads past the end and TMA zero-fills it (both operands)
    if (k % 8) return ssg_set_error(SSG_ERR_INVALID, "gemm: K=%d must be a multiple of 8 (16-byte row pitch)", k);
    CUtensorMap mapB;
    if (A.mode == 3) SSG_TRY(make_tmap_2d_bf16_sw64(&mapB, b, (uint64_t)n, (uint64_t)k, BN));
    else SSG_TRY(make_tmap_2d_bf16(&mapB, b, (uint64_t)n, (uint64_t)(b_cols > 0 ? b_cols : k), (uint64_t)(b_cols > 0 ? b_cols : k), BN));
    int sms = 0;
    SSG_TRY(tc_num_sms(&sms));
    const int tiles = ((m + BM - 1) / BM) * ((n + BN - 1) / BN);
    int grid = tiles < sms ? tiles : sms;
    if constexpr (VAR == VAR_BRES || VAR == VAR_BRESP) {
        if (epi.pool_out != nullptr) {                      // whole images per CTA
            if (m % (64 * BM)) return ssg_set_error(SSG_ERR_INVALID, "gemm: fused stem pool needs whole images (M=%d)", m);
            const int images = m / (64 * BM);
            grid = images < sms ? images : sms;
        }
    }
    auto kern = gemm_kernel<BN, Epi, STAGED, KHS, VAR, EPI2>;
    SSG_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL));
    // KHS: one K block per (kernel column, channel block), i.e. a third of the plain K blocks
    // VAR_BRESP: the whole K range of a tile rides in one stage
    const int nkb = L::PLANES ? 1 : KHS ? (k / BK) / 3 : (k + BK - 1) / BK;
#ifdef SSG_NO_LAUNCH_EX                 // CPU emulation (tests/cpu_cuda/stub_tc): plain launch
    kern<<<grid, NUM_THREADS, L::TOTAL, st>>>(A, mapB, m, n, nkb, epi);
#else
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = L::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    SSG_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, A, mapB, m, n, nkb, epi));
#endif
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// Plain C = A * B^T with A [m,k], B [n,k] bf16 row-major.
template <int BN, class Epi>
int launch_gemm(const void* a, int m, const void* b, int n, int k, const Epi& epi, cudaStream_t st) {
    AOperand A;
    memset(&A, 0, sizeof(A));
    A.mode = 0;
    A.cblks = k / BK;
    A.taps = 1;
    A.tiles_per_img = 1;
    A.hmul = 1;
    SSG_TRY(make_tmap_2d_bf16(&A.map[0], a, (uint64_t)m, (uint64_t)k, (uint64_t)k, BM));
    return launch_gemm_op<BN, Epi, false>(A, m, b, n, k, epi, st);
}

}  // namespace tc
}  // namespace ssg
