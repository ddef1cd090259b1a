// ResNet-50 trunk building blocks (reid/models/resnet.py:86-134 with torchvision's Bottleneck graph):
// every convolution runs on the tcgen05 GEMM of gemm_tc.cuh — 1x1 as a plain GEMM over NHWC pixels, 3x3 as an
// implicit GEMM whose A tiles are shifted TMA boxes (halo zero-filled by TMA), stride-2 through parity planes —
// with eval-mode BatchNorm folded into the bf16 weights and bias/residual/ReLU fused in the epilogue.
// The rest are small HBM-bound helpers: weight folding, stem im2col, max-pool, parity split, pooled tail.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "gemm_tc.cuh"
#include "gemm_tc2.cuh"
#include "conv.h"

namespace ssg {

// ---------------------------------------------------------------------------------------------------
// epilogue: out[row, col] = bf16( relu?( acc + bias[col] (+ residual[row, col]) ) ), NHWC ([M, Cout] row-major),
// staged through shared memory and written / prefetched by TMA (tc::StagedEpi in gemm_tc.cuh).
// ---------------------------------------------------------------------------------------------------
// stem kernel variant: 0 = plain (weights re-loaded per tile; SSG_STEM_BRES=0), 1 = resident weights, one TMA box per
// kernel row (SSG_STEM_PLANES=0), 2 = resident weights + whole-tile parity-plane operand staging (default)
static int stem_variant() {
    static int v = -1;
    if (v < 0) {
        const char* b = getenv("SSG_STEM_BRES");
        const char* p = getenv("SSG_STEM_PLANES");
        v = (b && !atoi(b)) ? 0 : ((p && !atoi(p)) ? 1 : 2);
    }
    return v;
}

static int gemm_dispatch(const tc::AOperand& A, int m, const void* w, int cout, int k, const float* bias,
                         const void* residual, int relu, void* y, cudaStream_t st, void* pool_out = nullptr) {
    if (cout % 64) return ssg_set_error(SSG_ERR_INVALID, "conv: Cout=%d must be a multiple of 64", cout);
    tc::StagedEpi epi;
    memset(&epi, 0, sizeof(epi));
    epi.bias = bias;
    epi.relu = relu;
    epi.has_res = residual != nullptr;
    epi.pool_out = pool_out;
    SSG_TRY(make_tmap_2d_bf16(&epi.mapC, y, (uint64_t)m, (uint64_t)cout, (uint64_t)cout, tc::BM));
    SSG_TRY(make_tmap_2d_bf16(&epi.mapR, residual ? residual : y, (uint64_t)m, (uint64_t)cout, (uint64_t)cout, tc::BM));
    // SSG_CONV_EPI2=1 (opt-in): the one-barrier, pipelined-load epilogue of gemm_tc.cuh for the generic kernels
    static int epi2 = -1;
    if (epi2 < 0) { const char* e = getenv("SSG_CONV_EPI2"); epi2 = e ? atoi(e) : 0; }
    const bool stem_kernel = stem_variant() && A.mode == 3 && cout == 64 && k == tc::BRES_K;
    // (the 128x256 tiles keep the default epilogue: with the doubled bias rows the EPI2 layout is 256 bytes over the
    // 227 KB limit -- first B200 run of round 2 -- and those launches are L2-bandwidth bound, not epilogue bound)
    // (SSG_WIDE_K: smallest K that takes the 256-wide tiles; 256 by default, A/B knob for the K = 128 conv3 of layer 2)
    static int wide_k = -1;
    if (wide_k < 0) { const char* e = getenv("SSG_WIDE_K"); wide_k = e ? atoi(e) : 256; }
    const bool wide = cout % 256 == 0 && k >= wide_k;
    if (epi2 && !stem_kernel && !pool_out && !wide) {
        if (cout % 128 == 0) return tc::launch_gemm_op<128, tc::StagedEpi, true, false, tc::VAR_NONE, true>(A, m, w, cout, k, epi, st);
        return tc::launch_gemm_op<64, tc::StagedEpi, true, false, tc::VAR_NONE, true>(A, m, w, cout, k, epi, st);
    }
    // two-CTA 256 x BN tiles (cta_group::2, gemm_tc2.cuh) for the plain / implicit GEMMs; SSG_CONV_PAIR=0 disables
    // bit 0: the 256-wide tiles; bit 1: the 128-wide 3x3 convolutions; bit 2: the 128-wide 1x1 convolutions.
    // Default 3 (round 2, B200, profiles/r02f_ab_pair*.json: embedding 601.9 -> 578.8 ms per cycle; the 128-wide 1x1
    // convolutions are HBM bound and lose 1 % as pairs).  Bit-identical to the single-CTA kernels (tests/c/embed_dump).
    static int pair = -1;
    if (pair < 0) { const char* e = getenv("SSG_CONV_PAIR"); pair = e ? atoi(e) : 3; }
#ifdef SSG_PAIR_KERNEL_UNAVAILABLE      // CPU emulation (tests/cpu_cuda): no thread-block clusters there
    pair = 0;
#endif
    if (pair && !stem_kernel && !pool_out && m >= 2 * tc::BM && A.mode != 2 && A.mode != 3) {
        if (wide && (pair & 1)) {
            if (residual) return tc::launch_gemm2_op<256, true>(A, m, w, cout, k, epi, st);
            return tc::launch_gemm2_op<256, false>(A, m, w, cout, k, epi, st);
        }
        if (!wide && cout % 128 == 0 && (pair & (A.taps == 9 ? 2 : 4))) {
            if (residual) return tc::launch_gemm2_op<128, true>(A, m, w, cout, k, epi, st);
            return tc::launch_gemm2_op<128, false>(A, m, w, cout, k, epi, st);
        }
    }
    // 128x256 tiles for the K-heavy convolutions without a residual (SSG_CONV_BN256=0 disables, for A/B runs)
    static int bn256 = -1;
    if (bn256 < 0) { const char* e = getenv("SSG_CONV_BN256"); bn256 = e ? atoi(e) : 1; }
    if (bn256 && !residual && wide)
        return tc::launch_gemm_op<256, tc::StagedEpi, true>(A, m, w, cout, k, epi, st);
    // ... and WITH a residual (the conv3 of layers 3 and 4): 128x256 tiles with a residual sub-tile ring
    // (SSG_CONV_BN256_RES=0 falls back to 128x128 tiles with a whole-tile residual double buffer)
    static int bn256_res = -1;
    if (bn256_res < 0) { const char* e = getenv("SSG_CONV_BN256_RES"); bn256_res = e ? atoi(e) : 1; }
    if (bn256_res && residual && wide)
        return tc::launch_gemm_op<256, tc::StagedEpi, true, false, tc::VAR_RRING>(A, m, w, cout, k, epi, st);
    // the stem (mode 3: N = 64, K = 256) keeps its weights resident in shared memory (SSG_STEM_BRES=0 disables)
    if (stem_variant() && A.mode == 3 && cout == 64 && k == tc::BRES_K) {
        if (stem_variant() == 2)
            return tc::launch_gemm_op<64, tc::StagedEpi, true, false, tc::VAR_BRESP>(A, m, w, cout, k, epi, st);
        return tc::launch_gemm_op<64, tc::StagedEpi, true, false, tc::VAR_BRES>(A, m, w, cout, k, epi, st);
    }
    if (pool_out) return ssg_set_error(SSG_ERR_UNSUPPORTED, "the fused stem max-pool needs the resident-weight stem kernel");
    // without a residual the residual buffers of the layout become operand stages (SSG_CONV_NORES=0 disables)
    static int nores = -1;
    if (nores < 0) { const char* e = getenv("SSG_CONV_NORES"); nores = e ? atoi(e) : 1; }
    if (nores && !residual) {
        if (cout % 128 == 0) return tc::launch_gemm_op<128, tc::StagedEpi, true, false, tc::VAR_NORES>(A, m, w, cout, k, epi, st);
        return tc::launch_gemm_op<64, tc::StagedEpi, true, false, tc::VAR_NORES>(A, m, w, cout, k, epi, st);
    }
    if (cout % 128 == 0) return tc::launch_gemm_op<128, tc::StagedEpi, true>(A, m, w, cout, k, epi, st);
    return tc::launch_gemm_op<64, tc::StagedEpi, true>(A, m, w, cout, k, epi, st);
}

// 1x1 convolution (or any [M,K] x [Cout,K]^T product) on NHWC pixels.
int conv1x1(const void* x, int m, int cin, const void* w, const float* bias, int cout, const void* residual,
            int relu, void* y, cudaStream_t st) {
    tc::AOperand A;
    memset(&A, 0, sizeof(A));
    A.mode = 0;
    A.cblks = (cin + tc::BK - 1) / tc::BK;
    A.taps = 1;
    A.tiles_per_img = 1;
    A.hmul = 1;
    SSG_TRY(make_tmap_2d_bf16(&A.map[0], x, (uint64_t)m, (uint64_t)cin, (uint64_t)cin, tc::BM));
    return gemm_dispatch(A, m, w, cout, cin, bias, residual, relu, y, st);
}

static int tile_geometry(int H, int W, int* bw, int* bh, int* bb, int* tiles_per_img);
static inline int tile_geometry_fwd(int H, int W, int* bw, int* bh, int* bb, int* t) { return tile_geometry(H, W, bw, bh, bb, t); }

// Bottleneck tail of a first block, fused: y = relu( conv3(t2) + downsample(x) ) as ONE GEMM whose K axis is the
// concatenation [mid | cin]: K blocks < mid/64 read t2 [M, mid], the rest read x — as a plain [M, cin] matrix
// (stride 1) or through stride-2 single-tap TMA boxes of x [B,2H,2W,cin] (stride 2).  w = [cout, mid + cin] (conv3 and
// downsample weights side by side), bias = bias3 + bias_ds.  The downsample output never touches memory.
int conv_fused_ds(const void* t2, const void* x, int B, int H, int W, int mid, int cin, int stride, const void* w,
                  const float* bias, int cout, void* y, cudaStream_t st) {
    if (mid % 64 || cin % 64) return ssg_set_error(SSG_ERR_INVALID, "conv_fused_ds: channels must be multiples of 64");
    const int m = B * H * W;
    tc::AOperand A;
    memset(&A, 0, sizeof(A));
    A.mode = 0;
    A.cblks = mid / tc::BK;
    A.taps = 1;
    A.tiles_per_img = 1;
    A.hmul = 1;
    A.kb_split = mid / tc::BK;
    SSG_TRY(make_tmap_2d_bf16(&A.map[0], t2, (uint64_t)m, (uint64_t)mid, (uint64_t)mid, tc::BM));
    if (stride == 1) {
        A.mode1 = 0;
        SSG_TRY(make_tmap_2d_bf16(&A.map[1], x, (uint64_t)m, (uint64_t)cin, (uint64_t)cin, tc::BM));
    } else {
        A.mode1 = 1;
        A.hmul = 2;
        int bw;
        SSG_TRY(tile_geometry_fwd(H, W, &bw, &A.bh, &A.bb, &A.tiles_per_img));
        SSG_TRY(make_tmap_nhwc_bf16(&A.map[1], x, B, 2 * H, 2 * W, cin, bw, A.bh, A.bb, 2));
    }
    return gemm_dispatch(A, m, w, cout, mid + cin, bias, nullptr, 1, y, st);
}

// Layer-1 bottleneck tail AND the next block's conv1 in one launch (tc::VAR_CHAIN in gemm_tc.cuh):
//   y  = relu( conv3(t2) + shortcut ),  shortcut = residual [M, 256]  or  downsample(x_ds) K-concatenated (w = [W3 | Wds])
//   t1 = relu( conv1_next(y) )          [M, n2], n2 = 64 (next block of layer 1) or 128 (first block of layer 2)
// t2 [M = B*H*W, mid], x_ds [M, cin_ds] (or NULL), w [256, mid (+ cin_ds)], w_next [n2, 256]; stride 1 only.
int conv_chain(const void* t2, int B, int H, int W, int mid, const void* x_ds, int cin_ds, const void* w,
               const float* bias, const void* residual, void* y, const void* w_next, const float* bias_next, int n2,
               void* t1_next, cudaStream_t st) {
    const int cout = 256, m = B * H * W;
    if (mid % 64 || (x_ds && cin_ds % 64) || (x_ds && residual))
        return ssg_set_error(SSG_ERR_INVALID, "conv_chain: bad arguments");
    tc::AOperand A;
    memset(&A, 0, sizeof(A));
    A.mode = 0;
    A.cblks = mid / tc::BK;
    A.taps = 1;
    A.tiles_per_img = 1;
    A.hmul = 1;
    SSG_TRY(make_tmap_2d_bf16(&A.map[0], t2, (uint64_t)m, (uint64_t)mid, (uint64_t)mid, tc::BM));
    int k = mid;
    if (x_ds) {
        A.kb_split = mid / tc::BK;
        A.mode1 = 0;
        SSG_TRY(make_tmap_2d_bf16(&A.map[1], x_ds, (uint64_t)m, (uint64_t)cin_ds, (uint64_t)cin_ds, tc::BM));
        k += cin_ds;
    }
    tc::StagedEpi epi;
    memset(&epi, 0, sizeof(epi));
    epi.bias = bias;
    epi.relu = 1;
    epi.has_res = residual != nullptr;
    epi.bias2 = bias_next;
    epi.n2 = n2;
    SSG_TRY(make_tmap_2d_bf16(&epi.mapC, y, (uint64_t)m, (uint64_t)cout, (uint64_t)cout, tc::BM));
    SSG_TRY(make_tmap_2d_bf16(&epi.mapR, residual ? residual : y, (uint64_t)m, (uint64_t)cout, (uint64_t)cout, tc::BM));
    SSG_TRY(make_tmap_2d_bf16(&epi.mapW2, w_next, (uint64_t)n2, (uint64_t)cout, (uint64_t)cout, (uint32_t)n2));
    SSG_TRY(make_tmap_2d_bf16(&epi.mapC2, t1_next, (uint64_t)m, (uint64_t)n2, (uint64_t)n2, tc::BM));
    return tc::launch_gemm_op<256, tc::StagedEpi, true, false, tc::VAR_CHAIN>(A, m, w, cout, k, epi, st);
}

// the chained kernel is used for the blocks of layer 1 unless SSG_CONV_CHAIN=0
bool conv_chain_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SSG_CONV_CHAIN"); v = e ? atoi(e) : 1; }
    return v != 0;
}

// tile geometry of a 128-pixel M tile on an [H, W] output map (W divides 128, H*W multiple or divisor of 128)
static int tile_geometry(int H, int W, int* bw, int* bh, int* bb, int* tiles_per_img) {
    if (W <= 0 || 128 % W) return ssg_set_error(SSG_ERR_INVALID, "conv3x3: width %d does not divide 128", W);
    *bw = W;
    if (H * W >= 128) {
        if ((H * W) % 128) return ssg_set_error(SSG_ERR_INVALID, "conv3x3: map %dx%d not tileable", H, W);
        *bh = 128 / W; *bb = 1; *tiles_per_img = H * W / 128;
    } else {
        if (128 % (H * W)) return ssg_set_error(SSG_ERR_INVALID, "conv3x3: map %dx%d not tileable", H, W);
        *bh = H; *bb = 128 / (H * W); *tiles_per_img = 1;
    }
    return SSG_OK;
}

// the stem's max-pool runs in the stem kernel's epilogue unless SSG_STEM_POOL=0 (or the resident-weight stem is off)
bool stem_pool_fused() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SSG_STEM_POOL");
        v = ((e && !atoi(e)) || stem_variant() == 0) ? 0 : 1;
    }
    return v != 0;
}

// stride-2 convolutions read the un-split input through element-strided TMA boxes unless SSG_S2_PLANES=1
bool s2_strided_tma() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SSG_S2_PLANES"); v = (e && atoi(e)) ? 0 : 1; }
    return v != 0;
}

// 1x1 stride-2 convolution straight from x [B,2H,2W,C] (H, W = output size): implicit mode with a single tap.
int conv1x1_s2(const void* x, int B, int H, int W, int cin, const void* w, const float* bias, int cout, int relu,
               void* y, cudaStream_t st) {
    if (cin % 64) return ssg_set_error(SSG_ERR_INVALID, "conv1x1_s2: Cin=%d must be a multiple of 64", cin);
    tc::AOperand A;
    memset(&A, 0, sizeof(A));
    A.mode = 1;
    A.cblks = cin / tc::BK;
    A.taps = 1;
    A.hmul = 2;
    int bw;
    SSG_TRY(tile_geometry(H, W, &bw, &A.bh, &A.bb, &A.tiles_per_img));
    SSG_TRY(make_tmap_nhwc_bf16(&A.map[0], x, B, 2 * H, 2 * W, cin, bw, A.bh, A.bb, 2));
    return gemm_dispatch(A, B * H * W, w, cout, cin, bias, nullptr, relu, y, st);
}

// 3x3 convolution, padding 1; H, W are the OUTPUT map size.  stride 1: x is [B,H,W,C] NHWC.
// stride 2: x is the input [B,2H,2W,C] read through stride-2 TMA boxes (default), or — SSG_S2_PLANES=1 — the four
// parity planes [4][B,H,W,C] produced by parity_split (plane = (h&1)*2 + (w&1)).
int conv3x3(const void* x, int B, int H, int W, int cin, int stride, const void* w, const float* bias, int cout,
            int relu, void* y, cudaStream_t st) {
    if (cin % 64) return ssg_set_error(SSG_ERR_INVALID, "conv3x3: Cin=%d must be a multiple of 64", cin);
    tc::AOperand A;
    memset(&A, 0, sizeof(A));
    A.mode = 1;
    A.cblks = cin / tc::BK;
    A.taps = 9;
    int bw;
    SSG_TRY(tile_geometry(H, W, &bw, &A.bh, &A.bb, &A.tiles_per_img));
    const size_t plane_elems = (size_t)B * H * W * cin;
    const bool strided = stride == 2 && s2_strided_tma();
    A.hmul = strided ? 2 : 1;
    if (strided) {
        SSG_TRY(make_tmap_nhwc_bf16(&A.map[0], x, B, 2 * H, 2 * W, cin, bw, A.bh, A.bb, 2));
    } else {
        const int nplanes = stride == 2 ? 4 : 1;
        for (int pl = 0; pl < nplanes; ++pl)
            SSG_TRY(make_tmap_nhwc_bf16(&A.map[pl], (const __nv_bfloat16*)x + pl * plane_elems, B, H, W, cin, bw,
                                        A.bh, A.bb));
    }
    for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
            const int t = kh * 3 + kw;
            if (stride == 1 || strided) {
                A.tap_plane[t] = 0; A.tap_dh[t] = (signed char)(kh - 1); A.tap_dw[t] = (signed char)(kw - 1);
            } else {
                // input row 2h+kh-1: kh=0 -> odd plane, h-1; kh=1 -> even plane, h; kh=2 -> odd plane, h
                const int ph = kh == 1 ? 0 : 1, pw = kw == 1 ? 0 : 1;
                A.tap_plane[t] = (signed char)(ph * 2 + pw);
                A.tap_dh[t] = (signed char)(kh == 0 ? -1 : 0);
                A.tap_dw[t] = (signed char)(kw == 0 ? -1 : 0);
            }
        }
    const int m = B * H * W;
    // few-channel stride-1 convolutions (layer 1: C = 64) are L2-bandwidth bound when every tap re-reads the map:
    // share the three kernel rows of a kernel column through one haloed TMA box (SSG_CONV_KHS=0 disables)
    static int khs = -1;
    if (khs < 0) { const char* e = getenv("SSG_CONV_KHS"); khs = e ? atoi(e) : 1; }
    if (khs && stride == 1 && cin == 64 && cout == 64 && A.bb == 1 && (A.bh + 2) * bw == 192 && (bw * 128) % 1024 == 0) {
        // (the haloed box must fill the 192-row stage buffer exactly: the TMA transaction count is fixed at compile time)
        A.khs_row_bytes = bw * 128;
        SSG_TRY(make_tmap_nhwc_bf16(&A.map[0], x, B, H, W, cin, bw, A.bh + 2, 1));
        tc::StagedEpi epi;
        memset(&epi, 0, sizeof(epi));
        epi.bias = bias;
        epi.relu = relu;
        epi.has_res = 0;
        SSG_TRY(make_tmap_2d_bf16(&epi.mapC, y, (uint64_t)m, (uint64_t)cout, (uint64_t)cout, tc::BM));
        SSG_TRY(make_tmap_2d_bf16(&epi.mapR, y, (uint64_t)m, (uint64_t)cout, (uint64_t)cout, tc::BM));
        // SSG_KHS_PAIR=1 (opt-in until measured): the two-CTA form (gemm_tc2.cuh): per MMA and CTA 4 KB of A + 1 KB of B
        // instead of 4 + 2 (the single-CTA kernel is bound by the UMMA's shared-memory operand bandwidth)
        static int khs_pair = -1;
        if (khs_pair < 0) { const char* e = getenv("SSG_KHS_PAIR"); khs_pair = e ? atoi(e) : 0; }
#ifdef SSG_PAIR_KERNEL_UNAVAILABLE
        khs_pair = 0;
#endif
        if (khs_pair && m >= 2 * tc::BM) return tc::launch_gemm2_op<64, false, true>(A, m, w, cout, 9 * cin, epi, st);
        // weights resident in shared memory (72 KB, loaded once per CTA) unless SSG_KHS_BRES=0
        static int khs_bres = -1;
        if (khs_bres < 0) { const char* e = getenv("SSG_KHS_BRES"); khs_bres = e ? atoi(e) : 1; }
        if (khs_bres) return tc::launch_gemm_op<64, tc::StagedEpi, true, true, tc::VAR_KHSB>(A, m, w, cout, 9 * cin, epi, st);
        return tc::launch_gemm_op<64, tc::StagedEpi, true, true>(A, m, w, cout, 9 * cin, epi, st);
    }
    return gemm_dispatch(A, m, w, cout, 9 * cin, bias, nullptr, relu, y, st);
}

// ---------------------------------------------------------------------------------------------------
// weight ingest: fold eval-mode BatchNorm into the convolution (torch: y = (conv(x) - mean)/sqrt(var+eps)*g + b)
//   w'[co][kh][kw][ci] = bf16( w[co][ci][kh][kw] * g/sqrt(var+eps) ),  bias'[co] = b - mean*g/sqrt(var+eps)
// K is padded with zeros up to kpad (stem: 147 -> 192).
// ---------------------------------------------------------------------------------------------------
__global__ void fold_bn_kernel(const float* __restrict__ w, int cout, int cin, int kh, int kw,
                               const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, float eps, int kpad,
                               __nv_bfloat16* __restrict__ wout, float* __restrict__ bout) {
    const int co = blockIdx.x;
    const float scale = gamma[co] / sqrtf(var[co] + eps);
    if (threadIdx.x == 0) bout[co] = beta[co] - mean[co] * scale;
    const int kk = kh * kw * cin;
    for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
        float v = 0.f;
        if (k < kk) {
            const int ci = k % cin, t = k / cin, x = t % kw, y = t / kw;
            v = w[(((size_t)co * cin + ci) * kh + y) * kw + x] * scale;
        }
        wout[(size_t)co * kpad + k] = __float2bfloat16_rn(v);
    }
}

__global__ void vec_add_f32_kernel(const float* a, const float* b, int n, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] + b[i];
}
int vec_add_f32(const float* a, const float* b, int n, float* out, cudaStream_t st) {
    vec_add_f32_kernel<<<ssg_cdiv(n, 256), 256, 0, st>>>(a, b, n, out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int fold_bn(const float* w, int cout, int cin, int kh, int kw, const float* gamma, const float* beta,
            const float* mean, const float* var, float eps, int kpad, void* wout, float* bout, cudaStream_t st) {
    fold_bn_kernel<<<cout, 256, 0, st>>>(w, cout, cin, kh, kw, gamma, beta, mean, var, eps, kpad,
                                         (__nv_bfloat16*)wout, bout);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// stem im2col: fp32 NCHW images [n,3,256,128] -> bf16 [2n*128*64, 192] patches of the 7x7 stride-2 pad-3 conv
// (K order (kh, kw, ci), zero padded 147 -> 192).  Rows [0, n*8192) are the images, rows [n*8192, 2n*8192) the
// horizontally flipped images (reid/evaluators.py:12-16 fliplr folded into the load index).
// One CTA per output row segment: (image, oh) -> 64 output pixels x 192.
// ---------------------------------------------------------------------------------------------------
constexpr int STEM_K = 147, STEM_KPAD = 192;   // 384-byte rows: every 128-byte TMA box row is cache-line aligned (a 152-wide
                                                // buffer is 20 % smaller but its misaligned rows doubled the GEMM's A traffic)

__global__ void __launch_bounds__(192)
stem_im2col_kernel(const float* __restrict__ img, int n, int flip_too, __nv_bfloat16* __restrict__ out) {
    constexpr int H = 256, W = 128, OW = 64, OH = 128, RW = W + 6;
    __shared__ float rows[3 * 7 * RW];             // 7 input rows x 3 channels, with the 3-pixel halo
    const int oh = blockIdx.x % OH;
    const int im = blockIdx.x / OH;                // 0 .. (flip_too ? 2n : n)
    const bool flipped = im >= n;
    const int src = flipped ? im - n : im;
    const float* base = img + (size_t)src * 3 * H * W;
    for (int e = threadIdx.x; e < 3 * 7 * RW; e += STEM_KPAD) {
        const int xw = e % RW, t = e / RW, ky = t % 7, c = t / 7;
        const int ih = oh * 2 - 3 + ky, iw = xw - 3;
        float v = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = base[((size_t)c * H + ih) * W + (flipped ? W - 1 - iw : iw)];
        rows[e] = v;
    }
    __syncthreads();
    // thread = (pair of K columns, output-pixel parity): the K -> (c, ky, kx) decode is done once per thread
    const int kp = threadIdx.x % (STEM_KPAD / 2), par = threadIdx.x / (STEM_KPAD / 2);
    int off[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int k = kp * 2 + h;
        off[h] = -1;
        if (k < STEM_K) {
            const int c = k % 3, t = k / 3, kx = t % 7, ky = t / 7;
            off[h] = (c * 7 + ky) * RW + kx;
        }
    }
    __nv_bfloat16* orow = out + ((size_t)im * OH + oh) * OW * STEM_KPAD + kp * 2;
    for (int ow = par; ow < OW; ow += 2) {
        const float v0 = off[0] >= 0 ? rows[off[0] + ow * 2] : 0.f;
        const float v1 = off[1] >= 0 ? rows[off[1] + ow * 2] : 0.f;
        *reinterpret_cast<__nv_bfloat162*>(orow + (size_t)ow * STEM_KPAD) = __floats2bfloat162_rn(v0, v1);
    }
}

// ---------------------------------------------------------------------------------------------------
// stem without an im2col buffer: (1) stem_prep: fp32 NCHW image -> bf16 P [images][256][144 px][4 ch] (3 zero px on
// the left, zeros on the right, zero 4th channel; mirror images appended when flip_too); (2) implicit GEMM whose K
// block kh is one overlapping-window TMA box of P (make_tmap_stem_windows): K = 7 x 64 (16 px x 4 ch per kernel row,
// of which 7 x 3 carry weights).
// ---------------------------------------------------------------------------------------------------
constexpr int STEM_WP = 144;

__global__ void __launch_bounds__(256)
stem_prep_kernel(const float* __restrict__ img, int n, int flip_too, __nv_bfloat16* __restrict__ P) {
    constexpr int H = 256, W = 128;
    const int images = flip_too ? 2 * n : n;
    const size_t total = (size_t)images * H * STEM_WP;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int p = (int)(e % STEM_WP);
        const size_t t = e / STEM_WP;
        const int ih = (int)(t % H), im = (int)(t / H);
        const bool flipped = im >= n;
        const float* base = img + (size_t)(flipped ? im - n : im) * 3 * H * W;
        const int iw = p - 3;
        float v[3] = {0.f, 0.f, 0.f};
        if (iw >= 0 && iw < W) {
            const int sw = flipped ? W - 1 - iw : iw;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = base[((size_t)c * H + ih) * W + sw];
        }
        uint2 pk;
        *reinterpret_cast<__nv_bfloat162*>(&pk.x) = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(&pk.y) = __floats2bfloat162_rn(v[2], 0.f);
        reinterpret_cast<uint2*>(P)[e] = pk;
    }
}

// The same for raw pixels: img uint8 HWC [n][256][128][3] as the decoder delivers them; the
// loader's ToTensor + Normalize (selftraining.py:36-45: x/255, then (x - mean)/std, IEEE fp32 in that order) happen
// here on the way to bf16, so a quarter of the bytes cross PCIe and no fp32 image is ever materialised.
struct StemNorm { float mean[3], std[3]; };

__global__ void __launch_bounds__(256)
stem_prep_u8_kernel(const uint8_t* __restrict__ img, int n, int flip_too, StemNorm nm, __nv_bfloat16* __restrict__ P) {
    constexpr int H = 256, W = 128;
    const int images = flip_too ? 2 * n : n;
    const size_t total = (size_t)images * H * STEM_WP;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int p = (int)(e % STEM_WP);
        const size_t t = e / STEM_WP;
        const int ih = (int)(t % H), im = (int)(t / H);
        const bool flipped = im >= n;
        const int iw = p - 3;
        float v[3] = {0.f, 0.f, 0.f};
        if (iw >= 0 && iw < W) {
            const int sw = flipped ? W - 1 - iw : iw;
            const uint8_t* px = img + (((size_t)(flipped ? im - n : im) * H + ih) * W + sw) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c)
                v[c] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[c], 255.f), nm.mean[c]), nm.std[c]);
        }
        uint2 pk;
        *reinterpret_cast<__nv_bfloat162*>(&pk.x) = __floats2bfloat162_rn(v[0], v[1]);
        *reinterpret_cast<__nv_bfloat162*>(&pk.y) = __floats2bfloat162_rn(v[2], 0.f);
        reinterpret_cast<uint2*>(P)[e] = pk;
    }
}

int stem_prep_u8(const uint8_t* img, int n, int flip_too, const float* mean, const float* std, void* P,
                 cudaStream_t st) {
    StemNorm nm;
    for (int c = 0; c < 3; ++c) { nm.mean[c] = mean[c]; nm.std[c] = std[c]; }
    const size_t total = (size_t)(flip_too ? 2 * n : n) * 256 * STEM_WP;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    stem_prep_u8_kernel<<<grid, 256, 0, st>>>(img, n, flip_too, nm, (__nv_bfloat16*)P);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

int stem_prep(const float* img, int n, int flip_too, void* P, cudaStream_t st) {
    const size_t total = (size_t)(flip_too ? 2 * n : n) * 256 * STEM_WP;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    stem_prep_kernel<<<grid, 256, 0, st>>>(img, n, flip_too, (__nv_bfloat16*)P);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// weights for the window GEMM: w'[co][kh][px 0..15][c 0..3] = w[co][c][kh][px] * bn_scale for px < 7, c < 3, else 0
__global__ void fold_bn_stem_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                    const float* __restrict__ beta, const float* __restrict__ mean,
                                    const float* __restrict__ var, float eps, __nv_bfloat16* __restrict__ wout,
                                    float* __restrict__ bout) {
    const int co = blockIdx.x;
    const float scale = gamma[co] / sqrtf(var[co] + eps);
    if (threadIdx.x == 0) bout[co] = beta[co] - mean[co] * scale;
    for (int k = threadIdx.x; k < 7 * 64; k += blockDim.x) {
        const int c = k & 3, px = (k >> 2) & 15, kh = k >> 6;
        float v = 0.f;
        if (c < 3 && px < 7) v = w[(((size_t)co * 3 + c) * 7 + kh) * 7 + px] * scale;
        wout[(size_t)co * 448 + k] = __float2bfloat16_rn(v);
    }
}
// 64-byte-row variant: w'[co][kh 0..7][px 0..7][c 0..3] (kh = 7, px = 7 and c = 3 are zero)  ->  K = 256
__global__ void fold_bn_stem64_kernel(const float* __restrict__ w, const float* __restrict__ gamma,
                                      const float* __restrict__ var, float eps, __nv_bfloat16* __restrict__ wout) {
    const int co = blockIdx.x;
    const float scale = gamma[co] / sqrtf(var[co] + eps);
    for (int k = threadIdx.x; k < 256; k += blockDim.x) {
        const int c = k & 3, px = (k >> 2) & 7, kh = k >> 5;
        float v = 0.f;
        if (c < 3 && px < 7 && kh < 7) v = w[(((size_t)co * 3 + c) * 7 + kh) * 7 + px] * scale;
        wout[(size_t)co * 256 + k] = __float2bfloat16_rn(v);
    }
}
int fold_bn_stem(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                 void* wout, float* bout, void* wout64, cudaStream_t st) {
    fold_bn_stem_kernel<<<64, 256, 0, st>>>(w, gamma, beta, mean, var, eps, (__nv_bfloat16*)wout, bout);
    SSG_CHECK_LAUNCH();
    fold_bn_stem64_kernel<<<64, 256, 0, st>>>(w, gamma, var, eps, (__nv_bfloat16*)wout64);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}
// pool_out != NULL: the 3x3/2 max-pool is fused behind the convolution (y is then only a placeholder for the tensor
// map; the full-resolution map is never written) -- pool_out is [images, 64, 32, 64] NHWC bf16.
int conv_stem_windows64(const void* P, int images, const void* w256, const float* bias, void* y, cudaStream_t st,
                        void* pool_out) {
    tc::AOperand A;
    memset(&A, 0, sizeof(A));
    A.mode = 3;
    A.cblks = 1;
    A.taps = 8;
    A.tiles_per_img = 64;
    A.hmul = 1;
    SSG_TRY(make_tmap_stem_windows64(&A.map[0], P, (uint64_t)images, stem_variant() == 2 ? 10 : 4));
    return gemm_dispatch(A, images * 8192, w256, 64, 256, bias, nullptr, 1, y, st, pool_out);
}

// the window GEMM: P [images][256][144][4] -> relu(conv7x7/2 + bias) as NHWC bf16 [images,128,64,64]
int conv_stem_windows(const void* P, int images, const void* w448, const float* bias, void* y, cudaStream_t st) {
    tc::AOperand A;
    memset(&A, 0, sizeof(A));
    A.mode = 2;
    A.cblks = 1;
    A.taps = 7;
    A.tiles_per_img = 64;
    A.hmul = 1;
    SSG_TRY(make_tmap_stem_windows(&A.map[0], P, (uint64_t)images));
    return gemm_dispatch(A, images * 8192, w448, 64, 448, bias, nullptr, 1, y, st);
}

int stem_im2col(const float* img, int n, int flip_too, void* out, cudaStream_t st) {
    const int images = flip_too ? 2 * n : n;
    stem_im2col_kernel<<<images * 128, STEM_KPAD, 0, st>>>(img, n, flip_too, (__nv_bfloat16*)out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// max-pool 3x3 stride 2 pad 1 on NHWC bf16: [B,H,W,C] -> [B,H/2,W/2,C]; 8 channels (16 bytes) per thread
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
maxpool3x3s2_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C, __nv_bfloat16* __restrict__ y) {
    // one thread = one output pixel x 8 channels.  Window coordinates are CLAMPED into the map instead of skipped: a
    // duplicate tap cannot change a maximum, and the nine 16-byte loads become unconditional and independent (all in
    // flight at once); the maximum is taken on packed bf16 pairs (exact, no float round trip).
    const int OH = H / 2, OW = W / 2, CV = C / 8;
    const size_t total = (size_t)B * OH * OW * CV;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(e % CV);
        size_t t = e / CV;
        const int ow = (int)(t % OW); t /= OW;
        const int oh = (int)(t % OH);
        const int b = (int)(t / OH);
        const __nv_bfloat16* base = x + (size_t)b * H * W * C + cv * 8;
        uint4 v[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int ih = min(max(oh * 2 - 1 + ky, 0), H - 1);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int iw = min(max(ow * 2 - 1 + kx, 0), W - 1);
                v[ky * 3 + kx] = *reinterpret_cast<const uint4*>(base + ((size_t)ih * W + iw) * C);
            }
        }
        uint4 o = v[0];
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 1; k < 9; ++k) {
            const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v[k]);
#pragma unroll
            for (int q = 0; q < 4; ++q) po[q] = __hmax2(po[q], p[q]);
        }
        *reinterpret_cast<uint4*>(y + (((size_t)b * OH + oh) * OW + ow) * C + cv * 8) = o;
    }
}

int maxpool3x3s2(const void* x, int B, int H, int W, int C, void* y, cudaStream_t st) {
    const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8);
    const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    maxpool3x3s2_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, B, H, W, C, (__nv_bfloat16*)y);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// parity split for the stride-2 convolutions: x [B,H,W,C] -> planes [P][B,H/2,W/2,C], plane = (h&1)*2 + (w&1);
// nplanes = 1 keeps only the (even, even) plane (the 1x1 stride-2 downsample branch).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
parity_split_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int C, int nplanes,
                    __nv_bfloat16* __restrict__ y) {
    const int OH = H / 2, OW = W / 2, CV = C / 8;
    const size_t plane = (size_t)B * OH * OW * CV;
    const size_t total = plane * nplanes;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int pl = (int)(e / plane);
        size_t t = e % plane;
        const int cv = (int)(t % CV); t /= CV;
        const int ow = (int)(t % OW); t /= OW;
        const int oh = (int)(t % OH);
        const int b = (int)(t / OH);
        const int ih = oh * 2 + (pl >> 1), iw = ow * 2 + (pl & 1);
        reinterpret_cast<uint4*>(y)[e] =
            *reinterpret_cast<const uint4*>(x + (((size_t)b * H + ih) * W + iw) * C + cv * 8);
    }
}

int parity_split(const void* x, int B, int H, int W, int C, int nplanes, void* y, cudaStream_t st) {
    const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 8) * nplanes;
    const int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    parity_split_kernel<<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, B, H, W, C, nplanes, (__nv_bfloat16*)y);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// pooled tail (resnet.py:93-111 + evaluators.py:29-46): layer4 map [2n,8,4,2048] (images then flipped images)
//   bank 0 = global average, bank 1+s = average of rows [h//S*s, h//S*(s+1)) (S = num_split > 1);
//   out = pool(img) + pool(flipped img); list mode: each bank / its own L2 norm -> feat[bank][row0+i][2048];
//   eval mode: one norm over the concatenated banks -> feat[row0+i][(S+1)*2048].
// One CTA per image, 256 threads x 8 channels.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pooled_tail_kernel(const __nv_bfloat16* __restrict__ x, int n, int num_split, int eval_mode, int flip_too,
                   float* __restrict__ feat, size_t bank_stride, int row0) {
    constexpr int H = 8, W = 4, C = 2048;
    const int i = blockIdx.x;
    const int c0 = threadIdx.x * 8;
    const int nb = num_split > 1 ? num_split + 1 : 1;
    const int step = num_split > 1 ? H / num_split : H;
    float acc[5][8];
#pragma unroll
    for (int b = 0; b < 5; ++b)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[b][q] = 0.f;
    // avg_pool2d = sum / area in float32; (pool(a) + pool(b)) as the reference adds the two passes
    for (int pass = 0; pass < (flip_too ? 2 : 1); ++pass) {
        const __nv_bfloat16* img = x + ((size_t)(pass * n + i) * H * W) * C + c0;
        float g[8], s[4][8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { g[q] = 0.f; s[0][q] = s[1][q] = s[2][q] = s[3][q] = 0.f; }
        // all 32 loads of the pass are independent: issue them before the first use (the kernel was latency bound with
        // one load in flight per thread: 76 us per 512-image batch for 67 MB, round-2 profile)
        uint4 px[H * W];
#pragma unroll
        for (int e = 0; e < H * W; ++e) px[e] = *reinterpret_cast<const uint4*>(img + (size_t)e * C);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            float r[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = 0.f;
#pragma unroll
            for (int w = 0; w < W; ++w) {
                const uint4 v = px[h * W + w];
                const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = __bfloat1622float2(p[q]);
                    r[2 * q] += f.x;
                    r[2 * q + 1] += f.y;
                }
            }
            const int sb = h / step;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                g[q] += r[q];
                if (num_split > 1 && sb < num_split && sb < 4) s[sb][q] += r[q];
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            acc[0][q] += g[q] / (float)(H * W);
            for (int b = 1; b < nb; ++b) acc[b][q] += s[b - 1][q] / (float)(step * W);
        }
    }
    // L2 norms
    __shared__ float red[5][8];
    float ss[5];
    for (int b = 0; b < nb; ++b) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += acc[b][q] * acc[b][q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        ss[b] = t;
    }
    if ((threadIdx.x & 31) == 0)
        for (int b = 0; b < nb; ++b) red[b][threadIdx.x >> 5] = ss[b];
    __syncthreads();
    float norm[5], all = 0.f;
    for (int b = 0; b < nb; ++b) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[b][w];
        all += t;
        norm[b] = sqrtf(t);
    }
    const float norm_all = sqrtf(all);
    const bool raw = (eval_mode & 2) != 0;       // bit 1: leave the pooled banks un-normalised (cnn.py:10-23)
    eval_mode &= 1;
    for (int b = 0; b < nb; ++b) {
        float* o = eval_mode ? feat + ((size_t)(row0 + i) * nb + b) * C + c0
                             : feat + (size_t)b * bank_stride + (size_t)(row0 + i) * C + c0;
        const float dv = raw ? 1.0f : (eval_mode ? norm_all : norm[b]);
        float4 o0, o1;
        o0.x = acc[b][0] / dv; o0.y = acc[b][1] / dv; o0.z = acc[b][2] / dv; o0.w = acc[b][3] / dv;
        o1.x = acc[b][4] / dv; o1.y = acc[b][5] / dv; o1.z = acc[b][6] / dv; o1.w = acc[b][7] / dv;
        *reinterpret_cast<float4*>(o) = o0;
        *reinterpret_cast<float4*>(o + 4) = o1;
    }
}

int pooled_tail(const void* x, int n, int num_split, int eval_mode, int flip_too, float* feat, size_t bank_stride,
                int row0, cudaStream_t st) {
    if (num_split < 1 || num_split > 4) return ssg_set_error(SSG_ERR_INVALID, "num_split=%d out of range", num_split);
    pooled_tail_kernel<<<n, 256, 0, st>>>((const __nv_bfloat16*)x, n, num_split, eval_mode, flip_too, feat, bank_stride,
                                          row0);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // namespace ssg
