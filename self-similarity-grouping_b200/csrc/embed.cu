// Embedding plan: the ResNet-50 trunk of reid/models/resnet.py:86-111 (torchvision Bottleneck graph, stride on
// the 3x3 conv) + flip test-time augmentation and L2 normalisation of reid/evaluators.py:18-60, as one
// device-resident forward over a batch of images.  Weights are ingested from the live torch module's
// state_dict (BatchNorm folded, bf16, [Cout][kh][kw][Cin]); activations are NHWC bf16, accumulation fp32.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "common.cuh"
#include "conv.h"
#include "gemm_tc.cuh"

using namespace ssg;

namespace {
struct LayerSpec {
    int cin, cout, k, stride;
    char conv_key[48], bn_key[48];
};

// canonical layer order: stem, then per block conv1, conv2, conv3, [downsample]
std::vector<LayerSpec> build_specs() {
    std::vector<LayerSpec> v;
    LayerSpec s{3, 64, 7, 2, "conv1", "bn1"};
    v.push_back(s);
    const int blocks[4] = {3, 4, 6, 3};
    int cin = 64;
    for (int L = 0; L < 4; ++L) {
        const int mid = 64 << L, out = mid * 4;
        for (int b = 0; b < blocks[L]; ++b) {
            const int stride = (b == 0 && L > 0) ? 2 : 1;
            LayerSpec c1{cin, mid, 1, 1, "", ""}, c2{mid, mid, 3, stride, "", ""}, c3{mid, out, 1, 1, "", ""};
            snprintf(c1.conv_key, 48, "layer%d.%d.conv1", L + 1, b); snprintf(c1.bn_key, 48, "layer%d.%d.bn1", L + 1, b);
            snprintf(c2.conv_key, 48, "layer%d.%d.conv2", L + 1, b); snprintf(c2.bn_key, 48, "layer%d.%d.bn2", L + 1, b);
            snprintf(c3.conv_key, 48, "layer%d.%d.conv3", L + 1, b); snprintf(c3.bn_key, 48, "layer%d.%d.bn3", L + 1, b);
            v.push_back(c1); v.push_back(c2); v.push_back(c3);
            if (b == 0) {
                LayerSpec ds{cin, out, 1, stride, "", ""};
                snprintf(ds.conv_key, 48, "layer%d.%d.downsample.0", L + 1, b);
                snprintf(ds.bn_key, 48, "layer%d.%d.downsample.1", L + 1, b);
                v.push_back(ds);
            }
            cin = out;
        }
    }
    return v;
}
const std::vector<LayerSpec>& specs() {
    static std::vector<LayerSpec> s = build_specs();
    return s;
}
int kpad_of(const LayerSpec& s) {
    const int k = s.k * s.k * s.cin;
    return (k + 63) / 64 * 64;   // whole K blocks (stem: 147 -> 192); the GEMM itself also accepts K % 8 == 0
}
}  // namespace

struct ssg_embed_plan {
    int device, batch_max;
    size_t bytes;
    std::vector<void*> w;        // bf16 [cout, kpad]
    std::vector<float*> b;       // fp32 [cout]
    std::vector<char> loaded;
    void* wf[4];                 // fused [conv3 | downsample] weights of the first block of each layer
    float* bf[4];                // fused bias
    bool fused_ready;
    void* w_stem448;             // stem weights for the overlapping-window GEMM [64, 7*64]
    void* w_stem256;             // ... 64-byte-row variant [64, 8*32]
    float* b_stem448;
    void* stemP;                 // padded 4-channel bf16 input [2*batch][256][144][4]
    int stem_windows;            // 1: window GEMM (no im2col), 0: im2col + GEMM, -1: undecided
    void *col, *stem, *x, *y, *ds, *t1, *t2, *planes, *xs;
    // SSG_L2_CHUNK: the chunk loop over layers 1-2 as a CUDA graph on a side stream (one graph per batch size)
    cudaStream_t side;
    cudaEvent_t ev_in, ev_out;
    cudaGraphExec_t l2_graph;
    int l2_graph_nb, l2_graph_chunk;
};

static int ealloc(void** p, size_t bytes, size_t* total) {
    SSG_CUDA_TRY(cudaMalloc(p, bytes ? bytes : 16));
    *total += bytes;
    return SSG_OK;
}

extern "C" int ssg_embed_num_layers(void) { return (int)specs().size(); }

extern "C" int ssg_embed_layer_info(int idx, int* cin, int* cout, int* ksize, int* stride, char* conv_key,
                                    char* bn_key, size_t cap) {
    if (idx < 0 || idx >= (int)specs().size()) return ssg_set_error(SSG_ERR_INVALID, "layer index %d", idx);
    const LayerSpec& s = specs()[idx];
    if (cin) *cin = s.cin;
    if (cout) *cout = s.cout;
    if (ksize) *ksize = s.k;
    if (stride) *stride = s.stride;
    if (conv_key && cap) { strncpy(conv_key, s.conv_key, cap - 1); conv_key[cap - 1] = 0; }
    if (bn_key && cap) { strncpy(bn_key, s.bn_key, cap - 1); bn_key[cap - 1] = 0; }
    return SSG_OK;
}

extern "C" int ssg_embed_plan_destroy(ssg_embed_plan* p) {
    if (!p) return SSG_OK;
    SsgDeviceGuard device_guard__(p->device);
    for (void* q : p->w) if (q) cudaFree(q);
    for (int L = 0; L < 4; ++L) { if (p->wf[L]) cudaFree(p->wf[L]); if (p->bf[L]) cudaFree(p->bf[L]); }
    for (float* q : p->b) if (q) cudaFree(q);
    void* bufs[] = {p->col, p->stem, p->x, p->y, p->ds, p->t1, p->t2, p->planes, p->xs, p->w_stem448, p->w_stem256, p->b_stem448,
                    p->stemP};
    for (void* q : bufs) if (q) cudaFree(q);
    if (p->l2_graph) cudaGraphExecDestroy(p->l2_graph);
    if (p->ev_in) cudaEventDestroy(p->ev_in);
    if (p->ev_out) cudaEventDestroy(p->ev_out);
    if (p->side) cudaStreamDestroy(p->side);
    delete p;
    return SSG_OK;
}

extern "C" int ssg_embed_plan_create(ssg_embed_plan** out, int device, int batch_max, int height, int width) {
    if (!out || batch_max <= 0) return ssg_set_error(SSG_ERR_INVALID, "embed_plan_create: bad arguments");
    if (height != 256 || width != 128)
        return ssg_set_error(SSG_ERR_UNSUPPORTED, "embed: only 256x128 inputs are supported (got %dx%d)", height, width);
    SSG_ON_DEVICE(device);
    ssg_embed_plan* p = new ssg_embed_plan();
    p->device = device; p->batch_max = batch_max; p->bytes = 0;
    p->col = p->stem = p->x = p->y = p->ds = p->t1 = p->t2 = p->planes = p->xs = nullptr;
    p->side = nullptr; p->ev_in = nullptr; p->ev_out = nullptr; p->l2_graph = nullptr; p->l2_graph_nb = 0; p->l2_graph_chunk = 0;
    const auto& sp = specs();
    p->w.assign(sp.size(), nullptr);
    p->b.assign(sp.size(), nullptr);
    p->loaded.assign(sp.size(), 0);
    p->fused_ready = false;
    for (int L = 0; L < 4; ++L) { p->wf[L] = nullptr; p->bf[L] = nullptr; }
    p->w_stem448 = nullptr; p->w_stem256 = nullptr; p->b_stem448 = nullptr; p->stemP = nullptr; p->stem_windows = -1;
    int rc = SSG_OK;
    for (size_t i = 0; i < sp.size() && rc == SSG_OK; ++i) {
        rc = ealloc(&p->w[i], (size_t)sp[i].cout * kpad_of(sp[i]) * 2, &p->bytes);
        if (rc == SSG_OK) rc = ealloc((void**)&p->b[i], sizeof(float) * sp[i].cout, &p->bytes);
    }
    {   // fused first-block tails: [outc, mid + cin]
        int cin = 64;
        for (int L = 0; L < 4 && rc == SSG_OK; ++L) {
            const int mid = 64 << L, outc = mid * 4;
            rc = ealloc(&p->wf[L], (size_t)outc * (mid + cin) * 2, &p->bytes);
            if (rc == SSG_OK) rc = ealloc((void**)&p->bf[L], sizeof(float) * outc, &p->bytes);
            cin = outc;
        }
    }
    const size_t nb = (size_t)batch_max * 2;       // images + flipped images
    if (rc == SSG_OK) rc = ealloc(&p->w_stem448, (size_t)64 * 448 * 2, &p->bytes);
    if (rc == SSG_OK) rc = ealloc(&p->w_stem256, (size_t)64 * 256 * 2, &p->bytes);
    if (rc == SSG_OK) rc = ealloc((void**)&p->b_stem448, sizeof(float) * 64, &p->bytes);
    if (rc == SSG_OK) rc = ealloc(&p->stemP, nb * 256 * 144 * 8, &p->bytes);
    const size_t px = 2;                           // bytes per bf16
#define A(ptr, elems) if (rc == SSG_OK) rc = ealloc(&(ptr), (elems) * px, &p->bytes)
    A(p->col, nb * 8192 * 192);
    A(p->stem, nb * 8192 * 64);
    A(p->x, nb * 2048 * 256);
    A(p->y, nb * 2048 * 256);
    A(p->ds, nb * 2048 * 256);
    A(p->t1, nb * 2048 * 128);
    A(p->t2, nb * 2048 * 64);
    A(p->planes, nb * 2048 * 128);
    A(p->xs, nb * 512 * 256);
#undef A
    if (rc != SSG_OK) { ssg_embed_plan_destroy(p); return rc; }
    *out = p;
    return SSG_OK;
}

extern "C" size_t ssg_embed_plan_bytes(const ssg_embed_plan* p) { return p ? p->bytes : 0; }

extern "C" int ssg_embed_load_layer(ssg_embed_plan* p, int idx, const float* d_w, const float* d_gamma,
                                    const float* d_beta, const float* d_mean, const float* d_var, float eps,
                                    void* stream) {
    if (!p || idx < 0 || idx >= (int)specs().size() || !d_w || !d_gamma || !d_beta || !d_mean || !d_var)
        return ssg_set_error(SSG_ERR_INVALID, "embed_load_layer: bad arguments");
    SSG_ON_DEVICE(p->device);
    const LayerSpec& s = specs()[idx];
    SSG_TRY(fold_bn(d_w, s.cout, s.cin, s.k, s.k, d_gamma, d_beta, d_mean, d_var, eps, kpad_of(s), p->w[idx],
                    p->b[idx], (cudaStream_t)stream));
    if (idx == 0)
        SSG_TRY(fold_bn_stem(d_w, d_gamma, d_beta, d_mean, d_var, eps, p->w_stem448, p->b_stem448, p->w_stem256,
                             (cudaStream_t)stream));
    p->loaded[idx] = 1;
    p->fused_ready = false;
    return SSG_OK;
}

// One bottleneck block (conv1 1x1 -> conv2 3x3 [stride on it] -> conv3 1x1 + shortcut, ReLU) over NB image-passes.
// x: input [NB,H,W,C], y: the other ping-pong buffer; on return x holds the output and H, W, C, li are advanced.
// out (chunked mode, last block of a chunk): the output goes there instead of y, and x / y are left alone.
// t1_ready (layer 1 with the chained kernel): on entry, true = the previous block's chained launch already wrote this
// block's conv1 output into p->t1; on return, true = this block did the same for the next one.
static int run_block(ssg_embed_plan* p, int L, int b, int NB, int fuse_ds, void*& x, void*& y, int& H, int& W, int& C,
                     int& li, void* out, cudaStream_t st, bool* t1_ready = nullptr) {
    const int mid = 64 << L, outc = mid * 4;
    const int stride = (b == 0 && L > 0) ? 2 : 1;
    const int OH = H / stride, OW = W / stride;
    const int i1 = li, i2 = li + 1, i3 = li + 2, id = li + 3;
    li += (b == 0) ? 4 : 3;
    const bool have_t1 = t1_ready && *t1_ready;
    if (t1_ready) *t1_ready = false;
    if (!have_t1) { SSG_PROF("conv1x1_tc", st); SSG_TRY(conv1x1(x, NB * H * W, C, p->w[i1], p->b[i1], mid, nullptr, 1, p->t1, st)); }
    // layer 1: conv3 (+ shortcut, ReLU) and the NEXT block's conv1 (index li after the advance above: the next block of
    // layer 1, or the first block of layer 2) as one launch; its conv2 below reads p->t1 before this block overwrites it
    const bool chain = L == 0 && t1_ready && !out && conv_chain_enabled() && (b > 0 || fuse_ds) &&
                       li < (int)specs().size() && specs()[li].k == 1 && specs()[li].cin == outc;
    if (stride == 2 && s2_strided_tma()) {
        { SSG_PROF("conv3x3_tc", st); SSG_TRY(conv3x3(p->t1, NB, OH, OW, mid, 2, p->w[i2], p->b[i2], mid, 1, p->t2, st)); }
    } else if (stride == 2) {
        { SSG_PROF("parity_split", st); SSG_TRY(parity_split(p->t1, NB, H, W, mid, 4, p->planes, st)); }
        { SSG_PROF("conv3x3_tc", st); SSG_TRY(conv3x3(p->planes, NB, OH, OW, mid, 2, p->w[i2], p->b[i2], mid, 1, p->t2, st)); }
    } else {
        { SSG_PROF("conv3x3_tc", st); SSG_TRY(conv3x3(p->t1, NB, H, W, mid, 1, p->w[i2], p->b[i2], mid, 1, p->t2, st)); }
    }
    if (chain) {
        const int n2 = specs()[li].cout;
        SSG_PROF("conv1x1_tc", st);
        if (b == 0)
            SSG_TRY(conv_chain(p->t2, NB, OH, OW, mid, x, C, p->wf[L], p->bf[L], nullptr, y, p->w[li], p->b[li], n2, p->t1, st));
        else
            SSG_TRY(conv_chain(p->t2, NB, OH, OW, mid, nullptr, 0, p->w[i3], p->b[i3], x, y, p->w[li], p->b[li], n2, p->t1, st));
        *t1_ready = true;
        void* t = x; x = y; y = t;
        H = OH; W = OW; C = outc;
        return SSG_OK;
    }
    if (b == 0 && fuse_ds) {
        if (out) return ssg_set_error(SSG_ERR_INVALID, "embed: a chunk cannot end on a downsample block");
        { SSG_PROF("conv1x1_tc", st); SSG_TRY(conv_fused_ds(p->t2, x, NB, OH, OW, mid, C, stride, p->wf[L], p->bf[L], outc, y, st)); }
        void* t = x; x = y; y = t;
        H = OH; W = OW; C = outc;
        return SSG_OK;
    }
    const void* res = x;
    if (b == 0) {
        if (stride == 2 && s2_strided_tma()) {
            { SSG_PROF("conv1x1_tc", st); SSG_TRY(conv1x1_s2(x, NB, OH, OW, C, p->w[id], p->b[id], outc, 0, p->ds, st)); }
        } else if (stride == 2) {
            { SSG_PROF("parity_split", st); SSG_TRY(parity_split(x, NB, H, W, C, 1, p->xs, st)); }
            { SSG_PROF("conv1x1_tc", st); SSG_TRY(conv1x1(p->xs, NB * OH * OW, C, p->w[id], p->b[id], outc, nullptr, 0, p->ds, st)); }
        } else {
            { SSG_PROF("conv1x1_tc", st); SSG_TRY(conv1x1(x, NB * H * W, C, p->w[id], p->b[id], outc, nullptr, 0, p->ds, st)); }
        }
        res = p->ds;
    }
    if (out) {
        { SSG_PROF("conv1x1_tc", st); SSG_TRY(conv1x1(p->t2, NB * OH * OW, mid, p->w[i3], p->b[i3], outc, res, 1, out, st)); }
        H = OH; W = OW; C = outc;
        return SSG_OK;
    }
    { SSG_PROF("conv1x1_tc", st); SSG_TRY(conv1x1(p->t2, NB * OH * OW, mid, p->w[i3], p->b[i3], outc, res, 1, y, st)); }
    void* t = x; x = y; y = t;
    H = OH; W = OW; C = outc;
    return SSG_OK;
}

// One forward over a batch; the images come either as fp32 NCHW (d_images) or as raw uint8 HWC pixels (d_u8 with the
// loader's per-channel mean / std).
static int embed_forward_impl(ssg_embed_plan* p, const float* d_images, const uint8_t* d_u8, const float* mean,
                              const float* stdv, int n, int num_split, int eval_mode, int flip, float* d_feat,
                              size_t bank_stride, int row0, void* stream) {
    if (!p || (!d_images && !d_u8) || !d_feat || n <= 0 || n > p->batch_max)
        return ssg_set_error(SSG_ERR_INVALID, "embed_forward: bad arguments (n=%d, batch_max=%d)", n, p ? p->batch_max : -1);
    // resnet.py:93-108 accepts any num_split up to the map height (8 rows); the pooled tail holds up to 4 stripes
    // (+ the global bank) in registers, which covers every configuration the drivers use (run.sh: 2; BASELINE: 2, 3)
    if (num_split < 1 || num_split > 4)
        return ssg_set_error(SSG_ERR_INVALID, "embed_forward: num_split=%d out of range (1..4)", num_split);
    for (size_t i = 0; i < p->loaded.size(); ++i)
        if (!p->loaded[i]) return ssg_set_error(SSG_ERR_INVALID, "embed_forward: layer %zu (%s) not loaded", i, specs()[i].conv_key);
    SSG_ON_DEVICE(p->device);
    cudaStream_t st = (cudaStream_t)stream;
    const auto& sp = specs();
    const int NB = flip ? 2 * n : n;
    static int fuse_ds = -1;
    if (fuse_ds < 0) { const char* e = getenv("SSG_FUSE_DS"); fuse_ds = e ? atoi(e) : 1; }
    if (fuse_ds && !p->fused_ready) {
        // [conv3 | downsample] weights side by side along K, biases added (first block of every layer)
        int idx = 1, cin = 64;
        const int nblk[4] = {3, 4, 6, 3};
        for (int L = 0; L < 4; ++L) {
            const int mid = 64 << L, outc = mid * 4, i3 = idx + 2, id = idx + 3;
            const size_t pitch = (size_t)(mid + cin) * 2;
            SSG_CUDA_TRY(cudaMemcpy2DAsync(p->wf[L], pitch, p->w[i3], (size_t)mid * 2, (size_t)mid * 2, outc,
                                           cudaMemcpyDeviceToDevice, st));
            SSG_CUDA_TRY(cudaMemcpy2DAsync((char*)p->wf[L] + (size_t)mid * 2, pitch, p->w[id], (size_t)cin * 2,
                                           (size_t)cin * 2, outc, cudaMemcpyDeviceToDevice, st));
            SSG_TRY(vec_add_f32(p->b[i3], p->b[id], outc, p->bf[L], st));
            idx += 4 + 3 * (nblk[L] - 1);
            cin = outc;
        }
        p->fused_ready = true;
    }
    int li = 0;
    // stem: 7x7/2 conv as im2col + GEMM (K 147 -> 192), then 3x3/2 max-pool
    if (p->stem_windows < 0) {
        // overlapping-window TMA view of the input (no im2col buffer) unless SSG_STEM_WINDOWS=0 or the driver refuses
        const char* e = getenv("SSG_STEM_WINDOWS");
        // 2 (default): 8-pixel windows, 64-byte swizzle; 1: 16-pixel windows, 128-byte swizzle; 0: im2col + GEMM
        p->stem_windows = e ? atoi(e) : 2;
        if (p->stem_windows) {
            CUtensorMap probe;
            if (make_tmap_stem_windows(&probe, p->stemP, 2) != SSG_OK) p->stem_windows = 0;
        }
    }
    bool pooled = false;       // the max-pool already ran inside the stem kernel
    if (d_u8 && !p->stem_windows)
        return ssg_set_error(SSG_ERR_UNSUPPORTED, "embed_forward_u8 needs the window stem (SSG_STEM_WINDOWS != 0)");
    // SSG_L2_CHUNK=<image-passes> (default 0 = off): run layers 1-2 -- whose 1x1 convolutions are HBM bound, DESIGN.md
    // 3.1 -- over chunks of that many image-passes that re-use the head of the ping-pong buffers, so that a chunk's
    // activations (2.5 MB per image-pass in layer 1: 32 passes = 80 MB) stay in the 126 MB L2 from the kernel that
    // writes them to the kernel that reads them and are overwritten there by the next chunk instead of being written
    // back.  The stem (whole images per CTA: it wants >= 148 of them) and layers 3-4 (tensor bound, they want the big
    // grids) run over the whole batch as before; the pooled stem map and the layer-2 outputs of all passes live in
    // full-batch buffers that are idle in this configuration (the unpooled-stem and the im2col buffer).  Original and
    // mirrored images are independent passes until the pooled tail, so a chunk is any contiguous range of passes.
    // Same kernels on the same per-image tiles: bit-identical features.
    static int l2_chunk = -1;
    if (l2_chunk < 0) { const char* e = getenv("SSG_L2_CHUNK"); l2_chunk = e ? atoi(e) : 0; }
    if (l2_chunk > 0 && NB > l2_chunk && p->stem_windows == 2 && stem_pool_fused() && fuse_ds) {
        {
            SSG_PROF("stem_prep", st);
            if (d_u8) SSG_TRY(stem_prep_u8(d_u8, n, flip, mean, stdv, p->stemP, st));
            else SSG_TRY(stem_prep(d_images, n, flip, p->stemP, st));
        }
        char* pooled_all = (char*)p->stem;                    // [NB][64][32][64] bf16
        char* l2_all = (char*)p->col;                         // [NB][32][16][512] bf16
        const size_t E0 = (size_t)64 * 32 * 64 * 2, E2 = (size_t)32 * 16 * 512 * 2;   // bytes per image-pass
        { SSG_PROF("conv_stem_tc", st); SSG_TRY(conv_stem_windows64(p->stemP, NB, p->w_stem256, p->b_stem448, /* tensor-map placeholder, never written */ p->x, st, pooled_all)); }
        const int blocks[4] = {3, 4, 6, 3};
        auto chunk_loop = [&](cudaStream_t cs) -> int {
            for (int q0 = 0; q0 < NB; q0 += l2_chunk) {
                const int NBc = NB - q0 < l2_chunk ? NB - q0 : l2_chunk;
                int lc = 1, H = 64, W = 32, C = 64;
                void *x = pooled_all + (size_t)q0 * E0, *y = p->y;
                for (int L = 0; L < 2; ++L)
                    for (int b = 0; b < blocks[L]; ++b) {
                        const bool last = L == 1 && b == blocks[1] - 1;
                        SSG_TRY(run_block(p, L, b, NBc, fuse_ds, x, y, H, W, C, lc, last ? l2_all + (size_t)q0 * E2 : nullptr, cs));
                        if (L == 0 && b == 0) y = p->x;       // the chunk's input slice is read-only: ping-pong on x / y
                    }
            }
            return SSG_OK;
        };
        // The loop issues (NB / chunk) x 21 launches whose host cost (tensor-map encodes + launch, ~8 us each) would
        // exceed their GPU time: record it ONCE per batch size as a CUDA graph on a side stream (every pointer in it is
        // plan-owned, so the graph is valid for every later batch of that size) and replay it.  SSG_L2_GRAPH=0: direct.
        static int l2_graph = -1;
        if (l2_graph < 0) { const char* e = getenv("SSG_L2_GRAPH"); l2_graph = e ? atoi(e) : 1; }
        if (l2_graph) {
            if (!p->side) {
                SSG_CUDA_TRY(cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking));
                SSG_CUDA_TRY(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
                SSG_CUDA_TRY(cudaEventCreateWithFlags(&p->ev_out, cudaEventDisableTiming));
            }
            if (!p->l2_graph || p->l2_graph_nb != NB || p->l2_graph_chunk != l2_chunk) {
                if (p->l2_graph) { cudaGraphExecDestroy(p->l2_graph); p->l2_graph = nullptr; }
                cudaGraph_t graph = nullptr;
                prof_suspend(true);
                cudaError_t ce = cudaStreamBeginCapture(p->side, cudaStreamCaptureModeThreadLocal);
                int rc = ce == cudaSuccess ? chunk_loop(p->side) : SSG_OK;
                if (ce == cudaSuccess) ce = cudaStreamEndCapture(p->side, &graph);
                prof_suspend(false);
                if (rc != SSG_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
                if (ce != cudaSuccess || !graph)
                    return ssg_set_error(SSG_ERR_CUDA, "embed: capturing the chunk loop failed: %s", cudaGetErrorString(ce));
                ce = cudaGraphInstantiate(&p->l2_graph, graph, 0);
                cudaGraphDestroy(graph);
                if (ce != cudaSuccess)
                    return ssg_set_error(SSG_ERR_CUDA, "embed: cudaGraphInstantiate: %s", cudaGetErrorString(ce));
                p->l2_graph_nb = NB; p->l2_graph_chunk = l2_chunk;
            }
            SSG_CUDA_TRY(cudaEventRecord(p->ev_in, st));
            SSG_CUDA_TRY(cudaStreamWaitEvent(p->side, p->ev_in, 0));
            { SSG_PROF("conv_l2_graph", p->side); SSG_CUDA_TRY(cudaGraphLaunch(p->l2_graph, p->side)); }
            SSG_CUDA_TRY(cudaEventRecord(p->ev_out, p->side));
            SSG_CUDA_TRY(cudaStreamWaitEvent(st, p->ev_out, 0));
        } else {
            SSG_TRY(chunk_loop(st));
        }
        int lc = 24, H = 32, W = 16, C = 512;                  // layer 3 starts at layer index 1 + 10 + 13
        void *x = l2_all, *y = p->y;
        for (int L = 2; L < 4; ++L)
            for (int b = 0; b < blocks[L]; ++b) {
                SSG_TRY(run_block(p, L, b, NB, fuse_ds, x, y, H, W, C, lc, nullptr, st));
                if (L == 2 && b == 0) y = p->x;               // keep the im2col buffer out of the ping-pong
            }
        { SSG_PROF("pooled_tail", st); SSG_TRY(pooled_tail(x, n, num_split, eval_mode, flip, d_feat, bank_stride, row0, st)); }
        return SSG_OK;
    }
    if (p->stem_windows) {
        {
            SSG_PROF("stem_prep", st);
            if (d_u8) SSG_TRY(stem_prep_u8(d_u8, n, flip, mean, stdv, p->stemP, st));
            else SSG_TRY(stem_prep(d_images, n, flip, p->stemP, st));
        }
        SSG_PROF("conv_stem_tc", st);
        pooled = p->stem_windows == 2 && stem_pool_fused();
        if (p->stem_windows == 2)
            SSG_TRY(conv_stem_windows64(p->stemP, NB, p->w_stem256, p->b_stem448, p->stem, st, pooled ? p->x : nullptr));
        else SSG_TRY(conv_stem_windows(p->stemP, NB, p->w_stem448, p->b_stem448, p->stem, st));
    } else {
        { SSG_PROF("stem_im2col", st); SSG_TRY(stem_im2col(d_images, n, flip, p->col, st)); }
        { SSG_PROF("conv_stem_tc", st); SSG_TRY(conv1x1(p->col, NB * 8192, 192, p->w[0], p->b[0], 64, nullptr, 1, p->stem, st)); }
    }
    if (!pooled) { SSG_PROF("maxpool", st); SSG_TRY(maxpool3x3s2(p->stem, NB, 128, 64, 64, p->x, st)); }
    li = 1;
    int H = 64, W = 32, C = 64;
    void *x = p->x, *y = p->y;
    const int blocks[4] = {3, 4, 6, 3};
    bool t1_ready = false;
    for (int L = 0; L < 4; ++L)
        for (int b = 0; b < blocks[L]; ++b)
            SSG_TRY(run_block(p, L, b, NB, fuse_ds, x, y, H, W, C, li, nullptr, st, &t1_ready));
    (void)sp;
    { SSG_PROF("pooled_tail", st); SSG_TRY(pooled_tail(x, n, num_split, eval_mode, flip, d_feat, bank_stride, row0, st)); }
    return SSG_OK;
}

extern "C" int ssg_embed_forward(ssg_embed_plan* p, const float* d_images, int n, int num_split, int eval_mode,
                                 int flip, float* d_feat, size_t bank_stride, int row0, void* stream) {
    if (!d_images) return ssg_set_error(SSG_ERR_INVALID, "embed_forward: null images");
    return embed_forward_impl(p, d_images, nullptr, nullptr, nullptr, n, num_split, eval_mode, flip, d_feat,
                              bank_stride, row0, stream);
}

extern "C" int ssg_embed_forward_u8(ssg_embed_plan* p, const uint8_t* d_images_u8, const float* h_mean,
                                    const float* h_std, int n, int num_split, int eval_mode, int flip, float* d_feat,
                                    size_t bank_stride, int row0, void* stream) {
    if (!d_images_u8 || !h_mean || !h_std) return ssg_set_error(SSG_ERR_INVALID, "embed_forward_u8: null argument");
    for (int c = 0; c < 3; ++c)
        if (!(h_std[c] != 0.f)) return ssg_set_error(SSG_ERR_INVALID, "embed_forward_u8: std[%d] is zero", c);
    return embed_forward_impl(p, nullptr, d_images_u8, h_mean, h_std, n, num_split, eval_mode, flip, d_feat,
                              bank_stride, row0, stream);
}

// ---------------------------------------------------------------------------------------- building blocks
// The individual operators of the trunk, exported so that each can be parity-tested in isolation
// (tests/test_gpu_embed.py) and reused by callers that bring their own graph.  NHWC bf16 activations.
extern "C" int ssg_op_conv(const void* d_x, int B, int H, int W, int cin, int ksize, int stride, const void* d_w,
                           const float* d_bias, int cout, const void* d_res, int relu, void* d_y, void* d_scratch,
                           void* stream) {
    if (!d_x || !d_w || !d_bias || !d_y || (stride != 1 && stride != 2) || (ksize != 1 && ksize != 3))
        return ssg_set_error(SSG_ERR_INVALID, "op_conv: bad arguments");
    if (stride == 2 && !d_scratch) return ssg_set_error(SSG_ERR_INVALID, "op_conv: stride 2 needs scratch");
    cudaStream_t st = (cudaStream_t)stream;
    if (ksize == 1) {
        if (stride == 1) return conv1x1(d_x, B * H * W, cin, d_w, d_bias, cout, d_res, relu, d_y, st);
        if (s2_strided_tma() && !d_res) return conv1x1_s2(d_x, B, H / 2, W / 2, cin, d_w, d_bias, cout, relu, d_y, st);
        SSG_TRY(parity_split(d_x, B, H, W, cin, 1, d_scratch, st));
        return conv1x1(d_scratch, B * (H / 2) * (W / 2), cin, d_w, d_bias, cout, d_res, relu, d_y, st);
    }
    if (d_res) return ssg_set_error(SSG_ERR_INVALID, "op_conv: residual is only fused into 1x1 convolutions");
    if (stride == 1) return conv3x3(d_x, B, H, W, cin, 1, d_w, d_bias, cout, relu, d_y, st);
    if (s2_strided_tma()) return conv3x3(d_x, B, H / 2, W / 2, cin, 2, d_w, d_bias, cout, relu, d_y, st);
    SSG_TRY(parity_split(d_x, B, H, W, cin, 4, d_scratch, st));
    return conv3x3(d_scratch, B, H / 2, W / 2, cin, 2, d_w, d_bias, cout, relu, d_y, st);
}

extern "C" int ssg_op_fold_bn(const float* d_w, int cout, int cin, int ksize, const float* d_gamma,
                              const float* d_beta, const float* d_mean, const float* d_var, float eps, int kpad,
                              void* d_wout, float* d_bout, void* stream) {
    if (!d_w || !d_wout || !d_bout || kpad < ksize * ksize * cin)
        return ssg_set_error(SSG_ERR_INVALID, "op_fold_bn: bad arguments");
    return fold_bn(d_w, cout, cin, ksize, ksize, d_gamma, d_beta, d_mean, d_var, eps, kpad, d_wout, d_bout,
                   (cudaStream_t)stream);
}

extern "C" int ssg_op_stem(const float* d_images, int n, int flip, const void* d_w, const float* d_bias,
                           void* d_col, void* d_conv_out, void* d_pool_out, void* stream) {
    if (!d_images || !d_w || !d_bias || !d_col || !d_conv_out || n <= 0)
        return ssg_set_error(SSG_ERR_INVALID, "op_stem: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int NB = flip ? 2 * n : n;
    SSG_TRY(stem_im2col(d_images, n, flip, d_col, st));
    SSG_TRY(conv1x1(d_col, NB * 8192, 192, d_w, d_bias, 64, nullptr, 1, d_conv_out, st));
    if (d_pool_out) SSG_TRY(maxpool3x3s2(d_conv_out, NB, 128, 64, 64, d_pool_out, st));
    return SSG_OK;
}

extern "C" int ssg_op_pooled_tail(const void* d_x, int n, int num_split, int eval_mode, int flip, float* d_feat,
                                  size_t bank_stride, int row0, void* stream) {
    if (!d_x || !d_feat || n <= 0) return ssg_set_error(SSG_ERR_INVALID, "op_pooled_tail: bad arguments");
    return pooled_tail(d_x, n, num_split, eval_mode, flip, d_feat, bank_stride, row0, (cudaStream_t)stream);
}
