// Training-side convolution operators of the fine-tune step: what loss.backward() of
// reid/trainers.py:204-271 (FinedTrainer2.train / _forward) asks of the ResNet-50 convolutions -- the data gradient and
// the weight gradient -- on the same tcgen05 GEMM kernels as the forward path (gemm_tc.cuh).  NHWC bf16 activations and
// gradients, fp32 master weights [cout, cin, k, k] as torch holds them, fp32 weight gradients.
//
//   forward  : ssg_op_conv with the weights packed by ssg_op_conv_pack_weight(transposed = 0) and a zero bias (the
//              BatchNorm that follows runs on batch statistics in training mode and cannot be folded).
//   dgrad    : dx = conv(dy, W') with W'[ci][kh][kw][co] = W[co][ci][k-1-kh][k-1-kw] -- the SAME implicit-GEMM kernels
//              with the roles of the channel axes swapped.  Stride 2: dy is first spread onto the even positions of a
//              zeroed map of the input's size (3x3), or the 1x1 result is (1x1: the gradient only reaches the even pixels).
//   wgrad    : dW[co][(kh,kw,ci)] = sum over the B*Ho*Wo output pixels m of dy[m][co] * x[pixel(m)+tap][ci]: a GEMM
//              whose K axis is the PIXEL axis.  Both operands are re-laid K-major by plain kernels (dy^T, and the
//              transposed im2col of x); the output is tiny (cout x k*k*cin) and K is huge, so the K range is split into
//              S batches stacked along M ([S][cout_pad][Kc] A operand, AOperand::ksplit_mblks) -- one launch of the
//              persistent kernel fills all SMs -- and the S fp32 partial products are summed in a fixed order.
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "conv.h"
#include "gemm_tc.cuh"

namespace ssg {

// ---- weight packing -----------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int cout, int cin, int k, int transposed,
                                   __nv_bfloat16* __restrict__ out) {
    const size_t total = (size_t)cout * cin * k * k;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        // i indexes the OUTPUT: [rows][kh][kw][cols]
        const int cols = transposed ? cout : cin, rows_ = transposed ? cin : cout;
        const int c = (int)(i % cols);
        const int t = (int)((i / cols) % (k * k));
        const int r = (int)(i / ((size_t)cols * k * k));
        (void)rows_;
        const int kh = t / k, kw = t % k;
        float v;
        if (!transposed) v = w[(((size_t)r * cin + c) * k + kh) * k + kw];                       // r = co, c = ci
        else v = w[(((size_t)c * cin + r) * k + (k - 1 - kh)) * k + (k - 1 - kw)];               // r = ci, c = co
        out[i] = __float2bfloat16_rn(v);
    }
}

// ---- stride-2 helpers -----------------------------------------------------------------------------------------
// y [B, 2H, 2W, C]: y[b, 2h, 2w, :] = x[b, h, w, :], zero elsewhere (16-byte chunks; C % 8 == 0)
__global__ void dilate2_kernel(const uint4* __restrict__ x, int B, int H, int W, int c8, uint4* __restrict__ y) {
    const size_t total = (size_t)B * 2 * H * 2 * W * c8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % c8);
        size_t p = i / c8;
        const int w2 = (int)(p % (2 * W)); p /= 2 * W;
        const int h2 = (int)(p % (2 * H));
        const int b = (int)(p / (2 * H));
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (!(w2 & 1) && !(h2 & 1)) v = x[(((size_t)b * H + (h2 >> 1)) * W + (w2 >> 1)) * c8 + c];
        y[i] = v;
    }
}

// ---- K-major re-layouts for the weight-gradient GEMM ------------------------------------------------------------
// out [S][cpad][kc] (one [S*cpad, kc] matrix): out[s][c][q] = dy[s*kc + q][c]  (0 for pixels >= m or channels >= C)
__global__ void __launch_bounds__(256)
transpose_split_kernel(const __nv_bfloat16* __restrict__ dy, int m, int C, int cpad, int kc,
                       __nv_bfloat16* __restrict__ out) {
    __shared__ __nv_bfloat16 tile[32][33];
    const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int mm = m0 + r, c = c0 + tx;
        tile[r][tx] = (mm < m && c < C) ? dy[(size_t)mm * C + c] : __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    const int s = m0 / kc, q0 = m0 - s * kc;            // kc is a multiple of 64: a tile never straddles two splits
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        if (c < cpad) out[((size_t)s * cpad + c) * kc + q0 + tx] = tile[tx][r];
    }
}

// out [k*k*cin][ktot]: out[(kh*k + kw)*cin + ci][mm] = x[b, oh*stride + kh - pad, ow*stride + kw - pad, ci] for the output
// pixel mm = (b, oh, ow) (0 in the halo and for mm >= B*Ho*Wo)
__global__ void __launch_bounds__(256)
im2col_t_kernel(const __nv_bfloat16* __restrict__ x, int B, int H, int W, int cin, int k, int stride, int pad, int Ho,
                int Wo, int ktot, __nv_bfloat16* __restrict__ out) {
    __shared__ __nv_bfloat16 tile[32][33];
    const int m0 = blockIdx.x * 32, c0 = blockIdx.y * 32, tap = blockIdx.z;
    const int kh = tap / k, kw = tap - kh * k;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int m = B * Ho * Wo;
    for (int r = ty; r < 32; r += 8) {
        const int mm = m0 + r, c = c0 + tx;
        __nv_bfloat16 v = __float2bfloat16_rn(0.f);
        if (mm < m && c < cin) {
            const int ow = mm % Wo, oh = (mm / Wo) % Ho, b = mm / (Wo * Ho);
            const int ih = oh * stride + kh - pad, iw = ow * stride + kw - pad;
            if (ih >= 0 && ih < H && iw >= 0 && iw < W) v = x[(((size_t)b * H + ih) * W + iw) * cin + c];
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r;
        if (c < cin && m0 + tx < ktot) out[((size_t)tap * cin + c) * ktot + m0 + tx] = tile[tx][r];
    }
}

// dw [cout][cin][k][k] = sum over the S partial products part[s][co][(kh*k + kw)*cin + ci], s ascending
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int S, int cpad, int cout, int cin, int k,
                                    float* __restrict__ dw) {
    const int n = k * k * cin;
    const size_t total = (size_t)cout * n;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int col = (int)(i % n), co = (int)(i / n);
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += part[((size_t)s * cpad + co) * n + col];
        const int ci = col % cin, tap = col / cin;
        dw[((size_t)co * cin + ci) * k * k + tap] = acc;
    }
}

// epilogue of the split-K GEMM: the fp32 accumulator chunk as it is
struct EpiStoreF32 {
    float* out;        // [M, ldc]
    size_t ldc;
    static constexpr bool kSkippable = false;
    __device__ __forceinline__ void operator()(int row, int col0, int ncols, const uint32_t (&acc)[32]) const {
        float* o = out + (size_t)row * ldc + col0;
        if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(o + 4 * q) =
                    make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]), __uint_as_float(acc[4 * q + 2]),
                                __uint_as_float(acc[4 * q + 3]));
        } else {
            for (int q = 0; q < ncols; ++q) o[q] = __uint_as_float(acc[q]);
        }
    }
};

static inline int grid_for(size_t total, int threads) {
    size_t g = (total + threads - 1) / threads;
    return (int)(g < 1 ? 1 : (g > 65535u * 16u ? 65535u * 16u : g));
}

// scratch memory of one operator call, released on the stream when the holder goes out of scope
struct Scratch {
    cudaStream_t st;
    void* p[6];
    int n;
    explicit Scratch(cudaStream_t s) : st(s), n(0) {}
    ~Scratch() { for (int i = 0; i < n; ++i) cudaFreeAsync(p[i], st); }
    int get(void** out, size_t bytes) {
        SSG_CUDA_TRY(cudaMallocAsync(out, bytes < 16 ? 16 : bytes, st));
        p[n++] = *out;
        return SSG_OK;
    }
};

static int conv_dgrad(const void* dy, int B, int H, int W, int cout, int k, int stride, const float* w, int cin, void* dx,
                      cudaStream_t st) {
    if ((k != 1 && k != 3) || (stride != 1 && stride != 2) || cin % 64 || cout % 64 || (stride == 2 && ((H | W) & 1)))
        return ssg_set_error(SSG_ERR_INVALID, "conv_dgrad: k=%d stride=%d cin=%d cout=%d map %dx%d not supported", k, stride,
                             cin, cout, H, W);
    Scratch sc(st);
    void* wt = nullptr;
    float* zero = nullptr;
    SSG_TRY(sc.get(&wt, (size_t)cin * k * k * cout * 2));
    SSG_TRY(sc.get((void**)&zero, sizeof(float) * cin));
    SSG_CUDA_TRY(cudaMemsetAsync(zero, 0, sizeof(float) * cin, st));
    const size_t wn = (size_t)cin * k * k * cout;
    pack_weight_kernel<<<grid_for(wn, 256), 256, 0, st>>>(w, cout, cin, k, 1, (__nv_bfloat16*)wt);
    SSG_CHECK_LAUNCH();
    const int Ho = H / stride, Wo = W / stride;
    if (stride == 1) {
        if (k == 1) return conv1x1(dy, B * H * W, cout, wt, zero, cin, nullptr, 0, dx, st);
        return conv3x3(dy, B, H, W, cout, 1, wt, zero, cin, 0, dx, st);
    }
    void* tmp = nullptr;
    if (k == 3) {
        // dy spread onto the even positions of an [H, W] map, then the stride-1 convolution with the mirrored taps
        SSG_TRY(sc.get(&tmp, (size_t)B * H * W * cout * 2));
        const size_t total = (size_t)B * H * W * (cout / 8);
        dilate2_kernel<<<grid_for(total, 256), 256, 0, st>>>((const uint4*)dy, B, Ho, Wo, cout / 8, (uint4*)tmp);
        SSG_CHECK_LAUNCH();
        return conv3x3(tmp, B, H, W, cout, 1, wt, zero, cin, 0, dx, st);
    }
    // 1x1 stride 2 reads the even pixels only: their gradient is dy * W, every other pixel gets zero
    SSG_TRY(sc.get(&tmp, (size_t)B * Ho * Wo * cin * 2));
    SSG_TRY(conv1x1(dy, B * Ho * Wo, cout, wt, zero, cin, nullptr, 0, tmp, st));
    const size_t total = (size_t)B * H * W * (cin / 8);
    dilate2_kernel<<<grid_for(total, 256), 256, 0, st>>>((const uint4*)tmp, B, Ho, Wo, cin / 8, (uint4*)dx);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

static int conv_wgrad(const void* x, int B, int H, int W, int cin, const void* dy, int cout, int k, int stride, float* dw,
                      cudaStream_t st) {
    if ((k != 1 && k != 3) || (stride != 1 && stride != 2) || cin < 1 || cout < 1 || (stride == 2 && ((H | W) & 1)))
        return ssg_set_error(SSG_ERR_INVALID, "conv_wgrad: k=%d stride=%d cin=%d cout=%d map %dx%d not supported", k, stride,
                             cin, cout, H, W);
    const int pad = k / 2, Ho = H / stride, Wo = W / stride;
    const long long m_ll = (long long)B * Ho * Wo;
    if (m_ll < 1 || m_ll > (1ll << 30)) return ssg_set_error(SSG_ERR_INVALID, "conv_wgrad: %lld output pixels", m_ll);
    const int m = (int)m_ll, n = k * k * cin;
    const int cpad = ssg_cdiv(cout, tc::BM) * tc::BM;
    int sms = 0;
    SSG_TRY(tc_num_sms(&sms));
    const int bn = n <= 64 ? 64 : (n <= 128 ? 128 : 256);
    const int tiles_per_split = (cpad / tc::BM) * ssg_cdiv(n, bn);
    const int kblocks = ssg_cdiv(m, tc::BK);
    int S = sms / tiles_per_split;                       // fill the SMs once ...
    if (S > kblocks / 4) S = kblocks / 4;                // ... with at least four K blocks per split
    if (S < 1) S = 1;
    static int force_s = -1;                             // SSG_WGRAD_SPLITS=<S>: tuning / test switch
    if (force_s < 0) { const char* e = getenv("SSG_WGRAD_SPLITS"); force_s = e ? atoi(e) : 0; }
    if (force_s > 0) S = force_s < kblocks ? force_s : kblocks;
    const int kc = ssg_cdiv(kblocks, S) * tc::BK;
    S = ssg_cdiv(kblocks * tc::BK, kc);
    const size_t ktot = (size_t)S * kc;
    Scratch sc(st);
    void *dyt = nullptr, *xt = nullptr;
    float* part = nullptr;
    SSG_TRY(sc.get(&dyt, (size_t)S * cpad * kc * 2));
    SSG_TRY(sc.get(&xt, (size_t)n * ktot * 2));
    SSG_TRY(sc.get((void**)&part, (size_t)S * cpad * n * sizeof(float)));
    transpose_split_kernel<<<dim3((unsigned)(ktot / 32), (unsigned)(cpad / 32)), 256, 0, st>>>(
        (const __nv_bfloat16*)dy, m, cout, cpad, kc, (__nv_bfloat16*)dyt);
    SSG_CHECK_LAUNCH();
    im2col_t_kernel<<<dim3((unsigned)(ktot / 32), (unsigned)ssg_cdiv(cin, 32), (unsigned)(k * k)), 256, 0, st>>>(
        (const __nv_bfloat16*)x, B, H, W, cin, k, stride, pad, Ho, Wo, (int)ktot, (__nv_bfloat16*)xt);
    SSG_CHECK_LAUNCH();
    tc::AOperand A;
    memset(&A, 0, sizeof(A));
    A.mode = 0;
    A.cblks = kc / tc::BK;
    A.taps = 1;
    A.tiles_per_img = 1;
    A.hmul = 1;
    A.ksplit_mblks = cpad / tc::BM;
    const int mrows = S * cpad;
    SSG_TRY(make_tmap_2d_bf16(&A.map[0], dyt, (uint64_t)mrows, (uint64_t)kc, (uint64_t)kc, tc::BM));
    EpiStoreF32 epi{part, (size_t)n};
    if (bn == 64) SSG_TRY((tc::launch_gemm_op<64, EpiStoreF32, false>(A, mrows, xt, n, kc, epi, st, (int)ktot)));
    else if (bn == 128) SSG_TRY((tc::launch_gemm_op<128, EpiStoreF32, false>(A, mrows, xt, n, kc, epi, st, (int)ktot)));
    else SSG_TRY((tc::launch_gemm_op<256, EpiStoreF32, false>(A, mrows, xt, n, kc, epi, st, (int)ktot)));
    wgrad_reduce_kernel<<<grid_for((size_t)cout * n, 256), 256, 0, st>>>(part, S, cpad, cout, cin, k, dw);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

}  // namespace ssg

using namespace ssg;

extern "C" int ssg_op_conv_pack_weight(const float* d_w, int cout, int cin, int ksize, int transposed, void* d_out,
                                       void* stream) {
    if (!d_w || !d_out || cout < 1 || cin < 1 || ksize < 1)
        return ssg_set_error(SSG_ERR_INVALID, "op_conv_pack_weight: bad arguments");
    const size_t total = (size_t)cout * cin * ksize * ksize;
    pack_weight_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(d_w, cout, cin, ksize, transposed ? 1 : 0,
                                                                               (__nv_bfloat16*)d_out);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

extern "C" int ssg_op_conv_dgrad(const void* d_dy, int B, int H, int W, int cout, int ksize, int stride, const float* d_w,
                                 int cin, void* d_dx, void* stream) {
    if (!d_dy || !d_w || !d_dx || B < 1 || H < 1 || W < 1) return ssg_set_error(SSG_ERR_INVALID, "op_conv_dgrad: bad arguments");
    return conv_dgrad(d_dy, B, H, W, cout, ksize, stride, d_w, cin, d_dx, (cudaStream_t)stream);
}

extern "C" int ssg_op_conv_wgrad(const void* d_x, int B, int H, int W, int cin, const void* d_dy, int cout, int ksize,
                                 int stride, float* d_dw, void* stream) {
    if (!d_x || !d_dy || !d_dw || B < 1 || H < 1 || W < 1) return ssg_set_error(SSG_ERR_INVALID, "op_conv_wgrad: bad arguments");
    return conv_wgrad(d_x, B, H, W, cin, d_dy, cout, ksize, stride, d_dw, (cudaStream_t)stream);
}

extern "C" int ssg_op_stem_im2col(const float* d_images, int n, int flip, void* d_col, void* stream) {
    if (!d_images || !d_col || n < 1) return ssg_set_error(SSG_ERR_INVALID, "op_stem_im2col: bad arguments");
    return stem_im2col(d_images, n, flip, d_col, (cudaStream_t)stream);
}
