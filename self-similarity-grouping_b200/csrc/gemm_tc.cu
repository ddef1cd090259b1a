// Persistent, warp-specialised tcgen05 GEMM for sm_100a:   C[M,N] = A[M,K] * B[N,K]^T
//   A, B : bf16, K-major (row-major with K contiguous), fed by TMA into 128B-swizzled shared memory
//   acc  : fp32 in TMEM (two accumulator stages so that the epilogue of tile t overlaps the MMAs of t+1)
//   roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM allocation), warps 2..5 = epilogue
// Epilogues (template): squared-distance (norms - 2*acc -> fp32) for the re-ranking path, and
// bias(+residual)(+ReLU) -> bf16 NHWC for the convolution path (conv.cu instantiates those).
//
// The squared-distance path feeds it bf16x3 split operands (A' = [hi|hi|lo], B' = [hi|lo|hi], K = 3d) so that
// the fp32 features are represented to ~2^-17 relative: the result is an approximation (|err| ~1e-5) that is
// only used to pick candidates; api.cu re-scores candidates exactly (DESIGN.md "tensor distance mode").
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"
#include "gemm_tc.cuh"

namespace ssg {

// ---------------------------------------------------------------------------------------------------
// host: cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency,
// so the library still loads on a machine without a GPU driver).
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode(PFN_encodeTiled* out) {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        SSG_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !p)
            return ssg_set_error(SSG_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled not available in this driver");
        fn = (PFN_encodeTiled)p;
    }
    *out = fn;
    return SSG_OK;
}

// 2-D bf16 tensor [rows, cols] (cols contiguous, row pitch `ld` elements), box = box_rows x 64 cols, SW128.
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows) {
    PFN_encodeTiled enc;
    SSG_TRY(get_encode(&enc));
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16)
        return ssg_set_error(SSG_ERR_INVALID, "tensor map: base/pitch must be 16-byte aligned");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ssg_set_error(SSG_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return SSG_OK;
}

// 4-D bf16 NHWC activation [B, H, W, C] seen as dims (C, W, H, B); box = (64, bw, bh, bb): one box is a
// [bb*bh*bw, 64] K-major tile of an implicit-GEMM A operand; out-of-bounds (halo) elements are zero-filled.
// `stride` (1 or 2) is the element stride of the W and H traversal: a stride-2 box of extent (2bw, 2bh) delivers
// every other pixel, i.e. the bw x bh input pixels a stride-2 convolution tap needs — no parity-split copy.
int make_tmap_nhwc_bf16(CUtensorMap* map, const void* base, uint64_t B, uint64_t H, uint64_t W, uint64_t C,
                        uint32_t bw, uint32_t bh, uint32_t bb, uint32_t stride) {
    PFN_encodeTiled enc;
    SSG_TRY(get_encode(&enc));
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (C * 2) % 16)
        return ssg_set_error(SSG_ERR_INVALID, "tensor map: NHWC base/channel pitch must be 16-byte aligned");
    cuuint64_t dims[4] = {C, W, H, B};
    cuuint64_t strides[3] = {C * 2, W * C * 2, H * W * C * 2};
    cuuint32_t box[4] = {64, bw * stride, bh * stride, bb};
    cuuint32_t estr[4] = {1, stride, stride, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ssg_set_error(SSG_ERR_CUDA, "cuTensorMapEncodeTiled(4d) failed (%d)", (int)r);
    return SSG_OK;
}

// Sliding-window view of the padded 4-channel stem input P [images][256 rows][144 px][4 ch] (bf16): dim0 = 64 elements
// (16 px x 4 ch) of one window, dim1 = output column (window start advances 2 px = 16 bytes: the windows OVERLAP),
// dim2 = input row (element stride 2 = one box row per output row), dim3 = image.  One box is the [2 x 64, 64] A tile
// of one kernel row of the 7x7/2 convolution: no im2col buffer.
int make_tmap_stem_windows(CUtensorMap* map, const void* base, uint64_t images) {
    PFN_encodeTiled enc;
    SSG_TRY(get_encode(&enc));
    cuuint64_t dims[4] = {64, 64, 256, images};
    cuuint64_t strides[3] = {16, 144 * 8, (cuuint64_t)256 * 144 * 8};
    cuuint32_t box[4] = {64, 64, 4, 1};
    cuuint32_t estr[4] = {1, 1, 2, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return ssg_set_error(SSG_ERR_UNSUPPORTED, "overlapping-window tensor map rejected by the driver (%d)", (int)r);
    return SSG_OK;
}

// 64-byte-swizzle variants for the stem: 8-pixel windows (32 elements = 64 B per row) halve the L2 traffic of the
// overlapping-window view; the weight matrix is cut into matching [rows, 32] boxes.
// box_rows = 4: two input rows (r, r+2) per box, one kernel row of a two-row output tile; box_rows = 10: five rows
// (r, r+2, .., r+8), one parity plane of a whole tile (VAR_BRESP)
int make_tmap_stem_windows64(CUtensorMap* map, const void* base, uint64_t images, uint32_t box_rows) {
    PFN_encodeTiled enc;
    SSG_TRY(get_encode(&enc));
    cuuint64_t dims[4] = {32, 64, 256, images};
    cuuint64_t strides[3] = {16, 144 * 8, (cuuint64_t)256 * 144 * 8};
    cuuint32_t box[4] = {32, 64, box_rows, 1};
    cuuint32_t estr[4] = {1, 1, 2, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return ssg_set_error(SSG_ERR_UNSUPPORTED, "overlapping-window (64B) tensor map rejected by the driver (%d)", (int)r);
    return SSG_OK;
}
int make_tmap_2d_bf16_sw64(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    PFN_encodeTiled enc;
    SSG_TRY(get_encode(&enc));
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ssg_set_error(SSG_ERR_CUDA, "cuTensorMapEncodeTiled(sw64) failed (%d)", (int)r);
    return SSG_OK;
}

bool pdl_enabled() {
    static int v = -1;
    // default off: bit-identical and measured at no gain on the B200 (627.0 vs 627.1 ms per cycle, profiles/r02j_ab_*.json:
    // a CTA holds > 200 KB of shared memory, so a dependent CTA can only start where a CTA of the previous kernel has left)
    if (v < 0) { const char* e = getenv("SSG_PDL"); v = e ? atoi(e) : 0; }
    return v != 0;
}

int tc_num_sms(int* out) {
    static int cached[64] = {0};
    int dev = 0;
    SSG_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && cached[dev]) { *out = cached[dev]; return SSG_OK; }
    int n = 0;
    SSG_CUDA_TRY(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    if (dev < 64) cached[dev] = n;
    *out = n;
    return SSG_OK;
}

// ---------------------------------------------------------------------------------------------------
// squared-distance front end
// ---------------------------------------------------------------------------------------------------
// x (fp32 [n,d]) -> bf16 hi/lo split laid out as [n, 3d]: which = 0: [hi|hi|lo] (A side), 1: [hi|lo|hi] (B side);
// also the squared norm (fp64 accumulate, rounded to fp32).
// column mean of x [n,d] (float64 partial sums over 64 interleaved row groups, deterministic)
constexpr int MEAN_GROUPS = 64;
__global__ void __launch_bounds__(256)
col_partial_kernel(const float* __restrict__ x, int n, int d, double* __restrict__ partial) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= d) return;
    double acc = 0.0;
    for (int r = blockIdx.y; r < n; r += MEAN_GROUPS) acc += (double)x[(size_t)r * d + c];
    partial[(size_t)blockIdx.y * d + c] = acc;
}
__global__ void __launch_bounds__(256)
col_mean_kernel(const double* __restrict__ partial, int n, int d, float* __restrict__ mean) {
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c >= d) return;
    double acc = 0.0;
    for (int g = 0; g < MEAN_GROUPS; ++g) acc += partial[(size_t)g * d + c];
    mean[c] = (float)(acc / (double)n);
}
int launch_col_mean(const float* x, int n, int d, double* partial, float* mean, cudaStream_t st) {
    dim3 grid(ssg_cdiv(d, 256), MEAN_GROUPS);
    col_partial_kernel<<<grid, 256, 0, st>>>(x, n, d, partial);
    SSG_CHECK_LAUNCH();
    col_mean_kernel<<<ssg_cdiv(d, 256), 256, 0, st>>>(partial, n, d, mean);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

// `centre` (optional, [d]) is subtracted first: distances are translation invariant, and centred operands make the
// approximation error proportional to the spread of the data instead of its absolute norm.
__global__ void __launch_bounds__(256)
split_bf16x3_kernel(const float* __restrict__ x, int n, int d, int which, const float* __restrict__ centre,
                    __nv_bfloat16* __restrict__ out, float* __restrict__ norm2) {
    const int i = blockIdx.x;
    const float* row = x + (size_t)i * d;
    __nv_bfloat16* o = out + (size_t)i * 3 * d;
    double acc = 0.0;
    for (int k = threadIdx.x; k < d; k += 256) {
        const float v = centre ? __fsub_rn(row[k], centre[k]) : row[k];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        o[k] = hi;
        o[d + k] = which == 0 ? hi : lo;
        o[2 * d + k] = which == 0 ? lo : hi;
        acc += (double)v * (double)v;
    }
    __shared__ double sh[256];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0 && norm2) norm2[i] = (float)sh[0];
}

int launch_split_bf16x3(const float* x, int n, int d, int which, const float* centre, void* out_bf16, float* norm2,
                        cudaStream_t st) {
    if (n <= 0) return SSG_OK;
    split_bf16x3_kernel<<<n, 256, 0, st>>>(x, n, d, which, centre, (__nv_bfloat16*)out_bf16, norm2);
    SSG_CHECK_LAUNCH();
    return SSG_OK;
}

struct EpiDist {
    const float* na;   // [M]
    const float* nb;   // [N]
    float* out;        // [M, ldc]
    size_t ldc;
    static constexpr bool kSkippable = false;
    __device__ __forceinline__ void operator()(int row, int col0, int ncols, const uint32_t (&acc)[32]) const {
        const float a = na[row];
        float* o = out + (size_t)row * ldc + col0;
        if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 b = *reinterpret_cast<const float4*>(nb + col0 + 4 * q);
                float4 r;
                r.x = fmaf(-2.0f, __uint_as_float(acc[4 * q + 0]), a + b.x);
                r.y = fmaf(-2.0f, __uint_as_float(acc[4 * q + 1]), a + b.y);
                r.z = fmaf(-2.0f, __uint_as_float(acc[4 * q + 2]), a + b.z);
                r.w = fmaf(-2.0f, __uint_as_float(acc[4 * q + 3]), a + b.w);
                *reinterpret_cast<float4*>(o + 4 * q) = r;
            }
        } else {
            for (int q = 0; q < ncols; ++q) o[q] = fmaf(-2.0f, __uint_as_float(acc[q]), a + nb[col0 + q]);
        }
    }
};

// Symmetric form (opt-in, SSG_DIST_SYM=1): A and B are the two splits of the SAME rows, M == N, and `out` holds the
// whole square matrix.  Tiles that lie entirely below the diagonal are not computed; element (i, j) with i <= j is
// stored by the tile that computes it, element (j, i) as its mirror -- every element is written exactly once, so the
// matrix is deterministic.  The mirrored value is A_i.B_j where a direct computation would give A_j.B_i: both are
// within the certified error of the true distance (api.cu, tensor_eps_rel), which is all the candidate selection
// relies on.
struct EpiDistSym {
    const float* na;   // [M] (== nb: the same rows)
    const float* nb;
    float* out;        // [M, ldc]
    size_t ldc;
    static constexpr bool kSkippable = true;
    __device__ __forceinline__ bool skip_tile(int m_blk, int n_blk, int bn) const {
        return (n_blk + 1) * bn <= m_blk * tc::BM;             // last column < first row
    }
    __device__ __forceinline__ void operator()(int row, int col0, int ncols, const uint32_t (&acc)[32]) const {
        const float a = na[row];
        float* o = out + (size_t)row * ldc + col0;
        const int r0 = row & ~31;                     // rows of this warp: [r0, r0 + 32); col0 is a multiple of 32
        if (col0 < r0) return;                        // below the diagonal: arrives as a mirror
        if (col0 > r0 && ncols == 32) {               // strictly above: direct row + mirrored column, unmasked
            float r[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) r[q] = fmaf(-2.0f, __uint_as_float(acc[q]), a + nb[col0 + q]);
            if ((reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    *reinterpret_cast<float4*>(o + 4 * q) = make_float4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
            } else {
#pragma unroll
                for (int q = 0; q < 32; ++q) o[q] = r[q];
            }
#pragma unroll
            for (int q = 0; q < 32; ++q) out[(size_t)(col0 + q) * ldc + row] = r[q];   // lanes = consecutive rows
            return;
        }
        for (int q = 0; q < ncols; ++q) {             // the diagonal chunk (or a ragged last chunk)
            const int j = col0 + q;
            const float r = fmaf(-2.0f, __uint_as_float(acc[q]), a + nb[j]);
            if (j >= row) o[q] = r;
            if (j > row) out[(size_t)j * ldc + row] = r;
        }
    }
};

// C = dist(A', B') with pre-split operands.  M x N output, K = 3d.
int launch_gemm_dist(const void* a_split, const float* na, int m, const void* b_split, const float* nb, int n,
                     int k, float* out, size_t ldc, cudaStream_t st, int sym) {
    if (sym) {
        if (m != n) return ssg_set_error(SSG_ERR_INVALID, "gemm_dist: the symmetric form needs M == N (%d, %d)", m, n);
        EpiDistSym es{na, nb, out, ldc};
        if (n < 256) return tc::launch_gemm<128, EpiDistSym>(a_split, m, b_split, n, k, es, st);
        return tc::launch_gemm<256, EpiDistSym>(a_split, m, b_split, n, k, es, st);
    }
    EpiDist epi{na, nb, out, ldc};
    // 128x256 tiles halve the shared-memory operand traffic per MMA (128x128 is smem-bandwidth bound); the
    // environment switch exists for A/B measurements only
    static int bn = 0;
    if (!bn) { const char* e = getenv("SSG_GEMM_BN"); bn = e ? atoi(e) : 256; }
    if (bn == 128 || n < 256) return tc::launch_gemm<128, EpiDist>(a_split, m, b_split, n, k, epi, st);
    return tc::launch_gemm<256, EpiDist>(a_split, m, b_split, n, k, epi, st);
}

// Stand-alone ssg_sqdist(mode = TENSOR): split both operands into scratch memory, run the GEMM.
int launch_sqdist_tensor(const float* X, int nx, const float* Y, int ny, int d, float* out, size_t ldo,
                         cudaStream_t st) {
    if (nx <= 0 || ny <= 0) return SSG_OK;
    if (d % 8) return ssg_set_error(SSG_ERR_INVALID, "sqdist tensor mode: d must be a multiple of 8");
    void *ax = nullptr, *by = nullptr;
    float *na = nullptr, *nb = nullptr;
    SSG_CUDA_TRY(cudaMallocAsync(&ax, (size_t)nx * 3 * d * 2, st));
    SSG_CUDA_TRY(cudaMallocAsync(&by, (size_t)ny * 3 * d * 2, st));
    SSG_CUDA_TRY(cudaMallocAsync((void**)&na, sizeof(float) * nx, st));
    SSG_CUDA_TRY(cudaMallocAsync((void**)&nb, sizeof(float) * ny, st));
    int rc = launch_split_bf16x3(X, nx, d, 0, nullptr, ax, na, st);
    if (rc == SSG_OK) rc = launch_split_bf16x3(Y, ny, d, 1, nullptr, by, nb, st);
    if (rc == SSG_OK) rc = launch_gemm_dist(ax, na, nx, by, nb, ny, 3 * d, out, ldo, st);
    cudaFreeAsync(ax, st);
    cudaFreeAsync(by, st);
    cudaFreeAsync(na, st);
    cudaFreeAsync(nb, st);
    return rc;
}

}  // namespace ssg
