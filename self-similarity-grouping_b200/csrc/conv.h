// Convolution-path launchers (conv.cu), used by embed.cu.  All activations are NHWC bf16 on the device.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace ssg {
int conv1x1(const void* x, int m, int cin, const void* w, const float* bias, int cout, const void* residual,
            int relu, void* y, cudaStream_t st);
bool s2_strided_tma();
int conv1x1_s2(const void* x, int B, int H, int W, int cin, const void* w, const float* bias, int cout, int relu,
               void* y, cudaStream_t st);
int conv_fused_ds(const void* t2, const void* x, int B, int H, int W, int mid, int cin, int stride, const void* w,
                  const float* bias, int cout, void* y, cudaStream_t st);
int conv3x3(const void* x, int B, int H, int W, int cin, int stride, const void* w, const float* bias, int cout,
            int relu, void* y, cudaStream_t st);
int vec_add_f32(const float* a, const float* b, int n, float* out, cudaStream_t st);
int fold_bn(const float* w, int cout, int cin, int kh, int kw, const float* gamma, const float* beta,
            const float* mean, const float* var, float eps, int kpad, void* wout, float* bout, cudaStream_t st);
int stem_prep(const float* img, int n, int flip_too, void* P, cudaStream_t st);
int stem_prep_u8(const uint8_t* img, int n, int flip_too, const float* mean, const float* std, void* P,
                 cudaStream_t st);
int fold_bn_stem(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                 void* wout, float* bout, void* wout64, cudaStream_t st);
int conv_stem_windows64(const void* P, int images, const void* w256, const float* bias, void* y, cudaStream_t st,
                        void* pool_out = nullptr);
bool stem_pool_fused();
int conv_chain(const void* t2, int B, int H, int W, int mid, const void* x_ds, int cin_ds, const void* w,
               const float* bias, const void* residual, void* y, const void* w_next, const float* bias_next, int n2,
               void* t1_next, cudaStream_t st);
bool conv_chain_enabled();
int conv_stem_windows(const void* P, int images, const void* w448, const float* bias, void* y, cudaStream_t st);
int stem_im2col(const float* img, int n, int flip_too, void* out, cudaStream_t st);
int maxpool3x3s2(const void* x, int B, int H, int W, int C, void* y, cudaStream_t st);
int parity_split(const void* x, int B, int H, int W, int C, int nplanes, void* y, cudaStream_t st);
int pooled_tail(const void* x, int n, int num_split, int eval_mode, int flip_too, float* feat, size_t bank_stride,
                int row0, cudaStream_t st);
}  // namespace ssg
