// Internal launcher declarations of the fp32 SAM2 kernels (sam2_kernels.cu) used by the forward schedule (sam2.cu).
#pragma once
#include "ap_internal.cuh"

enum { SAM_ACT_NONE = 0, SAM_ACT_GELU = 1, SAM_ACT_RELU = 2 };

int sam_linear(ap_ctx* ctx, const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int M, int N, int K, int act,
               int accumulate, cudaStream_t st);
// window_partition fused into the linear's operand staging: row r of the (virtual) A matrix is token (y, x) of the [H, W, lda] map with
// r = ((wy * nWx + wx) * ws + iy) * ws + ix, y = wy * ws + iy, x = wx * ws + ix; rows outside the map are zero (padded windows).
struct SamGather { int H, W, ws, nWx; };
int sam_linear_windows(ap_ctx* ctx, const float* A, int lda, const SamGather& g, const float* W, const float* bias, float* C, int ldc, int M, int N,
                       int K, cudaStream_t st);
int sam_layernorm(ap_ctx* ctx, const float* x, const float* g, const float* b, float* y, int rows, int D, float eps, int act, cudaStream_t st);
int sam_patch_embed(ap_ctx* ctx, const uint8_t* img, int H, int W, const float* w, const float* bias, const float* pos, float* out, int C,
                    const float* mean, const float* stdv, cudaStream_t st);
int sam_patch_im2col(ap_ctx* ctx, const uint8_t* img, int H, int W, float* cols, const float* mean, const float* stdv, cudaStream_t st);
int sam_window_gather(ap_ctx* ctx, const float* x, float* win, int H, int W, int C, int ws, int nWy, int nWx, cudaStream_t st);
int sam_window_scatter_add(ap_ctx* ctx, const float* win, const float* res, float* out, int H, int W, int C, int ws, int nWx, cudaStream_t st);
int sam_maxpool2(ap_ctx* ctx, const float* x, int ld, float* y, int nB, int H, int W, int C, cudaStream_t st);
int sam_attention(ap_ctx* ctx, const float* q, int q_stride, const float* k, const float* v, int kv_stride, float* out, int out_stride, int nB,
                  int Lq, int Lk, int heads, int hd, float scale, cudaStream_t st);
int sam_add(ap_ctx* ctx, const float* a, const float* b, float* y, int64_t n, int D, int b_rows, cudaStream_t st);
int sam_pixel_shuffle_add(ap_ctx* ctx, const float* lin, const float* skip, float* dst, int H, int W, int Co, int act, cudaStream_t st);
int sam_upsample2_add(ap_ctx* ctx, const float* src, float* dst, int H, int W, int C, cudaStream_t st);
int sam_bilinear(ap_ctx* ctx, const float* src, int Hs, int Ws, float* dst, int Hd, int Wd, cudaStream_t st);
