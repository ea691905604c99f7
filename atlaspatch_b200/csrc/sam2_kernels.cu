// Building blocks of the SAM2 image path (a4).  The SAM2 forward runs once per slide on a 1024 x 1024 thumbnail (~210 GFLOP for
// Hiera-T, 1.6 TFLOP for Hiera-L, against ~900 TFLOP for embedding a slide).  Activations stay fp32 end to end; the linear layers
// (78 % of the time) run on the tensor cores with split-fp16 operands (fp32-like accuracy, see sam_linear_tc_kernel), attention,
// LayerNorm, pooling and resampling are fp32 SIMT kernels.  Porting the whole path to the tcgen05 GEMM / attention of the ViT
// encoder (fp16 activation buffers, padded 144/288/432-wide layers) is left for a later round.
// Layout everywhere: tokens x channels, row-major ("NHWC").
#include "ap_internal.cuh"
#include "ptx.cuh"
#include <unordered_map>

#include "sam2_internal.cuh"

namespace {

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

// ---------------------------------------------------------------------------------------------------------------------
// C[M,N] = act(A[M,K] W[N,K]^T + bias[N]) (+ C if accumulate).  64x64 tile, BK = 16, 256 threads, 4x4 micro-tile.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int LT = 64, LK = 16;
__global__ void __launch_bounds__(256)
sam_linear_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ C,
                  int ldc, int M, int N, int K, int act, int accumulate) {
    __shared__ float sA[LK][LT + 4];
    __shared__ float sW[LK][LT + 4];
    const int m0 = blockIdx.y * LT, n0 = blockIdx.x * LT;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += LK) {
        for (int i = threadIdx.x; i < LT * LK; i += 256) {
            const int r = i / LK, c = i - r * LK;
            const int k = k0 + c;
            sA[c][r] = (m0 + r < M && k < K) ? A[static_cast<int64_t>(m0 + r) * lda + k] : 0.f;
            sW[c][r] = (n0 + r < N && k < K) ? W[static_cast<int64_t>(n0 + r) * K + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = sA[k][ty * 4 + i]; b[i] = sW[k][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j] + (bias ? bias[n] : 0.f);
            if (act == SAM_ACT_GELU) v = gelu_exact(v);
            else if (act == SAM_ACT_RELU) v = fmaxf(v, 0.f);
            float* dst = C + static_cast<int64_t>(m) * ldc + n;
            *dst = accumulate ? *dst + v : v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same contract on the tensor cores: fp32 operands are rounded to fp16 while they are staged in shared memory, products are
// accumulated in fp32 (mma.sync m16n8k16).  128 x 128 x 32 tiles, 8 warps (2 x 4, 64 x 32 each), register prefetch of the next
// k-tile, two smem buffers (one __syncthreads per k-tile).  The Hiera linears are 78 % of the fp32 forward (ncu launch list in
// profiles/); this kernel takes every linear with M >= 64 and 4-element-aligned operands, the SIMT kernel above the rest.
// (The ViT encoder's tcgen05 GEMM needs fp16 activations, K % 64 == 0 and N % 128 == 0; Hiera's 144/288/432-wide layers and fp32
// activation buffers do not fit it without re-laying out the whole SAM2 path, which is left for a later round.)
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TCM = 128, TCN = 128, TCK = 32, TCP = 40;  // TCP: padded smem row (halfs): 80 B stride keeps ldmatrix conflict-free
constexpr float TC_WSCALE = 256.0f;  // weights are staged as 256 w so that the low half of the split stays a normal fp16 number
constexpr int TC_TILE_HALFS = TCM * TCP;
constexpr int TC_SMEM_BYTES = 2 * 4 * TC_TILE_HALFS * 2;  // 2 buffers x (A_hi, A_lo, W_hi, W_lo)

// SPLIT: every operand is staged as an fp16 pair x = hi + lo (hi = fp16(x), lo = fp16(x - hi), ~22 significant bits) and the
// product is three MMAs, A_hi W_hi + A_lo W_hi + A_hi W_lo (the dropped A_lo W_lo term is ~2^-22 relative): fp32-like accuracy on
// the tensor cores.  The randomly initialised Hiera of the parity oracle amplifies a per-layer error ~1000x over its 48 blocks,
// so plain fp16 operands (2.3e-4 per layer) come out at 4 % / 22 % logit error (Hiera-T / -L) and cannot be pinned; the split can.
template <bool SPLIT>
__global__ void __launch_bounds__(256)
sam_linear_tc_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ C,
                     int ldc, int M, int N, int K, int act, int accumulate) {
    extern __shared__ __align__(16) __half tc_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    const int m0 = blockIdx.y * TCM, n0 = blockIdx.x * TCN;
    float acc[4][4][4] = {};
    float4 ra[4], rw[4];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + 256 * i, r = idx >> 3, k = k0 + (idx & 7) * 4;
            ra[i] = (m0 + r < M && k < K) ? __ldg(reinterpret_cast<const float4*>(A + static_cast<int64_t>(m0 + r) * lda + k))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
            rw[i] = (n0 + r < N && k < K) ? __ldg(reinterpret_cast<const float4*>(W + static_cast<int64_t>(n0 + r) * K + k))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto put = [&](__half* hi_tile, __half* lo_tile, int off, float4 v, float scale) {
        const float x[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2half_rn(x[j]);
            l[j] = __float2half_rn(x[j] - __half2float(h[j]));
        }
        *reinterpret_cast<uint2*>(hi_tile + off) = *reinterpret_cast<uint2*>(h);
        if (SPLIT) *reinterpret_cast<uint2*>(lo_tile + off) = *reinterpret_cast<uint2*>(l);
    };
    auto store_tile = [&](int buf) {
        __half* base = tc_smem + buf * 4 * TC_TILE_HALFS;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int idx = tid + 256 * i, off = (idx >> 3) * TCP + (idx & 7) * 4;
            put(base, base + TC_TILE_HALFS, off, ra[i], 1.0f);
            put(base + 2 * TC_TILE_HALFS, base + 3 * TC_TILE_HALFS, off, rw[i], TC_WSCALE);
        }
    };
    const int k_tiles = (K + TCK - 1) / TCK;
    load_tile(0);
    for (int kt = 0; kt < k_tiles; ++kt) {
        const int buf = kt & 1;
        store_tile(buf);
        __syncthreads();
        if (kt + 1 < k_tiles) load_tile((kt + 1) * TCK);
        const uint32_t a_hi = ptx::smem_u32(tc_smem + buf * 4 * TC_TILE_HALFS), a_lo = a_hi + TC_TILE_HALFS * 2;
        const uint32_t w_hi = a_hi + 2 * TC_TILE_HALFS * 2, w_lo = a_hi + 3 * TC_TILE_HALFS * 2;
#pragma unroll
        for (int ks = 0; ks < TCK; ks += 16) {
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int np = 0; np < 2; ++np) {   // two 8-wide n tiles per ldmatrix.x4
                const int n = wn * 32 + np * 16 + (lane & 7) + ((lane >> 4) << 3), k = ks + ((lane >> 3) & 1) * 8;
                ptx::ldmatrix_x4(w_hi + (n * TCP + k) * 2, bh[2 * np][0], bh[2 * np][1], bh[2 * np + 1][0], bh[2 * np + 1][1]);
                if (SPLIT) ptx::ldmatrix_x4(w_lo + (n * TCP + k) * 2, bl[2 * np][0], bl[2 * np][1], bl[2 * np + 1][0], bl[2 * np + 1][1]);
            }
#pragma unroll
            for (int mi = 0; mi < 4; ++mi) {
                uint32_t a0, a1, a2, a3;
                const int r = wm * 64 + mi * 16 + (lane & 15), k = ks + (lane >> 4) * 8;
                ptx::ldmatrix_x4(a_hi + (r * TCP + k) * 2, a0, a1, a2, a3);
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) {
                    if (SPLIT) ptx::mma_m16n8k16_f16(acc[mi][ni], a0, a1, a2, a3, bl[ni][0], bl[ni][1]);   // small terms first
                }
                if (SPLIT) {
                    uint32_t l0, l1, l2, l3;
                    ptx::ldmatrix_x4(a_lo + (r * TCP + k) * 2, l0, l1, l2, l3);
#pragma unroll
                    for (int ni = 0; ni < 4; ++ni) ptx::mma_m16n8k16_f16(acc[mi][ni], l0, l1, l2, l3, bh[ni][0], bh[ni][1]);
                }
#pragma unroll
                for (int ni = 0; ni < 4; ++ni) ptx::mma_m16n8k16_f16(acc[mi][ni], a0, a1, a2, a3, bh[ni][0], bh[ni][1]);
            }
        }
    }
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 4; ++ni)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = m0 + wm * 64 + mi * 16 + (lane >> 2) + (e >> 1) * 8;
                const int n = n0 + wn * 32 + ni * 8 + (lane & 3) * 2 + (e & 1);
                if (m >= M || n >= N) continue;
                float v = acc[mi][ni][e] * (1.0f / TC_WSCALE) + (bias ? bias[n] : 0.f);
                if (act == SAM_ACT_GELU) v = gelu_exact(v);
                else if (act == SAM_ACT_RELU) v = fmaxf(v, 0.f);
                float* dst = C + static_cast<int64_t>(m) * ldc + n;
                *dst = accumulate ? *dst + v : v;
            }
}

// LayerNorm over the last dim (any D), one warp per row, optional GELU afterwards.
__global__ void __launch_bounds__(256)
sam_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b, float* __restrict__ y, int rows,
                     int D, float eps, int act) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + static_cast<int64_t>(row) * D;
    float s = 0.f;
    for (int i = lane; i < D; i += 32) s += xr[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / D;
    float q = 0.f;
    for (int i = lane; i < D; i += 32) { const float d = xr[i] - mean; q += d * d; }
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / D + eps);
    float* yr = y + static_cast<int64_t>(row) * D;
    for (int i = lane; i < D; i += 32) {
        float v = (xr[i] - mean) * rstd * g[i] + b[i];
        if (act == SAM_ACT_GELU) v = gelu_exact(v);
        yr[i] = v;
    }
}

// 7x7 stride-4 pad-3 patch embedding on a uint8 HWC image with the ImageNet normalisation applied on the fly, + positional
// embedding.  One thread per (output pixel, output channel).
__global__ void __launch_bounds__(256)
sam_patch_embed_kernel(const uint8_t* __restrict__ img, int H, int W, const float* __restrict__ w, const float* __restrict__ bias,
                       const float* __restrict__ pos, float* __restrict__ out, int C, float3 mean, float3 inv_std) {
    const int Ho = H / 4, Wo = W / 4;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(Ho) * Wo * C) return;
    const int c = static_cast<int>(idx % C);
    const int p = static_cast<int>(idx / C);
    const int oy = p / Wo, ox = p - oy * Wo;
    float acc = bias[c];
    const float mu[3] = {mean.x, mean.y, mean.z}, is[3] = {inv_std.x, inv_std.y, inv_std.z};
    for (int ky = 0; ky < 7; ++ky) {
        const int y = oy * 4 - 3 + ky;
        if (y < 0 || y >= H) continue;
        for (int kx = 0; kx < 7; ++kx) {
            const int x = ox * 4 - 3 + kx;
            if (x < 0 || x >= W) continue;
            const uint8_t* px = img + (static_cast<int64_t>(y) * W + x) * 3;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float v = (static_cast<float>(px[ci]) * (1.0f / 255.0f) - mu[ci]) * is[ci];
                acc = fmaf(v, w[((c * 3 + ci) * 7 + ky) * 7 + kx], acc);
            }
        }
    }
    out[idx] = acc + pos[idx];
}

// im2col of the patch-embed convolution (7 x 7, stride 4, padding 3 on the NORMALISED image: out-of-range taps are 0, not "black"):
// cols[(oy * Wo + ox)][(ci * 7 + ky) * 7 + kx], 148 columns per row (147 taps + one zero column: the tensor-core linears want K % 4 == 0).
// The convolution then runs as a tcgen05 GEMM with split operands like every other linear (1.75 ms -> 0.15 ms for 1024 x 1024).
__global__ void __launch_bounds__(256)
sam_patch_im2col_kernel(const uint8_t* __restrict__ img, int H, int W, float* __restrict__ cols, float3 mean, float3 inv_std) {
    const int Ho = H / 4, Wo = W / 4;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(Ho) * Wo * 148) return;
    const int k = static_cast<int>(idx % 148);
    const int p = static_cast<int>(idx / 148);
    float v = 0.f;
    if (k < 147) {
        const int ci = k / 49, r = k - ci * 49, ky = r / 7, kx = r - ky * 7;
        const int oy = p / Wo, ox = p - oy * Wo;
        const int y = oy * 4 - 3 + ky, x = ox * 4 - 3 + kx;
        if (y >= 0 && y < H && x >= 0 && x < W) {
            const float mu = ci == 0 ? mean.x : ci == 1 ? mean.y : mean.z, is = ci == 0 ? inv_std.x : ci == 1 ? inv_std.y : inv_std.z;
            v = (static_cast<float>(img[(static_cast<int64_t>(y) * W + x) * 3 + ci]) * (1.0f / 255.0f) - mu) * is;
        }
    }
    cols[idx] = v;
}

// [H, W, C] -> windows [nWy * nWx, ws * ws, C], zero padded at the bottom / right (window_partition).
__global__ void sam_window_gather_kernel(const float* __restrict__ x, float* __restrict__ win, int H, int W, int C, int ws, int nWy, int nWx) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int64_t total = static_cast<int64_t>(nWy) * nWx * ws * ws * C;
    if (idx >= total) return;
    const int c = static_cast<int>(idx % C);
    int64_t t = idx / C;
    const int ix = static_cast<int>(t % ws); t /= ws;
    const int iy = static_cast<int>(t % ws); t /= ws;
    const int wx = static_cast<int>(t % nWx);
    const int wy = static_cast<int>(t / nWx);
    const int y = wy * ws + iy, xx = wx * ws + ix;
    win[idx] = (y < H && xx < W) ? x[(static_cast<int64_t>(y) * W + xx) * C + c] : 0.f;
}

// out[H, W, C] = res[H, W, C] + windows (window_unpartition, padding dropped)
__global__ void sam_window_scatter_add_kernel(const float* __restrict__ win, const float* __restrict__ res, float* __restrict__ out, int H, int W,
                                              int C, int ws, int nWx) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(H) * W * C) return;
    const int c = static_cast<int>(idx % C);
    const int p = static_cast<int>(idx / C);
    const int y = p / W, x = p - y * W;
    const int wy = y / ws, wx = x / ws, iy = y - wy * ws, ix = x - wx * ws;
    out[idx] = res[idx] + win[((static_cast<int64_t>(wy) * nWx + wx) * ws * ws + iy * ws + ix) * C + c];
}

// 2x2 max pool over [nB, H, W, (row stride ld)] taking C channels -> [nB, H/2, W/2, C] dense
__global__ void sam_maxpool2_kernel(const float* __restrict__ x, int ld, float* __restrict__ y, int nB, int H, int W, int C) {
    const int Ho = H / 2, Wo = W / 2;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(nB) * Ho * Wo * C) return;
    const int c = static_cast<int>(idx % C);
    int64_t t = idx / C;
    const int ox = static_cast<int>(t % Wo); t /= Wo;
    const int oy = static_cast<int>(t % Ho);
    const int b = static_cast<int>(t / Ho);
    const float* base = x + (static_cast<int64_t>(b) * H * W) * ld + c;
    const float a0 = base[(static_cast<int64_t>(2 * oy) * W + 2 * ox) * ld], a1 = base[(static_cast<int64_t>(2 * oy) * W + 2 * ox + 1) * ld];
    const float a2 = base[(static_cast<int64_t>(2 * oy + 1) * W + 2 * ox) * ld], a3 = base[(static_cast<int64_t>(2 * oy + 1) * W + 2 * ox + 1) * ld];
    y[idx] = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
}

// ---------------------------------------------------------------------------------------------------------------------
// Attention, fp32, any head_dim <= 128, any Lq / Lk: CTA = 64 queries of one (batch, head); keys streamed in tiles of 64 with an
// online softmax.  q/k/v are addressed as base + (b * L + i) * tok_stride + h * hd.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int AQ = 64, AK = 64;
__global__ void __launch_bounds__(256)
sam_attention_kernel(const float* __restrict__ q, int q_stride, const float* __restrict__ k, const float* __restrict__ v, int kv_stride,
                     float* __restrict__ out, int out_stride, int Lq, int Lk, int heads, int hd, float scale) {
    extern __shared__ float smf[];
    const int hdp = hd + 1;
    float* sQ = smf;                 // [AQ][hdp]
    float* sK = sQ + AQ * hdp;       // [AK][hdp]
    float* sV = sK + AK * hdp;       // [AK][hdp]
    float* sS = sV + AK * hdp;       // [AQ][AK + 1]
    float* sM = sS + AQ * (AK + 1);  // [AQ] running max
    float* sL = sM + AQ;             // [AQ] running sum
    float* sAl = sL + AQ;            // [AQ] rescale factor of this tile
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AQ;
    const int tid = threadIdx.x;
    const float* qb = q + static_cast<int64_t>(b) * Lq * q_stride + h * hd;
    const float* kb = k + static_cast<int64_t>(b) * Lk * kv_stride + h * hd;
    const float* vb = v + static_cast<int64_t>(b) * Lk * kv_stride + h * hd;
    for (int i = tid; i < AQ * hd; i += 256) {
        const int r = i / hd, c = i - r * hd;
        sQ[r * hdp + c] = (q0 + r < Lq) ? qb[static_cast<int64_t>(q0 + r) * q_stride + c] * scale : 0.f;
    }
    if (tid < AQ) { sM[tid] = -INFINITY; sL[tid] = 0.f; }
    // O accumulators: thread (ty, tx) owns rows ty*4..+3 and columns tx + 16 j  (j < 8 -> hd <= 128)
    const int ty = tid >> 4, tx = tid & 15;
    float o[4][8] = {};
    __syncthreads();
    for (int k0 = 0; k0 < Lk; k0 += AK) {
        for (int i = tid; i < AK * hd; i += 256) {
            const int r = i / hd, c = i - r * hd;
            const bool ok = k0 + r < Lk;
            sK[r * hdp + c] = ok ? kb[static_cast<int64_t>(k0 + r) * kv_stride + c] : 0.f;
            sV[r * hdp + c] = ok ? vb[static_cast<int64_t>(k0 + r) * kv_stride + c] : 0.f;
        }
        __syncthreads();
        {   // S tile: thread owns a 4x4 patch (rows ty*4.., cols tx*4..)
            float s[4][4] = {};
            for (int d = 0; d < hd; ++d) {
                float a[4], bb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { a[i] = sQ[(ty * 4 + i) * hdp + d]; bb[i] = sK[(tx * 4 + i) * hdp + d]; }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) s[i][j] = fmaf(a[i], bb[j], s[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) sS[(ty * 4 + i) * (AK + 1) + tx * 4 + j] = (k0 + tx * 4 + j < Lk) ? s[i][j] : -INFINITY;
        }
        __syncthreads();
        {   // online softmax: 4 threads per row
            const int r = tid >> 2, part = tid & 3;
            float mx = -INFINITY;
            for (int j = part; j < AK; j += 4) mx = fmaxf(mx, sS[r * (AK + 1) + j]);
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_old = sM[r];
            const float m_new = fmaxf(m_old, mx);
            float sum = 0.f;
            for (int j = part; j < AK; j += 4) {
                const float p = __expf(sS[r * (AK + 1) + j] - m_new);
                sS[r * (AK + 1) + j] = p;
                sum += p;
            }
            sum += __shfl_xor_sync(0xffffffffu, sum, 1);
            sum += __shfl_xor_sync(0xffffffffu, sum, 2);
            __syncwarp();
            if (part == 0) {
                const float al = __expf(m_old - m_new);   // m_old = -inf on the first tile -> 0
                sAl[r] = al;
                sL[r] = sL[r] * al + sum;
                sM[r] = m_new;
            }
        }
        __syncthreads();
        {   // O = alpha * O + P V
            float al[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) al[i] = sAl[ty * 4 + i];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) o[i][j] *= al[i];
            for (int kk = 0; kk < AK; ++kk) {
                float p[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = sS[(ty * 4 + i) * (AK + 1) + kk];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = tx + 16 * j;
                    if (c < hd) {
                        const float vv = sV[kk * hdp + c];
#pragma unroll
                        for (int i = 0; i < 4; ++i) o[i][j] = fmaf(p[i], vv, o[i][j]);
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = q0 + ty * 4 + i;
        if (r >= Lq) continue;
        const float inv = 1.0f / sL[ty * 4 + i];
        float* orow = out + (static_cast<int64_t>(b) * Lq + r) * out_stride + h * hd;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = tx + 16 * j;
            if (c < hd) orow[c] = o[i][j] * inv;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Attention on the tensor cores with fp32-like operands (round 2).  Same contract as sam_attention_kernel.  Flash-attention
// structure on mma.sync.m16n8k16: CTA = 64 queries of one (batch, head) = 4 warps x 16 rows; keys in tiles of 64 with an online
// softmax in the accumulator registers.  Every operand is an fp16 pair x = hi + lo (the oracle's random Hiera amplifies a per-layer
// error ~1000x, see sam_linear_tcgen05) and every product three MMAs:  S = Q_hi K_hi + Q_lo K_hi + Q_hi K_lo,
// O += P_hi V_hi + P_lo V_hi + P_hi V_lo.  q / k / v are fp32 in global memory and are split while they are staged into shared
// memory; the softmax scale (times log2 e) is folded into Q.  HDS = head_dim padded to 16, in k16 steps (72 -> 5, 96 -> 6, 32 -> 2,
// 16 -> 1); padded columns are zero.  head_dim 72 at 4096 x 4096 tokens: 2.7 ms (SIMT) -> see profiles/r02_sam2_hiera_l_launches.md.
// (tcgen05 would want a TMEM-resident O with online rescaling and fp16 Q / K / V buffers; this path runs once per slide.)
// ---------------------------------------------------------------------------------------------------------------------
template <int HDS>
__global__ void __launch_bounds__(128)
sam_attention_tc_kernel(const float* __restrict__ q, int q_stride, const float* __restrict__ k, const float* __restrict__ v, int kv_stride,
                        float* __restrict__ out, int out_stride, int Lq, int Lk, int heads, int hd, float scale_log2) {
    constexpr int HDP = HDS * 16, PITCH = HDP + 8;           // halfs per smem row; +16 B keeps ldmatrix conflict-free
    constexpr int NT_O = HDP / 8;                            // n8 tiles of the output
    extern __shared__ __align__(16) __half sm_att[];
    __half* sQh = sm_att;                    // [64][PITCH]
    __half* sQl = sQh + 64 * PITCH;
    __half* sKh = sQl + 64 * PITCH;
    __half* sKl = sKh + 64 * PITCH;
    __half* sVh = sKl + 64 * PITCH;
    __half* sVl = sVh + 64 * PITCH;
    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 64;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* qb = q + static_cast<int64_t>(b) * Lq * q_stride + h * hd;
    const float* kb = k + static_cast<int64_t>(b) * Lk * kv_stride + h * hd;
    const float* vb = v + static_cast<int64_t>(b) * Lk * kv_stride + h * hd;
    // stage a [64][hd] fp32 tile as hi / lo fp16, 4 columns per thread and step (hd % 4 == 0, 16-byte aligned rows: checked by the launcher)
    auto stage = [&](const float* base, int64_t stride, int row0, int nrows, float mul, __half* dh, __half* dl) {
        constexpr int C4 = HDP / 4;
        for (int i = tid; i < 64 * C4; i += 128) {
            const int r = i / C4, c = (i - r * C4) * 4;
            float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row0 + r < nrows && c < hd) x = __ldg(reinterpret_cast<const float4*>(base + static_cast<int64_t>(row0 + r) * stride + c));
            x.x *= mul; x.y *= mul; x.z *= mul; x.w *= mul;
            const __half2 h0 = __floats2half2_rn(x.x, x.y), h1 = __floats2half2_rn(x.z, x.w);
            const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
            const __half2 l0 = __floats2half2_rn(x.x - f0.x, x.y - f0.y), l1 = __floats2half2_rn(x.z - f1.x, x.w - f1.y);
            *reinterpret_cast<uint2*>(dh + r * PITCH + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            *reinterpret_cast<uint2*>(dl + r * PITCH + c) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
        }
    };
    stage(qb, q_stride, q0, Lq, scale_log2, sQh, sQl);
    __syncthreads();
    // Q fragments of this warp's 16 rows (A operand, row-major): ldmatrix x4 per k16 step
    uint32_t qh[HDS][4], ql[HDS][4];
    {
        const int r = warp * 16 + (lane & 15), cofs = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < HDS; ++ks) {
            ptx::ldmatrix_x4(ptx::smem_u32(sQh + r * PITCH + ks * 16 + cofs), qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3]);
            ptx::ldmatrix_x4(ptx::smem_u32(sQl + r * PITCH + ks * 16 + cofs), ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3]);
        }
    }
    float o[NT_O][4];
#pragma unroll
    for (int j = 0; j < NT_O; ++j) { o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;     // rows g and g + 8 of the warp's 16 (g = lane / 4)
    for (int k0 = 0; k0 < Lk; k0 += 64) {
        __syncthreads();                                            // the previous tile's K / V are no longer read
        stage(kb, kv_stride, k0, Lk, 1.f, sKh, sKl);
        stage(vb, kv_stride, k0, Lk, 1.f, sVh, sVl);
        __syncthreads();
        // ---- S = Q K^T for 64 keys: 8 n8 tiles ----
        float sc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f; }
#pragma unroll
        for (int ks = 0; ks < HDS; ++ks) {
#pragma unroll
            for (int jp = 0; jp < 4; ++jp) {       // two n8 tiles (16 keys) per ldmatrix x4: rows = keys, 16 consecutive k
                const int kr = jp * 16 + (lane & 7) + ((lane >> 4) << 3), kc = ks * 16 + ((lane >> 3) & 1) * 8;
                uint32_t bh[4], bl[4];
                ptx::ldmatrix_x4(ptx::smem_u32(sKh + kr * PITCH + kc), bh[0], bh[1], bh[2], bh[3]);
                ptx::ldmatrix_x4(ptx::smem_u32(sKl + kr * PITCH + kc), bl[0], bl[1], bl[2], bl[3]);
                ptx::mma_m16n8k16_f16(sc[2 * jp], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bh[0], bh[1]);
                ptx::mma_m16n8k16_f16(sc[2 * jp], ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3], bh[0], bh[1]);
                ptx::mma_m16n8k16_f16(sc[2 * jp], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bl[0], bl[1]);
                ptx::mma_m16n8k16_f16(sc[2 * jp + 1], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bh[2], bh[3]);
                ptx::mma_m16n8k16_f16(sc[2 * jp + 1], ql[ks][0], ql[ks][1], ql[ks][2], ql[ks][3], bh[2], bh[3]);
                ptx::mma_m16n8k16_f16(sc[2 * jp + 1], qh[ks][0], qh[ks][1], qh[ks][2], qh[ks][3], bl[2], bl[3]);
            }
        }
        // ---- online softmax (log2 domain): thread holds keys 8 j + 2 (lane % 4) + {0, 1} of rows g (c0, c1) and g + 8 (c2, c3) ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int key = k0 + 8 * j + 2 * (lane & 3);
            if (key >= Lk) { sc[j][0] = -INFINITY; sc[j][2] = -INFINITY; }
            if (key + 1 >= Lk) { sc[j][1] = -INFINITY; sc[j][3] = -INFINITY; }
            mx0 = fmaxf(mx0, fmaxf(sc[j][0], sc[j][1]));
            mx1 = fmaxf(mx1, fmaxf(sc[j][2], sc[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);     // finite: every tile holds >= 1 valid key
        const float al0 = exp2f(m0 - mn0), al1 = exp2f(m1 - mn1);   // 0 on the first tile
        m0 = mn0; m1 = mn1;
        float s0 = 0.f, s1 = 0.f;
        uint32_t ph[4][4], pl[4][4];                                // A fragments of P for the 4 k16 steps over the 64 keys
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = exp2f(sc[j][0] - mn0), p1 = exp2f(sc[j][1] - mn0), p2 = exp2f(sc[j][2] - mn1), p3 = exp2f(sc[j][3] - mn1);
            s0 += p0 + p1; s1 += p2 + p3;
            const __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
            const __half2 l01 = __floats2half2_rn(p0 - f01.x, p1 - f01.y), l23 = __floats2half2_rn(p2 - f23.x, p3 - f23.y);
            const int ks = j >> 1, hi_half = (j & 1) * 2;          // n8 tile 2 ks -> a0 / a1, tile 2 ks + 1 -> a2 / a3
            ph[ks][hi_half] = *reinterpret_cast<const uint32_t*>(&h01); ph[ks][hi_half + 1] = *reinterpret_cast<const uint32_t*>(&h23);
            pl[ks][hi_half] = *reinterpret_cast<const uint32_t*>(&l01); pl[ks][hi_half + 1] = *reinterpret_cast<const uint32_t*>(&l23);
        }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        l0 = l0 * al0 + s0; l1 = l1 * al1 + s1;
#pragma unroll
        for (int j = 0; j < NT_O; ++j) { o[j][0] *= al0; o[j][1] *= al0; o[j][2] *= al1; o[j][3] *= al1; }
        // ---- O += P V: B operand = V^T, taken from the [key][hd] tile with transposing ldmatrix ----
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
            for (int jp = 0; jp < NT_O / 2; ++jp) {     // two n8 tiles (16 head-dim columns) per ldmatrix x4.trans
                const int vr = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, vc = jp * 16 + (lane >> 4) * 8;
                uint32_t bh[4], bl[4];
                ptx::ldmatrix_x4_trans(ptx::smem_u32(sVh + vr * PITCH + vc), bh[0], bh[1], bh[2], bh[3]);
                ptx::ldmatrix_x4_trans(ptx::smem_u32(sVl + vr * PITCH + vc), bl[0], bl[1], bl[2], bl[3]);
                ptx::mma_m16n8k16_f16(o[2 * jp], ph[ks][0], ph[ks][1], ph[ks][2], ph[ks][3], bh[0], bh[1]);
                ptx::mma_m16n8k16_f16(o[2 * jp], pl[ks][0], pl[ks][1], pl[ks][2], pl[ks][3], bh[0], bh[1]);
                ptx::mma_m16n8k16_f16(o[2 * jp], ph[ks][0], ph[ks][1], ph[ks][2], ph[ks][3], bl[0], bl[1]);
                ptx::mma_m16n8k16_f16(o[2 * jp + 1], ph[ks][0], ph[ks][1], ph[ks][2], ph[ks][3], bh[2], bh[3]);
                ptx::mma_m16n8k16_f16(o[2 * jp + 1], pl[ks][0], pl[ks][1], pl[ks][2], pl[ks][3], bh[2], bh[3]);
                ptx::mma_m16n8k16_f16(o[2 * jp + 1], ph[ks][0], ph[ks][1], ph[ks][2], ph[ks][3], bl[2], bl[3]);
            }
        }
    }
    const int g = lane >> 2, r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
#pragma unroll
    for (int j = 0; j < NT_O; ++j) {
        const int c = 8 * j + 2 * (lane & 3);
        if (c < hd) {
            if (r0 < Lq) *reinterpret_cast<float2*>(out + (static_cast<int64_t>(b) * Lq + r0) * out_stride + h * hd + c) = make_float2(o[j][0] * i0, o[j][1] * i0);
            if (r1 < Lq) *reinterpret_cast<float2*>(out + (static_cast<int64_t>(b) * Lq + r1) * out_stride + h * hd + c) = make_float2(o[j][2] * i1, o[j][3] * i1);
        }
    }
}

// y = a + b (b broadcast over rows when b_rows == 1)
__global__ void sam_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, int64_t n, int D, int b_rows) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = a[i] + (b_rows == 1 ? b[i % D] : b[i]);
}

// dst[2y+dy, 2x+dx, co] = lin[(y*W + x), (dy*2+dx)*Co + co] + skip[...]   (ConvTranspose2d k=2 s=2 as a GEMM + pixel shuffle), optional GELU
__global__ void sam_pixel_shuffle_add_kernel(const float* __restrict__ lin, const float* __restrict__ skip, float* __restrict__ dst, int H, int W,
                                             int Co, int act) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(4) * H * W * Co) return;
    const int co = static_cast<int>(idx % Co);
    const int p = static_cast<int>(idx / Co);
    const int oy = p / (2 * W), ox = p - oy * (2 * W);
    const int y = oy >> 1, dy = oy & 1, x = ox >> 1, dx = ox & 1;
    float v = lin[(static_cast<int64_t>(y) * W + x) * (4 * Co) + (dy * 2 + dx) * Co + co] + (skip ? skip[idx] : 0.f);
    if (act == SAM_ACT_GELU) v = gelu_exact(v);
    dst[idx] = v;
}

// dst[2y.., 2x.., :] += src[y, x, :]  (nearest x2 top-down path of the FPN)
__global__ void sam_upsample2_add_kernel(const float* __restrict__ src, float* __restrict__ dst, int H, int W, int C) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx >= static_cast<int64_t>(4) * H * W * C) return;
    const int c = static_cast<int>(idx % C);
    const int p = static_cast<int>(idx / C);
    const int oy = p / (2 * W), ox = p - oy * (2 * W);
    dst[idx] += src[(static_cast<int64_t>(oy >> 1) * W + (ox >> 1)) * C + c];
}

// bilinear resize (align_corners = False, no antialias) of a single-channel map: torch F.interpolate semantics
__global__ void sam_bilinear_kernel(const float* __restrict__ src, int Hs, int Ws, float* __restrict__ dst, int Hd, int Wd) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Hd * Wd) return;
    const int oy = idx / Wd, ox = idx - oy * Wd;
    const float sy = fmaxf((oy + 0.5f) * (static_cast<float>(Hs) / Hd) - 0.5f, 0.f);
    const float sx = fmaxf((ox + 0.5f) * (static_cast<float>(Ws) / Wd) - 0.5f, 0.f);
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = min(y0 + 1, Hs - 1), x1 = min(x0 + 1, Ws - 1);
    const float ly = sy - y0, lx = sx - x0;
    const float v = (1.f - ly) * ((1.f - lx) * src[y0 * Ws + x0] + lx * src[y0 * Ws + x1]) + ly * ((1.f - lx) * src[y1 * Ws + x0] + lx * src[y1 * Ws + x1]);
    dst[idx] = v;
}

inline unsigned blocks_for(int64_t n, int t = 256) { return static_cast<unsigned>((n + t - 1) / t); }

}  // namespace

#define SAM_LAUNCH_CHECK(ctx, what) AP_CHECK_LAUNCH(ctx, what)

// Linears with a handful of rows (the 9 decoder tokens): one warp per output column, lanes split K, the rows' partial sums are
// reduced with shuffles.  The 64 x 64-tile SIMT kernel ran these on N / 64 CTAs with a serial K loop: 58 us per launch, 52 per forward.
template <int MAXM>
__global__ void __launch_bounds__(256)
sam_linear_rows_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, const float* __restrict__ bias, float* __restrict__ C,
                       int ldc, int M, int N, int K, int act, int accumulate) {
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (n >= N) return;
    float acc[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
    const float* wrow = W + static_cast<int64_t>(n) * K;
    for (int k = lane * 4; k < K; k += 128) {        // K % 4 == 0, 16-byte aligned rows (checked by the launcher)
        const float4 w = __ldg(reinterpret_cast<const float4*>(wrow + k));
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            if (m < M) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(A + static_cast<int64_t>(m) * lda + k));
                acc[m] = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, acc[m]))));
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
    }
    if (lane == 0) {
        const float bv = bias ? bias[n] : 0.f;
#pragma unroll
        for (int m = 0; m < MAXM; ++m) {
            if (m < M) {
                float v = acc[m] + bv;
                if (act == SAM_ACT_GELU) v = 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
                else if (act == SAM_ACT_RELU) v = fmaxf(v, 0.f);
                float* dst = C + static_cast<int64_t>(m) * ldc + n;
                *dst = accumulate ? *dst + v : v;
            }
        }
    }
}

// ---- linears on the tcgen05 GEMM (gemm_tcgen05.cu) -----------------------------------------------------------------------------------
// The activations of this path are fp32 and its oracle amplifies a per-layer error ~1000x over Hiera's 48 blocks, so
// every product keeps ~22 significant bits: x = hi + lo (two fp16), A_hi W_hi + A_lo W_hi + A_hi W_lo as ONE GEMM whose contraction
// runs over three K segments (GemmPlan: AP_SPLIT_AW), fp32 accumulation in TMEM.
//   1. sam_split_a_kernel: A fp32 [M, lda] -> [A_hi | A_lo] fp16 [M, 2 K_pad]   (K padded to the GEMM's 64-column TMA box with zeros)
//   2. gemm_tcgen05_kernel: x [W_hi | W_lo] fp16 [N_pad, 2 K_pad] (split once per weight, scaled by 256 so that lo stays normal;
//      the epilogue multiplies by 1/256), bias, fp32 out (+ residual when the output needs no de-padding / activation)
//   3. sam_finish_kernel: activation / accumulate / de-padding when N is not a multiple of 128 or an activation follows
namespace {

struct SamWSplit {
    __half* w = nullptr;      // [N_pad, 2 K_pad]
    float* bias = nullptr;    // [N_pad]
    int N_pad = 0, K_pad = 0;
};
struct SamState {
    std::unordered_map<const void*, SamWSplit> weights;
    void* a_buf = nullptr; size_t a_cap = 0;      // split A operand
    void* o_buf = nullptr; size_t o_cap = 0;      // padded fp32 output
};

constexpr float SAM_W_SCALE = 256.f;

__global__ void sam_split_w_kernel(const float* __restrict__ W, const float* __restrict__ bias, __half* __restrict__ out, float* __restrict__ bias_out,
                                   int N, int K, int N_pad, int K_pad) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (idx < N_pad) bias_out[idx] = (idx < N && bias) ? bias[idx] : 0.f;
    if (idx >= static_cast<int64_t>(N_pad) * K_pad) return;
    const int n = static_cast<int>(idx / K_pad), k = static_cast<int>(idx - static_cast<int64_t>(n) * K_pad);
    const float w = (n < N && k < K) ? W[static_cast<int64_t>(n) * K + k] * SAM_W_SCALE : 0.f;
    const __half hi = __float2half_rn(w);
    out[static_cast<int64_t>(n) * 2 * K_pad + k] = hi;
    out[static_cast<int64_t>(n) * 2 * K_pad + K_pad + k] = __float2half_rn(w - __half2float(hi));
}

// 8 consecutive k per thread: two float4 in, one uint4 of hi and one of lo out
__global__ void __launch_bounds__(256)
sam_split_a_kernel(const float* __restrict__ A, int lda, __half* __restrict__ out, int M, int K, int K_pad, SamGather g) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int kc = K_pad >> 3;
    if (idx >= static_cast<int64_t>(M) * kc) return;
    const int row = static_cast<int>(idx / kc), k = static_cast<int>(idx - static_cast<int64_t>(row) * kc) * 8;
    int m = row;
    if (g.ws > 0) {      // window_partition: (window, iy, ix) -> (y, x), zero rows for the padding of the last windows
        const int ix = row % g.ws, t1 = row / g.ws, iy = t1 % g.ws, wi = t1 / g.ws;
        const int wx = wi % g.nWx, wy = wi / g.nWx;
        const int y = wy * g.ws + iy, x = wx * g.ws + ix;
        m = (y < g.H && x < g.W) ? y * g.W + x : -1;
    }
    float v[8];
    if (m < 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
    } else if (k + 8 <= K) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(A + static_cast<int64_t>(m) * lda + k));
        const float4 b = __ldg(reinterpret_cast<const float4*>(A + static_cast<int64_t>(m) * lda + k + 4));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = k + e < K ? A[static_cast<int64_t>(m) * lda + k + e] : 0.f;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        const float2 hf = __half22float2(h);
        const __half2 l = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        hi[e] = *reinterpret_cast<const uint32_t*>(&h);
        lo[e] = *reinterpret_cast<const uint32_t*>(&l);
    }
    __half* orow = out + static_cast<int64_t>(row) * 2 * K_pad;
    *reinterpret_cast<uint4*>(orow + k) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(orow + K_pad + k) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__device__ __forceinline__ float sam_act(float v, int act) {
    if (act == SAM_ACT_GELU) return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
    if (act == SAM_ACT_RELU) return fmaxf(v, 0.f);
    return v;
}
__global__ void sam_finish_kernel(const float* __restrict__ tmp, int ldt, float* __restrict__ C, int ldc, int M, int N, int act, int accumulate) {
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int nc = N >> 2;   // N % 4 == 0 (checked by the launcher)
    if (idx >= static_cast<int64_t>(M) * nc) return;
    const int m = static_cast<int>(idx / nc), n = static_cast<int>(idx - static_cast<int64_t>(m) * nc) * 4;
    float4 v = *reinterpret_cast<const float4*>(tmp + static_cast<int64_t>(m) * ldt + n);
    v.x = sam_act(v.x, act); v.y = sam_act(v.y, act); v.z = sam_act(v.z, act); v.w = sam_act(v.w, act);
    float4* dst = reinterpret_cast<float4*>(C + static_cast<int64_t>(m) * ldc + n);
    if (accumulate) {
        const float4 c = *dst;
        v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
    }
    *dst = v;
}

int sam_scratch(ap_ctx* ctx, void** buf, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return AP_OK;
    // the stream that still uses the old buffer has to drain before it is freed; growth happens a handful of times per model
    AP_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    if (*buf) cudaFree(*buf);
    *buf = nullptr; *cap = 0;
    cudaError_t e = cudaMalloc(buf, bytes);
    if (e != cudaSuccess) return ap_set_error(ctx, AP_ENOMEM, "sam2: cudaMalloc(%zu) for the tensor-core linears failed: %s", bytes, cudaGetErrorString(e));
    *cap = bytes;
    return AP_OK;
}

int sam_linear_tcgen05(ap_ctx* ctx, const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int M, int N, int K, int act,
                       int accumulate, cudaStream_t st, const SamGather* gather = nullptr) {
    if (!ctx->sam_state) ctx->sam_state = new SamState();
    SamState* ss = static_cast<SamState*>(ctx->sam_state);
    const int K_pad = (K + 63) / 64 * 64, N_pad = (N + 127) / 128 * 128;
    auto it = ss->weights.find(W);
    if (it == ss->weights.end()) {     // first use of this weight: split it once
        SamWSplit ws;
        ws.N_pad = N_pad; ws.K_pad = K_pad;
        cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&ws.w), static_cast<size_t>(N_pad) * 2 * K_pad * sizeof(__half));
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&ws.bias), static_cast<size_t>(N_pad) * sizeof(float));
        if (e != cudaSuccess) return ap_set_error(ctx, AP_ENOMEM, "sam2: cudaMalloc for split weights failed: %s", cudaGetErrorString(e));
        const int64_t n = static_cast<int64_t>(N_pad) * K_pad;
        sam_split_w_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(W, bias, ws.w, ws.bias, N, K, N_pad, K_pad);
        SAM_LAUNCH_CHECK(ctx, "sam_split_w_kernel");
        it = ss->weights.emplace(W, ws).first;
    }
    const SamWSplit& ws = it->second;
    AP_REQUIRE(ctx, ws.N_pad == N_pad && ws.K_pad == K_pad, "sam2: weight %p reused with another shape", static_cast<const void*>(W));
    int rc = sam_scratch(ctx, &ss->a_buf, &ss->a_cap, static_cast<size_t>(M) * 2 * K_pad * sizeof(__half));
    if (rc) return rc;
    {
        const int64_t n = static_cast<int64_t>(M) * (K_pad >> 3);
        const SamGather g = gather ? *gather : SamGather{0, 0, 0, 0};
        sam_split_a_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(A, lda, static_cast<__half*>(ss->a_buf), M, K, K_pad, g);
        SAM_LAUNCH_CHECK(ctx, "sam_split_a_kernel");
    }
    // without accumulation the GEMM's fp32 epilogue writes C directly (its own row stride, column bound and activation); the
    // accumulating linears (mlp.proj_out) go through the residual epilogue when the shapes allow it, else through a padded temporary
    const bool direct_plain = !accumulate;
    const bool direct_resid = accumulate && N == N_pad && ldc == N && act == SAM_ACT_NONE;
    const bool direct = direct_plain || direct_resid;
    float* out = C;
    if (!direct) {
        rc = sam_scratch(ctx, &ss->o_buf, &ss->o_cap, static_cast<size_t>(M) * N_pad * sizeof(float));
        if (rc) return rc;
        out = static_cast<float*>(ss->o_buf);
    }
    GemmPlan plan;
    rc = ap_gemm_plan_split(ctx, &plan, ss->a_buf, ws.w, M, N_pad, K_pad, direct_resid ? AP_EPI_BIAS_RESID_F32 : AP_EPI_BIAS_F32, AP_SPLIT_AW);
    if (rc) return rc;
    GemmExtra ex;
    ex.alpha = 1.0f / SAM_W_SCALE;
    if (direct_plain) {
        ex.out_ld = ldc; ex.n_valid = N;
        ex.act = act == SAM_ACT_GELU ? 1 : act == SAM_ACT_RELU ? 2 : 0;
    }
    rc = ap_gemm_run(ctx, &plan, ws.bias, direct_resid ? C : nullptr, out, &ex, st);
    if (rc) return rc;
    if (!direct) {
        const int64_t n = static_cast<int64_t>(M) * (N >> 2);
        sam_finish_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(out, N_pad, C, ldc, M, N, act, accumulate);
        SAM_LAUNCH_CHECK(ctx, "sam_finish_kernel");
    }
    return AP_OK;
}

}  // namespace

void sam_state_free(ap_ctx* ctx) {
    if (!ctx || !ctx->sam_state) return;
    SamState* ss = static_cast<SamState*>(ctx->sam_state);
    cudaDeviceSynchronize();
    for (auto& kv : ss->weights) { cudaFree(kv.second.w); cudaFree(kv.second.bias); }
    if (ss->a_buf) cudaFree(ss->a_buf);
    if (ss->o_buf) cudaFree(ss->o_buf);
    delete ss;
    ctx->sam_state = nullptr;
}

// qkv linear of a windowed block with window_partition fused into the operand staging (tcgen05 path only; the caller checks the mode)
int sam_linear_windows(ap_ctx* ctx, const float* A, int lda, const SamGather& g, const float* W, const float* bias, float* C, int ldc, int M, int N,
                       int K, cudaStream_t st) {
    AP_REQUIRE(ctx, ctx->sam_tensor_cores == 3 && M >= 256 && lda % 4 == 0 && K % 4 == 0 && N % 4 == 0 && ldc % 4 == 0 &&
                        ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(C)) & 15) == 0,
               "sam_linear_windows: shape / alignment not supported by the tensor-core linear");
    return sam_linear_tcgen05(ctx, A, lda, W, bias, C, ldc, M, N, K, SAM_ACT_NONE, 0, st, &g);
}

int sam_linear(ap_ctx* ctx, const float* A, int lda, const float* W, const float* bias, float* C, int ldc, int M, int N, int K, int act,
               int accumulate, cudaStream_t st) {
    if (M == 0) return AP_OK;
    const bool aligned = ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(W)) & 15) == 0 && lda % 4 == 0 && K % 4 == 0;
    // tcgen05 GEMM: whole 128-row tiles are worth it from a few hundred rows on; the 9-token decoder linears stay on the SIMT kernel
    if (ctx->sam_tensor_cores == 3 && aligned && M >= 256 && N % 4 == 0 && ldc % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0)
        return sam_linear_tcgen05(ctx, A, lda, W, bias, C, ldc, M, N, K, act, accumulate, st);
    if (aligned && M <= 16) {
        sam_linear_rows_kernel<16><<<(N + 7) / 8, 256, 0, st>>>(A, lda, W, bias, C, ldc, M, N, K, act, accumulate);
        SAM_LAUNCH_CHECK(ctx, "sam_linear_rows_kernel");
        return AP_OK;
    }
    if (ctx->sam_tensor_cores && aligned && M >= 64) {
        static PerDeviceOnce attr;
        if (attr.need(ctx->device)) {
            AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(sam_linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
            AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(sam_linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
            attr.done(ctx->device);
        }
        dim3 grid((N + TCN - 1) / TCN, (M + TCM - 1) / TCM);
        if (ctx->sam_tensor_cores == 2) sam_linear_tc_kernel<false><<<grid, 256, TC_SMEM_BYTES, st>>>(A, lda, W, bias, C, ldc, M, N, K, act, accumulate);
        else sam_linear_tc_kernel<true><<<grid, 256, TC_SMEM_BYTES, st>>>(A, lda, W, bias, C, ldc, M, N, K, act, accumulate);
        SAM_LAUNCH_CHECK(ctx, "sam_linear_tc_kernel");
        return AP_OK;
    }
    dim3 grid((N + LT - 1) / LT, (M + LT - 1) / LT);
    sam_linear_kernel<<<grid, 256, 0, st>>>(A, lda, W, bias, C, ldc, M, N, K, act, accumulate);
    SAM_LAUNCH_CHECK(ctx, "sam_linear_kernel");
    return AP_OK;
}
int sam_layernorm(ap_ctx* ctx, const float* x, const float* g, const float* b, float* y, int rows, int D, float eps, int act, cudaStream_t st) {
    sam_layernorm_kernel<<<blocks_for(static_cast<int64_t>(rows) * 32), 256, 0, st>>>(x, g, b, y, rows, D, eps, act);
    SAM_LAUNCH_CHECK(ctx, "sam_layernorm_kernel");
    return AP_OK;
}
int sam_patch_embed(ap_ctx* ctx, const uint8_t* img, int H, int W, const float* w, const float* bias, const float* pos, float* out, int C,
                    const float* mean, const float* stdv, cudaStream_t st) {
    sam_patch_embed_kernel<<<blocks_for(static_cast<int64_t>(H / 4) * (W / 4) * C), 256, 0, st>>>(
        img, H, W, w, bias, pos, out, C, make_float3(mean[0], mean[1], mean[2]), make_float3(1.f / stdv[0], 1.f / stdv[1], 1.f / stdv[2]));
    SAM_LAUNCH_CHECK(ctx, "sam_patch_embed_kernel");
    return AP_OK;
}
int sam_patch_im2col(ap_ctx* ctx, const uint8_t* img, int H, int W, float* cols, const float* mean, const float* stdv, cudaStream_t st) {
    sam_patch_im2col_kernel<<<blocks_for(static_cast<int64_t>(H / 4) * (W / 4) * 148), 256, 0, st>>>(
        img, H, W, cols, make_float3(mean[0], mean[1], mean[2]), make_float3(1.f / stdv[0], 1.f / stdv[1], 1.f / stdv[2]));
    SAM_LAUNCH_CHECK(ctx, "sam_patch_im2col_kernel");
    return AP_OK;
}
int sam_window_gather(ap_ctx* ctx, const float* x, float* win, int H, int W, int C, int ws, int nWy, int nWx, cudaStream_t st) {
    sam_window_gather_kernel<<<blocks_for(static_cast<int64_t>(nWy) * nWx * ws * ws * C), 256, 0, st>>>(x, win, H, W, C, ws, nWy, nWx);
    SAM_LAUNCH_CHECK(ctx, "sam_window_gather_kernel");
    return AP_OK;
}
int sam_window_scatter_add(ap_ctx* ctx, const float* win, const float* res, float* out, int H, int W, int C, int ws, int nWx, cudaStream_t st) {
    sam_window_scatter_add_kernel<<<blocks_for(static_cast<int64_t>(H) * W * C), 256, 0, st>>>(win, res, out, H, W, C, ws, nWx);
    SAM_LAUNCH_CHECK(ctx, "sam_window_scatter_add_kernel");
    return AP_OK;
}
int sam_maxpool2(ap_ctx* ctx, const float* x, int ld, float* y, int nB, int H, int W, int C, cudaStream_t st) {
    sam_maxpool2_kernel<<<blocks_for(static_cast<int64_t>(nB) * (H / 2) * (W / 2) * C), 256, 0, st>>>(x, ld, y, nB, H, W, C);
    SAM_LAUNCH_CHECK(ctx, "sam_maxpool2_kernel");
    return AP_OK;
}
namespace {
template <int HDS>
int launch_sam_attention_tc(ap_ctx* ctx, const float* q, int q_stride, const float* k, const float* v, int kv_stride, float* out, int out_stride,
                            int nB, int Lq, int Lk, int heads, int hd, float scale, cudaStream_t st) {
    constexpr size_t smem = sizeof(__half) * 6 * 64 * (HDS * 16 + 8);
    static PerDeviceOnce attr;
    if (attr.need(ctx->device)) {
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(sam_attention_tc_kernel<HDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr.done(ctx->device);
    }
    dim3 grid((Lq + 63) / 64, heads, nB);
    sam_attention_tc_kernel<HDS><<<grid, 128, smem, st>>>(q, q_stride, k, v, kv_stride, out, out_stride, Lq, Lk, heads, hd,
                                                          scale * 1.44269504088896340736f);
    SAM_LAUNCH_CHECK(ctx, "sam_attention_tc_kernel");
    return AP_OK;
}
}  // namespace

int sam_attention(ap_ctx* ctx, const float* q, int q_stride, const float* k, const float* v, int kv_stride, float* out, int out_stride, int nB,
                  int Lq, int Lk, int heads, int hd, float scale, cudaStream_t st) {
    AP_REQUIRE(ctx, hd >= 1 && hd <= 128, "sam attention: head_dim %d unsupported", hd);
    const bool vec_ok = hd % 4 == 0 && q_stride % 4 == 0 && kv_stride % 4 == 0 && out_stride % 2 == 0 &&
                        ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v)) & 15) == 0 &&
                        (reinterpret_cast<uintptr_t>(out) & 7) == 0;
    if (ctx->sam_tensor_cores && vec_ok && Lk >= 1) {
        const int hds = (hd + 15) / 16;
        if (hds == 5) return launch_sam_attention_tc<5>(ctx, q, q_stride, k, v, kv_stride, out, out_stride, nB, Lq, Lk, heads, hd, scale, st);
        if (hds == 6) return launch_sam_attention_tc<6>(ctx, q, q_stride, k, v, kv_stride, out, out_stride, nB, Lq, Lk, heads, hd, scale, st);
        if (hds == 2) return launch_sam_attention_tc<2>(ctx, q, q_stride, k, v, kv_stride, out, out_stride, nB, Lq, Lk, heads, hd, scale, st);
        if (hds == 1) return launch_sam_attention_tc<1>(ctx, q, q_stride, k, v, kv_stride, out, out_stride, nB, Lq, Lk, heads, hd, scale, st);
    }
    const size_t smem = sizeof(float) * (static_cast<size_t>(AQ + 2 * AK) * (hd + 1) + AQ * (AK + 1) + 3 * AQ);
    static PerDeviceOnce attr;
    if (attr.need(ctx->device)) {
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(sam_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)(sizeof(float) * ((AQ + 2 * AK) * 129 + AQ * (AK + 1) + 3 * AQ))));
        attr.done(ctx->device);
    }
    dim3 grid((Lq + AQ - 1) / AQ, heads, nB);
    sam_attention_kernel<<<grid, 256, smem, st>>>(q, q_stride, k, v, kv_stride, out, out_stride, Lq, Lk, heads, hd, scale);
    SAM_LAUNCH_CHECK(ctx, "sam_attention_kernel");
    return AP_OK;
}
int sam_add(ap_ctx* ctx, const float* a, const float* b, float* y, int64_t n, int D, int b_rows, cudaStream_t st) {
    sam_add_kernel<<<blocks_for(n), 256, 0, st>>>(a, b, y, n, D, b_rows);
    SAM_LAUNCH_CHECK(ctx, "sam_add_kernel");
    return AP_OK;
}
int sam_pixel_shuffle_add(ap_ctx* ctx, const float* lin, const float* skip, float* dst, int H, int W, int Co, int act, cudaStream_t st) {
    sam_pixel_shuffle_add_kernel<<<blocks_for(static_cast<int64_t>(4) * H * W * Co), 256, 0, st>>>(lin, skip, dst, H, W, Co, act);
    SAM_LAUNCH_CHECK(ctx, "sam_pixel_shuffle_add_kernel");
    return AP_OK;
}
int sam_upsample2_add(ap_ctx* ctx, const float* src, float* dst, int H, int W, int C, cudaStream_t st) {
    sam_upsample2_add_kernel<<<blocks_for(static_cast<int64_t>(4) * H * W * C), 256, 0, st>>>(src, dst, H, W, C);
    SAM_LAUNCH_CHECK(ctx, "sam_upsample2_add_kernel");
    return AP_OK;
}
int sam_bilinear(ap_ctx* ctx, const float* src, int Hs, int Ws, float* dst, int Hd, int Wd, cudaStream_t st) {
    sam_bilinear_kernel<<<blocks_for(static_cast<int64_t>(Hd) * Wd), 256, 0, st>>>(src, Hs, Ws, dst, Hd, Wd);
    SAM_LAUNCH_CHECK(ctx, "sam_bilinear_kernel");
    return AP_OK;
}
