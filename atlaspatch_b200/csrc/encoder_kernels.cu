// Non-GEMM kernels of the encoder forward: patch crop -> fp16 im2col (preprocess), class-token rows,
// LayerNorm, multi-head attention.  All fp32 statistics / softmax; fp16 only as tensor-core operands.
#include <algorithm>
#include <cmath>
#include <vector>

#include "ap_internal.cuh"
#include "ptx.cuh"

namespace {

// =====================================================================================================
// a11 + a12: patch read + encoder preprocess, fused with the im2col of conv_proj.
//   reference: feature_embedding.py:81-96 (wsi.extract at coords), models/patch/base.py:42-45 +
//   torchvision ImageClassification preset (centre crop `image` out of `input_patch`, /255, normalise).
//   Normalisation is folded into the conv_proj weights at ap_encoder_finalize, so this kernel is a pure
//   byte gather: out[(b*g*g + ty*g + tx), c*p*p + ky*p + kx] = fp16((pixel - centre_c) / 256)  (exact in fp16;
//   centre_c = round(255 mean_c) keeps the operand small so weight rounding does not see the common mode).
//   With `dup` the row is written twice ([A | A], K = 2*3*p*p) to meet the hi/lo split conv_proj weights.
//   Pixels outside the slide read as 0 (reference backends pad with black, openslide_wsi.py:198).
// One CTA per (patch b, token row ty): stages p image rows x (image*3) bytes in smem with coalesced
// byte loads, then writes g token rows of 3*p*p halfs with 16-byte stores.
// =====================================================================================================
template <int P>  // conv patch edge (16)
__global__ void __launch_bounds__(256)
preprocess_kernel(const uint8_t* __restrict__ slide, int64_t W, int64_t H, int64_t pitch,
                  const int32_t* __restrict__ coords, int input_patch, int image, __half* __restrict__ out,
                  int64_t out_row_stride, int3 centre, int dup, const int32_t* __restrict__ lin_s, const int16_t* __restrict__ lin_w) {
    extern __shared__ uint8_t s_rows[];  // [P][image*3]
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int g = image / P;
    const int b = blockIdx.x / g;
    const int ty = blockIdx.x % g;
    const int off = (input_patch - image) / 2;  // centre crop: int(round((256-224)/2)) = 16
    const int row_bytes = image * 3;
    if (lin_s == nullptr) {
        const int64_t x0 = static_cast<int64_t>(coords[b * 5 + 0]) + off;
        const int64_t y0 = static_cast<int64_t>(coords[b * 5 + 1]) + off + ty * P;
        for (int i = threadIdx.x; i < P * row_bytes; i += blockDim.x) {
            const int r = i / row_bytes;
            const int bx = i - r * row_bytes;
            const int64_t y = y0 + r;
            const int64_t x = x0 + bx / 3;
            uint8_t v = 0;
            if (y >= 0 && y < H && x >= 0 && x < W) v = __ldg(slide + y * pitch + x0 * 3 + bx);
            s_rows[i] = v;
        }
    } else {
        // read size > patch size (e.g. a 40x slide read for 20x patches): the reference resizes the R x R read with
        // cv2.resize(patch, (P, P)) (feature_embedding.py:93-95) = OpenCV's 8-bit INTER_LINEAR: two taps per axis with 11-bit
        // coefficients from the tables built by ap_build_linear_tables (same float arithmetic as cv2), horizontal pass in int32,
        // vertical pass ((b0 (h0 >> 4)) >> 16) + ((b1 (h1 >> 4)) >> 16) + 2) >> 2.  Pixels outside the slide count as 0 (the read is
        // zero-padded before the resize).  oracle/patch_filter.py: resize_linear_u8 restates it and is pinned against cv2.
        const int64_t x0 = coords[b * 5 + 0], y0 = coords[b * 5 + 1];
        for (int i = threadIdx.x; i < P * row_bytes; i += blockDim.x) {
            const int r = i / row_bytes;
            const int bx = i - r * row_bytes;
            const int px = bx / 3, c = bx - px * 3;
            const int ox = off + px, oy = off + ty * P + r;
            const int a0 = lin_w[2 * ox], a1 = lin_w[2 * ox + 1], b0 = lin_w[2 * oy], b1 = lin_w[2 * oy + 1];
            const int64_t xa = x0 + lin_s[2 * ox], xb = x0 + lin_s[2 * ox + 1], ya = y0 + lin_s[2 * oy], yb = y0 + lin_s[2 * oy + 1];
            auto px_at = [&](int64_t y, int64_t x) -> int {
                return (y >= 0 && y < H && x >= 0 && x < W) ? static_cast<int>(__ldg(slide + y * pitch + x * 3 + c)) : 0;
            };
            const int h0 = px_at(ya, xa) * a0 + px_at(ya, xb) * a1, h1 = px_at(yb, xa) * a0 + px_at(yb, xb) * a1;
            s_rows[i] = static_cast<uint8_t>((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
        }
    }
    __syncthreads();
    // output: g tokens x (3*P*P) halfs; each thread produces 8 consecutive kx (16 B)
    const int kdim = 3 * P * P;
    const int chunks_per_token = kdim / 8;
    for (int i = threadIdx.x; i < g * chunks_per_token; i += blockDim.x) {
        const int tx = i / chunks_per_token;
        const int ch = i - tx * chunks_per_token;
        const int k0 = ch * 8;
        const int c = k0 / (P * P);
        const int ky = (k0 - c * P * P) / P;
        const int kx0 = k0 % P;
        const uint8_t* src = s_rows + ky * row_bytes + (tx * P + kx0) * 3 + c;
        const int cen = c == 0 ? centre.x : (c == 1 ? centre.y : centre.z);
        uint32_t w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const __half2 hh = __floats2half2_rn(static_cast<float>(static_cast<int>(src[(2 * j) * 3]) - cen) * (1.0f / 256.0f),
                                                 static_cast<float>(static_cast<int>(src[(2 * j + 1) * 3]) - cen) * (1.0f / 256.0f));
            w[j] = *reinterpret_cast<const uint32_t*>(&hh);
        }
        const int64_t orow = static_cast<int64_t>(b) * g * g + ty * g + tx;
        const uint4 u = make_uint4(w[0], w[1], w[2], w[3]);
        *reinterpret_cast<uint4*>(out + orow * out_row_stride + k0) = u;
        if (dup) *reinterpret_cast<uint4*>(out + orow * out_row_stride + kdim + k0) = u;
    }
}

// x[b*tokens + 0, :] = class_token + pos[0, :]; with LayerNorm folding also the fp16 copy of the row and its statistics partials
// (one warp per block of D / parts columns, fixed shuffle tree: same layout and reproducibility as the GEMM producers').
__global__ void cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls, const float* __restrict__ pos,
                                const float* __restrict__ regs, int lead, int n_images, int tokens, int D, __half* __restrict__ xh,
                                float2* __restrict__ stats, int parts) {
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    // one CTA per leading row: r = 0 the class token + its position, r >= 1 register token r - 1 (no position: transformers
    // Dinov2WithRegistersEmbeddings inserts them after the position embedding has been added)
    const int b = blockIdx.x / lead, r = blockIdx.x - b * lead, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (b >= n_images || w >= parts) return;
    const int slot = D / parts;
    const int64_t row = static_cast<int64_t>(b) * tokens + r;
    float s = 0.f, q = 0.f;
    for (int d = w * slot + lane; d < (w + 1) * slot; d += 32) {
        const float v = r == 0 ? cls[d] + pos[d] : regs[static_cast<int64_t>(r - 1) * D + d];
        x[row * D + d] = v;
        if (xh != nullptr) xh[row * D + d] = __float2half_rn(v);
        s += v;
        q += v * v;
    }
    if (stats != nullptr) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        if (lane == 0) stats[row * parts + w] = make_float2(s, q);
    }
}

// =====================================================================================================
// LayerNorm over the last dim (biased variance, eps inside the sqrt: torch.nn.LayerNorm, eps 1e-6 in
// torchvision's EncoderBlock).  One warp per row, the row lives in registers, two-pass mean / variance.
// =====================================================================================================
template <int VEC>  // D = VEC * 128
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, int64_t x_row_stride, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __half* __restrict__ y16, float* __restrict__ y32, int rows, int y_ld,
                 int split_lo) {
    constexpr int D = VEC * 128;
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(warp) * x_row_stride);
    float4 v[VEC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        v[i] = xr[lane + 32 * i];
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / D) + eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const float4 g = __ldg(g4 + lane + 32 * i);
        const float4 bb = __ldg(b4 + lane + 32 * i);
        const float o0 = (v[i].x - mean) * rstd * g.x + bb.x;
        const float o1 = (v[i].y - mean) * rstd * g.y + bb.y;
        const float o2 = (v[i].z - mean) * rstd * g.z + bb.z;
        const float o3 = (v[i].w - mean) * rstd * g.w + bb.w;
        if (y16) {
            __half2 h0 = __floats2half2_rn(o0, o1), h1 = __floats2half2_rn(o2, o3);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0);
            u.y = *reinterpret_cast<uint32_t*>(&h1);
            reinterpret_cast<uint2*>(y16 + static_cast<int64_t>(warp) * y_ld)[lane + 32 * i] = u;
            if (split_lo) {   // A-operand split (DESIGN.md "precision"): the row holds [hi | lo], lo = fp16(y - hi)
                const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
                __half2 l0 = __floats2half2_rn(o0 - f0.x, o1 - f0.y), l1 = __floats2half2_rn(o2 - f1.x, o3 - f1.y);
                uint2 ul;
                ul.x = *reinterpret_cast<uint32_t*>(&l0);
                ul.y = *reinterpret_cast<uint32_t*>(&l1);
                reinterpret_cast<uint2*>(y16 + static_cast<int64_t>(warp) * y_ld + D)[lane + 32 * i] = ul;
            }
        } else {
            reinterpret_cast<float4*>(y32 + static_cast<int64_t>(warp) * D)[lane + 32 * i] = make_float4(o0, o1, o2, o3);
        }
    }
}

// =====================================================================================================
// [CLS || mean(patch tokens)] head (atlas_patch/models/patch/midnight.py:57-61, virchow.py:57-61): the final LayerNorm applied to
// EVERY token of the residual stream, then  out[b] = [ LN(x[b, 0]) , mean_{t >= 1} LN(x[b, t]) ]  (2 D floats per image).
// One CTA per image; each warp normalises whole rows held in registers (as layernorm_kernel) and sums the normalised values of its
// rows; gamma / beta are applied to the mean once (the LayerNorm's affine map commutes with it).  The 8 warp partials are added
// in warp order, so the result does not depend on scheduling.  HBM-bound: tokens * D * 4 bytes read once per image.
// =====================================================================================================
template <int VEC>  // D = VEC * 128
__global__ void __launch_bounds__(256)
cls_mean_pool_kernel(const float* __restrict__ x, int tokens1, int lead, const float* __restrict__ gamma, const float* __restrict__ beta,
                     float eps, float* __restrict__ out) {
    constexpr int D = VEC * 128;
    __shared__ float4 s_acc[VEC * 32];
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* xb = x + static_cast<int64_t>(blockIdx.x) * tokens1 * D;
    float* ob = out + static_cast<int64_t>(blockIdx.x) * 2 * D;
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
    float4 acc[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    // warp 0 starts at the class token (t = 0), which is written out instead of accumulated; register tokens (rows 1 .. lead - 1,
    // virchow.py:110-114 `output[:, 5:]`) are skipped; patch tokens are dealt round-robin
    for (int t = warp == 0 ? 0 : lead - 1 + warp; t < tokens1; t = (t == 0 ? lead + 7 : t + 8)) {
        const float4* xr = reinterpret_cast<const float4*>(xb + static_cast<int64_t>(t) * D);
        float4 v[VEC];
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            v[i] = xr[lane + 32 * i];
            sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum * (1.0f / D);
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
            sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq * (1.0f / D) + eps);
        if (t == 0) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float4 g = __ldg(g4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
                reinterpret_cast<float4*>(ob)[lane + 32 * i] =
                    make_float4(v[i].x * rstd * g.x + bb.x, v[i].y * rstd * g.y + bb.y, v[i].z * rstd * g.z + bb.z, v[i].w * rstd * g.w + bb.w);
            }
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                acc[i].x += v[i].x * rstd; acc[i].y += v[i].y * rstd; acc[i].z += v[i].z * rstd; acc[i].w += v[i].w * rstd;
            }
        }
    }
    for (int w = 0; w < 8; ++w) {
        if (warp == w) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float4 a = acc[i];
                if (w > 0) {
                    const float4 p = s_acc[lane + 32 * i];
                    a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
                }
                s_acc[lane + 32 * i] = a;
            }
        }
        __syncthreads();
    }
    const float inv = 1.0f / static_cast<float>(tokens1 - lead);
    for (int j = threadIdx.x; j < VEC * 32; j += 256) {
        const float4 a = s_acc[j], g = __ldg(g4 + j), bb = __ldg(b4 + j);
        reinterpret_cast<float4*>(ob + D)[j] = make_float4(a.x * inv * g.x + bb.x, a.y * inv * g.y + bb.y, a.z * inv * g.z + bb.z, a.w * inv * g.w + bb.w);
    }
}

// =====================================================================================================
// Multi-head self-attention, head_dim 64, sequence S <= 272 (197 for ViT/16@224, 257 for ViT/14@224).
// nn.MultiheadAttention semantics (torchvision EncoderBlock.self_attention): softmax((q/sqrt(d)) k^T) v.
// One CTA per (head, image): Q, K, V head slices staged once in (XOR-swizzled) smem with cp.async,
// each warp owns 16-query tiles and streams keys in chunks of 64 with an online softmax (fp32).
// v1 uses warp-level mma.sync m16n8k16 (4 % of the model FLOPs); the tcgen05 version is the next step.
// =====================================================================================================
constexpr int HD = 64;
constexpr int ATT_WARPS = 7;
constexpr int ATT_THREADS = ATT_WARPS * 32;
constexpr int KCHUNK = 64;

__device__ __forceinline__ uint32_t swz_off(int row, int chunk16) {  // byte offset in a [rows][128 B] tile
    return static_cast<uint32_t>(row * 128 + ((chunk16 ^ (row & 7)) << 4));
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int S, int S_pad, int heads) {
    extern __shared__ __align__(128) uint8_t smem_att[];
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int D = heads * HD;
    const int h = blockIdx.x;
    const int b = blockIdx.y;
    uint8_t* sQ = smem_att;
    uint8_t* sK = sQ + S_pad * 128;
    uint8_t* sV = sK + S_pad * 128;
    const uint32_t sQ_u = ptx::smem_u32(sQ), sK_u = ptx::smem_u32(sK), sV_u = ptx::smem_u32(sV);

    // ---- stage Q, K, V (zero-filled beyond S) ----
    const __half* base = qkv + static_cast<int64_t>(b) * S * 3 * D + h * HD;
    for (int i = threadIdx.x; i < S_pad * 8 * 3; i += ATT_THREADS) {
        const int mat = i / (S_pad * 8);
        const int rem = i - mat * (S_pad * 8);
        const int row = rem >> 3, ch = rem & 7;
        const uint32_t dst = (mat == 0 ? sQ_u : (mat == 1 ? sK_u : sV_u)) + swz_off(row, ch);
        const int srow = row < S ? row : S - 1;
        const __half* src = base + static_cast<int64_t>(srow) * 3 * D + mat * D + ch * 8;
        ptx::cp_async_16(dst, src, row < S ? 16u : 0u);
    }
    ptx::cp_async_commit();
    ptx::cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const float scale_log2 = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)
    const int n_qtiles = S_pad / 16;

    for (int qt = warp; qt < n_qtiles; qt += ATT_WARPS) {
        const int q0 = qt * 16;
        uint32_t qf[4][4];  // A fragments for the 4 k-steps of head_dim 64
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int row = q0 + (lane & 7) + ((lane >> 3) & 1) * 8;
            const int ch = ks * 2 + (lane >> 4);
            ptx::ldmatrix_x4(sQ_u + swz_off(row, ch), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
        }
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

        for (int kc = 0; kc < S_pad; kc += KCHUNK) {
            float s[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
            // ---- S = Q K^T for up to 64 keys (8 n-tiles), two n-tiles per ldmatrix.x4 ----
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const int key0 = kc + np * 16;
                if (key0 < S_pad) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        uint32_t b0, b1, b2, b3;
                        const int row = key0 + (lane & 7) + (lane >> 4) * 8;
                        const int ch = ks * 2 + ((lane >> 3) & 1);
                        ptx::ldmatrix_x4(sK_u + swz_off(row, ch), b0, b1, b2, b3);
                        ptx::mma_m16n8k16_f16(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
                        ptx::mma_m16n8k16_f16(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
                    }
                }
            }
            // ---- mask keys >= S, chunk max ----
            float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int key = kc + i * 8 + 2 * t;
                if (key >= S) { s[i][0] = -INFINITY; s[i][2] = -INFINITY; }
                if (key + 1 >= S) { s[i][1] = -INFINITY; s[i][3] = -INFINITY; }
                cm0 = fmaxf(cm0, fmaxf(s[i][0], s[i][1]));
                cm1 = fmaxf(cm1, fmaxf(s[i][2], s[i][3]));
            }
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1));
            cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1));
            cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
            const float nm0 = fmaxf(m0, cm0), nm1 = fmaxf(m1, cm1);  // finite: every chunk holds >= 1 valid key
            const float corr0 = exp2f((m0 - nm0) * scale_log2), corr1 = exp2f((m1 - nm1) * scale_log2);
            m0 = nm0; m1 = nm1;
            l0 *= corr0; l1 *= corr1;
#pragma unroll
            for (int i = 0; i < 8; ++i) { o[i][0] *= corr0; o[i][1] *= corr0; o[i][2] *= corr1; o[i][3] *= corr1; }
            // ---- P = exp2((s - m) * scale), row sums, pack to fp16 A fragments ----
            uint32_t pf[4][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float p0 = exp2f((s[i][0] - m0) * scale_log2), p1 = exp2f((s[i][1] - m0) * scale_log2);
                const float p2 = exp2f((s[i][2] - m1) * scale_log2), p3 = exp2f((s[i][3] - m1) * scale_log2);
                l0 += p0 + p1; l1 += p2 + p3;
                __half2 h01 = __floats2half2_rn(p0, p1), h23 = __floats2half2_rn(p2, p3);
                pf[i >> 1][(i & 1) * 2 + 0] = *reinterpret_cast<uint32_t*>(&h01);
                pf[i >> 1][(i & 1) * 2 + 1] = *reinterpret_cast<uint32_t*>(&h23);
            }
            // ---- O += P V ----
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int key0 = kc + ks * 16;
                if (key0 < S_pad) {
#pragma unroll
                    for (int np = 0; np < 4; ++np) {
                        uint32_t b0, b1, b2, b3;
                        const int row = key0 + (lane & 7) + ((lane >> 3) & 1) * 8;
                        const int ch = np * 2 + (lane >> 4);
                        ptx::ldmatrix_x4_trans(sV_u + swz_off(row, ch), b0, b1, b2, b3);
                        ptx::mma_m16n8k16_f16(o[2 * np], pf[ks][0], pf[ks][1], pf[ks][2], pf[ks][3], b0, b1);
                        ptx::mma_m16n8k16_f16(o[2 * np + 1], pf[ks][0], pf[ks][1], pf[ks][2], pf[ks][3], b2, b3);
                    }
                }
            }
        }
        // ---- finalise: O / l, store fp16 ----
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
        const int r0 = q0 + g, r1 = q0 + g + 8;
        __half* obase = out + static_cast<int64_t>(b) * S * D + h * HD;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int col = i * 8 + 2 * t;
            if (r0 < S) *reinterpret_cast<__half2*>(obase + static_cast<int64_t>(r0) * D + col) = __floats2half2_rn(o[i][0] * inv0, o[i][1] * inv0);
            if (r1 < S) *reinterpret_cast<__half2*>(obase + static_cast<int64_t>(r1) * D + col) = __floats2half2_rn(o[i][2] * inv1, o[i][3] * inv1);
        }
    }
}

// =====================================================================================================
// Last encoder layer: only the class-token row feeds the output (x[:, 0] after the final LayerNorm), so attention is
// evaluated for that single query per (image, head) and out_proj / MLP run on B rows instead of B * S.
// One warp per (image, head): lane j scores keys j, j+32, ...; softmax over the warp; lane pair d accumulates P.V.
// =====================================================================================================
// Parallelism: one warp per (image, head), 4 warps per CTA, so B * heads / 4 CTAs spread over all SMs (the first version ran one
// 12-warp CTA per image: 127 CTAs, 2 heads in sequence on a third of the warps -- fine once per forward, but DINOv2's 257-token
// attention calls it in every layer, where it cost as much as the tensor-core kernel beside it).
__global__ void __launch_bounds__(128)
cls_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int S, int heads, int out_row_stride, int n_jobs) {
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int lane = threadIdx.x & 31;
    const int D = heads * HD;
    __shared__ float s_p[4][288];
    const int job = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (job < n_jobs) {
        const int b = job / heads, h = job - b * heads;
        float* sp = s_p[threadIdx.x >> 5];
        const __half* base = qkv + static_cast<int64_t>(b) * S * 3 * D + h * HD;
        // q (class-token row) -> every lane holds all 64 values as 32 half2
        __half2 q2[32];
        const uint4* qv = reinterpret_cast<const uint4*>(base);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint4 u = __ldg(qv + i);
            q2[4 * i + 0] = *reinterpret_cast<const __half2*>(&u.x); q2[4 * i + 1] = *reinterpret_cast<const __half2*>(&u.y);
            q2[4 * i + 2] = *reinterpret_cast<const __half2*>(&u.z); q2[4 * i + 3] = *reinterpret_cast<const __half2*>(&u.w);
        }
        float sc[9];
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int key = lane + 32 * i;
            float acc = -INFINITY;
            if (key < S) {
                const uint4* kv = reinterpret_cast<const uint4*>(base + static_cast<int64_t>(key) * 3 * D + D);
                acc = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint4 u = __ldg(kv + c);
                    const __half2 k0 = *reinterpret_cast<const __half2*>(&u.x), k1 = *reinterpret_cast<const __half2*>(&u.y);
                    const __half2 k2 = *reinterpret_cast<const __half2*>(&u.z), k3 = *reinterpret_cast<const __half2*>(&u.w);
                    const float2 a0 = __half22float2(q2[4 * c]), b0 = __half22float2(k0);
                    const float2 a1 = __half22float2(q2[4 * c + 1]), b1 = __half22float2(k1);
                    const float2 a2 = __half22float2(q2[4 * c + 2]), b2 = __half22float2(k2);
                    const float2 a3 = __half22float2(q2[4 * c + 3]), b3 = __half22float2(k3);
                    acc += a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
                }
            }
            sc[i] = acc;
            m = fmaxf(m, acc);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float scale_log2 = 0.125f * 1.44269504088896340736f;
        float l = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int key = lane + 32 * i;
            // P is rounded to fp16 before P.V exactly like the tensor-core kernels do
            const float p = key < S ? __half2float(__float2half_rn(exp2f((sc[i] - m) * scale_log2))) : 0.f;
            const float pe = key < S ? exp2f((sc[i] - m) * scale_log2) : 0.f;
            l += pe;
            if (key < 288) sp[key] = p;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
        __syncwarp();
        // O[d] = sum_j p_j V[j][d]; lane owns d = 2 lane, 2 lane + 1
        float o0 = 0.f, o1 = 0.f;
        const __half* vbase = base + 2 * D + 2 * lane;
        int key = 0;
        for (; key + 8 <= S; key += 8) {   // 8 independent 128 B row loads in flight per warp; the sum stays in key order
            __half2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const __half2*>(vbase + static_cast<int64_t>(key + u) * 3 * D));
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float2 f = __half22float2(v[u]);
                const float p = sp[key + u];
                o0 = fmaf(p, f.x, o0);
                o1 = fmaf(p, f.y, o1);
            }
        }
        for (; key < S; ++key) {
            const float2 v = __half22float2(*reinterpret_cast<const __half2*>(vbase + static_cast<int64_t>(key) * 3 * D));
            const float p = sp[key];
            o0 = fmaf(p, v.x, o0);
            o1 = fmaf(p, v.y, o1);
        }
        const float inv = 1.0f / l;
        *reinterpret_cast<__half2*>(out + static_cast<int64_t>(b) * out_row_stride * D + h * HD + 2 * lane) = __floats2half2_rn(o0 * inv, o1 * inv);
        __syncwarp();
    }
}

// dst[b, :] = src[b * row_stride_rows, :]   (fp32 rows of D floats; picks the class-token row of every image)
__global__ void gather_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int n_rows, int64_t src_row_stride, int D) {
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * (D / 4)) return;
    const int b = i / (D / 4), c = i - b * (D / 4);
    reinterpret_cast<float4*>(dst + static_cast<int64_t>(b) * D)[c] = reinterpret_cast<const float4*>(src + b * src_row_stride)[c];
}

}  // namespace

int ap_cls_attention_run(ap_ctx* ctx, const __half* qkv, __half* out, int B, int S, int heads, int out_row_stride, cudaStream_t stream) {
    AP_REQUIRE(ctx, S >= 1 && S <= 288, "cls attention: S=%d unsupported (<= 288)", S);
    if (B == 0) return AP_OK;
    ProfScope prof(ctx, stream, AP_K_ATTENTION);
    const int n_jobs = B * heads;
    AP_CHECK_CUDA(ctx, ap_launch_pdl(cls_attention_kernel, dim3((n_jobs + 3) / 4), dim3(128), 0, stream, 1, ctx->pdl != 0, qkv, out, S, heads,
                                     out_row_stride, n_jobs));
    AP_CHECK_LAUNCH(ctx, "cls_attention_kernel");
    return AP_OK;
}

int ap_gather_rows_run(ap_ctx* ctx, const float* src, float* dst, int n_rows, int64_t src_row_stride, int D, cudaStream_t stream) {
    if (n_rows == 0) return AP_OK;
    const int total = n_rows * (D / 4);
    ProfScope prof(ctx, stream, AP_K_OTHER);
    AP_CHECK_CUDA(ctx, ap_launch_pdl(gather_rows_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, 1, ctx->pdl != 0, src, dst, n_rows,
                                     src_row_stride, D));
    AP_CHECK_LAUNCH(ctx, "gather_rows_kernel");
    return AP_OK;
}

int ap_preprocess_run(ap_ctx* ctx, const uint8_t* slide, int64_t W, int64_t H, int64_t pitch, const int32_t* coords,
                      int64_t n, int input_patch, int image, int patch, __half* out, int64_t out_row_stride,
                      const int* centre, int dup, const int32_t* lin_s, const int16_t* lin_w, cudaStream_t stream) {
    AP_REQUIRE(ctx, (lin_s == nullptr) == (lin_w == nullptr), "preprocess: resize tables must be given together");
    AP_REQUIRE(ctx, patch == 16 || patch == 32, "preprocess: conv patch %d unsupported (16 or 32)", patch);
    AP_REQUIRE(ctx, image % patch == 0 && input_patch >= image, "preprocess: bad geometry input %d image %d patch %d",
               input_patch, image, patch);
    if (n == 0) return AP_OK;
    const int g = image / patch;
    const size_t smem = static_cast<size_t>(patch) * image * 3;
    ProfScope prof(ctx, stream, AP_K_PREPROCESS);
    if (patch == 16)
        AP_CHECK_CUDA(ctx, ap_launch_pdl(preprocess_kernel<16>, dim3(static_cast<unsigned>(n * g)), dim3(256), smem, stream, 1, ctx->pdl != 0, slide,
                                         W, H, pitch, coords, input_patch, image, out, out_row_stride,
                                         make_int3(centre[0], centre[1], centre[2]), dup, lin_s, lin_w));
    else   // vit_b_32 / vit_l_32: 7 x 7 tokens of 32 x 32 pixels
        AP_CHECK_CUDA(ctx, ap_launch_pdl(preprocess_kernel<32>, dim3(static_cast<unsigned>(n * g)), dim3(256), smem, stream, 1, ctx->pdl != 0, slide,
                                         W, H, pitch, coords, input_patch, image, out, out_row_stride,
                                         make_int3(centre[0], centre[1], centre[2]), dup, lin_s, lin_w));
    AP_CHECK_LAUNCH(ctx, "preprocess_kernel");
    return AP_OK;
}

// OpenCV's INTER_LINEAR tap tables for an n_src -> n_dst downscale (resize.cpp: resizeGeneric_ / HResizeLinear, 8-bit path):
//   fx = (float)((d + 0.5) * scale - 0.5), scale = 1 / ((double)n_dst / n_src);  s = floor(fx);  fx -= s;
//   coefficients saturate_cast<short>((1 - fx) * 2048), saturate_cast<short>(fx * 2048)   (cvRound = round half to even)
// taps[2 d] = s, taps[2 d + 1] = s + 1; weights[2 d], weights[2 d + 1].  For n_src > n_dst no tap leaves [0, n_src), so the x rule
// (fx zeroed at the border) and the y rule (row index clamped) coincide and one table serves both axes.
int ap_build_linear_tables(ap_ctx* ctx, int n_src, int n_dst, std::vector<int32_t>& taps, std::vector<int16_t>& weights) {
    AP_REQUIRE(ctx, n_dst > 0 && n_src > n_dst, "linear resize tables: only down-scaling is built (%d -> %d)", n_src, n_dst);
    const double scale = 1.0 / (static_cast<double>(n_dst) / n_src);
    taps.resize(2 * static_cast<size_t>(n_dst));
    weights.resize(2 * static_cast<size_t>(n_dst));
    for (int d = 0; d < n_dst; ++d) {
        float fx = static_cast<float>((d + 0.5) * scale - 0.5);
        int sx = static_cast<int>(floorf(fx));
        fx -= sx;
        AP_REQUIRE(ctx, sx >= 0 && sx + 1 <= n_src - 1, "linear resize tables: tap %d of %d -> %d leaves the source", d, n_src, n_dst);
        taps[2 * d] = sx;
        taps[2 * d + 1] = sx + 1;
        weights[2 * d] = static_cast<int16_t>(lrintf((1.f - fx) * 2048.f));
        weights[2 * d + 1] = static_cast<int16_t>(lrintf(fx * 2048.f));
    }
    return AP_OK;
}

int ap_cls_rows_run(ap_ctx* ctx, float* x, const float* cls, const float* pos, const float* regs, int lead, int n_images, int tokens, int D, __half* xh,
                    float2* stats, int parts, cudaStream_t stream) {
    if (n_images == 0) return AP_OK;
    AP_REQUIRE(ctx, lead >= 1 && (lead == 1 || regs != nullptr), "cls rows: %d leading rows need register tokens", lead);
    if (parts <= 0) parts = D % 128 == 0 ? D / 128 : 1;
    AP_REQUIRE(ctx, D % parts == 0 && parts <= 32, "cls rows: D=%d parts=%d unsupported", D, parts);
    ProfScope prof(ctx, stream, AP_K_OTHER);
    AP_CHECK_CUDA(ctx, ap_launch_pdl(cls_rows_kernel, dim3(n_images * lead), dim3(parts * 32), 0, stream, 1, ctx->pdl != 0, x, cls, pos, regs, lead, n_images,
                                     tokens, D, xh, stats, parts));
    AP_CHECK_LAUNCH(ctx, "cls_rows_kernel");
    return AP_OK;
}

int ap_layernorm_run(ap_ctx* ctx, const float* x, int64_t x_row_stride, const float* gamma, const float* beta, float eps,
                     __half* y_f16, float* y_f32, int rows, int D, cudaStream_t stream, int y_ld, int split_lo) {
    if (y_ld <= 0) y_ld = D;
    AP_REQUIRE(ctx, y_ld % 4 == 0 && y_ld >= (split_lo ? 2 * D : D), "layernorm: output row stride %d too small", y_ld);
    AP_REQUIRE(ctx, D % 128 == 0 && D <= 1536, "layernorm: D=%d unsupported (multiple of 128, <= 1536)", D);
    AP_REQUIRE(ctx, x_row_stride % 4 == 0, "layernorm: row stride %lld not a multiple of 4", (long long)x_row_stride);
    if (rows == 0) return AP_OK;
    const int blocks = (rows + 7) / 8;
    ProfScope prof(ctx, stream, AP_K_LAYERNORM);
#define AP_LN_CASE(V)                                                                                               \
    case V:                                                                                                         \
        AP_CHECK_CUDA(ctx, ap_launch_pdl(layernorm_kernel<V>, dim3(blocks), dim3(256), 0, stream, 1, ctx->pdl != 0, x, x_row_stride, gamma, \
                                         beta, eps, y_f16, y_f32, rows, y_ld, split_lo));                                  \
        break;
    switch (D / 128) {
        AP_LN_CASE(1) AP_LN_CASE(2) AP_LN_CASE(3) AP_LN_CASE(4) AP_LN_CASE(5) AP_LN_CASE(6) AP_LN_CASE(8) AP_LN_CASE(10)
        AP_LN_CASE(12)
        default: return ap_set_error(ctx, AP_EINVAL, "layernorm: D=%d not instantiated", D);
    }
#undef AP_LN_CASE
    AP_CHECK_LAUNCH(ctx, "layernorm_kernel");
    return AP_OK;
}

int ap_cls_mean_pool_run(ap_ctx* ctx, const float* x, int n_images, int tokens1, int lead, int D, const float* gamma, const float* beta, float eps,
                         float* out, cudaStream_t stream) {
    AP_REQUIRE(ctx, D % 128 == 0 && D <= 1536, "cls_mean pool: D=%d unsupported (multiple of 128, <= 1536)", D);
    AP_REQUIRE(ctx, lead >= 1 && tokens1 > lead, "cls_mean pool: needs at least one patch token (tokens %d, %d leading)", tokens1, lead);
    if (n_images == 0) return AP_OK;
    ProfScope prof(ctx, stream, AP_K_LAYERNORM);
#define AP_POOL_CASE(V)                                                                                                          \
    case V:                                                                                                                      \
        AP_CHECK_CUDA(ctx, ap_launch_pdl(cls_mean_pool_kernel<V>, dim3(n_images), dim3(256), 0, stream, 1, ctx->pdl != 0, x, tokens1, lead, gamma, \
                                         beta, eps, out));                                                                       \
        break;
    switch (D / 128) {
        AP_POOL_CASE(1) AP_POOL_CASE(2) AP_POOL_CASE(3) AP_POOL_CASE(4) AP_POOL_CASE(5) AP_POOL_CASE(6) AP_POOL_CASE(8) AP_POOL_CASE(10)
        AP_POOL_CASE(12)
        default: return ap_set_error(ctx, AP_EINVAL, "cls_mean pool: D=%d not instantiated", D);
    }
#undef AP_POOL_CASE
    AP_CHECK_LAUNCH(ctx, "cls_mean_pool_kernel");
    return AP_OK;
}

int ap_attention_run(ap_ctx* ctx, const __half* qkv, __half* out, int B, int S, int heads, cudaStream_t stream) {
    AP_REQUIRE(ctx, S >= 1 && S <= 272, "attention: S=%d unsupported (<= 272)", S);
    if (B == 0) return AP_OK;
    const int S_pad = (S + 15) / 16 * 16;
    const size_t smem = static_cast<size_t>(S_pad) * 128 * 3;
    static PerDeviceOnce attr;
    if (attr.need(ctx->device)) {
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 272 * 128 * 3));
        attr.done(ctx->device);
    }
    ProfScope prof(ctx, stream, AP_K_ATTENTION);
    AP_CHECK_CUDA(ctx, ap_launch_pdl(attention_kernel, dim3(heads, B), dim3(ATT_THREADS), smem, stream, 1, ctx->pdl != 0, qkv, out, S, S_pad,
                                     heads));
    AP_CHECK_LAUNCH(ctx, "attention_kernel");
    return AP_OK;
}

extern "C" int ap_layernorm_f16(ap_ctx* ctx, const float* x_dev, int64_t x_row_stride, const float* gamma_dev,
                                const float* beta_dev, float eps, void* y_dev, int rows, int D, void* stream) {
    if (!ctx) return AP_EINVAL;
    DeviceGuard guard(ctx);
    return ap_layernorm_run(ctx, x_dev, x_row_stride, gamma_dev, beta_dev, eps, static_cast<__half*>(y_dev), nullptr, rows, D,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int ap_attention_f16(ap_ctx* ctx, const void* qkv_dev, void* out_dev, int B, int S, int heads, void* stream) {
    if (!ctx) return AP_EINVAL;
    DeviceGuard guard(ctx);
    if (ctx->attn_mode == 2 && S >= 1 && S <= 257 && B > 0) {
        AttnPlan plan;
        int rc = ap_attention_tc_plan(ctx, &plan, static_cast<const __half*>(qkv_dev), B * S, S, heads);
        if (rc) return rc;
        return ap_attention_tc_run(ctx, &plan, static_cast<__half*>(out_dev), B, S, heads, static_cast<cudaStream_t>(stream));
    }
    return ap_attention_run(ctx, static_cast<const __half*>(qkv_dev), static_cast<__half*>(out_dev), B, S, heads,
                            static_cast<cudaStream_t>(stream));
}

extern "C" int ap_linear_tap_tables(int n_src, int n_dst, int32_t* taps, int16_t* weights) {
    if (!taps || !weights) return AP_EINVAL;
    std::vector<int32_t> t;
    std::vector<int16_t> w;
    int rc = ap_build_linear_tables(nullptr, n_src, n_dst, t, w);
    if (rc) return rc;
    std::copy(t.begin(), t.end(), taps);
    std::copy(w.begin(), w.end(), weights);
    return AP_OK;
}
