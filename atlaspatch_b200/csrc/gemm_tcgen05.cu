// Persistent, warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   out[M,N] = epilogue( A[M,K] (fp16, K contiguous) x W[N,K]^T (fp16, K contiguous) )     fp32 accumulate in TMEM
//
// This is the dense contraction behind every Linear of the encoder forward the reference delegates to
// torch (atlas_patch/models/patch/base.py:100 -> torchvision VisionTransformer): conv_proj (as an
// im2col GEMM), in_proj (QKV), out_proj, mlp.0 (+GELU), mlp.3 (+residual).
//
// Two instantiations of one kernel:
//   CG = 2 (default when N % 256 == 0): a CTA PAIR (cluster 2x1x1, tcgen05 cta_group::2) owns a 256 x 256 output tile.
//       Each CTA TMA-loads its 128 rows of A and its 128 rows of W per 64-wide K block (32 KB / stage, 6 stages); the
//       leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16) which reads both
//       CTAs' shared memory and writes 128 accumulator rows into each CTA's TMEM.  L2 -> SM operand traffic per FLOP
//       is 1.5x lower than with 128 x 256 single-CTA tiles, which is what bounded the first version.
//   CG = 1: one CTA owns a 128 x BN tile (BN = 256 or 128), 4 stages.  Used when N % 256 != 0.
// CTA = 384 threads, one CTA per SM, persistent over tiles (static round-robin over CTAs / CTA pairs):
//   warp 0     TMA producer (128B-swizzled boxes, mbarrier ring)
//   warp 1     MMA issuer: one thread; tcgen05.commit frees the smem stage / publishes the accumulator
//   warp 2     TMEM allocator (2 x BN fp32 columns: double-buffered accumulator)
//   warps 4-11 epilogue: tcgen05.ld 32 lanes x 32 columns (software-pipelined) -> bias / GELU / residual /
//              positional embedding -> vectorised global stores; overlaps the next tile's MMAs
#include "ap_internal.cuh"
#include "ptx.cuh"

namespace {

constexpr int BK = 64;  // 64 fp16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;
constexpr int NUM_EPI_WARPS = 8;

template <int CG, int BN>
struct SmemLayout {
    static constexpr int STAGES = CG == 2 ? 6 : 4;
    static constexpr int A_STAGE = 128 * BK * 2;        // 16 KB: this CTA's 128 rows of A
    static constexpr int B_ROWS = BN / CG;              // rows of W this CTA loads
    static constexpr int B_STAGE = B_ROWS * BK * 2;
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = STAGES * A_STAGE;
    static constexpr int BAR_OFF = B_OFF + STAGES * B_STAGE;
    static constexpr int NUM_BARS = 2 * STAGES + 4;
    static constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
    static constexpr int STAGING_OFF = (TMEM_PTR_OFF + 16 + 1023) / 1024 * 1024;  // 8 epilogue warps x (32 rows x 128 B); 1 KB-aligned
                                                                                   // tiles: the TMA store's 128B swizzle is address-based
    static constexpr int STAGING_BYTES = 8 * 32 * 128;
    static constexpr int TOTAL = STAGING_OFF + STAGING_BYTES;
    static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024 B alignment
};

struct EpiParams {
    const float* bias;
    const float* resid;
    void* out;
    const float* pos;
    int tokens_per_image;
    int lead_tokens;   // class + register tokens in front of each image's patch tokens (1 without registers)
    float alpha;
    int seg_blocks;  // 64-wide K blocks per segment; block kb of the contraction is block kb % seg_blocks of segment kb / seg_blocks
    int a_map, w_map;  // 2 bits per segment: which Kseg-wide column range of A / W that segment reads (split operands, GemmPlan)
    int debug;  // diagnostics only (ap_set_option "gemm_debug"): 1 = no epilogue math/stores, 2 = no MMA issue, 4 = no TMA loads
    // LayerNorm folding (GemmExtra in ap_internal.cuh)
    __half* out_h;
    float2* stats_out;
    const float2* stats_in;
    int ln_parts;
    float ln_inv_dim, ln_eps;
    // AP_EPI_BIAS_F32 only (the SAM2 linears): output row stride, number of real columns (<= N), activation (0 none, 1 GELU erf, 2 ReLU)
    int out_ld, n_valid, act;
};

// Per-row scale of the consumer epilogues: with gamma folded into the weights AND their rows centred (sum_k W''[n, k] = 0, which
// absorbs the mean: x W''^T = (x - mean) W'^T), the folded LayerNorm is v = acc * rstd + bias'.
struct LnRow {
    float rstd;
};
__device__ __forceinline__ LnRow ln_row_coeffs(const EpiParams& ep, int row, int M) {
    LnRow c{ep.alpha};
    if (ep.stats_in == nullptr || row >= M) return c;
    float s = 0.f, q = 0.f;
    const float2* p = ep.stats_in + static_cast<int64_t>(row) * ep.ln_parts;
    for (int j = 0; j < ep.ln_parts; ++j) {   // fixed order: bitwise reproducible
        const float2 t = __ldg(p + j);
        s += t.x;
        q += t.y;
    }
    const float mean = s * ep.ln_inv_dim;
    const float var = fmaxf(q * ep.ln_inv_dim - mean * mean, 0.f);
    c.rstd = rsqrtf(var + ep.ln_eps);
    return c;
}

// ---- packed fp32x2 arithmetic (sm_100 FFMA2 / FMUL2): halves the issue slots of the epilogue math ----------------
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Exact-erf GELU (torch.nn.GELU() default, as in torchvision's MLPBlock) for two values at once.
//   gelu(x) = x Phi(x),  Phi(x) = 1 - t for x >= 0,  t for x < 0,  t = 0.5 erfc(|x| / sqrt 2)
//           = max(x, 0) - |x| t
//   t = 2^Q(|x|): Q = degree-6 polynomial fit of log2(0.5 erfc(|x|/sqrt 2)) on [0, 6] (|x| clamped to 6, t(6) = 1e-9),
//   weighted by the sensitivity |x| t of the result.  Evaluated in fp32 against float64 erf on [-6, 6] (fit script in
//   DESIGN.md): max abs error of gelu 9.2e-8, never more than 0.11 of an fp16 half-ulp of the result (+1e-7).
// Cost per element: 7 FMA-pipe ops (packed as FFMA2) + 1 MUFU.EX2 + 4 ALU-pipe ops, against ~30 issue slots for libdevice
// erff, which made the mlp.0 GEMM epilogue-bound (166 us with erff, 110 us with an A&S 7.1.28 variant, 75 us without
// any epilogue).
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
    const float a0 = __uint_as_float(__float_as_uint(x0) & 0x7fffffffu), a1 = __uint_as_float(__float_as_uint(x1) & 0x7fffffffu);
    const uint64_t z = pk2(fminf(a0, 6.0f), fminf(a1, 6.0f));
    uint64_t q = fma2(pk2(3.4089207474607974e-05f, 3.4089207474607974e-05f), z, pk2(-0.0007762229652144015f, -0.0007762229652144015f));
    q = fma2(q, z, pk2(0.008098662830889225f, 0.008098662830889225f));
    q = fma2(q, z, pk2(-0.05343286693096161f, -0.05343286693096161f));
    q = fma2(q, z, pk2(-0.4587600827217102f, -0.4587600827217102f));
    q = fma2(q, z, pk2(-1.151203989982605f, -1.151203989982605f));
    q = fma2(q, z, pk2(-0.9999929070472717f, -0.9999929070472717f));
    float q0, q1;
    upk2(q, q0, q1);
    const uint64_t nabs = pk2(__uint_as_float(__float_as_uint(x0) | 0x80000000u), __uint_as_float(__float_as_uint(x1) | 0x80000000u));
    upk2(fma2(nabs, pk2(ex2_approx(q0), ex2_approx(q1)), pk2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f))), x0, x1);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// Epilogue of one warp = 32 accumulator rows x (BN/2) columns, processed in 128-byte-per-row groups.
// tcgen05.ld hands every lane ONE ROW (32 consecutive fp32 columns); storing that directly would make each warp
// store instruction touch 32 different 128 B lines (the first version did, and ncu showed the GEMMs epilogue-bound:
// tensor pipe 20-54 % active).  So each group is transposed through a per-warp 4 KB shared-memory tile
// (32 rows x 128 B, 16-byte pieces XOR-swizzled by row -> conflict-free both ways) and written / residual-read with
// fully coalesced accesses: 8 lanes cover one 128 B row segment, one warp instruction covers 4 rows.
__device__ __forceinline__ uint32_t stg_off(int row, int piece) { return static_cast<uint32_t>(row * 128 + ((piece ^ (row & 7)) << 4)); }

// The bias / column-sum values of a warp's COLS_PER_WARP columns are the same for all of its 32 rows.  Loading them with one
// LDG per lane and chunk put a global-load latency (long scoreboard) in front of every chunk's math (ncu: ~1/3 of the epilogue
// warps' stall samples); instead each lane fetches 4 of the 128 values once per tile -- before the accumulator is ready -- and the
// chunks pick them up with warp shuffles.
__device__ __forceinline__ float4 bcast4(const float4& v, int src_lane) {
    return make_float4(__shfl_sync(0xffffffffu, v.x, src_lane), __shfl_sync(0xffffffffu, v.y, src_lane),
                       __shfl_sync(0xffffffffu, v.z, src_lane), __shfl_sync(0xffffffffu, v.w, src_lane));
}

// fp32 output (residual add / positional embedding): one 32-column chunk = 128 B per row.
// The residual rows are fetched (coalesced) one chunk AHEAD -- the first chunk even before the accumulator is ready, i.e.
// under the tile's MMAs -- because the epilogue of out_proj / mlp.3 is bound by the latency of these loads, not by
// bandwidth.  In-place use (resid == out) is fine: every element is read and written by the same lane, reads first.
template <int EPI>
__device__ __forceinline__ void resid_prefetch(float4 (&rr)[8], const EpiParams& ep, int M, int N, int row_base, int col0, int lane) {
    if (EPI != AP_EPI_BIAS_RESID_F32) return;
    const int p = lane & 7;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int row = row_base + it * 4 + (lane >> 3);
        rr[it] = row < M ? *(reinterpret_cast<const float4*>(ep.resid + static_cast<int64_t>(row) * N + col0) + p)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {   // explicit shared-space load: the generic LD the compiler emitted for the
    float4 v;                                               // staging reads is a long-scoreboard access
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// ncu (source view) showed this epilogue bound by single-warp instruction latency: ~43 instructions per 4-row step in eight
// separate basic blocks (row < M and "fold" branches), each starting with a staging load whose latency nothing overlapped.  Now all
// eight staged rows are fetched first, the arithmetic is branch-free and only the stores are predicated.
template <int EPI>
__device__ __forceinline__ void epilogue_group_f32(const uint32_t (&r)[32], const float4 (&rr)[8], uint8_t* stg, const EpiParams& ep,
                                                   int M, int N, int row_base, int col0, int lane, float (&rs)[8], float (&rq)[8],
                                                   const float4& bias_reg, int c) {
    const int p = lane & 7;
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(stg + stg_off(lane, q)) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    __syncwarp();
    const float4 bb = bcast4(bias_reg, c * 8 + p);   // this lane's 4 columns of chunk c
    const uint32_t stg_u = ptx::smem_u32(stg);
    float4 v[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) v[it] = lds128(stg_u + stg_off(it * 4 + (lane >> 3), p));
    const bool fold = ep.out_h != nullptr;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int row = row_base + it * 4 + (lane >> 3);
        v[it] = make_float4(fmaf(v[it].x, ep.alpha, bb.x), fmaf(v[it].y, ep.alpha, bb.y), fmaf(v[it].z, ep.alpha, bb.z),
                            fmaf(v[it].w, ep.alpha, bb.w));
        int64_t out_row = row;
        if (EPI == AP_EPI_BIAS_RESID_F32) {
            v[it].x += rr[it].x; v[it].y += rr[it].y; v[it].z += rr[it].z; v[it].w += rr[it].w;
        } else if (ep.tokens_per_image > 0) {   // conv_proj: token row -> sequence row (class [+ register] tokens lead each image), + pos
            const int b = row / ep.tokens_per_image;
            const int tk = row - b * ep.tokens_per_image;
            out_row = static_cast<int64_t>(b) * (ep.tokens_per_image + ep.lead_tokens) + ep.lead_tokens + tk;
            if (ep.pos != nullptr && row < M) {
                const float4 pp = __ldg(reinterpret_cast<const float4*>(ep.pos + static_cast<int64_t>(1 + tk) * N + col0) + p);
                v[it].x += pp.x; v[it].y += pp.y; v[it].z += pp.z; v[it].w += pp.w;
            }
        }
        if (EPI == AP_EPI_BIAS_F32 && ep.act != 0) {   // fp32 activations of the SAM2 path (exact erf GELU / ReLU)
            if (ep.act == 1) {      // the packed-FMA GELU of the fp16 epilogues: max abs error 9.2e-8 against float64 erf
                gelu_erf2(v[it].x, v[it].y);
                gelu_erf2(v[it].z, v[it].w);
            } else {
                v[it] = make_float4(fmaxf(v[it].x, 0.f), fmaxf(v[it].y, 0.f), fmaxf(v[it].z, 0.f), fmaxf(v[it].w, 0.f));
            }
        }
        if (EPI == AP_EPI_BIAS_F32) {     // row stride and column bound of the destination may differ from the (padded) GEMM N
            if (row < M && col0 + p * 4 < ep.n_valid)
                *(reinterpret_cast<float4*>(static_cast<float*>(ep.out) + out_row * ep.out_ld + col0) + p) = v[it];
        } else if (row < M) {
            *(reinterpret_cast<float4*>(static_cast<float*>(ep.out) + out_row * N + col0) + p) = v[it];
        }
        if (fold) {   // the next GEMM's A operand: raw x in fp16 (its LayerNorm is finished in that GEMM's epilogue) + row statistics
            if (row < M)
                *reinterpret_cast<uint2*>(ep.out_h + out_row * N + col0 + p * 4) =
                    make_uint2(pack_half2(v[it].x, v[it].y), pack_half2(v[it].z, v[it].w));
            rs[it] += (v[it].x + v[it].y) + (v[it].z + v[it].w);   // rows >= M accumulate values that are never written out
            rq[it] += (v[it].x * v[it].x + v[it].y * v[it].y) + (v[it].z * v[it].z + v[it].w * v[it].w);
        }
    }
    __syncwarp();
}

// After the last chunk of a tile: the 8 lanes that share a row combine their partial sums (fixed shuffle tree: reproducible) and
// one of them writes the (sum, sum of squares) of this warp's COLS_PER_WARP columns of that row.
__device__ __forceinline__ void epilogue_write_stats(const EpiParams& ep, int M, int row_base, int col_block, int lane, float (&rs)[8],
                                                     float (&rq)[8]) {
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        float s = rs[it], q = rq[it];
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, o);
            q += __shfl_xor_sync(0xffffffffu, q, o);
        }
        const int row = row_base + it * 4 + (lane >> 3);
        if ((lane & 7) == 0 && row < M) {
            int64_t out_row = row;
            if (ep.tokens_per_image > 0) {
                const int b = row / ep.tokens_per_image;
                out_row = static_cast<int64_t>(b) * (ep.tokens_per_image + ep.lead_tokens) + ep.lead_tokens + (row - b * ep.tokens_per_image);
            }
            ep.stats_out[out_row * ep.ln_parts + col_block] = make_float2(s, q);
        }
        rs[it] = 0.f;
        rq[it] = 0.f;
    }
}

// fp16 output: bias (+GELU) in the row-per-lane layout, packed halves staged; `half_sel` = which 64 B half of the
// 128 B row this 32-column chunk fills.  Call flush after both halves.
template <int EPI>
__device__ __forceinline__ void epilogue_stage_f16(const uint32_t (&r)[32], uint8_t* stg, int c, int lane, int half_sel, const LnRow& ln,
                                                   const float4& bias_reg) {
    const uint64_t alpha2 = pk2(ln.rstd, ln.rstd);   // the constant alpha, or this row's 1 / sigma (folded LayerNorm)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 b0 = bcast4(bias_reg, c * 8 + 2 * j), b1 = bcast4(bias_reg, c * 8 + 2 * j + 1);
        float v[8];
        upk2(fma2(pk2(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1])), alpha2, pk2(b0.x, b0.y)), v[0], v[1]);
        upk2(fma2(pk2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3])), alpha2, pk2(b0.z, b0.w)), v[2], v[3]);
        upk2(fma2(pk2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5])), alpha2, pk2(b1.x, b1.y)), v[4], v[5]);
        upk2(fma2(pk2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7])), alpha2, pk2(b1.z, b1.w)), v[6], v[7]);
        if (EPI == AP_EPI_BIAS_GELU_F16) {
            gelu_erf2(v[0], v[1]); gelu_erf2(v[2], v[3]); gelu_erf2(v[4], v[5]); gelu_erf2(v[6], v[7]);
        }
        if (EPI == AP_EPI_BIAS_QGELU_F16) {   // CLIP's QuickGELU: x sigmoid(1.702 x) = x / (1 + 2^(-1.702 log2(e) x))
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = v[k] * rcp_approx(1.0f + ex2_approx(-2.4554669595930157f * v[k]));
        }
        *reinterpret_cast<uint4*>(stg + stg_off(lane, half_sel * 4 + j)) =
            make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
    }
}
// SwiGLU (Dinov2SwiGLUFFN: hidden = silu(x1) * x2): the chunk's first 16 columns are gates, the last 16 the matching values
// (weights_in rows are interleaved on the host), so 32 accumulator columns give 16 fp16 outputs = pieces 2c, 2c+1 of the row.
__device__ __forceinline__ void epilogue_stage_swiglu(const uint32_t (&r)[32], uint8_t* stg, int cc, int lane, int c, const LnRow& ln,
                                                      const float4& bias_reg) {
    const float scale = ln.rstd;
#pragma unroll
    for (int j = 0; j < 2; ++j) {   // cc = chunk index inside the warp's 128 columns: gates at cc*32 + 0..15, values at cc*32 + 16..31
        const float4 g0 = bcast4(bias_reg, cc * 8 + 2 * j), g1 = bcast4(bias_reg, cc * 8 + 2 * j + 1);
        const float4 v0 = bcast4(bias_reg, cc * 8 + 4 + 2 * j), v1 = bcast4(bias_reg, cc * 8 + 5 + 2 * j);
        const float gb[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float vb[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        float h[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float g = fmaf(__uint_as_float(r[8 * j + k]), scale, gb[k]);
            const float v = fmaf(__uint_as_float(r[16 + 8 * j + k]), scale, vb[k]);
            h[k] = g * v * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * g));  // silu(g) * v
        }
        *reinterpret_cast<uint4*>(stg + stg_off(lane, c * 2 + j)) =
            make_uint4(pack_half2(h[0], h[1]), pack_half2(h[2], h[3]), pack_half2(h[4], h[5]), pack_half2(h[6], h[7]));
    }
}

// The staged 32 rows x 128 B tile has exactly the layout of a 128B-swizzled TMA box (16-byte chunk index XOR row & 7), so the
// tile goes to global memory with ONE bulk tensor store issued by one lane instead of 8 x (LDS.128 + guarded STG.128) per thread;
// rows >= M are clipped by the tensor map.  The staging tile may be rewritten once the bulk group has been READ (wait_group.read).
__device__ __forceinline__ void epilogue_flush_tma(uint8_t* stg, const CUtensorMap* map_out, int row_base, int col0, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this lane's st.shared -> visible to the async proxy
    __syncwarp();
    if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map_out), "r"(ptx::smem_u32(stg)),
                     "r"(col0), "r"(row_base)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
}

template <int CG, int BN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_out, int M, int N, int K, EpiParams ep) {
    using L = SmemLayout<CG, BN>;
    constexpr int STAGES = L::STAGES;
    constexpr int TILE_M = 128 * CG;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA      (CG=2: the leader CTA's copy is used)
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA      (CG=2: multicast commit to both CTAs)
    uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]       MMA -> epilogue (CG=2: multicast commit to both CTAs)
    uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA (CG=2: both CTAs arrive on the leader's)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::TMEM_PTR_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int cta_rank = CG == 2 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
    const int worker = blockIdx.x / CG;
    const int num_workers = gridDim.x / CG;
    const int tiles_n = N / BN;
    const int tiles_m = (M + TILE_M - 1) / TILE_M;
    const int num_tiles = tiles_m * tiles_n;
    const int k_blocks = K / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&map_a);
        ptx::prefetch_tmap(&map_w);
        ptx::prefetch_tmap(&map_out);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], NUM_EPI_WARPS * CG);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc<CG>(tmem_ptr_smem, 2 * BN);
        ptx::tmem_relinquish<CG>();
    }
    ptx::tc_fence_before();
    if (CG == 2) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    ptx::pdl_wait();                 // everything above overlapped the previous kernel's tail; its outputs are visible from here on
    ptx::pdl_launch_dependents();

    if (warp == 0) {
        // ================= TMA producer (every CTA loads its own 128 rows of A and BN/CG rows of W) =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = worker; tile < num_tiles; tile += num_workers) {
                const int m0 = (tile / tiles_n) * TILE_M + cta_rank * 128;
                const int n0 = (tile % tiles_n) * BN + cta_rank * L::B_ROWS;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    const int seg = kb / ep.seg_blocks, blk = kb - seg * ep.seg_blocks;
                    const int ka = (((ep.a_map >> (2 * seg)) & 3) * ep.seg_blocks + blk) * BK;
                    const int kw = (((ep.w_map >> (2 * seg)) & 3) * ep.seg_blocks + blk) * BK;
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    if (ep.debug & 4) {
                        if (cta_rank == 0) ptx::mbar_arrive(&full_bar[stage]);
                    } else if (CG == 1) {
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], L::A_STAGE + L::B_STAGE);
                        ptx::tma_load_2d(smem + L::A_OFF + stage * L::A_STAGE, &map_a, &full_bar[stage], ka, m0);
                        ptx::tma_load_2d(smem + L::B_OFF + stage * L::B_STAGE, &map_w, &full_bar[stage], kw, n0);
                    } else {
                        // both CTAs' bytes are accounted on the leader's barrier (peer bit cleared in the address)
                        if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * (L::A_STAGE + L::B_STAGE));
                        ptx::tma_load_2d_2sm(smem + L::A_OFF + stage * L::A_STAGE, &map_a, &full_bar[stage], ka, m0);
                        ptx::tma_load_2d_2sm(smem + L::B_OFF + stage * L::B_STAGE, &map_w, &full_bar[stage], kw, n0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA only) =================
        if (lane == 0 && cta_rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(TILE_M, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                ptx::mbar_wait(&tempty_bar[as], aphase ^ 1, 2);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase, 3);
                    ptx::tc_fence_after();
                    const uint64_t a_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + L::A_OFF + stage * L::A_STAGE));
                    const uint64_t b_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem + L::B_OFF + stage * L::B_STAGE));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        // advance 32 B (16 fp16) inside the 128 B swizzle row: +2 in the (addr >> 4) field
                        if (!(ep.debug & 2)) ptx::tc_mma_f16<CG>(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::tc_commit<CG>(&empty_bar[stage]);  // smem stage reusable once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::tc_commit<CG>(&tfull_bar[as]);  // accumulator complete
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ================= epilogue (each CTA drains its own 128 accumulator rows) =================
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int half_idx = (warp - 4) >> 2;    // which half of the BN columns
        constexpr int COLS_PER_WARP = BN / 2;
        constexpr int NCH = COLS_PER_WARP / 32;
        uint8_t* stg = smem + L::STAGING_OFF + (warp - 4) * 4096;
        uint32_t tempty_remote[2] = {0, 0};
        if (CG == 2 && cta_rank != 0) {
            tempty_remote[0] = ptx::mapa_u32(ptx::smem_u32(&tempty_bar[0]), 0);
            tempty_remote[1] = ptx::mapa_u32(ptx::smem_u32(&tempty_bar[1]), 0);
        }
        int it = 0;
        for (int tile = worker; tile < num_tiles; tile += num_workers, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / tiles_n) * TILE_M + cta_rank * 128;
            const int n0 = (tile % tiles_n) * BN;
            const int row_base = m0 + q * 32;
            const int col_base = n0 + half_idx * COLS_PER_WARP;
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + half_idx * COLS_PER_WARP;
            auto release_tmem = [&]() {  // all of this warp's accumulator reads are complete: hand the TMEM stage back early
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2 && cta_rank != 0) ptx::mbar_arrive_remote_relaxed(tempty_remote[as]);
                    else ptx::mbar_arrive_relaxed(&tempty_bar[as]);
                }
            };
            if (EPI == AP_EPI_BIAS_F16 || EPI == AP_EPI_BIAS_GELU_F16 || EPI == AP_EPI_BIAS_QGELU_F16 || EPI == AP_EPI_BIAS_SWIGLU_F16) {
                // lane = accumulator row; its LayerNorm statistics are fetched (L2) while this tile's MMAs are still running
                const LnRow ln = ln_row_coeffs(ep, row_base + lane, M);
                const bool owns = lane * 4 < COLS_PER_WARP;   // BN = 128: 64 columns per warp, lanes 16..31 hold nothing
                const float4 bias_reg = owns ? __ldg(reinterpret_cast<const float4*>(ep.bias + col_base) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
                ptx::mbar_wait(&tfull_bar[as], aphase, 4);
                ptx::tc_fence_after();
                uint32_t r[2][32];
                ptx::tmem_ld_32x32(taddr0, r[0]);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    ptx::tc_wait_ld();
                    if (c + 1 < NCH) ptx::tmem_ld_32x32(taddr0 + (c + 1) * 32, r[(c + 1) & 1]);  // overlaps the math below
                    else release_tmem();
                    if (ep.debug & 1) continue;
                    if (EPI == AP_EPI_BIAS_SWIGLU_F16) {
                        epilogue_stage_swiglu(r[c & 1], stg, c, lane, c & 3, ln, bias_reg);
                        if ((c & 3) == 3) epilogue_flush_tma(stg, &map_out, row_base, (col_base + (c - 3) * 32) / 2, lane);
                        continue;
                    }
                    epilogue_stage_f16<EPI>(r[c & 1], stg, c, lane, c & 1, ln, bias_reg);
                    if (c & 1) epilogue_flush_tma(stg, &map_out, row_base, col_base + (c - 1) * 32, lane);
                }
            } else {
                float4 rr[2][8];
                float rs[8] = {}, rq[8] = {};   // LayerNorm statistics of this warp's rows (when the output feeds a folded LayerNorm)
                resid_prefetch<EPI>(rr[0], ep, M, N, row_base, col_base, lane);  // in flight while the MMAs of this tile run
                const float4 bias_reg = lane * 4 < COLS_PER_WARP ? __ldg(reinterpret_cast<const float4*>(ep.bias + col_base) + lane)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                ptx::mbar_wait(&tfull_bar[as], aphase, 4);
                ptx::tc_fence_after();
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    uint32_t r[32];
                    ptx::tmem_ld_32x32(taddr0 + c * 32, r);
                    if (c + 1 < NCH) resid_prefetch<EPI>(rr[(c + 1) & 1], ep, M, N, row_base, col_base + (c + 1) * 32, lane);
                    ptx::tc_wait_ld();
                    if (c + 1 == NCH) release_tmem();
                    if (ep.debug & 1) continue;
                    epilogue_group_f32<EPI>(r, rr[c & 1], stg, ep, M, N, row_base, col_base + c * 32, lane, rs, rq, bias_reg, c);
                }
                if (ep.stats_out != nullptr && !(ep.debug & 1)) epilogue_write_stats(ep, M, row_base, col_base / COLS_PER_WARP, lane, rs, rq);
            }
        }
    }

    ptx::tc_fence_before();
    if (CG == 2) ptx::cluster_sync(); else __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<CG>(tmem_base, 2 * BN);
    }
}

template <int CG, int BN, int EPI>
int launch(ap_ctx* ctx, const GemmPlan* p, const EpiParams& ep, cudaStream_t stream) {
    // fp16 outputs leave through TMA stores: box = 32 rows x 64 halfs (one staged tile), 128B swizzle, rows clipped at M
    CUtensorMap map_out{};
    if (EPI == AP_EPI_BIAS_F16 || EPI == AP_EPI_BIAS_GELU_F16 || EPI == AP_EPI_BIAS_QGELU_F16 || EPI == AP_EPI_BIAS_SWIGLU_F16) {
        const int n_out = EPI == AP_EPI_BIAS_SWIGLU_F16 ? p->N / 2 : p->N;
        int rc = ap_make_tmap_f16_2d(ctx, &map_out, ep.out, (uint64_t)p->M, (uint64_t)n_out, (uint64_t)n_out, 32, 64);
        if (rc) return rc;
    }
    using L = SmemLayout<CG, BN>;
    auto kern = gemm_tcgen05_kernel<CG, BN, EPI>;
    static PerDeviceOnce attr;  // per instantiation
    if (attr.need(ctx->device)) {
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
        attr.done(ctx->device);
    }
    const int tile_m = 128 * CG;
    const int tiles = ((p->M + tile_m - 1) / tile_m) * (p->N / BN);
    const int max_workers = ctx->sm_count / CG;
    const int workers = tiles < max_workers ? tiles : max_workers;
    ProfScope prof(ctx, stream, AP_K_GEMM, (static_cast<int64_t>(p->N) << 32) | (static_cast<int64_t>(p->K) << 4) | EPI);
    const CUtensorMap& mw = CG == 2 ? p->map_w_half : p->map_w;
    AP_CHECK_CUDA(ctx, ap_launch_pdl(kern, dim3(workers * CG), dim3(NUM_THREADS), L::DYN_BYTES, stream, CG, ctx->pdl != 0, p->map_a, mw,
                                     map_out, p->M, p->N, p->K, ep));
    AP_CHECK_LAUNCH(ctx, "gemm_tcgen05_kernel");
    return AP_OK;
}

template <int CG, int BN>
int dispatch_epi(ap_ctx* ctx, const GemmPlan* p, const EpiParams& ep, cudaStream_t stream) {
    switch (p->epilogue) {
        case AP_EPI_BIAS_F16: return launch<CG, BN, AP_EPI_BIAS_F16>(ctx, p, ep, stream);
        case AP_EPI_BIAS_GELU_F16: return launch<CG, BN, AP_EPI_BIAS_GELU_F16>(ctx, p, ep, stream);
        case AP_EPI_BIAS_QGELU_F16: return launch<CG, BN, AP_EPI_BIAS_QGELU_F16>(ctx, p, ep, stream);
        case AP_EPI_BIAS_RESID_F32: return launch<CG, BN, AP_EPI_BIAS_RESID_F32>(ctx, p, ep, stream);
        case AP_EPI_BIAS_F32: return launch<CG, BN, AP_EPI_BIAS_F32>(ctx, p, ep, stream);
        case AP_EPI_BIAS_SWIGLU_F16:
            if constexpr (BN == 256) return launch<CG, BN, AP_EPI_BIAS_SWIGLU_F16>(ctx, p, ep, stream);
            break;
    }
    return ap_set_error(ctx, AP_EINVAL, "gemm: unknown epilogue %d", p->epilogue);
}

}  // namespace

namespace {
int plan_common(ap_ctx* ctx, GemmPlan* plan, const void* A, const void* W, int M, int N, int epilogue, int Kw);
}

// K = width of W; Ka = width of A (K % Ka == 0): K > Ka means W holds K / Ka column segments (hi/lo split weights) that all
// multiply the same A.
int ap_gemm_plan(ap_ctx* ctx, GemmPlan* plan, const void* A, const void* W, int M, int N, int K, int epilogue, int Ka) {
    if (Ka <= 0) Ka = K;
    AP_REQUIRE(ctx, M > 0 && N > 0 && K > 0, "gemm: empty problem %dx%dx%d", M, N, K);
    AP_REQUIRE(ctx, K % BK == 0 && Ka % BK == 0 && K % Ka == 0 && K / Ka <= 3, "gemm: K=%d (A width %d) must be multiples of %d, K = 1..3 x Ka", K, Ka, BK);
    plan->segs = K / Ka; plan->Kseg = Ka; plan->K = K; plan->Ka = Ka;
    for (int i = 0; i < 3; ++i) { plan->a_seg[i] = 0; plan->w_seg[i] = i < plan->segs ? i : 0; }
    return plan_common(ctx, plan, A, W, M, N, epilogue, K);
}

int ap_gemm_plan_split(ap_ctx* ctx, GemmPlan* plan, const void* A, const void* W, int M, int N, int Kbase, int epilogue, int split) {
    AP_REQUIRE(ctx, M > 0 && N > 0 && Kbase > 0 && Kbase % BK == 0, "gemm: bad problem %dx%dx%d", M, N, Kbase);
    AP_REQUIRE(ctx, split >= AP_SPLIT_NONE && split <= AP_SPLIT_AW, "gemm: unknown split mode %d", split);
    static const int segs[4] = {1, 2, 2, 3};
    static const int am[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 1, 0}, {0, 1, 0}};   // A_hi W_hi + A_lo W_hi (+ A_hi W_lo)
    static const int wm[4][3] = {{0, 0, 0}, {0, 1, 0}, {0, 0, 0}, {0, 0, 1}};
    plan->segs = segs[split]; plan->Kseg = Kbase; plan->K = segs[split] * Kbase;
    plan->Ka = (split & AP_SPLIT_A) ? 2 * Kbase : Kbase;
    for (int i = 0; i < 3; ++i) { plan->a_seg[i] = am[split][i]; plan->w_seg[i] = wm[split][i]; }
    return plan_common(ctx, plan, A, W, M, N, epilogue, (split & AP_SPLIT_W) ? 2 * Kbase : Kbase);
}

namespace {
int plan_common(ap_ctx* ctx, GemmPlan* plan, const void* A, const void* W, int M, int N, int epilogue, int Kw) {
    const int K = plan->K, Ka = plan->Ka;
    AP_REQUIRE(ctx, N % 128 == 0, "gemm: N=%d must be a multiple of 128", N);
    AP_REQUIRE(ctx, epilogue >= 0 && epilogue <= 5, "gemm: unknown epilogue %d", epilogue);
    AP_REQUIRE(ctx, epilogue != AP_EPI_BIAS_SWIGLU_F16 || N % 256 == 0, "gemm: SwiGLU epilogue needs N %% 256 == 0 (N=%d)", N);
    (void)K;
    plan->M = M; plan->N = N; plan->epilogue = epilogue;
    plan->bn = (N % 256 == 0) ? 256 : 128;
    plan->cta_group = (plan->bn == 256 && ctx->gemm_cta_group == 2) ? 2 : 1;
    int rc = ap_make_tmap_f16_2d(ctx, &plan->map_a, A, (uint64_t)M, (uint64_t)Ka, (uint64_t)Ka, 128, BK);
    if (rc) return rc;
    rc = ap_make_tmap_f16_2d(ctx, &plan->map_w, W, (uint64_t)N, (uint64_t)Kw, (uint64_t)Kw, plan->bn, BK);
    if (rc) return rc;
    return ap_make_tmap_f16_2d(ctx, &plan->map_w_half, W, (uint64_t)N, (uint64_t)Kw, (uint64_t)Kw, plan->bn / 2, BK);
}
}  // namespace

int ap_gemm_run(ap_ctx* ctx, const GemmPlan* plan, const float* bias, const float* resid, void* out,
                const GemmExtra* extra, cudaStream_t stream) {
    AP_REQUIRE(ctx, bias != nullptr && out != nullptr, "gemm: bias/out must not be NULL");
    AP_REQUIRE(ctx, plan->epilogue != AP_EPI_BIAS_RESID_F32 || resid != nullptr, "gemm: residual epilogue needs resid");
    EpiParams ep;
    ep.bias = bias; ep.resid = resid; ep.out = out;
    ep.pos = extra ? extra->pos : nullptr;
    ep.tokens_per_image = extra ? extra->tokens_per_image : 0;
    ep.lead_tokens = extra ? extra->lead_tokens : 1;
    ep.alpha = extra ? extra->alpha : 1.0f;
    ep.debug = ctx->gemm_debug;
    ep.seg_blocks = plan->Kseg / BK;
    ep.a_map = plan->a_seg[0] | (plan->a_seg[1] << 2) | (plan->a_seg[2] << 4);
    ep.w_map = plan->w_seg[0] | (plan->w_seg[1] << 2) | (plan->w_seg[2] << 4);
    ep.out_h = extra ? extra->out_h : nullptr;
    ep.stats_out = extra ? extra->stats_out : nullptr;
    ep.stats_in = extra ? extra->stats_in : nullptr;
    ep.ln_parts = extra ? extra->ln_parts : 0;
    ep.ln_inv_dim = extra && extra->ln_dim > 0 ? 1.0f / static_cast<float>(extra->ln_dim) : 0.f;
    ep.ln_eps = extra ? extra->ln_eps : 0.f;
    ep.out_ld = extra && extra->out_ld > 0 ? extra->out_ld : plan->N;
    ep.n_valid = extra && extra->n_valid > 0 ? extra->n_valid : plan->N;
    ep.act = extra ? extra->act : 0;
    AP_REQUIRE(ctx, (ep.out_ld == plan->N && ep.n_valid == plan->N && ep.act == 0) || (plan->epilogue == AP_EPI_BIAS_F32 && ep.out_h == nullptr &&
                        ep.tokens_per_image == 0 && ep.n_valid % 4 == 0 && ep.out_ld % 4 == 0 && ep.n_valid <= plan->N),
               "gemm: out_ld / n_valid / act belong to the plain fp32-output epilogue");
    const bool f32_out = plan->epilogue == AP_EPI_BIAS_RESID_F32 || plan->epilogue == AP_EPI_BIAS_F32;
    AP_REQUIRE(ctx, (ep.out_h == nullptr) == (ep.stats_out == nullptr) && (ep.out_h == nullptr || f32_out),
               "gemm: out_h / stats_out belong together and to the fp32-output epilogues");
    AP_REQUIRE(ctx, ep.stats_out == nullptr || ep.ln_parts == plan->N / (plan->bn / 2),
               "gemm: producer ln_parts %d must be N / (bn / 2) = %d", ep.ln_parts, plan->N / (plan->bn / 2));
    AP_REQUIRE(ctx, ep.stats_in == nullptr || (!f32_out && ep.ln_parts > 0 && extra->ln_dim > 0),
               "gemm: folded LayerNorm needs an fp16-output epilogue, ln_parts and ln_dim");
    if (plan->cta_group == 2) return dispatch_epi<2, 256>(ctx, plan, ep, stream);
    if (plan->bn == 256) return dispatch_epi<1, 256>(ctx, plan, ep, stream);
    return dispatch_epi<1, 128>(ctx, plan, ep, stream);
}

extern "C" int ap_gemm_f16(ap_ctx* ctx, const void* A_dev, const void* W_dev, const float* bias_dev,
                           const float* resid_dev, void* out_dev, int M, int N, int K, int epilogue, void* stream) {
    if (!ctx) return AP_EINVAL;
    DeviceGuard guard(ctx);
    GemmPlan plan;
    int rc = ap_gemm_plan(ctx, &plan, A_dev, W_dev, M, N, K, epilogue, K);
    if (rc) return rc;
    return ap_gemm_run(ctx, &plan, bias_dev, resid_dev, out_dev, nullptr, static_cast<cudaStream_t>(stream));
}

// The same GEMM with split fp16 operands (DESIGN.md "precision"): split 1: W_dev is [N, 2K] = [W_hi | W_lo]; split 2: A_dev is
// [M, 2K] = [A_hi | A_lo]; split 3: both (three products per term: A_hi W_hi + A_lo W_hi + A_hi W_lo, ~22 significant bits).
extern "C" int ap_gemm_f16_split(ap_ctx* ctx, const void* A_dev, const void* W_dev, const float* bias_dev, const float* resid_dev,
                                 void* out_dev, int M, int N, int K, int epilogue, int split, void* stream) {
    if (!ctx) return AP_EINVAL;
    DeviceGuard guard(ctx);
    GemmPlan plan;
    int rc = ap_gemm_plan_split(ctx, &plan, A_dev, W_dev, M, N, K, epilogue, split);
    if (rc) return rc;
    AP_REQUIRE(ctx, N % 128 == 0, "gemm: N=%d must be a multiple of 128", N);
    return ap_gemm_run(ctx, &plan, bias_dev, resid_dev, out_dev, nullptr, static_cast<cudaStream_t>(stream));
}
