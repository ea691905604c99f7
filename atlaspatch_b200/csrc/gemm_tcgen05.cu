// Persistent, warp-specialised tcgen05 GEMM with fused epilogues (sm_100a).
//
//   out[M,N] = epilogue( A[M,K] (fp16, K contiguous) x W[N,K]^T (fp16, K contiguous) )     fp32 accumulate in TMEM
//
// This is the dense contraction behind every Linear of the encoder forward the reference delegates to
// torch (atlas_patch/models/patch/base.py:100 -> torchvision VisionTransformer): conv_proj (as an
// im2col GEMM), in_proj (QKV), out_proj, mlp.0 (+GELU), mlp.3 (+residual).
//
// CTA = 384 threads, one CTA per SM, persistent over 128 x BN output tiles (static round-robin):
//   warp 0     TMA producer: A (128x64) and W (BNx64) boxes, 128B swizzle, 4-stage mbarrier ring
//   warp 1     MMA issuer:   one thread issues tcgen05.mma 128xBNx16 (4 per 64-wide K block),
//                            tcgen05.commit frees the smem stage / publishes the accumulator
//   warp 2     TMEM allocator (2 x BN fp32 columns: double-buffered accumulator)
//   warps 4-11 epilogue:     tcgen05.ld 32 lanes x 32 columns -> registers -> bias / GELU / residual /
//                            positional embedding -> vectorised global stores; overlaps the next tile's MMAs
#include "ap_internal.cuh"
#include "ptx.cuh"

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 fp16 = 128 B = one swizzle row
constexpr int STAGES = 4;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;
constexpr int NUM_EPI_WARPS = 8;

template <int BN>
struct SmemLayout {
    static constexpr int A_STAGE = BM * BK * 2;  // 16 KB
    static constexpr int B_STAGE = BN * BK * 2;  // 32 KB (BN=256)
    static constexpr int A_OFF = 0;
    static constexpr int B_OFF = STAGES * A_STAGE;
    static constexpr int BAR_OFF = B_OFF + STAGES * B_STAGE;
    static constexpr int NUM_BARS = 2 * STAGES + 4;
    static constexpr int TMEM_PTR_OFF = BAR_OFF + NUM_BARS * 8;
    static constexpr int TOTAL = TMEM_PTR_OFF + 16;
    static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024 B alignment
};

struct EpiParams {
    const float* bias;
    const float* resid;
    void* out;
    const float* pos;
    int tokens_per_image;
    float alpha;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, int M, int N,
                    int K, EpiParams ep) {
    using L = SmemLayout<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
    uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
    uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA -> TMA
    uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]       MMA -> epilogue
    uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]       epilogue -> MMA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::TMEM_PTR_OFF);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_n = N / BN;
    const int tiles_m = (M + BM - 1) / BM;
    const int num_tiles = tiles_m * tiles_n;
    const int k_blocks = K / BK;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&map_a);
        ptx::prefetch_tmap(&map_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tfull_bar[a], 1);
            ptx::mbar_init(&tempty_bar[a], NUM_EPI_WARPS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc(tmem_ptr_smem, 2 * BN);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM;
                const int n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1, 1);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], L::A_STAGE + L::B_STAGE);
                    ptx::tma_load_2d(smem + L::A_OFF + stage * L::A_STAGE, &map_a, &full_bar[stage], kb * BK, m0);
                    ptx::tma_load_2d(smem + L::B_OFF + stage * L::B_STAGE, &map_w, &full_bar[stage], kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_f16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
                        ptx::tc_mma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::tc_commit(&empty_bar[stage]);  // smem stage reusable once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                ptx::tc_commit(&tfull_bar[as]);  // accumulator complete
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ================= epilogue =================
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int half_idx = (warp - 4) >> 2;    // which half of the BN columns
        constexpr int COLS_PER_WARP = BN / 2;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / tiles_n) * BM;
            const int n0 = (tile % tiles_n) * BN;
            ptx::mbar_wait(&tfull_bar[as], aphase, 4);
            ptx::tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < M;
            int64_t out_row = row;
            int pos_row = 0;
            if (EPI == AP_EPI_BIAS_F32 && ep.tokens_per_image > 0) {
                const int b = row / ep.tokens_per_image;
                const int t = row - b * ep.tokens_per_image;
                out_row = static_cast<int64_t>(b) * (ep.tokens_per_image + 1) + 1 + t;
                pos_row = 1 + t;
            }
#pragma unroll 1
            for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                const int col_local = half_idx * COLS_PER_WARP + c * 32;
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + col_local;
                uint32_t r[32];
                ptx::tmem_ld_32x32(taddr, r);
                ptx::tc_wait_ld();
                const int col0 = n0 + col_local;
                float v[32];
                const float4* b4 = reinterpret_cast<const float4*>(ep.bias + col0);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bb = __ldg(b4 + j);
                    v[4 * j + 0] = __uint_as_float(r[4 * j + 0]) * ep.alpha + bb.x;
                    v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) * ep.alpha + bb.y;
                    v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) * ep.alpha + bb.z;
                    v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) * ep.alpha + bb.w;
                }
                if (row_ok) {
                    if (EPI == AP_EPI_BIAS_F16 || EPI == AP_EPI_BIAS_GELU_F16) {
                        if (EPI == AP_EPI_BIAS_GELU_F16) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                        }
                        uint4* o = reinterpret_cast<uint4*>(static_cast<__half*>(ep.out) + out_row * N + col0);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 u;
                            u.x = pack_half2(v[8 * j + 0], v[8 * j + 1]);
                            u.y = pack_half2(v[8 * j + 2], v[8 * j + 3]);
                            u.z = pack_half2(v[8 * j + 4], v[8 * j + 5]);
                            u.w = pack_half2(v[8 * j + 6], v[8 * j + 7]);
                            o[j] = u;
                        }
                    } else {
                        if (EPI == AP_EPI_BIAS_RESID_F32) {
                            const float4* r4 = reinterpret_cast<const float4*>(ep.resid + out_row * N + col0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 rr = r4[j];
                                v[4 * j + 0] += rr.x; v[4 * j + 1] += rr.y; v[4 * j + 2] += rr.z; v[4 * j + 3] += rr.w;
                            }
                        } else if (ep.pos != nullptr) {
                            const float4* p4 = reinterpret_cast<const float4*>(ep.pos + static_cast<int64_t>(pos_row) * N + col0);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 pp = __ldg(p4 + j);
                                v[4 * j + 0] += pp.x; v[4 * j + 1] += pp.y; v[4 * j + 2] += pp.z; v[4 * j + 3] += pp.w;
                            }
                        }
                        float4* o = reinterpret_cast<float4*>(static_cast<float*>(ep.out) + out_row * N + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tempty_bar[as]);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 2 * BN);
    }
}

template <int BN, int EPI>
int launch(ap_ctx* ctx, const GemmPlan* p, const EpiParams& ep, cudaStream_t stream) {
    using L = SmemLayout<BN>;
    auto kern = gemm_tcgen05_kernel<BN, EPI>;
    static bool attr_set = false;  // per (BN, EPI) instantiation
    if (!attr_set) {
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
        attr_set = true;
    }
    const int tiles = ((p->M + BM - 1) / BM) * (p->N / BN);
    const int grid = tiles < ctx->sm_count ? tiles : ctx->sm_count;
    ProfScope prof(ctx, stream, AP_K_GEMM);
    kern<<<grid, NUM_THREADS, L::DYN_BYTES, stream>>>(p->map_a, p->map_w, p->M, p->N, p->K, ep);
    AP_CHECK_LAUNCH(ctx, "gemm_tcgen05_kernel");
    return AP_OK;
}

template <int BN>
int dispatch_epi(ap_ctx* ctx, const GemmPlan* p, const EpiParams& ep, cudaStream_t stream) {
    switch (p->epilogue) {
        case AP_EPI_BIAS_F16: return launch<BN, AP_EPI_BIAS_F16>(ctx, p, ep, stream);
        case AP_EPI_BIAS_GELU_F16: return launch<BN, AP_EPI_BIAS_GELU_F16>(ctx, p, ep, stream);
        case AP_EPI_BIAS_RESID_F32: return launch<BN, AP_EPI_BIAS_RESID_F32>(ctx, p, ep, stream);
        case AP_EPI_BIAS_F32: return launch<BN, AP_EPI_BIAS_F32>(ctx, p, ep, stream);
    }
    return ap_set_error(ctx, AP_EINVAL, "gemm: unknown epilogue %d", p->epilogue);
}

}  // namespace

int ap_gemm_plan(ap_ctx* ctx, GemmPlan* plan, const void* A, const void* W, int M, int N, int K, int epilogue) {
    AP_REQUIRE(ctx, M > 0 && N > 0 && K > 0, "gemm: empty problem %dx%dx%d", M, N, K);
    AP_REQUIRE(ctx, K % BK == 0, "gemm: K=%d must be a multiple of %d", K, BK);
    AP_REQUIRE(ctx, N % 128 == 0, "gemm: N=%d must be a multiple of 128", N);
    AP_REQUIRE(ctx, epilogue >= 0 && epilogue <= 3, "gemm: unknown epilogue %d", epilogue);
    plan->M = M; plan->N = N; plan->K = K; plan->epilogue = epilogue;
    plan->bn = (N % 256 == 0) ? 256 : 128;
    int rc = ap_make_tmap_f16_2d(ctx, &plan->map_a, A, (uint64_t)M, (uint64_t)K, (uint64_t)K, BM, BK);
    if (rc) return rc;
    return ap_make_tmap_f16_2d(ctx, &plan->map_w, W, (uint64_t)N, (uint64_t)K, (uint64_t)K, plan->bn, BK);
}

int ap_gemm_run(ap_ctx* ctx, const GemmPlan* plan, const float* bias, const float* resid, void* out,
                const GemmExtra* extra, cudaStream_t stream) {
    AP_REQUIRE(ctx, bias != nullptr && out != nullptr, "gemm: bias/out must not be NULL");
    AP_REQUIRE(ctx, plan->epilogue != AP_EPI_BIAS_RESID_F32 || resid != nullptr, "gemm: residual epilogue needs resid");
    EpiParams ep;
    ep.bias = bias; ep.resid = resid; ep.out = out;
    ep.pos = extra ? extra->pos : nullptr;
    ep.tokens_per_image = extra ? extra->tokens_per_image : 0;
    ep.alpha = extra ? extra->alpha : 1.0f;
    if (plan->bn == 256) return dispatch_epi<256>(ctx, plan, ep, stream);
    return dispatch_epi<128>(ctx, plan, ep, stream);
}

extern "C" int ap_gemm_f16(ap_ctx* ctx, const void* A_dev, const void* W_dev, const float* bias_dev,
                           const float* resid_dev, void* out_dev, int M, int N, int K, int epilogue, void* stream) {
    if (!ctx) return AP_EINVAL;
    GemmPlan plan;
    int rc = ap_gemm_plan(ctx, &plan, A_dev, W_dev, M, N, K, epilogue);
    if (rc) return rc;
    return ap_gemm_run(ctx, &plan, bias_dev, resid_dev, out_dev, nullptr, static_cast<cudaStream_t>(stream));
}
