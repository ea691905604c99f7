// Multi-head self-attention on the 5th-gen tensor cores (tcgen05 + TMEM), head_dim 64, 16 <= S <= 256.
//
// nn.MultiheadAttention semantics of torchvision's EncoderBlock (the forward the reference runs at
// atlas_patch/models/patch/base.py:100):  out = softmax((q / sqrt(d)) k^T) v  per (image, head).
//
// One persistent CTA per SM walks (image, head) jobs.  Per job, entirely on chip:
//   TMA      Q (two 128-row tiles), K, V head slices of the packed QKV activations -> 128B-swizzled smem (2 stages)
//   MMA      S_g = Q_g K^T        tcgen05.mma SS, M = 128, N = S_pad, K = 64     -> TMEM (fp32), g = query tile 0/1
//   softmax  one thread per query row: tcgen05.ld S, row max, exp2, row sum (fp32); P rounded to fp16 and written
//            back with tcgen05.st INTO THE SAME TMEM COLUMNS (P aliases the first half of S)
//   MMA      O_g = P_g V          tcgen05.mma TS: A = P from TMEM, B = V from smem (MN-major descriptor), N = 64
//   epilogue tcgen05.ld O, scale by 1 / row sum, fp16, 128 B per row to global
// Keys >= S (padding up to S_pad, a multiple of 16) get probability 0; query rows >= S are computed but never stored.
// Warps: 0 = TMA producer (Q, K), 3 = TMA producer (V), 1 = MMA issuer, 2 = TMEM allocator, 4-7 / 8-11 = softmax + epilogue of query tile 0 / 1
// (warp % 4 selects the TMEM lane quarter).  TMEM columns per tile g: S at [256 g, 256 g + S_pad), P at
// [256 g, 256 g + S_pad / 2), O at [256 g + 128, 256 g + 192).
#include "ap_internal.cuh"
#include "ptx.cuh"

namespace {

constexpr int ATC_THREADS = 384;
constexpr int Q_TILE_BYTES = 128 * 128;  // 128 rows x 64 fp16

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// NC > 0: S_pad = 16 * NC is a compile-time constant and the softmax makes ONE pass over TMEM: the score row is processed
//   as two register-resident halves (A = first ceil(NC/2) 16-key chunks, B = the rest), each normalised by its own
//   maximum; P.V is accumulated separately for the two halves (O_a, O_b in TMEM) and the epilogue combines them,
//   O = (alpha O_a + beta O_b) / (alpha l_a + beta l_b), alpha = 2^((m_a - m) scale), beta = 2^((m_b - m) scale).
//   (A whole 208-score row in registers needs setmaxnreg; ptxas could not fit it.)
// NC == 0: generic two-pass softmax (max pass, then exp pass) for any 16 <= S_pad <= 256.
template <int NC>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                    __half* __restrict__ out, int B, int S, int S_pad_rt, int heads, int variant) {
    const int S_pad = NC > 0 ? NC * 16 : S_pad_rt;
    extern __shared__ uint8_t smem_raw_att[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_att) + 1023) & ~uintptr_t(1023));
    const int D = heads * 64;
    const int kv_bytes = S_pad * 128;
    const int stage_bytes = 2 * Q_TILE_BYTES + 2 * kv_bytes;   // per stage: [Q0 | Q1 | K | V]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);
    uint64_t* qk_full = bars;        // [2] TMA -> MMA   (Q tiles + K)
    uint64_t* qk_empty = bars + 2;   // [2] MMA -> TMA   (released as soon as the S MMAs retire)
    uint64_t* v_full = bars + 4;     // [2] TMA -> MMA   (V)
    uint64_t* v_empty = bars + 6;    // [2] MMA -> TMA   (released when the PV MMAs retire)
    uint64_t* s_full = bars + 8;     // [2 tiles] MMA -> softmax
    uint64_t* p_full = bars + 10;    // [2 tiles] softmax -> MMA
    uint64_t* o_full = bars + 12;    // [2 tiles] MMA -> epilogue
    uint64_t* o_empty = bars + 14;   // [2 tiles] epilogue -> MMA (TMEM tile reusable)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_jobs = B * heads;
    const int n_qt = S > 128 ? 2 : 1;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&map_q);
        ptx::prefetch_tmap(&map_kv);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&qk_full[i], 1);
            ptx::mbar_init(&qk_empty[i], 1);
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&v_empty[i], 1);
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 4);
            ptx::mbar_init(&o_full[i], 1);
            ptx::mbar_init(&o_empty[i], 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        ptx::tmem_alloc<1>(tmem_ptr_smem, 512);
        ptx::tmem_relinquish<1>();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();

    if (warp < 4) {
      if (warp == 0 || warp == 3) {
        if (lane == 0) {
            int it = 0;
            for (int job = blockIdx.x; job < n_jobs; job += gridDim.x, ++it) {
                const int st = it & 1;
                const int b = job / heads, h = job - b * heads;
                uint8_t* sb = smem + st * stage_bytes;
                if (warp == 0) {
                    ptx::mbar_wait(&qk_empty[st], ((it >> 1) & 1) ^ 1, 11);
                    ptx::mbar_arrive_expect_tx(&qk_full[st], n_qt * Q_TILE_BYTES + kv_bytes);
                    for (int g = 0; g < n_qt; ++g)
                        ptx::tma_load_2d(sb + g * Q_TILE_BYTES, &map_q, &qk_full[st], h * 64, b * S + g * 128);
                    ptx::tma_load_2d(sb + 2 * Q_TILE_BYTES, &map_kv, &qk_full[st], D + h * 64, b * S);
                } else {
                    ptx::mbar_wait(&v_empty[st], ((it >> 1) & 1) ^ 1, 17);
                    ptx::mbar_arrive_expect_tx(&v_full[st], kv_bytes);
                    ptx::tma_load_2d(sb + 2 * Q_TILE_BYTES + kv_bytes, &map_kv, &v_full[st], 2 * D + h * 64, b * S);
                }
            }
        }
        __syncwarp();
      } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = ptx::make_idesc_f16(128, S_pad);
            const uint32_t idesc_o = ptx::make_idesc_f16(128, 64, false, true);
            const int k_steps_pv = S_pad / 16;
            int it = 0;
            for (int job = blockIdx.x; job < n_jobs; job += gridDim.x, ++it) {
                const int st = it & 1;
                const uint32_t ph = it & 1;
                uint8_t* sb = smem + st * stage_bytes;
                ptx::mbar_wait(&qk_full[st], (it >> 1) & 1, 12);
                ptx::tc_fence_after();
                const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + 2 * Q_TILE_BYTES));
                for (int g = 0; g < n_qt; ++g) {
                    ptx::mbar_wait(&o_empty[g], ph ^ 1, 13);   // previous job's O (same TMEM columns) has been drained
                    ptx::tc_fence_after();
                    const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + g * Q_TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::tc_mma_f16<1>(tmem_base + g * 256, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                    ptx::tc_commit<1>(&s_full[g]);
                }
                ptx::tc_commit<1>(&qk_empty[st]);           // Q and K of this stage are free once the S MMAs retire
                ptx::mbar_wait(&v_full[st], (it >> 1) & 1, 18);
                ptx::tc_fence_after();
                const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + 2 * Q_TILE_BYTES + kv_bytes), 64);
                for (int g = 0; g < n_qt; ++g) {
                    ptx::mbar_wait(&p_full[g], ph, 14);
                    ptx::tc_fence_after();
                    constexpr int NA = (NC + 1) / 2;
                    for (int ks = 0; ks < k_steps_pv; ++ks) {   // 16 keys per step: 8 TMEM columns of P, two 8-key groups (2 KB) of V
                        const bool half_b = NC > 0 && ks >= NA;   // second half of the keys accumulates into O_b
                        ptx::tc_mma_f16_ts(tmem_base + g * 256 + (half_b ? 192 : 128), tmem_base + g * 256 + ks * 8, v_desc + ks * 128,
                                           idesc_o, (ks != 0 && !(NC > 0 && ks == NA)) ? 1u : 0u);
                    }
                    ptx::tc_commit<1>(&o_full[g]);
                }
                ptx::tc_commit<1>(&v_empty[st]);
            }
        }
        __syncwarp();
      }
    } else {
        const int g = (warp - 4) >> 2;
        const int q = warp & 3;
        if (g < n_qt) {
            const int row_in_tile = q * 32 + lane;
            const int q_row = g * 128 + row_in_tile;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * 256;
            const float scale = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)
            const bool swap = (variant & 1) != 0;
            int it = 0;
            for (int job = blockIdx.x; job < n_jobs; job += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                const int b = job / heads, h = job - b * heads;
                ptx::mbar_wait(&s_full[g], ph, 15);
                ptx::tc_fence_after();
                float l = 0.f, alpha = 1.f, beta = 0.f;
                // tcgen05.ld is the scarce resource here (~64 B/clk/SM): warps whose 32 query rows are all padding skip the
                // softmax and the O read-out entirely (their P rows stay garbage; those O rows are never stored)
                const bool warp_has_rows = g * 128 + q * 32 < S;
                if (!warp_has_rows) {
                } else if (NC > 0) {
                    // ---- one pass: the whole score row (16 * NC fp32) lives in registers ----
                    constexpr int NA = (NC + 1) / 2, NB = NC - NA, NAA = NA > 0 ? NA : 1, NBA = NB > 0 ? NB : 1;
                    float m_a, m_b = -INFINITY, l_a, l_b = 0.f;
                    {   // ---- half A: keys [0, 16 NA), never padded ----
                        uint32_t sc[NAA][16];
#pragma unroll
                        for (int c = 0; c < NA; ++c) ptx::tmem_ld_32x16(t_row + c * 16, sc[c]);
                        ptx::tc_wait_ld();
                        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                        for (int c = 0; c < NA; ++c) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const int k0 = c * 16 + j;
                                if (NB > 0 || c < NA - 1) {
                                    m0 = fmaxf(m0, __uint_as_float(sc[c][j])); m1 = fmaxf(m1, __uint_as_float(sc[c][j + 1]));
                                    m2 = fmaxf(m2, __uint_as_float(sc[c][j + 2])); m3 = fmaxf(m3, __uint_as_float(sc[c][j + 3]));
                                } else {
                                    if (k0 < S) m0 = fmaxf(m0, __uint_as_float(sc[c][j]));
                                    if (k0 + 1 < S) m1 = fmaxf(m1, __uint_as_float(sc[c][j + 1]));
                                    if (k0 + 2 < S) m2 = fmaxf(m2, __uint_as_float(sc[c][j + 2]));
                                    if (k0 + 3 < S) m3 = fmaxf(m3, __uint_as_float(sc[c][j + 3]));
                                }
                            }
                        }
                        m_a = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                        const float msa = m_a * scale;
                        float l0 = 0.f, l1 = 0.f;
#pragma unroll
                        for (int c = 0; c < NA; ++c) {          // 16 keys -> 8 packed TMEM columns
                            uint32_t pk[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int k0 = c * 16 + 2 * j;
                                float p0 = exp2f(fmaf(__uint_as_float(sc[c][2 * j]), scale, -msa));
                                float p1 = exp2f(fmaf(__uint_as_float(sc[c][2 * j + 1]), scale, -msa));
                                if (NB == 0 && c == NA - 1) {
                                    p0 = (k0 < S) ? p0 : 0.f;
                                    p1 = (k0 + 1 < S) ? p1 : 0.f;
                                }
                                l0 += p0; l1 += p1;
                                pk[j] = pack_h2(p0, p1);
                            }
                            ptx::tmem_st_32x8(t_row + c * 8, pk);
                        }
                        l_a = l0 + l1;
                    }
                    if (NB > 0) {   // ---- half B: keys [16 NA, 16 NC), padding possible in the last chunk ----
                        uint32_t sc[NBA][16];
#pragma unroll
                        for (int c = 0; c < NB; ++c) ptx::tmem_ld_32x16(t_row + (NA + c) * 16, sc[c]);
                        ptx::tc_wait_ld();
                        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                        for (int c = 0; c < NB; ++c) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) {
                                const int k0 = (NA + c) * 16 + j;
                                if (c < NB - 1) {
                                    m0 = fmaxf(m0, __uint_as_float(sc[c][j])); m1 = fmaxf(m1, __uint_as_float(sc[c][j + 1]));
                                    m2 = fmaxf(m2, __uint_as_float(sc[c][j + 2])); m3 = fmaxf(m3, __uint_as_float(sc[c][j + 3]));
                                } else {
                                    if (k0 < S) m0 = fmaxf(m0, __uint_as_float(sc[c][j]));
                                    if (k0 + 1 < S) m1 = fmaxf(m1, __uint_as_float(sc[c][j + 1]));
                                    if (k0 + 2 < S) m2 = fmaxf(m2, __uint_as_float(sc[c][j + 2]));
                                    if (k0 + 3 < S) m3 = fmaxf(m3, __uint_as_float(sc[c][j + 3]));
                                }
                            }
                        }
                        m_b = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));   // finite: 16 NA < S, so half B holds >= 1 valid key
                        const float msb = m_b * scale;
                        float l0 = 0.f, l1 = 0.f;
#pragma unroll
                        for (int c = 0; c < NB; ++c) {
                            uint32_t pk[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int k0 = (NA + c) * 16 + 2 * j;
                                float p0 = exp2f(fmaf(__uint_as_float(sc[c][2 * j]), scale, -msb));
                                float p1 = exp2f(fmaf(__uint_as_float(sc[c][2 * j + 1]), scale, -msb));
                                if (c == NB - 1) {
                                    p0 = (k0 < S) ? p0 : 0.f;
                                    p1 = (k0 + 1 < S) ? p1 : 0.f;
                                }
                                l0 += p0; l1 += p1;
                                pk[j] = pack_h2(p0, p1);
                            }
                            ptx::tmem_st_32x8(t_row + (NA + c) * 8, pk);
                        }
                        l_b = l0 + l1;
                    }
                    const float m_all = fmaxf(m_a, m_b);
                    alpha = exp2f((m_a - m_all) * scale);
                    beta = NB > 0 ? exp2f((m_b - m_all) * scale) : 0.f;
                    l = alpha * l_a + beta * l_b;
                } else {
                // ---- pass 1: row max over the valid keys ----
                float m = -INFINITY;
                for (int c0 = 0; c0 < ((variant & 2) ? 0 : S_pad); c0 += 32) {
                    if (c0 + 32 <= S_pad) {
                        uint32_t r[32];
                        ptx::tmem_ld_32x32(t_row + c0, r);
                        ptx::tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + j < S) m = fmaxf(m, __uint_as_float(r[j]));
                    } else {
                        uint32_t r[16];
                        ptx::tmem_ld_32x16(t_row + c0, r);
                        ptx::tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < S) m = fmaxf(m, __uint_as_float(r[j]));
                    }
                }
                if (variant & 2) m = 0.f;
                const float ms = m * scale;
                // ---- pass 2: P = exp2(s * scale - m * scale), row sum, fp16 P back into TMEM (aliases S) ----
                for (int c0 = 0; c0 < ((variant & 8) ? 0 : S_pad); c0 += 32) {
                    if (c0 + 32 <= S_pad) {
                        uint32_t r[32];
                        ptx::tmem_ld_32x32(t_row + c0, r);
                        ptx::tc_wait_ld();
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float p0 = fmaf(__uint_as_float(r[2 * j]), scale, -ms), p1 = fmaf(__uint_as_float(r[2 * j + 1]), scale, -ms);
                            if (!(variant & 4)) { p0 = exp2f(p0); p1 = exp2f(p1); }
                            p0 = (c0 + 2 * j < S) ? p0 : 0.f;
                            p1 = (c0 + 2 * j + 1 < S) ? p1 : 0.f;
                            l += p0 + p1;
                            pk[j] = swap ? pack_h2(p1, p0) : pack_h2(p0, p1);
                        }
                        ptx::tmem_st_32x16(t_row + (c0 >> 1), pk);
                    } else {
                        uint32_t r[16];
                        ptx::tmem_ld_32x16(t_row + c0, r);
                        ptx::tc_wait_ld();
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float p0 = (c0 + 2 * j < S) ? exp2f(fmaf(__uint_as_float(r[2 * j]), scale, -ms)) : 0.f;
                            const float p1 = (c0 + 2 * j + 1 < S) ? exp2f(fmaf(__uint_as_float(r[2 * j + 1]), scale, -ms)) : 0.f;
                            l += p0 + p1;
                            pk[j] = swap ? pack_h2(p1, p0) : pack_h2(p0, p1);
                        }
                        ptx::tmem_st_32x8(t_row + (c0 >> 1), pk);
                    }
                }
                }  // NC == 0
                ptx::tc_wait_st();
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&p_full[g]);
                // ---- epilogue: O / l -> fp16 -> global ----
                ptx::mbar_wait(&o_full[g], ph, 16);
                ptx::tc_fence_after();
                const float inv = 1.0f / l;
                const float wa = alpha * inv, wb = beta * inv;
                uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(b) * S + q_row) * D + h * 64);
                if (!warp_has_rows) {
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&o_empty[g]);
                    continue;
                }
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {       // 32 output columns at a time
                    uint32_t oa[32], ob[32];
                    ptx::tmem_ld_32x32(t_row + 128 + hh * 32, oa);
                    if (NC > 0) ptx::tmem_ld_32x32(t_row + 192 + hh * 32, ob);
                    ptx::tc_wait_ld();
                    if (hh == 1) {                     // all TMEM reads of this tile are done: hand it back to the MMA warp
                        ptx::tc_fence_before();
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&o_empty[g]);
                    }
                    if (q_row < S) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float v[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                v[e] = NC > 0 ? fmaf(__uint_as_float(oa[8 * j + e]), wa, __uint_as_float(ob[8 * j + e]) * wb)
                                              : __uint_as_float(oa[8 * j + e]) * inv;
                            dst[hh * 4 + j] = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                        }
                    }
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem_base, 512);
    }
}

}  // namespace

int ap_attention_tc_plan(ap_ctx* ctx, AttnPlan* plan, const __half* qkv, int rows, int S, int heads) {
    const int D = heads * 64;
    plan->S_pad = (S + 15) / 16 * 16;
    AP_REQUIRE(ctx, plan->S_pad >= 16 && plan->S_pad <= 256, "attention(tcgen05): S=%d unsupported (16..256 after padding)", S);
    int rc = ap_make_tmap_f16_2d(ctx, &plan->map_q, qkv, (uint64_t)rows, (uint64_t)3 * D, (uint64_t)3 * D, 128, 64);
    if (rc) return rc;
    return ap_make_tmap_f16_2d(ctx, &plan->map_kv, qkv, (uint64_t)rows, (uint64_t)3 * D, (uint64_t)3 * D, plan->S_pad, 64);
}

int ap_attention_tc_run(ap_ctx* ctx, const AttnPlan* plan, __half* out, int B, int S, int heads, cudaStream_t stream) {
    if (B == 0) return AP_OK;
    const int S_pad = plan->S_pad;
    const size_t smem = 2 * (2 * (size_t)Q_TILE_BYTES + 2 * (size_t)S_pad * 128) + 17 * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        const int max_smem = 2 * (2 * Q_TILE_BYTES + 2 * 256 * 128) + 17 * 8 + 16 + 1024;
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(attention_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(attention_tc_kernel<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr_set = true;
    }
    const int jobs = B * heads;
    const int grid = jobs < ctx->sm_count ? jobs : ctx->sm_count;
    ProfScope prof(ctx, stream, AP_K_ATTENTION);
    if (S_pad == 208 && !(ctx->attn_variant & 16))   // 197 tokens (ViT/16 @ 224): register-resident single-pass softmax
        AP_CHECK_CUDA(ctx, ap_launch_pdl(attention_tc_kernel<13>, dim3(grid), dim3(ATC_THREADS), smem, stream, 1, ctx->pdl != 0, plan->map_q,
                                         plan->map_kv, out, B, S, S_pad, heads, ctx->attn_variant));
    else
        AP_CHECK_CUDA(ctx, ap_launch_pdl(attention_tc_kernel<0>, dim3(grid), dim3(ATC_THREADS), smem, stream, 1, ctx->pdl != 0, plan->map_q,
                                         plan->map_kv, out, B, S, S_pad, heads, ctx->attn_variant));
    AP_CHECK_LAUNCH(ctx, "attention_tc_kernel");
    return AP_OK;
}
