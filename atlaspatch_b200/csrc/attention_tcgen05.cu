// Multi-head self-attention on the 5th-gen tensor cores (tcgen05 + TMEM), head_dim 64, up to 256 MMA keys (+1 extra key).
//
// nn.MultiheadAttention semantics of torchvision's EncoderBlock (the forward the reference runs at
// atlas_patch/models/patch/base.py:100) and of transformers' Dinov2SelfAttention (models/patch/dinov2.py:60):
//     out = softmax((q / sqrt(d)) k^T) v          per (image, head).
//
// One persistent CTA per SM walks (image, head) jobs; a job is cut into query TILES of 128 rows and the tiles of all the
// CTA's jobs form one stream that is software-pipelined over two TMEM buffers (tile i lives in buffer i & 1):
//   TMA      Q tiles, K, V head slices of the packed QKV activations -> 128B-swizzled smem (2 job stages)
//   MMA      S_i = Q_i K^T        tcgen05.mma SS, M = 128, N = S_pad, K = 64      -> TMEM buffer (fp32)
//   softmax  one thread per query row: tcgen05.ld S, row max, exp2, row sum (fp32); P rounded to fp16 and written back
//            with tcgen05.st INTO THE SAME TMEM COLUMNS (P aliases the first half of S)
//   MMA      O_i = P_i V          tcgen05.mma TS: A = P from TMEM, B = V from smem (MN-major descriptor), N = 64
//   epilogue tcgen05.ld O, scale by 1 / row sum, fp16, 128 B per row to global
// Round 1 ran the two tiles of a job in lock step, so the tensor pipe and the TMEM read port (tcgen05.ld, the scarce resource:
// tools/microbench/ldtm_bench.cu) took turns: S-MMA -> softmax -> PV-MMA -> read-out was one serial chain per job (tensor pipe
// 18 % active).  Now the single MMA thread issues S_{i+1} as soon as buffer (i + 1) & 1 has been drained and P.V_i as soon as
// softmax warpgroup i & 1 has published P_i -- whichever comes first -- so one warpgroup's tcgen05.ld traffic overlaps the
// other's MMAs and waits.
//
// Variants measured and dropped this round (B200, B = 127, 12 heads, S = 197; profiles/r02_attention_notes.md has the ncu stall
// tables): this kernel 56.9 us; the same with one blocking MMA-issuer thread per warpgroup instead of the polling thread 63.9 us;
// every tile cut into two key-half units (S_a | S_b in separate 128-column buffers, four units in flight, O_a read out while
// P_b V_b runs) 62.3-68.5 us with blocking issuers, 78.5 us with one event-loop issuer (a failed mbarrier test costs ~150 cycles, so
// a loop over four pending events reacts ~450 cycles late); a two-pass streaming softmax (no half row in registers) +1-2 us.  In all
// of them the softmax warps compute for ~35 % of their time, wait for P.V / the next S for ~35 % (mostly for the slowest sibling
// warp: two warps share each scheduler's MUFU) and spend ~17 % in the O read-out; MUFU is 31-34 % busy, the tensor pipe 17-18 %.
//
// Key / query windows: queries are tokens [q0, q0 + nq) of every image, MMA keys tokens [k0, k0 + nk) (nk <= 256, padded to
// S_pad = a multiple of 16; padded keys get probability 0).  XK: one EXTRA key token `xkey` outside that window is folded in
// as a rank-1 update (its score is a 64-term dot product per row on the CUDA cores, its value row is added to O in the
// epilogue).  That is how the 257-token DINOv2 sequence fits: 256 patch keys through the MMA (2 x 256 fp32 columns = all of
// TMEM) + the class token as the extra key; the 257th query row rides in a third, almost empty tile.
// Warps: 0 = TMA producer (Q, K), 3 = TMA producer (V), 1 = MMA issuer, 2 = TMEM allocator, 4-7 / 8-11 = softmax + epilogue
// warpgroups 0 / 1 (warp % 4 selects the TMEM lane quarter).  TMEM columns of buffer b: S at [256 b, 256 b + S_pad), P at
// [256 b, 256 b + S_pad / 2), O_a at [256 b + 128, +192), O_b at [256 b + 192, +256).
#include "ap_internal.cuh"
#include "ptx.cuh"
#include "attention_common.cuh"

namespace {

constexpr int ATC_THREADS = 384;
// Bounded polling on two conditions at once (the MMA thread): a protocol bug must become a trap, never a hung GPU.
__device__ __forceinline__ void spin_guard(long long t0, int tag) {
    if (clock64() - t0 > 6000000000LL) {
        printf("atlaspatch_b200: attention MMA thread timed out (tag %d, block %d)\n", tag, (int)blockIdx.x);
        __trap();
    }
}

// NC > 0: S_pad = 16 * NC is a compile-time constant and the softmax makes ONE pass over TMEM: the score row is processed
//   as two register-resident halves (A = first ceil(NC/2) 16-key chunks, B = the rest), each normalised by its own
//   maximum; P.V is accumulated separately for the two halves (O_a, O_b in TMEM) and the epilogue combines them,
//   O = (alpha O_a + beta O_b) / (alpha l_a + beta l_b), alpha = 2^((m_a - m) scale), beta = 2^((m_b - m) scale).
// NC == 0: generic two-pass softmax (max pass, then exp pass) for any 16 <= S_pad <= 256.
// XK: extra key (see above); its score joins half A.
template <int NC, bool XK>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                    const __half* __restrict__ qkv, __half* __restrict__ out, const AttnArgs a) {
    const int S_pad = NC > 0 ? NC * 16 : a.S_pad;
    extern __shared__ uint8_t smem_raw_att[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_att) + 1023) & ~uintptr_t(1023));
    const int heads = a.heads, S = a.S;
    const int D = heads * 64;
    const int kv_bytes = S_pad * 128;
    const int n_qt = (a.nq + 127) / 128;                        // query tiles per job: 1, 2 or 3 (257 queries = 128 + 128 + 1)
    const int q_bytes = n_qt * Q_TILE_BYTES;
    const int stage_bytes = q_bytes + 2 * kv_bytes;             // per job stage: [Q tiles | K | V]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);
    uint64_t* qk_full = bars;        // [2 stages] TMA -> MMA   (Q tiles + K)
    uint64_t* qk_empty = bars + 2;   // [2 stages] MMA -> TMA   (released as soon as the job's S MMAs retire)
    uint64_t* v_full = bars + 4;     // [2 stages] TMA -> MMA   (V)
    uint64_t* v_empty = bars + 6;    // [2 stages] MMA -> TMA   (released when the job's PV MMAs retire)
    uint64_t* s_full = bars + 8;     // [2 buffers] MMA -> softmax
    uint64_t* p_full = bars + 10;    // [2 buffers] softmax -> MMA
    uint64_t* o_full = bars + 12;    // [2 buffers] MMA -> epilogue
    uint64_t* o_empty = bars + 14;   // [2 buffers] epilogue -> MMA (TMEM buffer reusable)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_jobs = a.B * heads;
    const int my_jobs = blockIdx.x < n_jobs ? (n_jobs - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int n_tiles = my_jobs * n_qt;

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
            for (int jt = 0; jt < my_jobs; ++jt) {
                const int job = blockIdx.x + jt * gridDim.x;
                const int st = jt & 1;
                const uint32_t sph = (jt >> 1) & 1;
                const int b = job / heads, h = job - b * heads;
                uint8_t* sb = smem + st * stage_bytes;
                if (warp == 0) {
                    ptx::mbar_wait(&qk_empty[st], sph ^ 1, 11);
                    ptx::mbar_arrive_expect_tx(&qk_full[st], n_qt * Q_TILE_BYTES + kv_bytes);
                    for (int g = 0; g < n_qt; ++g)
                        ptx::tma_load_2d(sb + g * Q_TILE_BYTES, &map_q, &qk_full[st], h * 64, b * S + a.q0 + g * 128);
                    ptx::tma_load_2d(sb + q_bytes, &map_kv, &qk_full[st], D + h * 64, b * S + a.k0);
                } else {
                    ptx::mbar_wait(&v_empty[st], sph ^ 1, 17);
                    ptx::mbar_arrive_expect_tx(&v_full[st], kv_bytes);
                    ptx::tma_load_2d(sb + q_bytes + kv_bytes, &map_kv, &v_full[st], 2 * D + h * 64, b * S + a.k0);
                }
            }
        }
        __syncwarp();
      } else if (warp == 1) {
        // The whole warp polls and one ELECTED lane issues: warp-uniform control flow keeps the descriptor arithmetic on the uniform
        // datapath (issued from a `lane == 0` branch every tcgen05.mma operand needed an R2UR round trip: ~70 clk per MMA, see
        // attention_tc6_kernel).  The polled results are made uniform with a vote, since lanes may observe a barrier flip at different times.
        if (n_tiles > 0) {
            const uint32_t idesc_s = ptx::make_idesc_f16(128, S_pad);
            const uint32_t idesc_o = ptx::make_idesc_f16(128, 64, false, true);
            const int k_steps_pv = S_pad / 16;
            // ---- tile i: S = Q_g K^T into buffer i & 1 ----
            auto s_ready = [&](int i) -> bool {
                const int jt = i / n_qt, g = i - jt * n_qt, st = jt & 1, buf = i & 1;
                bool ok = true;
                if (g == 0) ok = ptx::mbar_try_wait(&qk_full[st], (jt >> 1) & 1);
                if (ok) ok = ptx::mbar_try_wait(&o_empty[buf], (((i >> 1) & 1) ^ 1));   // the buffer's previous tile has been drained
                return __all_sync(0xffffffffu, ok);
            };
            auto issue_s = [&](int i) {
                const int jt = i / n_qt, g = i - jt * n_qt, st = jt & 1, buf = i & 1;
                uint8_t* sb = smem + st * stage_bytes;
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + q_bytes));
                    const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + g * Q_TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::tc_mma_f16<1>(tmem_base + buf * 256, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                    ptx::tc_commit<1>(&s_full[buf]);
                    if (g == n_qt - 1) ptx::tc_commit<1>(&qk_empty[st]);   // Q and K of this stage are free once the S MMAs retire
                }
                __syncwarp();
            };
            // ---- tile i: O = P V ----
            auto pv_ready = [&](int i) -> bool {
                const int jt = i / n_qt, g = i - jt * n_qt, st = jt & 1, buf = i & 1;
                bool ok = true;
                if (g == 0) ok = ptx::mbar_try_wait(&v_full[st], (jt >> 1) & 1);
                if (ok) ok = ptx::mbar_try_wait(&p_full[buf], (i >> 1) & 1);
                return __all_sync(0xffffffffu, ok);
            };
            auto issue_pv = [&](int i) {
                const int jt = i / n_qt, g = i - jt * n_qt, st = jt & 1, buf = i & 1;
                uint8_t* sb = smem + st * stage_bytes;
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + q_bytes + kv_bytes), 64);
                    constexpr int NA = (NC + 1) / 2;
                    if (NC > 0) {
#pragma unroll
                        for (int ks = 0; ks < (NC > 0 ? NC : 1); ++ks) {   // 16 keys per step: 8 TMEM columns of P, two 8-key groups (2 KB) of V
                            const bool half_b = ks >= NA;                   // second half of the keys accumulates into O_b
                            ptx::tc_mma_f16_ts(tmem_base + buf * 256 + (half_b ? 192 : 128), tmem_base + buf * 256 + ks * 8, v_desc + ks * 128,
                                               idesc_o, (ks != 0 && ks != NA) ? 1u : 0u);
                        }
                    } else {
                        for (int ks = 0; ks < k_steps_pv; ++ks)
                            ptx::tc_mma_f16_ts(tmem_base + buf * 256 + 128, tmem_base + buf * 256 + ks * 8, v_desc + ks * 128, idesc_o,
                                               ks != 0 ? 1u : 0u);
                    }
                    ptx::tc_commit<1>(&o_full[buf]);
                    if (g == n_qt - 1) ptx::tc_commit<1>(&v_empty[st]);
                }
                __syncwarp();
            };
            long long t0 = clock64();
            while (!s_ready(0)) spin_guard(t0, 21);
            issue_s(0);
            for (int i = 0; i < n_tiles; ++i) {
                bool s_done = i + 1 >= n_tiles, pv_done = false;
                t0 = clock64();
                while (!pv_done || !s_done) {
                    if (!s_done && s_ready(i + 1)) { issue_s(i + 1); s_done = true; }
                    if (!pv_done && pv_ready(i)) { issue_pv(i); pv_done = true; }
                    if (lane == 0) spin_guard(t0, 22);
                }
            }
        }
      }
    } else {
        const int wg = (warp - 4) >> 2;   // softmax warpgroup = TMEM buffer
        const int q = warp & 3;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + wg * 256;
        const float scale = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)
        const bool swap = (a.variant & 1) != 0;
        for (int i = wg; i < n_tiles; i += 2) {
            const int jt = i / n_qt, g = i - jt * n_qt;
            const int job = blockIdx.x + jt * gridDim.x;
            const uint32_t ph = (i >> 1) & 1;
            const int b = job / heads, h = job - b * heads;
            const int q_row = g * 128 + q * 32 + lane;          // row inside the query window
            // tcgen05.ld is the scarce resource: warps whose 32 query rows are all padding skip the softmax and the O read-out
            // entirely (their P rows stay garbage; those O rows are never stored)
            const bool warp_has_rows = g * 128 + q * 32 < a.nq;
            float s_x = 0.f;
            const __half* xrow = nullptr;
            if (XK && warp_has_rows) {   // score of the extra key: 64-term dot product on the CUDA cores (L2-resident rows)
                const int qr = q_row < a.nq ? q_row : a.nq - 1;
                const uint4* qp = reinterpret_cast<const uint4*>(qkv + (static_cast<int64_t>(b) * S + a.q0 + qr) * 3 * D + h * 64);
                xrow = qkv + (static_cast<int64_t>(b) * S + a.xkey) * 3 * D + h * 64;
                const uint4* kp = reinterpret_cast<const uint4*>(xrow + D);
#pragma unroll
                for (int c = 0; c < 8; ++c) s_x = dot8_h(__ldg(qp + c), __ldg(kp + c), s_x);
            }
            ptx::mbar_wait(&s_full[wg], ph, 15);
            ptx::tc_fence_after();
            float l = 0.f, alpha = 1.f, beta = 0.f, p_x = 0.f;
            if (!warp_has_rows) {
            } else if (NC > 0) {
                // ---- one pass: the whole score row (16 * NC fp32) lives in registers, half at a time ----
                constexpr int NA = (NC + 1) / 2, NB = NC - NA, NAA = NA > 0 ? NA : 1, NBA = NB > 0 ? NB : 1;
                float m_a, m_b = -INFINITY, l_a, l_b = 0.f;
                {   // ---- half A: keys [0, 16 NA), never padded when NB > 0 ----
                    uint32_t sc[NAA][16];
#pragma unroll
                    for (int c = 0; c < NA; ++c) ptx::tmem_ld_32x16(t_row + c * 16, sc[c]);
                    ptx::tc_wait_ld();
                    float m0 = XK ? s_x : -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                    for (int c = 0; c < NA; ++c) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const int k0 = c * 16 + j;
                            if (NB > 0 || c < NA - 1) {
                                m0 = fmaxf(m0, __uint_as_float(sc[c][j])); m1 = fmaxf(m1, __uint_as_float(sc[c][j + 1]));
                                m2 = fmaxf(m2, __uint_as_float(sc[c][j + 2])); m3 = fmaxf(m3, __uint_as_float(sc[c][j + 3]));
                            } else {
                                if (k0 < a.nk) m0 = fmaxf(m0, __uint_as_float(sc[c][j]));
                                if (k0 + 1 < a.nk) m1 = fmaxf(m1, __uint_as_float(sc[c][j + 1]));
                                if (k0 + 2 < a.nk) m2 = fmaxf(m2, __uint_as_float(sc[c][j + 2]));
                                if (k0 + 3 < a.nk) m3 = fmaxf(m3, __uint_as_float(sc[c][j + 3]));
                            }
                        }
                    }
                    m_a = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    const float msa = m_a * scale;
                    float l0 = 0.f, l1 = 0.f;
                    if (XK) { p_x = exp2f(fmaf(s_x, scale, -msa)); l0 = p_x; }
#pragma unroll
                    for (int c = 0; c < NA; ++c) {          // 16 keys -> 8 packed TMEM columns
                        uint32_t pk[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int k0 = c * 16 + 2 * j;
                            float p0 = exp2f(fmaf(__uint_as_float(sc[c][2 * j]), scale, -msa));
                            float p1 = exp2f(fmaf(__uint_as_float(sc[c][2 * j + 1]), scale, -msa));
                            if (NB == 0 && c == NA - 1) {
                                p0 = (k0 < a.nk) ? p0 : 0.f;
                                p1 = (k0 + 1 < a.nk) ? p1 : 0.f;
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
                                if (k0 < a.nk) m0 = fmaxf(m0, __uint_as_float(sc[c][j]));
                                if (k0 + 1 < a.nk) m1 = fmaxf(m1, __uint_as_float(sc[c][j + 1]));
                                if (k0 + 2 < a.nk) m2 = fmaxf(m2, __uint_as_float(sc[c][j + 2]));
                                if (k0 + 3 < a.nk) m3 = fmaxf(m3, __uint_as_float(sc[c][j + 3]));
                            }
                        }
                    }
                    m_b = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));   // finite: 16 NA < nk, so half B holds >= 1 valid key
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
                                p0 = (k0 < a.nk) ? p0 : 0.f;
                                p1 = (k0 + 1 < a.nk) ? p1 : 0.f;
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
                float m = XK ? s_x : -INFINITY;
                for (int c0 = 0; c0 < ((a.variant & 2) ? 0 : S_pad); c0 += 32) {
                    if (c0 + 32 <= S_pad) {
                        uint32_t r[32];
                        ptx::tmem_ld_32x32(t_row + c0, r);
                        ptx::tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c0 + j < a.nk) m = fmaxf(m, __uint_as_float(r[j]));
                    } else {
                        uint32_t r[16];
                        ptx::tmem_ld_32x16(t_row + c0, r);
                        ptx::tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (c0 + j < a.nk) m = fmaxf(m, __uint_as_float(r[j]));
                    }
                }
                if (a.variant & 2) m = 0.f;
                const float ms = m * scale;
                if (XK) { p_x = exp2f(fmaf(s_x, scale, -ms)); l = p_x; }
                // ---- pass 2: P = exp2(s * scale - m * scale), row sum, fp16 P back into TMEM (aliases S) ----
                for (int c0 = 0; c0 < ((a.variant & 8) ? 0 : S_pad); c0 += 32) {
                    if (c0 + 32 <= S_pad) {
                        uint32_t r[32];
                        ptx::tmem_ld_32x32(t_row + c0, r);
                        ptx::tc_wait_ld();
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float p0 = fmaf(__uint_as_float(r[2 * j]), scale, -ms), p1 = fmaf(__uint_as_float(r[2 * j + 1]), scale, -ms);
                            if (!(a.variant & 4)) { p0 = exp2f(p0); p1 = exp2f(p1); }
                            p0 = (c0 + 2 * j < a.nk) ? p0 : 0.f;
                            p1 = (c0 + 2 * j + 1 < a.nk) ? p1 : 0.f;
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
                            const float p0 = (c0 + 2 * j < a.nk) ? exp2f(fmaf(__uint_as_float(r[2 * j]), scale, -ms)) : 0.f;
                            const float p1 = (c0 + 2 * j + 1 < a.nk) ? exp2f(fmaf(__uint_as_float(r[2 * j + 1]), scale, -ms)) : 0.f;
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
            if (lane == 0) ptx::mbar_arrive(&p_full[wg]);
            // ---- epilogue: O / l -> fp16 -> global ----
            ptx::mbar_wait(&o_full[wg], ph, 16);
            ptx::tc_fence_after();
            if (!warp_has_rows) {
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&o_empty[wg]);
                continue;
            }
            const float inv = 1.0f / l;
            const float wa = alpha * inv, wb = beta * inv;
            const float wx = p_x * wa;   // weight of the extra key's value row (its score was normalised with half A's maximum)
            uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(b) * S + a.q0 + q_row) * a.out_ld + h * 64);
            const uint4* vx = XK ? reinterpret_cast<const uint4*>(xrow + 2 * D) : nullptr;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {       // 32 output columns at a time
                uint32_t oa[32], ob[32];
                ptx::tmem_ld_32x32(t_row + 128 + hh * 32, oa);
                if (NC > 0) ptx::tmem_ld_32x32(t_row + 192 + hh * 32, ob);
                ptx::tc_wait_ld();
                if (hh == 1) {                     // all TMEM reads of this tile are done: hand the buffer back to the MMA warp
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&o_empty[wg]);
                }
                if (q_row < a.nq) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            v[e] = NC > 0 ? fmaf(__uint_as_float(oa[8 * j + e]), wa, __uint_as_float(ob[8 * j + e]) * wb)
                                          : __uint_as_float(oa[8 * j + e]) * inv;
                        if (XK) {
                            const uint4 u = __ldg(vx + hh * 4 + j);
                            const __half2* hv = reinterpret_cast<const __half2*>(&u);
                            const float wxx = NC > 0 ? wx : p_x * inv;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(hv[e]);
                                v[2 * e] = fmaf(f.x, wxx, v[2 * e]);
                                v[2 * e + 1] = fmaf(f.y, wxx, v[2 * e + 1]);
                            }
                        }
                        const uint4 hi4 = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                        dst[hh * 4 + j] = hi4;
                        if (a.split_lo) {
                            const uint32_t hw[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
                            uint32_t lw[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                                lw[e] = pack_h2(v[2 * e] - f.x, v[2 * e + 1] - f.y);
                            }
                            dst[(D >> 3) + hh * 4 + j] = make_uint4(lw[0], lw[1], lw[2], lw[3]);   // + D halfs = D / 8 uint4
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


// =====================================================================================================================================
// Round 2, "tile stream with a free-standing O": the pipeline used whenever two score tiles AND one output tile fit tensor memory
// (2 * S_pad + 64 + 16 <= 512 columns, i.e. up to 208 keys: the 197-token ViT/16 sequence, the 50-token ViT/32 one).
//
// What the profiles of the kernel above said (profiles/r02_attention_notes.md): nothing is saturated, every softmax warpgroup walks
// the serial chain  S-MMA -> softmax -> P.V -> O read-out -> S-MMA  (9 500 clk per tile) because O lives INSIDE the score buffer
// (dead upper half of S), so the next S cannot be issued into a buffer before its O has been read out.  Here
//   * O gets its own 64 columns [416, 480) and is drained by a DEDICATED epilogue warpgroup (warps 12-15), so a score buffer is
//     free again as soon as P.V has been issued: the MMA thread issues  P.V_i, L_i, S_{i+2}  back to back (tcgen05.mma executes in
//     issue order, so S_{i+2} may overwrite P_i behind P.V_i without a barrier in between);
//   * the MMA warp runs its loop warp-uniformly and one ELECTED lane issues, so descriptor arithmetic stays on the uniform datapath
//     (from a `lane == 0` branch the 30 MMAs of a tile took 2 850 clk to issue: an R2UR round trip per operand);
//   * the softmax is two passes over TMEM in 64-column loads: a tcgen05.wait::ld costs ~120 clk of warp time however little it has to
//     wait for (tools/attn6_trace.py: 7 x 32-column pieces took 1 340 clk for pass 1 alone), so prefetching does not help, only fewer
//     waits do; packed fma.rn.f32x2 for the scale / shift and the row sum, FMNMX3 for the maximum; the row sums reach the epilogue
//     warps through shared memory (a row-sum MMA against a tile of ones was tried: 13 N = 16 MMAs cost the tensor pipe ~600 clk / tile);
//   * a fraction of the exponentials (EMU of every 32) is evaluated on the FMA pipe instead of the MUFU (Cody-Waite: 2^x =
//     2^floor(x) * p(x - floor(x)), p = degree-3 minimax, exponent spliced in with one shift-add): MUFU does 4 lanes / clk / scheduler,
//     i.e. 8 clk per warp instruction, and two softmax warps share each scheduler's unit;
//   * the second query tile of a 197-token job holds 69 rows: on odd jobs it is loaded as tokens [nq - 128, nq) instead of
//     [128, 256), which moves its valid rows (and the MUFU work) from lane quarters 0-2 to quarters 1-3.
// Warps: 0 TMA (Q, K), 3 TMA (V), 1 MMA issuer, 2 TMEM allocator, 4-7 / 8-11 softmax warpgroups (score buffer 0 / 1), 12-15 epilogue.
constexpr int ATC6_THREADS = 512;
// diagnostics (variant bit 128): block 0 records clock64() at the hand-over points of its first tiles; tools/attn6_trace.py prints them
constexpr int ATC6_TRACE_N = 256;
__device__ long long g_attn6_trace[4][ATC6_TRACE_N];
#define ATC6_TRACE(role, slot)                                                                               \
    do {                                                                                                     \
        if (trace && (slot) < ATC6_TRACE_N) g_attn6_trace[role][slot] = clock64();                           \
    } while (0)
constexpr int ATC6_O_COL = 416;

// pass 1 over W score columns starting at key c0: four running maxima (FMNMX3 chains)
template <int W, bool MASK>
__device__ __forceinline__ void sm6_max_piece(uint32_t t_src, int c0, int nk, float (&m)[4]) {
    uint32_t r[W];
    tmem_ld_w<W>(t_src, r);
    ptx::tc_wait_ld();
#pragma unroll
    for (int j = 0; j < W; j += 8) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (!MASK) {
                m[e] = fmaxf(m[e], fmaxf(__uint_as_float(r[j + 2 * e]), __uint_as_float(r[j + 2 * e + 1])));
            } else {
                if (c0 + j + 2 * e < nk) m[e] = fmaxf(m[e], __uint_as_float(r[j + 2 * e]));
                if (c0 + j + 2 * e + 1 < nk) m[e] = fmaxf(m[e], __uint_as_float(r[j + 2 * e + 1]));
            }
        }
    }
}
// pass 2 over W score columns: P = 2^(s * scale - m * scale) -> fp16 pairs -> W / 2 TMEM columns at t_dst; the fp32 values are summed
// into lsum (packed pair of partial sums).  EMU of every 16 pairs take the FMA-pipe exponential, spread evenly between the MUFU ones.
template <int W, bool MASK, int EMU>
__device__ __forceinline__ void sm6_exp_piece(uint32_t t_src, uint32_t t_dst, int c0, int nk, uint64_t scale2, uint64_t negms2, uint64_t& lsum) {
    uint32_t r[W];
    tmem_ld_w<W>(t_src, r);
    ptx::tc_wait_ld();
#pragma unroll
    for (int g = 0; g < W / 16; ++g) {
        uint32_t pk[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = g * 8 + jj;          // pair index inside the piece
            const uint64_t x = fma2f(pk2f(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), scale2, negms2);
            float p0, p1;
            if ((((j & 15) + 1) * EMU) / 16 > ((j & 15) * EMU) / 16) {
                ex2_emu2(x, p0, p1);
            } else {
                float x0, x1;
                upk2f(x, x0, x1);
                p0 = ex2_mufu(x0);
                p1 = ex2_mufu(x1);
            }
            if (MASK) {
                p0 = (c0 + 2 * j < nk) ? p0 : 0.f;
                p1 = (c0 + 2 * j + 1 < nk) ? p1 : 0.f;
            }
            lsum = add2f(lsum, pk2f(p0, p1));
            pk[jj] = pack_h2(p0, p1);
        }
        ptx::tmem_st_32x8(t_dst + g * 8, pk);
    }
}

// ONE pass over W score columns: piece maximum, then P = 2^((s - m_ref) * scale) against the row's running REFERENCE maximum m_ref.
// m_ref is the exact maximum of the first piece; a later piece only moves it when its own maximum exceeds m_ref by more than 8 in
// the log2 domain (P would leave [0, 2^8], still far inside fp16), and then the P columns already written are rescaled in TMEM
// (rare: needs a 256-fold jump of the row's largest probability after the first 64 keys).  Reading the scores once instead of twice
// is what matters: a warp moves ~51 B/clk TMEM -> registers (2.5 clk per column) and pays ~123 clk per tcgen05.wait::ld
// (profiles/r02_ldtm_microbench.log), so a second pass over 208 columns costs ~1 000 clk per tile before any arithmetic.
template <int W, bool MASK, bool FIRST, int EMU>
__device__ __forceinline__ void sm6_piece(uint32_t t_row, int c0, int nk, float scale, uint64_t scale2, float& m_ref, uint64_t& lsum) {
    uint32_t r[W];
    tmem_ld_w<W>(t_row + c0, r);
    ptx::tc_wait_ld();
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < W; j += 8) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (!MASK) {
                m[e] = fmaxf(m[e], fmaxf(__uint_as_float(r[j + 2 * e]), __uint_as_float(r[j + 2 * e + 1])));
            } else {
                if (c0 + j + 2 * e < nk) m[e] = fmaxf(m[e], __uint_as_float(r[j + 2 * e]));
                if (c0 + j + 2 * e + 1 < nk) m[e] = fmaxf(m[e], __uint_as_float(r[j + 2 * e + 1]));
            }
        }
    }
    const float pm = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
    if (FIRST) {
        m_ref = pm;
    } else {
        const bool need = pm * scale > fmaf(m_ref, scale, 8.0f);
        if (__any_sync(0xffffffffu, need)) {      // rare path: move the reference, rescale what this row has written so far
            ptx::tc_wait_st();                    // the P columns written so far must have landed before they are read back
            const float new_ref = need ? pm : m_ref;
            const float f = ex2_mufu((m_ref - new_ref) * scale);          // 1 for the rows that keep their reference
            const __half2 f2 = __float2half2_rn(f);
            for (int pc = 0; pc < (c0 >> 1); pc += 16) {                   // c0 is a multiple of 32: whole 16-column groups of packed P
                uint32_t pr[16];
                ptx::tmem_ld_32x16(t_row + pc, pr);
                ptx::tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const __half2 v = __hmul2(*reinterpret_cast<const __half2*>(&pr[j]), f2);
                    pr[j] = *reinterpret_cast<const uint32_t*>(&v);
                }
                ptx::tmem_st_32x16(t_row + pc, pr);
            }
            float l0, l1;
            upk2f(lsum, l0, l1);
            lsum = pk2f(l0 * f, l1 * f);
            m_ref = new_ref;
        }
    }
    const float nms = -m_ref * scale;
    const uint64_t negms2 = pk2f(nms, nms);
#pragma unroll
    for (int g = 0; g < W / 16; ++g) {
        uint32_t pk[8];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
            const int j = g * 8 + jj;          // pair index inside the piece
            const uint64_t x = fma2f(pk2f(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])), scale2, negms2);
            float p0, p1;
            if ((((j & 15) + 1) * EMU) / 16 > ((j & 15) * EMU) / 16) {
                ex2_emu2(x, p0, p1);
            } else {
                float x0, x1;
                upk2f(x, x0, x1);
                p0 = ex2_mufu(x0);
                p1 = ex2_mufu(x1);
            }
            if (MASK) {
                p0 = (c0 + 2 * j < nk) ? p0 : 0.f;
                p1 = (c0 + 2 * j + 1 < nk) ? p1 : 0.f;
            }
            lsum = add2f(lsum, pk2f(p0, p1));
            pk[jj] = pack_h2(p0, p1);
        }
        ptx::tmem_st_32x8(t_row + (c0 >> 1) + g * 8, pk);
    }
}

template <int EMU, int NC>   // NC > 0: S_pad = 16 * NC at compile time (fully unrolled MMA issue), 0: run-time S_pad
__global__ void __launch_bounds__(ATC6_THREADS, 1)
attention_tc6_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv, __half* __restrict__ out,
                     const AttnArgs a) {
    const int S_pad = NC > 0 ? NC * 16 : a.S_pad;               // <= 208
    extern __shared__ uint8_t smem_raw_att[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_att) + 1023) & ~uintptr_t(1023));
    const int heads = a.heads, S = a.S;
    const int D = heads * 64;
    const int kv_bytes = S_pad * 128;
    const int n_qt = (a.nq + 127) / 128;                        // 1 or 2 query tiles per job
    const int q_bytes = n_qt * Q_TILE_BYTES;
    const int stage_bytes = q_bytes + 2 * kv_bytes;             // per job stage: [Q tiles | K | V], a multiple of 1024
    float* lsm = reinterpret_cast<float*>(smem + 2 * stage_bytes);   // [4 tiles in flight][128 rows] row sums, softmax -> epilogue
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes + 2048);
    uint64_t* qk_full = bars;        // [2 stages] TMA -> MMA   (Q tiles + K)
    uint64_t* qk_empty = bars + 2;   // [2 stages] MMA -> TMA   (the job's S MMAs have retired)
    uint64_t* v_full = bars + 4;     // [2 stages] TMA -> MMA   (V)
    uint64_t* v_empty = bars + 6;    // [2 stages] MMA -> TMA   (the job's P.V MMAs have retired)
    uint64_t* s_full = bars + 8;     // [2 buffers] MMA -> softmax
    uint64_t* p_full = bars + 10;    // [2 buffers] softmax -> MMA
    uint64_t* o_full = bars + 12;    // MMA -> epilogue
    uint64_t* o_empty = bars + 13;   // epilogue -> MMA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 14);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_jobs = a.B * heads;
    const int my_jobs = blockIdx.x < n_jobs ? (n_jobs - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int n_tiles = my_jobs * n_qt;
    const bool flip = n_qt == 2 && !(a.variant & 64);
    const bool trace = (a.variant & 128) && blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 4 || warp == 8 || warp == 12);
    // first token (inside the query window) of tile g of this CTA's jt-th job, and the first row of the tile that is this tile's to
    // compute (rows below it repeat tokens of tile 0)
    auto tile_tok0 = [&](int jt, int g) -> int { return (g == 1 && flip && (jt & 1)) ? a.nq - 128 : g * 128; };

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
        }
        ptx::mbar_init(o_full, 1);
        ptx::mbar_init(o_empty, 4);
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
            for (int jt = 0; jt < my_jobs; ++jt) {
                const int job = blockIdx.x + jt * gridDim.x;
                const int st = jt & 1;
                const uint32_t sph = (jt >> 1) & 1;
                const int b = job / heads, h = job - b * heads;
                uint8_t* sb = smem + st * stage_bytes;
                if (warp == 0) {
                    ptx::mbar_wait(&qk_empty[st], sph ^ 1, 61);
                    ptx::mbar_arrive_expect_tx(&qk_full[st], n_qt * Q_TILE_BYTES + kv_bytes);
                    ptx::tma_load_2d(sb + q_bytes, &map_kv, &qk_full[st], D + h * 64, b * S + a.k0);
                    for (int g = 0; g < n_qt; ++g)
                        ptx::tma_load_2d(sb + g * Q_TILE_BYTES, &map_q, &qk_full[st], h * 64, b * S + a.q0 + tile_tok0(jt, g));
                } else {
                    ptx::mbar_wait(&v_empty[st], sph ^ 1, 62);
                    ptx::mbar_arrive_expect_tx(&v_full[st], kv_bytes);
                    ptx::tma_load_2d(sb + q_bytes + kv_bytes, &map_kv, &v_full[st], 2 * D + h * 64, b * S + a.k0);
                }
            }
        }
        __syncwarp();
      } else if (warp == 1) {
        // The WHOLE warp walks the loop (waits included) and one elected lane issues: with warp-uniform control flow the descriptor
        // arithmetic stays on the uniform datapath.  Issued from a `lane == 0` branch every operand of every tcgen05.mma went through an
        // R2UR first and the 30 MMAs of a tile took 2 850 clk to ISSUE (tools/attn6_trace.py) -- longer than they take to execute.
        if (n_tiles > 0) {
            const uint32_t idesc_s = ptx::make_idesc_f16(128, S_pad);
            const uint32_t idesc_o = ptx::make_idesc_f16(128, 64, false, true);
            const int k_steps_pv = S_pad / 16;
            auto issue_s = [&](int i) {       // S_i = Q_g K^T into score buffer i & 1
                const int jt = i / n_qt, g = i - jt * n_qt, st = jt & 1, buf = i & 1;
                uint8_t* sb = smem + st * stage_bytes;
                if (g == 0) ptx::mbar_wait(&qk_full[st], (jt >> 1) & 1, 63);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + q_bytes));
                    const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + g * Q_TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::tc_mma_f16<1>(tmem_base + buf * S_pad, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                    ptx::tc_commit<1>(&s_full[buf]);
                    if (g == n_qt - 1) ptx::tc_commit<1>(&qk_empty[st]);   // Q and K of this stage are free once the S MMAs retire
                }
                __syncwarp();
            };
            issue_s(0);
            if (n_tiles > 1) issue_s(1);
            for (int i = 0; i < n_tiles; ++i) {
                const int jt = i / n_qt, g = i - jt * n_qt, st = jt & 1, buf = i & 1;
                uint8_t* sb = smem + st * stage_bytes;
                ATC6_TRACE(2, 4 * i);
                if (g == 0) ptx::mbar_wait(&v_full[st], (jt >> 1) & 1, 64);
                ptx::mbar_wait(&p_full[buf], (i >> 1) & 1, 65);
                ATC6_TRACE(2, 4 * i + 1);
                ptx::mbar_wait(o_empty, (i & 1) ^ 1, 66);              // O / L of tile i - 1 have been read out
                ATC6_TRACE(2, 4 * i + 2);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + q_bytes + kv_bytes), 64);
                    const uint32_t p_addr = tmem_base + buf * S_pad;
                    if (NC > 0) {
#pragma unroll
                        for (int ks = 0; ks < (NC > 0 ? NC : 1); ++ks) {   // 16 keys per step: 8 TMEM columns of P, two 8-key groups (2 KB) of V
                            ptx::tc_mma_f16_ts(tmem_base + ATC6_O_COL, p_addr + ks * 8, v_desc + ks * 128, idesc_o, ks != 0 ? 1u : 0u);
                        }
                    } else {
                        for (int ks = 0; ks < k_steps_pv; ++ks) {
                            ptx::tc_mma_f16_ts(tmem_base + ATC6_O_COL, p_addr + ks * 8, v_desc + ks * 128, idesc_o, ks != 0 ? 1u : 0u);
                        }
                    }
                    ptx::tc_commit<1>(o_full);
                    if (g == n_qt - 1) ptx::tc_commit<1>(&v_empty[st]);
                }
                __syncwarp();
                // the score buffer is free as soon as P.V_i is in the pipe: tcgen05.mma executes in issue order
                if (i + 2 < n_tiles) issue_s(i + 2);
                ATC6_TRACE(2, 4 * i + 3);
            }
        }
      }
    } else if (warp < 12) {
        // ---------------- softmax warpgroups ----------------
        const int wg = (warp - 4) >> 2;   // = score buffer
        const int q = warp & 3;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + wg * S_pad;
        const float scale = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)
        const uint64_t scale2 = pk2f(scale, scale);
        for (int i = wg; i < n_tiles; i += 2) {
            const int jt = i / n_qt, g = i - jt * n_qt;
            const uint32_t ph = (i >> 1) & 1;
            const int tok0 = tile_tok0(jt, g);
            // rows of this warp: tokens tok0 + 32 q .. + 31; the tile's own tokens are [g * 128, nq)
            const bool warp_has_rows = tok0 + q * 32 + 31 >= g * 128 && tok0 + q * 32 < a.nq;
            ATC6_TRACE(wg, 4 * (i >> 1));
            ptx::mbar_wait(&s_full[wg], ph, 67);
            ATC6_TRACE(wg, 4 * (i >> 1) + 1);
            ptx::tc_fence_after();
            if (warp_has_rows) {
                float m_ref;
                uint64_t lsum = pk2f(0.f, 0.f);
                int c0 = 0;
                if (a.variant & 512) {
                    // ---- two passes (diagnostics): exact row maximum first ----
                    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                    for (; c0 + 64 <= S_pad; c0 += 64) {
                        if (c0 + 64 <= a.nk) sm6_max_piece<64, false>(t_row + c0, c0, a.nk, m);
                        else sm6_max_piece<64, true>(t_row + c0, c0, a.nk, m);
                    }
                    if (S_pad - c0 >= 32) {
                        if (c0 + 32 <= a.nk) sm6_max_piece<32, false>(t_row + c0, c0, a.nk, m);
                        else sm6_max_piece<32, true>(t_row + c0, c0, a.nk, m);
                        c0 += 32;
                    }
                    if (S_pad - c0 >= 16) {
                        if (c0 + 16 <= a.nk) sm6_max_piece<16, false>(t_row + c0, c0, a.nk, m);
                        else sm6_max_piece<16, true>(t_row + c0, c0, a.nk, m);
                    }
                    ATC6_TRACE(wg, 4 * (i >> 1) + 2);
                    const float nms = -fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])) * scale;
                    const uint64_t negms2 = pk2f(nms, nms);
                    c0 = 0;
                    for (; c0 + 64 <= S_pad; c0 += 64) {
                        if (c0 + 64 <= a.nk) sm6_exp_piece<64, false, EMU>(t_row + c0, t_row + (c0 >> 1), c0, a.nk, scale2, negms2, lsum);
                        else sm6_exp_piece<64, true, EMU>(t_row + c0, t_row + (c0 >> 1), c0, a.nk, scale2, negms2, lsum);
                    }
                    if (S_pad - c0 >= 32) {
                        if (c0 + 32 <= a.nk) sm6_exp_piece<32, false, EMU>(t_row + c0, t_row + (c0 >> 1), c0, a.nk, scale2, negms2, lsum);
                        else sm6_exp_piece<32, true, EMU>(t_row + c0, t_row + (c0 >> 1), c0, a.nk, scale2, negms2, lsum);
                        c0 += 32;
                    }
                    if (S_pad - c0 >= 16) {
                        if (c0 + 16 <= a.nk) sm6_exp_piece<16, false, EMU>(t_row + c0, t_row + (c0 >> 1), c0, a.nk, scale2, negms2, lsum);
                        else sm6_exp_piece<16, true, EMU>(t_row + c0, t_row + (c0 >> 1), c0, a.nk, scale2, negms2, lsum);
                    }
                } else {
                    // ---- one pass with a lazily moved reference maximum (sm6_piece) ----
                    if (S_pad >= 64) {
                        if (64 <= a.nk) sm6_piece<64, false, true, EMU>(t_row, 0, a.nk, scale, scale2, m_ref, lsum);
                        else sm6_piece<64, true, true, EMU>(t_row, 0, a.nk, scale, scale2, m_ref, lsum);
                        c0 = 64;
                        for (; c0 + 64 <= S_pad; c0 += 64) {
                            if (c0 + 64 <= a.nk) sm6_piece<64, false, false, EMU>(t_row, c0, a.nk, scale, scale2, m_ref, lsum);
                            else sm6_piece<64, true, false, EMU>(t_row, c0, a.nk, scale, scale2, m_ref, lsum);
                        }
                        if (S_pad - c0 >= 32) {
                            if (c0 + 32 <= a.nk) sm6_piece<32, false, false, EMU>(t_row, c0, a.nk, scale, scale2, m_ref, lsum);
                            else sm6_piece<32, true, false, EMU>(t_row, c0, a.nk, scale, scale2, m_ref, lsum);
                            c0 += 32;
                        }
                        if (S_pad - c0 >= 16) {
                            if (c0 + 16 <= a.nk) sm6_piece<16, false, false, EMU>(t_row, c0, a.nk, scale, scale2, m_ref, lsum);
                            else sm6_piece<16, true, false, EMU>(t_row, c0, a.nk, scale, scale2, m_ref, lsum);
                        }
                    } else if (S_pad >= 32) {
                        if (32 <= a.nk) sm6_piece<32, false, true, EMU>(t_row, 0, a.nk, scale, scale2, m_ref, lsum);
                        else sm6_piece<32, true, true, EMU>(t_row, 0, a.nk, scale, scale2, m_ref, lsum);
                        if (S_pad >= 48) {
                            if (48 <= a.nk) sm6_piece<16, false, false, EMU>(t_row, 32, a.nk, scale, scale2, m_ref, lsum);
                            else sm6_piece<16, true, false, EMU>(t_row, 32, a.nk, scale, scale2, m_ref, lsum);
                        }
                    } else {
                        if (16 <= a.nk) sm6_piece<16, false, true, EMU>(t_row, 0, a.nk, scale, scale2, m_ref, lsum);
                        else sm6_piece<16, true, true, EMU>(t_row, 0, a.nk, scale, scale2, m_ref, lsum);
                    }
                }
                float l0, l1;
                upk2f(lsum, l0, l1);
                lsm[(i & 3) * 128 + q * 32 + lane] = l0 + l1;      // read by the epilogue warp of the same lane quarter after o_full
                ptx::tc_wait_st();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[wg]);
            ATC6_TRACE(wg, 4 * (i >> 1) + 3);
        }
    } else {
        // ---------------- epilogue warpgroup: O / L -> fp16 -> global ----------------
        const int q = warp & 3;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        for (int i = 0; i < n_tiles; ++i) {
            const int jt = i / n_qt, g = i - jt * n_qt;
            const int job = blockIdx.x + jt * gridDim.x;
            const int b = job / heads, h = job - b * heads;
            const int tok = tile_tok0(jt, g) + q * 32 + lane;     // token inside the query window
            ATC6_TRACE(3, 4 * i);
            ptx::mbar_wait(o_full, i & 1, 68);
            ATC6_TRACE(3, 4 * i + 1);
            ptx::tc_fence_after();
            uint32_t o[64];
            ptx::tmem_ld_32x32(t_row + ATC6_O_COL, *reinterpret_cast<uint32_t(*)[32]>(&o[0]));
            ptx::tmem_ld_32x32(t_row + ATC6_O_COL + 32, *reinterpret_cast<uint32_t(*)[32]>(&o[32]));
            const float lsum = lsm[(i & 3) * 128 + q * 32 + lane];
            ptx::tc_wait_ld();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_relaxed(o_empty);    // the values are in registers: P.V of the next tile may overwrite O
            ATC6_TRACE(3, 4 * i + 2);
            if (tok >= g * 128 && tok < a.nq) {
                const float inv = 1.0f / lsum;
                uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(b) * S + a.q0 + tok) * a.out_ld + h * 64);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * j + e]) * inv;
                    const uint4 hi4 = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                    dst[j] = hi4;
                    if (a.split_lo) {
                        const uint32_t hw[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
                        uint32_t lw[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
                            lw[e] = pack_h2(v[2 * e] - f.x, v[2 * e + 1] - f.y);
                        }
                        dst[(D >> 3) + j] = make_uint4(lw[0], lw[1], lw[2], lw[3]);   // + D halfs = D / 8 uint4
                    }
                }
            }
            ATC6_TRACE(3, 4 * i + 3);
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem_base, 512);
    }
}

template <int EMU, int NC>
int launch_attn6(ap_ctx* ctx, const AttnPlan* plan, __half* out, const AttnArgs& a, int grid, cudaStream_t stream) {
    auto kern = attention_tc6_kernel<EMU, NC>;
    const size_t n_qt = (a.nq + 127) / 128;
    const size_t smem = 2 * (n_qt * (size_t)Q_TILE_BYTES + 2 * (size_t)a.S_pad * 128) + 2048 + 15 * 8 + 16 + 1024;
    static PerDeviceOnce attr;   // per instantiation
    if (attr.need(ctx->device)) {
        const int max_smem = 2 * (2 * Q_TILE_BYTES + 2 * 208 * 128) + 2048 + 15 * 8 + 16 + 1024;
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr.done(ctx->device);
    }
    AP_CHECK_CUDA(ctx, ap_launch_pdl(kern, dim3(grid), dim3(ATC6_THREADS), smem, stream, 1, ctx->pdl != 0, plan->map_q, plan->map_kv, out, a));
    return AP_OK;
}

template <int NC, bool XK>
int launch_attn(ap_ctx* ctx, const AttnPlan* plan, const __half* qkv, __half* out, const AttnArgs& a, int grid, size_t smem,
                cudaStream_t stream) {
    auto kern = attention_tc_kernel<NC, XK>;
    static PerDeviceOnce attr;   // per instantiation
    if (attr.need(ctx->device)) {
        const int max_smem = 2 * (3 * Q_TILE_BYTES + 2 * 256 * 128) + 17 * 8 + 16 + 1024;
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr.done(ctx->device);
    }
    AP_CHECK_CUDA(ctx, ap_launch_pdl(kern, dim3(grid), dim3(ATC_THREADS), smem, stream, 1, ctx->pdl != 0, plan->map_q, plan->map_kv, qkv, out, a));
    return AP_OK;
}

}  // namespace

// S tokens per image; the key window is the whole sequence when it fits 256 MMA keys, else tokens [1, S) with token 0 (the class
// token) as the extra key.  All S query rows are covered: 257 = two full 128-row tiles + a third tile with one valid row (a
// separate class-token kernel re-read every K and V from L2 and cost half as much again as this kernel).
int ap_attention_tc_plan(ap_ctx* ctx, AttnPlan* plan, const __half* qkv, int rows, int S, int heads) {
    const int D = heads * 64;
    AP_REQUIRE(ctx, S >= 1 && S <= 257, "attention(tcgen05): S=%d unsupported (1..257)", S);
    plan->qkv = qkv;
    plan->xkey = S > 256 ? 0 : -1;
    plan->k0 = S > 256 ? 1 : 0;
    plan->nk = S > 256 ? S - 1 : S;
    plan->q0 = 0;
    plan->nq = S;       // 257 queries: two full tiles + a third tile that holds the last token only (its other warps idle)
    plan->S_pad = (plan->nk + 15) / 16 * 16;
    int rc = ap_make_tmap_f16_2d(ctx, &plan->map_q, qkv, (uint64_t)rows, (uint64_t)3 * D, (uint64_t)3 * D, 128, 64);
    if (rc) return rc;
    rc = ap_make_tmap_f16_2d(ctx, &plan->map_q16, qkv, (uint64_t)rows, (uint64_t)3 * D, (uint64_t)3 * D, 16, 64);
    if (rc) return rc;
    return ap_make_tmap_f16_2d(ctx, &plan->map_kv, qkv, (uint64_t)rows, (uint64_t)3 * D, (uint64_t)3 * D, plan->S_pad, 64);
}

int ap_attention_tc_run(ap_ctx* ctx, const AttnPlan* plan, __half* out, int B, int S, int heads, cudaStream_t stream, int out_ld,
                        int split_lo) {
    if (B == 0) return AP_OK;
    const int S_pad = plan->S_pad;
    const size_t n_qt = (plan->nq + 127) / 128;
    const size_t smem = 2 * (n_qt * (size_t)Q_TILE_BYTES + 2 * (size_t)S_pad * 128) + 17 * 8 + 16 + 1024;
    const int jobs = B * heads;
    const int grid = jobs < ctx->sm_count ? jobs : ctx->sm_count;
    AttnArgs a;
    a.B = B; a.S = S; a.heads = heads; a.q0 = plan->q0; a.nq = plan->nq; a.k0 = plan->k0; a.nk = plan->nk; a.xkey = plan->xkey;
    a.S_pad = S_pad; a.variant = ctx->attn_variant;
    a.out_ld = out_ld > 0 ? out_ld : heads * 64;
    a.split_lo = split_lo;
    AP_REQUIRE(ctx, a.out_ld >= (split_lo ? 2 : 1) * heads * 64 && a.out_ld % 8 == 0, "attention: output row stride %d too small", a.out_ld);
    int rc;
    {
        ProfScope prof(ctx, stream, AP_K_ATTENTION);
        const bool generic = (ctx->attn_variant & 16) != 0;
        // which pipeline: key-block units (attention_units.cu) where they win -- one key block (<= 128 keys: 46.8 us against 53.3 us at 50
        // tokens) and 209..256 keys (69.7 us against 90.6 us at 256) -- whole tiles with a free-standing O for 129..208 keys (197 tokens:
        // 46.6 us against 56.9 us: two 128 / 80-key units per tile cost more hand-overs than they hide); attn_variant 2048 / 1024 force one
        // the 257-token DINOv2 sequence on the units kernel (extra key + a short third query tile): correct, but its six units per job make
        // it slower than the round-1 pipeline (146.5 us against 125.5 us at 127 images x 16 heads) -- on request only
        const bool units_xk = plan->xkey >= 0 && S_pad == 256 && plan->nq == 257 && plan->q0 == 0 && (ctx->attn_variant & 2048);
        const bool units_ok = (plan->xkey < 0 && S_pad <= 256 && plan->nq <= 256) || units_xk;
        const bool tc6_ok = plan->xkey < 0 && S_pad <= 208 && plan->nq <= 256;
        const bool want_units = (ctx->attn_variant & 2048) || units_xk || (!(ctx->attn_variant & 1024) && (S_pad <= 128 || S_pad > 208));
        if (units_ok && !(ctx->attn_variant & 32) && (want_units || !tc6_ok)) {
            rc = ap_attention_units_run(ctx, plan, out, a, grid, stream);
        } else if (plan->xkey < 0 && S_pad <= 208 && plan->nq <= 256 && !(ctx->attn_variant & 32)) {   // two score tiles + a free-standing O fit TMEM
            const int emu = ctx->attn_emu;
            if (S_pad == 208 && !generic) {     // 197 tokens (ViT/16 @ 224)
                rc = emu == 0 ? launch_attn6<0, 13>(ctx, plan, out, a, grid, stream) : emu <= 4 ? launch_attn6<4, 13>(ctx, plan, out, a, grid, stream)
                     : emu <= 6 ? launch_attn6<6, 13>(ctx, plan, out, a, grid, stream) : launch_attn6<8, 13>(ctx, plan, out, a, grid, stream);
            } else {
                rc = emu == 0 ? launch_attn6<0, 0>(ctx, plan, out, a, grid, stream) : launch_attn6<6, 0>(ctx, plan, out, a, grid, stream);
            }
        } else if (plan->xkey >= 0) {
            if (S_pad == 256 && !generic) rc = launch_attn<16, true>(ctx, plan, plan->qkv, out, a, grid, smem, stream);
            else rc = launch_attn<0, true>(ctx, plan, plan->qkv, out, a, grid, smem, stream);
        } else if (S_pad == 208 && !generic) {   // 197 tokens (ViT/16 @ 224): register-resident single-pass softmax
            rc = launch_attn<13, false>(ctx, plan, plan->qkv, out, a, grid, smem, stream);
        } else {
            rc = launch_attn<0, false>(ctx, plan, plan->qkv, out, a, grid, smem, stream);
        }
        if (rc) return rc;
        AP_CHECK_LAUNCH(ctx, "attention_tc_kernel");
    }
    return AP_OK;
}

// diagnostics, not part of the public header: the clock64() trace block 0 of attention_tc6_kernel records with attn_variant bit 128
extern "C" int ap_debug_attn6_trace(ap_ctx* ctx, long long* host_out) {
    DeviceGuard guard(ctx);
    AP_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    AP_CHECK_CUDA(ctx, cudaMemcpyFromSymbol(host_out, g_attn6_trace, sizeof(long long) * 4 * ATC6_TRACE_N));
    return AP_OK;
}
