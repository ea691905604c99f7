// Multi-head self-attention on tcgen05 / TMEM, "key-block units" pipeline (head_dim 64, up to 256 MMA keys).
//
// Same contract and warp roles as attention_tc6_kernel (attention_tcgen05.cu).  What that kernel cannot do is keep a third score tile in
// flight: at 208 fp32 columns per tile, 2 tiles + O fill tensor memory, and a softmax warpgroup idles between publishing P_i and receiving
// S_{i+2} (~1 900 of the ~6 100 clk a job takes, profiles/r02_attention_notes.md section 4).  Here the unit of work is a (query tile,
// KEY BLOCK of <= 128 keys) pair:
//   * three rotating score buffers of 128 columns (unit c lives in buffer c % 3) + ONE O TILE PER SOFTMAX WARPGROUP = 3 x 128 + 2 x 64 =
//     512 columns.  The spare buffer always holds the next unit's scores, so a warpgroup that publishes P finds its next S waiting;
//   * the two key blocks of a query tile accumulate into the warpgroup's own O tile (P_a V_a with accumulate = 0, P_b V_b on top), which
//     a dedicated epilogue warpgroup drains -- the key-halves attempt of this round had one O tile for both warpgroups and that tile was
//     occupied from P_a V_a to the read-out (profiles/r02_attention_v7_keyhalves_timings.log);
//   * the one-pass softmax keeps ONE reference maximum per row across both key blocks.  Moving it in block b (its maximum exceeds the
//     reference by > 2^8) happens after P_a V_a has been accumulated, so that rare path also rescales the row's O accumulator in TMEM.
// Units are issued in the order  (t, a) (t+1, a) (t, b) (t+1, b)  for every pair of query tiles (tile i belongs to warpgroup i & 1), so the
// two warpgroups alternate on the tensor pipe.  Warps: 0 TMA (Q, K), 3 TMA (V), 1 MMA issuer (whole warp walks the loop, an elected lane
// issues), 2 TMEM allocator, 4-7 / 8-11 softmax warpgroups, 12-15 epilogue.
#include "ap_internal.cuh"
#include "ptx.cuh"
#include "attention_common.cuh"

namespace {

constexpr int ATU_THREADS = 512;
constexpr int ATU_O_COL = 384;       // O tile of warpgroup w at columns [384 + 64 w, +64)
constexpr int ATU_KB = 128;          // keys per block = columns per score buffer

// diagnostics (attn_variant bit 128): block 0 records clock64() at the hand-over points of its first units (tools/attn_units_trace.py)
constexpr int ATU_TRACE_N = 256;
__device__ long long g_attn_units_trace[4][ATU_TRACE_N];
#define ATU_TRACE(role, slot)                                                                                  \
    do {                                                                                                       \
        if (trace && (slot) < ATU_TRACE_N) g_attn_units_trace[role][slot] = clock64();                         \
    } while (0)

struct UnitRare {
    uint64_t* oa_done;     // P_a.V_a of this tile has retired (only waited for on the rare path)
    uint32_t phase;
    uint32_t t_o;          // this warp's lanes of the warpgroup's O tile
};

// One piece of W score columns at column c of the unit's buffer (absolute key index key0 + c): maximum, lazily moved reference, P.
// (Moving the rare path out of line halves the kernel's 8 900 instructions -- 27 % of the softmax warps' stall samples are instruction
// fetches -- but the call makes ptxas spill 240 B around it and the kernel got slower: 75.8 -> 96.7 us at 256 keys.)
template <int W, bool MASK, bool FIRST, bool BLOCK_B, int EMU>
__device__ __forceinline__ void unit_piece(uint32_t t_buf, int c, int key0, int nk, float scale, uint64_t scale2, float& m_ref, uint64_t& lsum,
                                           const UnitRare& rare, float extra, float& p_x) {
    uint32_t r[W];
    tmem_ld_w<W>(t_buf + c, r);
    ptx::tc_wait_ld();
    const int k0 = key0 + c;
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < W; j += 8) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (!MASK) {
                m[e] = fmaxf(m[e], fmaxf(__uint_as_float(r[j + 2 * e]), __uint_as_float(r[j + 2 * e + 1])));
            } else {
                if (k0 + j + 2 * e < nk) m[e] = fmaxf(m[e], __uint_as_float(r[j + 2 * e]));
                if (k0 + j + 2 * e + 1 < nk) m[e] = fmaxf(m[e], __uint_as_float(r[j + 2 * e + 1]));
            }
        }
    }
    const float pm = fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3]));
    if (FIRST) {
        m_ref = fmaxf(pm, extra);          // extra: score of the key outside the MMA window (-inf without one)
        p_x = ex2_mufu(fmaf(extra, scale, -m_ref * scale));
        lsum = pk2f(p_x, 0.f);
    } else {
        const bool need = pm * scale > fmaf(m_ref, scale, 8.0f);
        if (__any_sync(0xffffffffu, need)) {      // rare path: move the reference, rescale what this row has produced so far
            ptx::tc_wait_st();
            const float new_ref = need ? pm : m_ref;
            const float f = ex2_mufu((m_ref - new_ref) * scale);          // 1 for the rows that keep their reference
            const __half2 f2 = __float2half2_rn(f);
            for (int pc = 0; pc < (c >> 1); pc += 8) {                     // packed P of THIS unit written so far
                uint32_t pr[8];
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(pr[0]), "=r"(pr[1]), "=r"(pr[2]), "=r"(pr[3]), "=r"(pr[4]), "=r"(pr[5]), "=r"(pr[6]), "=r"(pr[7])
                             : "r"(t_buf + pc)
                             : "memory");
                ptx::tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const __half2 v = __hmul2(*reinterpret_cast<const __half2*>(&pr[j]), f2);
                    pr[j] = *reinterpret_cast<const uint32_t*>(&v);
                }
                ptx::tmem_st_32x8(t_buf + pc, pr);
            }
            if (BLOCK_B) {                         // block a is already inside O: rescale the accumulator rows
                ptx::mbar_wait(rare.oa_done, rare.phase, 95);
                ptx::tc_fence_after();
                for (int oc = 0; oc < 64; oc += 16) {
                    uint32_t orr[16];
                    ptx::tmem_ld_32x16(rare.t_o + oc, orr);
                    ptx::tc_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) orr[j] = __float_as_uint(__uint_as_float(orr[j]) * f);
                    ptx::tmem_st_32x16(rare.t_o + oc, orr);
                }
            }
            float l0, l1;
            upk2f(lsum, l0, l1);
            lsum = pk2f(l0 * f, l1 * f);
            p_x *= f;
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
                p0 = (k0 + 2 * j < nk) ? p0 : 0.f;
                p1 = (k0 + 2 * j + 1 < nk) ? p1 : 0.f;
            }
            lsum = add2f(lsum, pk2f(p0, p1));
            pk[jj] = pack_h2(p0, p1);
        }
        ptx::tmem_st_32x8(t_buf + (c >> 1) + g * 8, pk);
    }
}

// all pieces of one unit (len score columns at t_buf, first key key0): 64-column loads, the remainder in 16-column loads
template <bool BLOCK_B, int EMU>
__device__ __forceinline__ void unit_softmax(uint32_t t_buf, int len, int key0, int nk, float scale, uint64_t scale2, float& m_ref, uint64_t& lsum,
                                             const UnitRare& rare, float extra, float& p_x) {
    int c = 0;
    if (!BLOCK_B) {     // the row's first piece defines the reference maximum
        if (len >= 64) {
            if (key0 + 64 <= nk) unit_piece<64, false, true, false, EMU>(t_buf, 0, key0, nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
            else unit_piece<64, true, true, false, EMU>(t_buf, 0, key0, nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
            c = 64;
        } else {
            if (key0 + 16 <= nk) unit_piece<16, false, true, false, EMU>(t_buf, 0, key0, nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
            else unit_piece<16, true, true, false, EMU>(t_buf, 0, key0, nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
            c = 16;
        }
    }
    for (; c + 64 <= len; c += 64) {
        if (key0 + c + 64 <= nk) unit_piece<64, false, false, BLOCK_B, EMU>(t_buf, c, key0, nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
        else unit_piece<64, true, false, BLOCK_B, EMU>(t_buf, c, key0, nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
    }
    for (; c + 16 <= len; c += 16) {
        if (key0 + c + 16 <= nk) unit_piece<16, false, false, BLOCK_B, EMU>(t_buf, c, key0, nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
        else unit_piece<16, true, false, BLOCK_B, EMU>(t_buf, c, key0, nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
    }
}

// XK: one EXTRA key token `xkey` outside the MMA key window is folded in as a rank-1 update (its score is a 64-term dot product per row
// on the CUDA cores and joins the first piece's maximum, its value row is added to O in the epilogue): the 257-token DINOv2 sequence =
// 256 patch keys through the MMA + the class token.  Its 257th query rides in a third query tile that is loaded as a 16-row box.
template <int EMU, bool XK>
__global__ void __launch_bounds__(ATU_THREADS, 1)
attention_units_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_q16,
                       const __grid_constant__ CUtensorMap map_kv, const __half* __restrict__ qkv, __half* __restrict__ out, const AttnArgs a) {
    const int S_pad = a.S_pad;                                   // <= 256
    const int ka = S_pad < ATU_KB ? S_pad : ATU_KB, kb = S_pad - ka;   // keys of block a / block b (multiples of 16; kb may be 0)
    const int KBLK = kb > 0 ? 2 : 1;                             // key blocks = units per query tile
    extern __shared__ uint8_t smem_raw_att[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_att) + 1023) & ~uintptr_t(1023));
    const int heads = a.heads, S = a.S;
    const int D = heads * 64;
    const int kv_bytes = S_pad * 128;
    const int n_qt = (a.nq + 127) / 128;                        // 1 or 2 query tiles per job, 3 for the 257 queries of XK
    const int q_bytes = n_qt == 3 ? 2 * Q_TILE_BYTES + 2048 : n_qt * Q_TILE_BYTES;   // third tile: 16 rows (one is real)
    const int stage_bytes = q_bytes + 2 * kv_bytes;             // per job stage: [Q tiles | K | V], a multiple of 1024
    // row sums, softmax -> epilogue: [8 tiles][128 rows].  Eight slots, not four: with one key block per tile (<= 128 keys) the MMA warp
    // may be blocked at P.V of tile T (waiting for the read-out of tile T - 2) while S of tile T + 2 is already out, so the softmax of
    // T + 2 can finish before the epilogue has read the sums of T - 2 -- (T + 2) & 3 == (T - 2) & 3 was a real race for 50-token sequences
    float* lsm = reinterpret_cast<float*>(smem + 2 * stage_bytes);
    float* pxs = lsm + 8 * 128;                                  // [8 tiles][128 rows] probability of the extra key (XK)
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes + 8192);
    uint64_t* qk_full = bars;        // [2 stages] TMA -> MMA   (Q tiles + K)
    uint64_t* qk_empty = bars + 2;   // [2 stages] MMA -> TMA   (the job's S MMAs have retired)
    uint64_t* v_full = bars + 4;     // [2 stages] TMA -> MMA   (V)
    uint64_t* v_empty = bars + 6;    // [2 stages] MMA -> TMA   (the job's P.V MMAs have retired)
    uint64_t* s_full = bars + 8;     // [3 buffers] MMA -> softmax
    uint64_t* p_full = bars + 11;    // [3 buffers] softmax -> MMA
    uint64_t* oa_done = bars + 14;   // [2 warpgroups] P_a.V_a has retired (rare path of the softmax only)
    uint64_t* o_full = bars + 16;    // [2 warpgroups] MMA -> epilogue: the tile's last P.V has retired
    uint64_t* o_empty = bars + 18;   // [2 warpgroups] epilogue -> MMA
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 20);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_jobs = a.B * heads;
    const int my_jobs = blockIdx.x < n_jobs ? (n_jobs - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    const int n_tiles = my_jobs * n_qt;
    const int n_units = n_tiles * KBLK;
    const bool flip = n_qt == 2 && !(a.variant & 64);
    const bool trace = (a.variant & 128) && blockIdx.x == 0 && lane == 0 && (warp == 1 || warp == 4 || warp == 8 || warp == 12);
    auto tile_tok0 = [&](int jt, int g) -> int { return (g == 1 && flip && (jt & 1)) ? a.nq - 128 : g * 128; };
    // unit c of the stream -> (query tile, key block): pairs of tiles interleave their blocks, a trailing single tile runs a, b
    const int full_units = KBLK == 2 ? (n_tiles >> 1) * 4 : n_units;
    auto unit_tile = [&](int c) -> int { return KBLK == 1 ? c : (c < full_units ? 2 * (c >> 2) + (c & 1) : n_tiles - 1); };
    auto unit_blk = [&](int c) -> int { return KBLK == 1 ? 0 : (c < full_units ? (c >> 1) & 1 : c - full_units); };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&map_q);
        ptx::prefetch_tmap(&map_q16);
        ptx::prefetch_tmap(&map_kv);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < 2; ++i) {
            ptx::mbar_init(&qk_full[i], 1);
            ptx::mbar_init(&qk_empty[i], 1);
            ptx::mbar_init(&v_full[i], 1);
            ptx::mbar_init(&v_empty[i], 1);
            ptx::mbar_init(&oa_done[i], 1);
            ptx::mbar_init(&o_full[i], 1);
            ptx::mbar_init(&o_empty[i], 4);
        }
        for (int i = 0; i < 3; ++i) {
            ptx::mbar_init(&s_full[i], 1);
            ptx::mbar_init(&p_full[i], 4);
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
                    ptx::mbar_wait(&qk_empty[st], sph ^ 1, 91);
                    ptx::mbar_arrive_expect_tx(&qk_full[st], q_bytes + kv_bytes);
                    ptx::tma_load_2d(sb + q_bytes, &map_kv, &qk_full[st], D + h * 64, b * S + a.k0);
                    for (int g = 0; g < n_qt && g < 2; ++g)
                        ptx::tma_load_2d(sb + g * Q_TILE_BYTES, &map_q, &qk_full[st], h * 64, b * S + a.q0 + tile_tok0(jt, g));
                    if (n_qt == 3) ptx::tma_load_2d(sb + 2 * Q_TILE_BYTES, &map_q16, &qk_full[st], h * 64, b * S + a.q0 + 256);
                } else {
                    ptx::mbar_wait(&v_empty[st], sph ^ 1, 92);
                    ptx::mbar_arrive_expect_tx(&v_full[st], kv_bytes);
                    ptx::tma_load_2d(sb + q_bytes + kv_bytes, &map_kv, &v_full[st], 2 * D + h * 64, b * S + a.k0);
                }
            }
        }
        __syncwarp();
      } else if (warp == 1) {
        // the whole warp walks the loop, one elected lane issues (uniform-datapath descriptors; see attention_tc6_kernel)
        if (n_units > 0) {
            const uint32_t idesc_sa = ptx::make_idesc_f16(128, ka);
            const uint32_t idesc_sb = ptx::make_idesc_f16(128, kb > 0 ? kb : 16);
            const uint32_t idesc_o = ptx::make_idesc_f16(128, 64, false, true);
            int qk_waited = 0, v_waited = 0;       // jobs whose Q / K (V) are known to have landed
            auto issue_s = [&](int c) {            // scores of unit c into buffer c % 3
                const int tile = unit_tile(c), blk = unit_blk(c);
                const int jt = tile / n_qt, g = tile - jt * n_qt, st = jt & 1, buf = c % 3;
                uint8_t* sb = smem + st * stage_bytes;
                if (jt >= qk_waited) {
                    ptx::mbar_wait(&qk_full[st], (jt >> 1) & 1, 93);
                    qk_waited = jt + 1;
                }
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + q_bytes + blk * ATU_KB * 128));
                    const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + g * Q_TILE_BYTES));
                    const uint32_t idesc = blk ? idesc_sb : idesc_sa;
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ptx::tc_mma_f16<1>(tmem_base + buf * ATU_KB, q_desc + 2 * k, k_desc + 2 * k, idesc, k != 0 ? 1u : 0u);
                    ptx::tc_commit<1>(&s_full[buf]);
                    // the job's last unit in stream order is (its last tile, its last block): Q and K are free once these MMAs retire
                    if (g == n_qt - 1 && blk == KBLK - 1) ptx::tc_commit<1>(&qk_empty[st]);
                }
                __syncwarp();
            };
            for (int c = 0; c < 3 && c < n_units; ++c) issue_s(c);
            for (int c = 0; c < n_units; ++c) {
                const int tile = unit_tile(c), blk = unit_blk(c);
                const int jt = tile / n_qt, g = tile - jt * n_qt, st = jt & 1, buf = c % 3, wg = tile & 1;
                uint8_t* sb = smem + st * stage_bytes;
                if (jt >= v_waited) {
                    ptx::mbar_wait(&v_full[st], (jt >> 1) & 1, 94);
                    v_waited = jt + 1;
                }
                ATU_TRACE(2, 4 * c);
                ptx::mbar_wait(&p_full[buf], (c / 3) & 1, 96);
                ATU_TRACE(2, 4 * c + 1);
                if (blk == 0 && tile >= 2) ptx::mbar_wait(&o_empty[wg], ((tile >> 1) - 1) & 1, 97);   // this warpgroup's previous O has been read out
                ATU_TRACE(2, 4 * c + 2);
                ptx::tc_fence_after();
                if (ptx::elect_one()) {
                    const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(sb + q_bytes + kv_bytes + blk * ATU_KB * 128), 64);
                    const uint32_t p_addr = tmem_base + buf * ATU_KB;
                    const int steps = (blk ? kb : ka) >> 4;      // 16 keys per step: 8 TMEM columns of P, two 8-key groups (2 KB) of V
                    for (int ks = 0; ks < steps; ++ks)
                        ptx::tc_mma_f16_ts(tmem_base + ATU_O_COL + 64 * wg, p_addr + ks * 8, v_desc + ks * 128, idesc_o, (blk != 0 || ks != 0) ? 1u : 0u);
                    if (blk == KBLK - 1) ptx::tc_commit<1>(&o_full[wg]);
                    else ptx::tc_commit<1>(&oa_done[wg]);
                    if (g == n_qt - 1 && blk == KBLK - 1) ptx::tc_commit<1>(&v_empty[st]);
                }
                __syncwarp();
                if (c + 3 < n_units) issue_s(c + 3);     // behind P.V of unit c in the pipe: may overwrite its P
                ATU_TRACE(2, 4 * c + 3);
            }
        }
      }
    } else if (warp < 12) {
        // ---------------- softmax warpgroups: warpgroup w takes the query tiles i with i & 1 == w ----------------
        const int wg = (warp - 4) >> 2;
        const int q = warp & 3;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        const float scale = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)
        const uint64_t scale2 = pk2f(scale, scale);
        float m_ref = 0.f, p_x = 0.f;
        uint64_t lsum = pk2f(0.f, 0.f);
        for (int c = 0; c < n_units; ++c) {
            const int tile = unit_tile(c);
            if ((tile & 1) != wg) continue;
            const int blk = unit_blk(c), buf = c % 3;
            const int jt = tile / n_qt, g = tile - jt * n_qt;
            const int tok0 = tile_tok0(jt, g);
            // rows of this warp: tokens tok0 + 32 q .. + 31; the tile's own tokens are [g * 128, nq)
            const bool warp_has_rows = tok0 + q * 32 + 31 >= g * 128 && tok0 + q * 32 < a.nq;
            const uint32_t t_buf = t_lane + buf * ATU_KB;
            UnitRare rare;
            rare.oa_done = &oa_done[wg];
            rare.phase = (tile >> 1) & 1;
            rare.t_o = t_lane + ATU_O_COL + 64 * wg;
            float extra = -INFINITY;
            if (XK && blk == 0 && warp_has_rows) {   // score of the extra key: 64-term dot product on the CUDA cores (L2-resident rows)
                const int job = blockIdx.x + jt * gridDim.x;
                const int b = job / heads, h = job - b * heads;
                const int qr = tok0 + q * 32 + lane < a.nq ? tok0 + q * 32 + lane : a.nq - 1;
                const uint4* qp = reinterpret_cast<const uint4*>(qkv + (static_cast<int64_t>(b) * S + a.q0 + qr) * 3 * D + h * 64);
                const uint4* kp = reinterpret_cast<const uint4*>(qkv + (static_cast<int64_t>(b) * S + a.xkey) * 3 * D + D + h * 64);
                float s_x = 0.f;
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) s_x = dot8_h(__ldg(qp + cc), __ldg(kp + cc), s_x);
                extra = s_x;
            }
            ATU_TRACE(wg, 4 * c);
            ptx::mbar_wait(&s_full[buf], (c / 3) & 1, 98);
            ATU_TRACE(wg, 4 * c + 1);
            ptx::tc_fence_after();
            if (warp_has_rows) {
                if (blk == 0) {
                    unit_softmax<false, EMU>(t_buf, ka, 0, a.nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
                } else {
                    unit_softmax<true, EMU>(t_buf, kb, ATU_KB, a.nk, scale, scale2, m_ref, lsum, rare, extra, p_x);
                }
                if (blk == KBLK - 1) {
                    float l0, l1;
                    upk2f(lsum, l0, l1);
                    lsm[(tile & 7) * 128 + q * 32 + lane] = l0 + l1;      // read by the epilogue warp of the same lane quarter after o_full
                    if (XK) pxs[(tile & 7) * 128 + q * 32 + lane] = p_x;
                }
                ptx::tc_wait_st();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_full[buf]);
            ATU_TRACE(wg, 4 * c + 2);
        }
    } else {
        // ---------------- epilogue warpgroup: O / row sum -> fp16 -> global ----------------
        const int q = warp & 3;
        const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        for (int i = 0; i < n_tiles; ++i) {
            const int jt = i / n_qt, g = i - jt * n_qt, wg = i & 1;
            const int job = blockIdx.x + jt * gridDim.x;
            const int b = job / heads, h = job - b * heads;
            const int tok = tile_tok0(jt, g) + q * 32 + lane;     // token inside the query window
            ATU_TRACE(3, 4 * i);
            ptx::mbar_wait(&o_full[wg], (i >> 1) & 1, 99);
            ATU_TRACE(3, 4 * i + 1);
            ptx::tc_fence_after();
            uint32_t o[64];
            ptx::tmem_ld_32x64(t_lane + ATU_O_COL + 64 * wg, o);
            const float lsum = lsm[(i & 7) * 128 + q * 32 + lane];
            const float px = XK ? pxs[(i & 7) * 128 + q * 32 + lane] : 0.f;
            ptx::tc_wait_ld();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_relaxed(&o_empty[wg]);   // the values are in registers: the warpgroup's next tile may overwrite O
            ATU_TRACE(3, 4 * i + 2);
            if (tok >= g * 128 && tok < a.nq) {
                const float inv = 1.0f / lsum;
                uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(b) * S + a.q0 + tok) * a.out_ld + h * 64);
                const uint4* vx = XK ? reinterpret_cast<const uint4*>(qkv + (static_cast<int64_t>(b) * S + a.xkey) * 3 * D + 2 * D + h * 64) : nullptr;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float v[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(o[8 * j + e]);
                    if (XK) {                   // + p_x * (value row of the extra key)
                        const uint4 u = __ldg(vx + j);
                        const __half2* hv = reinterpret_cast<const __half2*>(&u);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float2 f = __half22float2(hv[e]);
                            v[2 * e] = fmaf(f.x, px, v[2 * e]);
                            v[2 * e + 1] = fmaf(f.y, px, v[2 * e + 1]);
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] *= inv;
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
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc<1>(tmem_base, 512);
    }
}

template <int EMU, bool XK>
int launch_units(ap_ctx* ctx, const AttnPlan* plan, __half* out, const AttnArgs& a, int grid, cudaStream_t stream) {
    auto kern = attention_units_kernel<EMU, XK>;
    const size_t n_qt = (a.nq + 127) / 128;
    const size_t q_bytes = n_qt == 3 ? 2 * (size_t)Q_TILE_BYTES + 2048 : n_qt * (size_t)Q_TILE_BYTES;
    const size_t smem = 2 * (q_bytes + 2 * (size_t)a.S_pad * 128) + 8192 + 21 * 8 + 16 + 1024;
    static PerDeviceOnce attr;   // per instantiation
    if (attr.need(ctx->device)) {
        const int max_smem = 2 * (2 * Q_TILE_BYTES + 2048 + 2 * 256 * 128) + 8192 + 21 * 8 + 16 + 1024;
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        attr.done(ctx->device);
    }
    AP_CHECK_CUDA(ctx, ap_launch_pdl(kern, dim3(grid), dim3(ATU_THREADS), smem, stream, 1, ctx->pdl != 0, plan->map_q, plan->map_q16, plan->map_kv, plan->qkv, out, a));
    return AP_OK;
}

}  // namespace

// plan->xkey < 0: S_pad <= 256 keys, at most two query tiles (nq <= 256).  plan->xkey >= 0: the 257-token case (256 MMA keys + the extra
// key, q0 = 0, nq = 257).
int ap_attention_units_run(ap_ctx* ctx, const AttnPlan* plan, __half* out, const AttnArgs& a, int grid, cudaStream_t stream) {
    const int emu = ctx->attn_emu;
    if (plan->xkey >= 0) {
        AP_REQUIRE(ctx, a.S_pad == 256 && a.nq == 257 && a.q0 == 0, "attention(units): extra key needs 256 MMA keys and 257 queries");
        return emu == 0 ? launch_units<0, true>(ctx, plan, out, a, grid, stream) : launch_units<4, true>(ctx, plan, out, a, grid, stream);
    }
    if (emu == 0) return launch_units<0, false>(ctx, plan, out, a, grid, stream);
    if (emu <= 2) return launch_units<2, false>(ctx, plan, out, a, grid, stream);
    if (emu <= 4) return launch_units<4, false>(ctx, plan, out, a, grid, stream);
    return launch_units<6, false>(ctx, plan, out, a, grid, stream);
}

// diagnostics, not part of the public header: the clock64() trace block 0 of attention_units_kernel records with attn_variant bit 128
extern "C" int ap_debug_attn_units_trace(ap_ctx* ctx, long long* host_out) {
    DeviceGuard guard(ctx);
    AP_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    AP_CHECK_CUDA(ctx, cudaMemcpyFromSymbol(host_out, g_attn_units_trace, sizeof(long long) * 4 * ATU_TRACE_N));
    return AP_OK;
}
