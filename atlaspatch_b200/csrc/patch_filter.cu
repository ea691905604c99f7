// a9b: the --no-fast-mode content filter of the coordinate path.
//
//   reference: PatchExtractionService._iter_patch_entries with fast_mode off reads every candidate patch through
//   IWSI.extract, cv2.resize()s it to patch_size when the read size differs, and drops it when is_black_patch or
//   is_white_patch says so (atlas_patch/services/extraction.py:105-119, atlas_patch/utils/image.py:7-41).
//
// Here the slide is resident in HBM, so the filter is one streaming pass over the candidates' pixels
// (patch^2 * scale^2 * 3 bytes per candidate, the kernel's algorithmic bytes) followed by an ordered compaction:
//
//   filter_count_kernel<SCALE>   grid = candidates x row bands; a CTA stages its band of slide rows in shared memory with
//                                aligned 16-byte loads, then every thread classifies 4 consecutive output pixels per step with
//                                OpenCV's 8-bit fixed-point definitions (restated and pinned in oracle/patch_filter.py):
//                                  gray = (9798 R + 19235 G + 3735 B + 2^14) >> 15           (COLOR_RGB2GRAY)
//                                  v = max, s = ((v - min) * sdiv[v] + 2^11) >> 12           (COLOR_RGB2HSV)
//                                  SCALE 2: pixel = (a + b + c + d + 2) >> 2                 (cv2.resize INTER_LINEAR at 2:1)
//                                the saturation/value test is folded into one table: white <=> min >= lo[v].
//   filter_compact_kernel        one CTA: keep = !(black/N >= f || white/N >= f) in float64 (numpy's bool mean), stable
//                                ballot/popc compaction of the kept rows.
#include <algorithm>
#include <vector>

#include "ap_internal.cuh"
#include "ptx.cuh"

namespace {
using namespace ptx;

constexpr int FILTER_THREADS = 256;
constexpr int FILTER_WARPS = FILTER_THREADS / 32;
constexpr int FILTER_CHUNKS = 4;             // bulk-copy pipeline depth: a CTA's band arrives as 4 groups of rows
constexpr int FILTER_MAX_BAND_ROWS = 64;     // slide rows per CTA (multiple of 8)
constexpr int FILTER_MAX_SMEM = 100 * 1024;  // staging bytes per CTA (2 CTAs / SM at the largest reads)

struct FilterParams {
    const uint8_t* slide;
    long long W, H, pitch;
    const int32_t* rows;  // n x 5 (x, y, read_w, read_h, level)
    int patch;            // output patch size P
    int band_rows;        // slide rows staged per CTA (multiple of 4 * SCALE)
    int bands;            // ceil(read / band_rows)
    int stride;           // bytes between staged rows (multiple of 16, odd number of 16-byte units)
    int gray_limit;       // black <=> 9798 R + 19235 G + 3735 B + 16384 < gray_limit  (= black_thresh << 15)
    int vec_ok;           // slide base and pitch are 16-byte aligned
    int32_t* counts;      // n x 2, zeroed
    uint16_t lo[256];     // white <=> min(R,G,B) >= lo[max(R,G,B)]   (256 = never)
};

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 9798 = 38 * 256 + 70, 19235 = 75 * 256 + 35, 3735 = 14 * 256 + 151: the gray sum of a packed pixel word [R G B x] is two
// byte dot products, no channel extraction
constexpr uint32_t GRAY_LO = 70u | (35u << 8) | (151u << 16), GRAY_HI = 38u | (75u << 8) | (14u << 16);

// Both tests are accumulated as sign bits: black <=> y - limit < 0 (the limit rides in the dot product's addend),
// white <=> (lo[v] - 1) - min < 0.
__device__ __forceinline__ void classify_word(uint32_t pix, uint32_t gray_bias, const int* lo_m1, uint32_t& nb, uint32_t& nw) {
    const uint32_t y = __dp4a(pix, GRAY_HI, 0u) * 256u + __dp4a(pix, GRAY_LO, gray_bias);  // (gray sum + 2^14) - limit
    nb += y >> 31;
    const uint32_t r = pix & 255u, g = __byte_perm(pix, 0u, 0x4441u), b = __byte_perm(pix, 0u, 0x4442u);
    const uint32_t v = __vimax3_u32(r, g, b), mn = __vimin3_u32(r, g, b);
    int d;  // opaque subtraction: keeps the sign-bit accumulation (IADD + LEA.HI) instead of a compare + select + add
    asm("sub.s32 %0, %1, %2;" : "=r"(d) : "r"(lo_m1[v]), "r"((int)mn));
    nw += (uint32_t)d >> 31;
}

// Same test with the table addressed through a hoisted shared-window address (interior fast path).
__device__ __forceinline__ void classify_fast(uint32_t pix, uint32_t gray_bias, uint32_t lo_addr, uint32_t& nb, uint32_t& nw) {
    const uint32_t y = __dp4a(pix, GRAY_HI, 0u) * 256u + __dp4a(pix, GRAY_LO, gray_bias);
    nb += y >> 31;
    const uint32_t r = pix & 255u, g = __byte_perm(pix, 0u, 0x4441u), b = __byte_perm(pix, 0u, 0x4442u);
    const uint32_t v = __vimax3_u32(r, g, b), mn = __vimin3_u32(r, g, b);
    int t, d;
    asm("ld.shared.s32 %0, [%1];" : "=r"(t) : "r"(lo_addr + v * 4u));
    asm("sub.s32 %0, %1, %2;" : "=r"(d) : "r"(t), "r"((int)mn));
    nw += (uint32_t)d >> 31;
}

template <int SCALE>
__global__ void __launch_bounds__(FILTER_THREADS)
filter_count_kernel(const __grid_constant__ FilterParams p) {
    extern __shared__ __align__(128) uint8_t stage[];
    __shared__ __align__(8) uint64_t bars[FILTER_CHUNKS];
    __shared__ int lo_s[256];  // lo[v] - 1
    __shared__ int tot[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    lo_s[tid & 255] = (int)p.lo[tid & 255] - 1;
    const uint32_t gray_bias = 16384u - (uint32_t)p.gray_limit;
    if (tid < 2) tot[tid] = 0;
    if (tid == 0) {
#pragma unroll
        for (int c = 0; c < FILTER_CHUNKS; ++c) mbar_init(&bars[c], 1);
        fence_barrier_init();
    }

    const long long cand = blockIdx.x / p.bands;
    const int band = blockIdx.x % p.bands;
    const int32_t* row = p.rows + cand * 5;
    const long long x = row[0], y = row[1];
    const int P = p.patch, read = P * SCALE, stride = p.stride;
    const int in_r0 = band * p.band_rows, in_r1 = min(read, in_r0 + p.band_rows);  // slide rows of this CTA, patch-relative
    const int chunk_rows = p.band_rows / FILTER_CHUNKS;

    // bytes of one slide row this patch touches, clipped to the slide; staged from the 16-byte boundary below it
    const bool inside = x >= 0 && y >= 0 && x < p.W;  // extraction never emits anything else; such a patch reads as all zero
    const long long sb = x * 3;
    const int row_bytes = inside ? (int)((min(x + read, p.W) - x) * 3) : 0;
    const int mis = p.vec_ok ? (int)(sb & 15) : 0;
    const int valid_w = row_bytes / 3;
    const uint32_t copy_bytes = (uint32_t)((mis + row_bytes + 15) & ~15);
    const bool bulk = inside && p.vec_ok;
    const bool interior = bulk && x + read <= p.W && y + in_r1 <= p.H && (P & 3) == 0;
    __syncthreads();

    // ---- producer: one bulk copy per slide row, one mbarrier per chunk of rows ------------------------------------------
    if (bulk && warp == 0) {
#pragma unroll
        for (int c = 0; c < FILTER_CHUNKS; ++c) {
            const int r0 = in_r0 + c * chunk_rows, r1 = min(in_r1, r0 + chunk_rows);
            const int live = (int)max(0ll, min((long long)r1, p.H - y) - r0);  // rows of the chunk that exist in the slide
            if (live > 0) {
                if (lane == 0) mbar_arrive_expect_tx(&bars[c], (uint32_t)live * copy_bytes);
                __syncwarp();
                for (int r = r0 + lane; r < r0 + live; r += 32)
                    bulk_copy_g2s(stage + (r - in_r0) * stride, p.slide + (y + r) * p.pitch + (sb - mis), copy_bytes, &bars[c]);
            }
        }
    }

    uint32_t nb = 0, nw = 0;
    if (SCALE == 1 && interior && (P & 7) == 0) {
        // ---- fast path: the band lies inside the slide; 8 pixels (24 bytes) per lane and step, no masks ---------------------
        const uint32_t lo_addr = smem_u32(lo_s);
        const int sh = (mis & 3) * 8, groups8 = P >> 3;
        const uint8_t* base = stage + (mis & ~3);
        for (int c = 0; c < FILTER_CHUNKS; ++c) {
            const int r0 = c * chunk_rows, r1 = min(in_r1 - in_r0, r0 + chunk_rows);
            if (r1 <= r0) break;
            mbar_wait(&bars[c], 0, 80 + c);
            for (int r = r0 + warp; r < r1; r += FILTER_WARPS) {
                const uint32_t* w = reinterpret_cast<const uint32_t*>(base + r * stride) + lane * 6;
                for (int g = lane; g < groups8; g += 32, w += 32 * 6) {
                    const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3], a4 = w[4], a5 = w[5], a6 = w[6];
                    const uint32_t w0 = __funnelshift_r(a0, a1, sh), w1 = __funnelshift_r(a1, a2, sh), w2 = __funnelshift_r(a2, a3, sh);
                    const uint32_t w3 = __funnelshift_r(a3, a4, sh), w4 = __funnelshift_r(a4, a5, sh), w5 = __funnelshift_r(a5, a6, sh);
                    classify_fast(w0, gray_bias, lo_addr, nb, nw);
                    classify_fast(__funnelshift_r(w0, w1, 24), gray_bias, lo_addr, nb, nw);
                    classify_fast(__funnelshift_r(w1, w2, 16), gray_bias, lo_addr, nb, nw);
                    classify_fast(w2 >> 8, gray_bias, lo_addr, nb, nw);
                    classify_fast(w3, gray_bias, lo_addr, nb, nw);
                    classify_fast(__funnelshift_r(w3, w4, 24), gray_bias, lo_addr, nb, nw);
                    classify_fast(__funnelshift_r(w4, w5, 16), gray_bias, lo_addr, nb, nw);
                    classify_fast(w5 >> 8, gray_bias, lo_addr, nb, nw);
                }
            }
        }
    } else
    for (int c = 0; c < FILTER_CHUNKS; ++c) {
        const int r0 = in_r0 + c * chunk_rows, r1 = min(in_r1, r0 + chunk_rows);
        if (r1 <= r0) break;
        const int live = (int)max(0ll, min((long long)r1, p.H - y) - r0);
        if (!interior) {
            // rows below the slide read as zero; a slide without 16-byte alignment is staged with plain loads
            for (int r = r0 + (bulk ? live : 0) + warp; r < r1; r += FILTER_WARPS) {
                uint8_t* d = stage + (r - in_r0) * stride + mis;
                const bool have = inside && y + r < p.H;
                const uint8_t* src = p.slide + (y + r) * p.pitch + sb;
                for (int u = lane; u < row_bytes; u += 32) d[u] = have ? __ldg(src + u) : (uint8_t)0;
            }
            __syncthreads();
        }
        if (bulk && live > 0) mbar_wait(&bars[c], 0, 70 + c);

        const int groups = (P + 3) >> 2;  // 4 consecutive output pixels per thread step
        for (int orow = r0 / SCALE + warp; orow < r1 / SCALE; orow += FILTER_WARPS) {
            const uint8_t* srow = stage + (orow * SCALE - in_r0) * stride + mis;
            if (SCALE == 1) {
                const uint32_t* wrow = reinterpret_cast<const uint32_t*>(srow - (mis & 3));
                const int sh = (mis & 3) * 8;
                for (int gq = lane; gq < groups; gq += 32) {
                    const uint32_t* w = wrow + gq * 3;
                    const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3];
                    const uint32_t w0 = __funnelshift_r(a0, a1, sh), w1 = __funnelshift_r(a1, a2, sh), w2 = __funnelshift_r(a2, a3, sh);
                    uint32_t pix[4] = {w0, __funnelshift_r(w0, w1, 24), __funnelshift_r(w1, w2, 16), w2 >> 8};
                    if (interior) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) classify_word(pix[k], gray_bias, lo_s, nb, nw);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int px = gq * 4 + k;
                            if (px < P) classify_word(px < valid_w ? pix[k] : 0u, gray_bias, lo_s, nb, nw);
                        }
                    }
                }
            } else {
                // cv2.resize at 2:1 = rounded mean of each 2 x 2 block; 4 output pixels = 8 slide pixels of 2 rows
                const uint8_t* s1 = srow + stride;
                for (int gq = lane; gq < groups; gq += 32) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int px = gq * 4 + k;
                        if (px < P) {
                            const bool in0 = px * 2 < valid_w, in1 = px * 2 + 1 < valid_w;
                            uint32_t pix = 0;
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) {
                                const int o = px * 6 + ch;
                                const uint32_t a = in0 ? srow[o] : 0u, b = in1 ? srow[o + 3] : 0u;
                                const uint32_t cc = in0 ? s1[o] : 0u, d = in1 ? s1[o + 3] : 0u;
                                pix |= ((a + b + cc + d + 2u) >> 2) << (8 * ch);
                            }
                            classify_word(pix, gray_bias, lo_s, nb, nw);
                        }
                    }
                }
            }
        }
    }
    nb = __reduce_add_sync(0xffffffffu, nb);
    nw = __reduce_add_sync(0xffffffffu, nw);
    if (lane == 0) {
        atomicAdd(&tot[0], (int)nb);
        atomicAdd(&tot[1], (int)nw);
    }
    __syncthreads();
    if (tid < 2 && tot[tid] != 0) atomicAdd(p.counts + cand * 2 + tid, tot[tid]);
}

// Any other read size (e.g. an 80x slide filtered at 20x, or 40x at 15x): the patch is cv2.resize()d to patch_size first, i.e.
// OpenCV's 8-bit INTER_LINEAR with two taps per axis (tables from ap_build_linear_tables; same arithmetic as preprocess_kernel).
// Only 4 source pixels per output pixel are touched, so staging whole rows would mostly move unused bytes: direct loads, one
// CTA per (candidate, 32 output rows).
__global__ void __launch_bounds__(FILTER_THREADS)
filter_count_direct_kernel(const __grid_constant__ FilterParams p, const int32_t* __restrict__ lin_s, const int16_t* __restrict__ lin_w) {
    __shared__ uint16_t lo_s[256];
    __shared__ int tot[2];
    const int tid = threadIdx.x, lane = tid & 31;
    lo_s[tid & 255] = p.lo[tid & 255];
    if (tid < 2) tot[tid] = 0;
    __syncthreads();
    const long long cand = blockIdx.x / p.bands;
    const int band = blockIdx.x % p.bands;
    const long long x0 = p.rows[cand * 5 + 0], y0 = p.rows[cand * 5 + 1];
    const int P = p.patch, r0 = band * 32, r1 = min(P, r0 + 32);
    int nb = 0, nw = 0;
    for (int i = tid; i < (r1 - r0) * P; i += FILTER_THREADS) {
        const int oy = r0 + i / P, ox = i % P;
        const int a0 = lin_w[2 * ox], a1 = lin_w[2 * ox + 1], b0 = lin_w[2 * oy], b1 = lin_w[2 * oy + 1];
        const long long xa = x0 + lin_s[2 * ox], xb = x0 + lin_s[2 * ox + 1], ya = y0 + lin_s[2 * oy], yb = y0 + lin_s[2 * oy + 1];
        int h0[3] = {0, 0, 0}, h1[3] = {0, 0, 0};
        auto add = [&](long long y, long long x, int w, int (&h)[3]) {
            if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
                const uint8_t* s = p.slide + y * p.pitch + x * 3;
                h[0] += w * __ldg(s); h[1] += w * __ldg(s + 1); h[2] += w * __ldg(s + 2);
            }
        };
        add(ya, xa, a0, h0); add(ya, xb, a1, h0); add(yb, xa, a0, h1); add(yb, xb, a1, h1);
        uint32_t ch[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) ch[c] = static_cast<uint32_t>((((b0 * (h0[c] >> 4)) >> 16) + ((b1 * (h1[c] >> 4)) >> 16) + 2) >> 2);
        const int y = (int)(ch[0] * 9798u + ch[1] * 19235u + ch[2] * 3735u + 16384u);
        nb += (y < p.gray_limit);
        const uint32_t v = max(max(ch[0], ch[1]), ch[2]), mn = min(min(ch[0], ch[1]), ch[2]);
        nw += (mn >= lo_s[v]);
    }
    nb = __reduce_add_sync(0xffffffffu, nb);
    nw = __reduce_add_sync(0xffffffffu, nw);
    if (lane == 0) {
        atomicAdd(&tot[0], nb);
        atomicAdd(&tot[1], nw);
    }
    __syncthreads();
    if (tid < 2 && tot[tid] != 0) atomicAdd(p.counts + cand * 2 + tid, tot[tid]);
}

// keep flags + stable compaction, one CTA of 1024 threads, 4 consecutive candidates per thread and iteration
__global__ void __launch_bounds__(1024)
filter_compact_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ counts, long long n, double n_pixels,
                      double min_fraction, int32_t* __restrict__ out_rows, unsigned long long* __restrict__ out_count) {
    __shared__ int warp_tot[32];
    __shared__ long long base_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (long long i0 = 0; i0 < n; i0 += 4096) {
        const long long i = i0 + tid * 4;
        unsigned keep = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (i + k < n) {
                const double fb = (double)counts[(i + k) * 2] / n_pixels, fw = (double)counts[(i + k) * 2 + 1] / n_pixels;
                keep |= (unsigned)(!(fb >= min_fraction || fw >= min_fraction)) << k;
            }
        }
        const int mine = __popc(keep);
        int incl = mine;  // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const int t = warp_tot[w];
            before += w < wid ? t : 0;
            total += t;
        }
        const long long base = base_s;
        long long o = base + before + incl - mine;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (keep & (1u << k)) {
#pragma unroll
                for (int j = 0; j < 5; ++j) out_rows[o * 5 + j] = rows[(i + k) * 5 + j];
                ++o;
            }
        }
        __syncthreads();
        if (tid == 0) base_s = base + total;
    }
    __syncthreads();
    if (tid == 0) *out_count = (unsigned long long)base_s;
}

}  // namespace

extern "C" int ap_filter_patches(ap_ctx* ctx, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch,
                                 const int32_t* rows_dev, int64_t n, int read_size, int patch_size, int black_thresh,
                                 int white_thresh, double min_fraction, int32_t* out_rows_dev, int32_t* out_rows_host,
                                 int64_t* out_count, int32_t* counts_dev, void* stream) {
    if (!ctx) return AP_EINVAL;
    DeviceGuard guard(ctx);
    AP_REQUIRE(ctx, out_count, "filter_patches: out_count is NULL");
    *out_count = 0;
    if (n == 0) return AP_OK;
    AP_REQUIRE(ctx, slide_dev && rows_dev && n > 0, "filter_patches: NULL pointer / negative n");
    AP_REQUIRE(ctx, out_rows_dev || out_rows_host, "filter_patches: no output buffer");
    AP_REQUIRE(ctx, patch_size > 0 && read_size >= patch_size,
               "filter_patches: read_size %d is smaller than patch_size %d (up-sampling reads do not occur in the reference)", read_size,
               patch_size);
    AP_REQUIRE(ctx, pitch >= 3 * W && W > 0 && H > 0, "filter_patches: bad slide geometry");
    const int scale = read_size == patch_size ? 1 : (read_size == 2 * patch_size ? 2 : 0);  // 0: general cv2.resize, direct kernel
    const int staged_read = scale != 0 ? read_size : patch_size;   // the direct kernel stages nothing
    const int stride = ((15 + staged_read * 3 + 15) & ~15) + 16;  // +16 keeps the number of 16-byte units odd: rows spread over the banks
    int band_rows = std::min(FILTER_MAX_BAND_ROWS, FILTER_MAX_SMEM / stride) & ~7;
    AP_REQUIRE(ctx, band_rows >= 8, "filter_patches: read_size %d too large for the staging buffer", read_size);
    while (band_rows > 8 && band_rows - 8 >= staged_read) band_rows -= 8;  // small patches: do not stage rows that do not exist
    const int bands = scale != 0 ? (read_size + band_rows - 1) / band_rows : (patch_size + 31) / 32;
    AP_REQUIRE(ctx, n * (int64_t)bands < (1ll << 31), "filter_patches: too many candidates");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    FilterParams p{};
    p.slide = slide_dev; p.W = W; p.H = H; p.pitch = pitch; p.rows = rows_dev; p.patch = patch_size;
    p.band_rows = band_rows; p.bands = bands; p.stride = stride;
    // gray < T  <=>  9798 R + 19235 G + 3735 B + 2^14 < T << 15  (the sum is < 2^23, so clamping T to [0, 256] is exact)
    p.gray_limit = (black_thresh < 0 ? 0 : black_thresh > 256 ? 256 : black_thresh) << 15;
    p.vec_ok = ((reinterpret_cast<uintptr_t>(slide_dev) | (uintptr_t)pitch) & 15) == 0;
    for (int v = 0; v < 256; ++v) {
        int lo = 256;  // never white
        if (v >= 200 && white_thresh > 0) {  // utils/image.py:24 value_thresh default; s < T
            // s = (diff * sdiv + 2048) >> 12 < T  <=>  diff * sdiv <= (T << 12) - 2049
            const long long sdiv = llrint((double)(255 << 12) / (double)v);
            long long dmax = (((long long)white_thresh << 12) - 2049) / sdiv;
            if (dmax > v) dmax = v;
            lo = (int)(v - dmax);
        }
        p.lo[v] = (uint16_t)lo;
    }

    const bool own_counts = counts_dev == nullptr, own_rows = out_rows_dev == nullptr;
    const size_t o_rows = ((size_t)n * 8 + 255) & ~(size_t)255;
    std::vector<int32_t> lin_s;
    std::vector<int16_t> lin_w;
    if (scale == 0) {
        int rc = ap_build_linear_tables(ctx, read_size, patch_size, lin_s, lin_w);
        if (rc) return rc;
    }
    const size_t o_tab = o_rows + (own_rows ? (((size_t)n * 20 + 255) & ~(size_t)255) : 0);
    const size_t tab_bytes = scale == 0 ? (size_t)patch_size * 16 : 0;   // 2 int32 taps + 2 int16 weights per index, padded
    const size_t total = o_tab + tab_bytes + 256;
    uint8_t* scratch = nullptr;
    AP_CHECK_CUDA(ctx, cudaMallocAsync((void**)&scratch, total, st));
    auto fail = [&](int rc) { cudaFreeAsync(scratch, st); return rc; };
#define AP_TRY(call)                                                                                          \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) return fail(ap_set_error(ctx, AP_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__))); \
    } while (0)
    int32_t* counts = own_counts ? reinterpret_cast<int32_t*>(scratch) : counts_dev;
    int32_t* rows_out = own_rows ? reinterpret_cast<int32_t*>(scratch + o_rows) : out_rows_dev;
    unsigned long long* count_dev = reinterpret_cast<unsigned long long*>(scratch + total - 256);
    AP_TRY(cudaMemsetAsync(counts, 0, (size_t)n * 8, st));
    p.counts = counts;
    {
        ProfScope prof(ctx, st, AP_K_COORDS);
        const unsigned grid = (unsigned)(n * p.bands);
        const size_t smem = (size_t)band_rows * stride;
        if (scale == 1) {
            AP_TRY(cudaFuncSetAttribute(filter_count_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILTER_MAX_SMEM));
            filter_count_kernel<1><<<grid, FILTER_THREADS, smem, st>>>(p);
        } else if (scale == 2) {
            AP_TRY(cudaFuncSetAttribute(filter_count_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FILTER_MAX_SMEM));
            filter_count_kernel<2><<<grid, FILTER_THREADS, smem, st>>>(p);
        } else {
            int32_t* d_s = reinterpret_cast<int32_t*>(scratch + o_tab);
            int16_t* d_w = reinterpret_cast<int16_t*>(scratch + o_tab + (size_t)patch_size * 8);
            AP_TRY(cudaMemcpyAsync(d_s, lin_s.data(), lin_s.size() * 4, cudaMemcpyHostToDevice, st));
            AP_TRY(cudaMemcpyAsync(d_w, lin_w.data(), lin_w.size() * 2, cudaMemcpyHostToDevice, st));
            filter_count_direct_kernel<<<grid, FILTER_THREADS, 0, st>>>(p, d_s, d_w);
        }
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        AP_TRY(cudaGetLastError());
        filter_compact_kernel<<<1, 1024, 0, st>>>(rows_dev, counts, n, (double)patch_size * (double)patch_size, min_fraction,
                                                  rows_out, count_dev);
        ctx->launches.fetch_add(1, std::memory_order_relaxed);
        AP_TRY(cudaGetLastError());
    }
    unsigned long long count = 0;
    AP_TRY(cudaMemcpyAsync(&count, count_dev, 8, cudaMemcpyDeviceToHost, st));
    AP_TRY(cudaStreamSynchronize(st));
    if (out_rows_host && count > 0) {
        AP_TRY(cudaMemcpyAsync(out_rows_host, rows_out, (size_t)count * 20, cudaMemcpyDeviceToHost, st));
        AP_TRY(cudaStreamSynchronize(st));
    }
#undef AP_TRY
    cudaFreeAsync(scratch, st);
    *out_count = (int64_t)count;
    return AP_OK;
}
