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
#include "ap_internal.cuh"

namespace {

constexpr int FILTER_THREADS = 256;
constexpr int FILTER_BAND = 32;          // output rows per CTA
constexpr int FILTER_SMEM_BYTES = 24576; // staging buffer

struct FilterParams {
    const uint8_t* slide;
    long long W, H, pitch;
    const int32_t* rows;  // n x 5 (x, y, read_w, read_h, level)
    int patch;            // output patch size P
    int bands;            // ceil(P / FILTER_BAND)
    int gray_limit;       // black <=> 9798 R + 19235 G + 3735 B + 16384 < gray_limit  (= black_thresh << 15)
    int vec_ok;           // slide base and pitch are 16-byte aligned
    int32_t* counts;      // n x 2, zeroed
    uint16_t lo[256];     // white <=> min(R,G,B) >= lo[max(R,G,B)]   (256 = never)
};

__device__ __forceinline__ void classify(uint32_t r, uint32_t g, uint32_t b, int gray_limit, const uint16_t* lo, int& nb, int& nw) {
    const int y = (int)(r * 9798u + g * 19235u + b * 3735u + 16384u);
    nb += (y < gray_limit);
    const uint32_t v = max(max(r, g), b), mn = min(min(r, g), b);
    nw += (mn >= lo[v]);
}

template <int SCALE>
__global__ void __launch_bounds__(FILTER_THREADS)
filter_count_kernel(const __grid_constant__ FilterParams p) {
    extern __shared__ __align__(16) uint8_t stage[];
    __shared__ uint16_t lo_s[256];
    __shared__ int tot[2];
    const int tid = threadIdx.x;
    lo_s[tid & 255] = p.lo[tid & 255];
    if (tid < 2) tot[tid] = 0;

    const long long cand = blockIdx.x / p.bands;
    const int band = blockIdx.x % p.bands;
    const int32_t* row = p.rows + cand * 5;
    const long long x = row[0], y = row[1];
    const int P = p.patch, read = P * SCALE;
    const int out_r0 = band * FILTER_BAND, out_r1 = min(P, out_r0 + FILTER_BAND);

    // bytes of one slide row this patch touches, clipped to the slide; staged from the 16-byte boundary below it
    const long long sb = x * 3;
    const long long eb = min((x + read), p.W) * 3;
    const int mis = p.vec_ok ? (int)(sb & 15) : 0;
    const int row_bytes = (int)(eb - sb);                 // > 0: candidates start inside the slide
    const int stride = ((mis + read * 3 + 15) & ~15) + 16;  // +16: odd number of 16-byte units -> rows spread over the banks
    const bool inside = x >= 0 && y >= 0 && x < p.W;  // extraction never emits anything else; such a patch reads as all zero
    const int valid_w = inside ? (int)(min((long long)read, p.W - x)) : 0;
    const int rows_per_chunk = max(SCALE, (FILTER_SMEM_BYTES / stride) / SCALE * SCALE);

    int nb = 0, nw = 0;
    for (int ir0 = out_r0 * SCALE; ir0 < out_r1 * SCALE; ir0 += rows_per_chunk) {
        const int nrows = min(rows_per_chunk, out_r1 * SCALE - ir0);
        __syncthreads();
        if (!inside) {
        } else if (p.vec_ok) {
            const int units = (mis + row_bytes + 15) >> 4;
            for (int i = tid; i < nrows * units; i += FILTER_THREADS) {
                const int r = i / units, u = i - r * units;
                const long long gy = y + ir0 + r;
                uint4 val = make_uint4(0, 0, 0, 0);
                if (gy < p.H) val = __ldg(reinterpret_cast<const uint4*>(p.slide + gy * p.pitch + (sb - mis)) + u);
                *reinterpret_cast<uint4*>(stage + r * stride + u * 16) = val;
            }
        } else {
            for (int i = tid; i < nrows * row_bytes; i += FILTER_THREADS) {
                const int r = i / row_bytes, u = i - r * row_bytes;
                const long long gy = y + ir0 + r;
                stage[r * stride + u] = gy < p.H ? __ldg(p.slide + gy * p.pitch + sb + u) : (uint8_t)0;
            }
        }
        __syncthreads();

        // 4 consecutive output pixels per thread step
        const int groups = (P + 3) >> 2;
        const int out_rows = nrows / SCALE;
        for (int i = tid; i < out_rows * groups; i += FILTER_THREADS) {
            const int r = i / groups, gq = i - r * groups;
            const int px0 = gq * 4;
            if (SCALE == 1) {
                const uint8_t* s = stage + r * stride + mis + px0 * 3;
                const uint32_t* w = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(s) & ~(uintptr_t)3);
                const int sh = (int)(reinterpret_cast<uintptr_t>(s) & 3) * 8;
                const uint32_t a0 = w[0], a1 = w[1], a2 = w[2], a3 = w[3];
                const uint32_t w0 = __funnelshift_r(a0, a1, sh), w1 = __funnelshift_r(a1, a2, sh), w2 = __funnelshift_r(a2, a3, sh);
                const uint32_t rr[4] = {w0 & 255u, w0 >> 24, (w1 >> 16) & 255u, (w2 >> 8) & 255u};
                const uint32_t gg[4] = {(w0 >> 8) & 255u, w1 & 255u, w1 >> 24, (w2 >> 16) & 255u};
                const uint32_t bb[4] = {(w0 >> 16) & 255u, (w1 >> 8) & 255u, w2 & 255u, w2 >> 24};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int px = px0 + k;
                    if (px < P) {
                        const bool in = px < valid_w;
                        classify(in ? rr[k] : 0u, in ? gg[k] : 0u, in ? bb[k] : 0u, p.gray_limit, lo_s, nb, nw);
                    }
                }
            } else {
                // 2:1 bilinear = rounded mean of the 2 x 2 block; 4 output pixels need 8 input pixels (24 bytes) of 2 rows
                const uint8_t* s0 = stage + (r * 2) * stride + mis + px0 * 6;
                const uint8_t* s1 = s0 + stride;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int px = px0 + k;
                    if (px < P) {
                        uint32_t c[3];
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            const int cx0 = px * 2, cx1 = px * 2 + 1;
                            const uint32_t a = cx0 < valid_w ? s0[k * 6 + ch] : 0u, b = cx1 < valid_w ? s0[k * 6 + 3 + ch] : 0u;
                            const uint32_t cc = cx0 < valid_w ? s1[k * 6 + ch] : 0u, d = cx1 < valid_w ? s1[k * 6 + 3 + ch] : 0u;
                            c[ch] = (a + b + cc + d + 2u) >> 2;
                        }
                        classify(c[0], c[1], c[2], p.gray_limit, lo_s, nb, nw);
                    }
                }
            }
        }
    }
    nb = __reduce_add_sync(0xffffffffu, nb);
    nw = __reduce_add_sync(0xffffffffu, nw);
    if ((tid & 31) == 0) {
        atomicAdd(&tot[0], nb);
        atomicAdd(&tot[1], nw);
    }
    __syncthreads();
    if (tid < 2 && tot[tid] != 0) atomicAdd(p.counts + cand * 2 + tid, tot[tid]);
}

// keep flags + stable compaction, one CTA of 1024 threads
__global__ void __launch_bounds__(1024)
filter_compact_kernel(const int32_t* __restrict__ rows, const int32_t* __restrict__ counts, long long n, double n_pixels,
                      double min_fraction, int32_t* __restrict__ out_rows, unsigned long long* __restrict__ out_count) {
    __shared__ int warp_tot[32];
    __shared__ long long base_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
    for (long long i0 = 0; i0 < n; i0 += 1024) {
        const long long i = i0 + tid;
        bool keep = false;
        if (i < n) {
            const double fb = (double)counts[i * 2] / n_pixels, fw = (double)counts[i * 2 + 1] / n_pixels;
            keep = !(fb >= min_fraction || fw >= min_fraction);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[wid] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 32; ++w) {
            const int t = warp_tot[w];
            before += w < wid ? t : 0;
            total += t;
        }
        const long long base = base_s;
        if (keep) {
            const long long o = base + before + __popc(bal & ((1u << lane) - 1u));
#pragma unroll
            for (int k = 0; k < 5; ++k) out_rows[o * 5 + k] = rows[i * 5 + k];
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
    AP_REQUIRE(ctx, out_count, "filter_patches: out_count is NULL");
    *out_count = 0;
    if (n == 0) return AP_OK;
    AP_REQUIRE(ctx, slide_dev && rows_dev && n > 0, "filter_patches: NULL pointer / negative n");
    AP_REQUIRE(ctx, out_rows_dev || out_rows_host, "filter_patches: no output buffer");
    AP_REQUIRE(ctx, patch_size > 0 && (read_size == patch_size || read_size == 2 * patch_size),
               "filter_patches: read_size %d must be patch_size %d or exactly twice it (general cv2.resize is not implemented)",
               read_size, patch_size);
    AP_REQUIRE(ctx, pitch >= 3 * W && W > 0 && H > 0, "filter_patches: bad slide geometry");
    AP_REQUIRE(ctx, read_size * 3 + 48 <= FILTER_SMEM_BYTES / 2, "filter_patches: read_size %d too large for the staging buffer", read_size);
    AP_REQUIRE(ctx, n * (int64_t)((patch_size + FILTER_BAND - 1) / FILTER_BAND) < (1ll << 31), "filter_patches: too many candidates");
    cudaStream_t st = static_cast<cudaStream_t>(stream);

    FilterParams p{};
    p.slide = slide_dev; p.W = W; p.H = H; p.pitch = pitch; p.rows = rows_dev; p.patch = patch_size;
    p.bands = (patch_size + FILTER_BAND - 1) / FILTER_BAND;
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
    const size_t total = o_rows + (own_rows ? (size_t)n * 20 : 0) + 256;
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
        if (read_size == patch_size) filter_count_kernel<1><<<grid, FILTER_THREADS, FILTER_SMEM_BYTES, st>>>(p);
        else filter_count_kernel<2><<<grid, FILTER_THREADS, FILTER_SMEM_BYTES, st>>>(p);
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
