// Slide-sized byte kernels: the synthetic-slide generator and the whole-level thumbnail area reduction (a1).
#include <vector>

#include "ap_internal.cuh"

namespace {

// =====================================================================================================
// Synthetic slide: identical integer function to atlaspatch_b200/synthetic.py (render_region_host).
// =====================================================================================================
constexpr int MAX_BLOBS = 16, MAX_HOLES = 16;
struct SynthParams {
    int n_blobs, n_holes;
    int blobs[MAX_BLOBS][6];
    int holes[MAX_HOLES][3];
};

__device__ __forceinline__ uint32_t mix32(uint32_t u) {
    u ^= u >> 16; u *= 0x7FEB352Du; u ^= u >> 15; u *= 0x846CA68Bu; u ^= u >> 16;
    return u;
}

__device__ __forceinline__ bool tissue_cell(const SynthParams& sp, long long X, long long Y) {
    bool inside = false;
    for (int i = 0; i < sp.n_blobs; ++i) {
        const long long dx = X - sp.blobs[i][0], dy = Y - sp.blobs[i][1];
        const long long a = sp.blobs[i][2], b = sp.blobs[i][3], c = sp.blobs[i][4], s = sp.blobs[i][5];
        const long long u = (dx * c + dy * s) >> 6, v = (dy * c - dx * s) >> 6;
        inside |= (u * u * (b * b) + v * v * (a * a)) <= (a * a) * (b * b);
    }
    for (int i = 0; i < sp.n_holes; ++i) {
        const long long dx = X - sp.holes[i][0], dy = Y - sp.holes[i][1], r = sp.holes[i][2];
        if (dx * dx + dy * dy <= r * r) inside = false;
    }
    return inside;
}

// one thread = 4 consecutive pixels of one row (12 bytes); tissue membership is constant over the 4 pixels
// when x0 % 4 == 0 (cells are 16 px), but is evaluated per pixel to stay valid for any region origin.
__global__ void __launch_bounds__(256)
synth_kernel(uint8_t* __restrict__ out, long long pitch, long long W, long long H, uint32_t seed, SynthParams sp,
             long long x0, long long y0, long long w, long long h) {
    const long long gx = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
    const long long gy = blockIdx.y;
    if (gx >= w || gy >= h) return;
    const long long y = y0 + gy;
    uint8_t* row = out + gy * pitch;
    for (int i = 0; i < 4 && gx + i < w; ++i) {
        const long long x = x0 + gx + i;
        uint8_t r = 0, g = 0, b = 0;
        if (x >= 0 && x < W && y >= 0 && y < H) {
            const uint32_t xu = static_cast<uint32_t>(x), yu = static_cast<uint32_t>(y);
            const uint32_t hpx = mix32(xu * 0x9E3779B1u + yu * 0x85EBCA77u + seed * 0xC2B2AE3Du);
            const uint32_t gc = mix32((xu >> 3) * 0x9E3779B1u + (yu >> 3) * 0x85EBCA77u + (seed + 1u) * 0xC2B2AE3Du);
            if (tissue_cell(sp, x >> 4, y >> 4)) {
                r = 150 + (gc & 63) + (hpx & 15);
                g = 60 + ((gc >> 8) & 63) + ((hpx >> 4) & 15);
                b = 130 + ((gc >> 16) & 63) + ((hpx >> 8) & 15);
            } else {
                r = 232 + (hpx & 7);
                g = 232 + ((hpx >> 4) & 7);
                b = 232 + ((hpx >> 8) & 7);
            }
        }
        row[(gx + i) * 3 + 0] = r;
        row[(gx + i) * 3 + 1] = g;
        row[(gx + i) * 3 + 2] = b;
    }
}

// =====================================================================================================
// a1: thumbnail = INTER_AREA reduction by an integer factor F of the whole level-0 image.
//   reference: IWSI.get_thumbnail_at_power reads the whole level and cv2.resize(INTER_AREA)s it
//   (core/wsi/iwsi.py:293-321).  OpenCV's integer-factor uint8 path = saturate(rint(sum * (1/F^2))) in fp32.
// HBM-bound: W*H*3 bytes are read exactly once.  One thread per output pixel; its F rows x 3F bytes are
// read as 16-byte vectors (consecutive lanes read consecutive 3F-byte spans, so a warp covers one
// contiguous 96F-byte run per row and every fetched sector is fully used through L1).  Per-channel sums
// use dp4a with byte masks (0.75 instructions per input byte).
// Requires (3F) % 16 == 0, i.e. F % 16 == 0, and pitch % 16 == 0 for the vector path; other integer factors
// take the scalar path.
// =====================================================================================================
// Byte j of a 32-bit word whose first byte has channel `ph` carries channel (ph + j) % 3; the mask has a 1 in every
// byte that belongs to channel c (little-endian: byte 0 = lowest address).
__host__ __device__ constexpr unsigned chan_mask(int ph, int c) {
    return ((c - ph + 3) % 3) == 0 ? 0x01000001u : (((c - ph + 3) % 3) == 1 ? 0x00000100u : 0x00010000u);
}

template <int F>
__global__ void __launch_bounds__(128)
thumb_vec_kernel(const uint8_t* __restrict__ src, long long pitch, int out_w, int out_h, uint8_t* __restrict__ dst) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (ox >= out_w) return;
    constexpr int VPR = 3 * F / 16;  // uint4 per row per output pixel
    const uint8_t* base = src + static_cast<long long>(oy) * F * pitch + static_cast<long long>(ox) * 3 * F;
    unsigned s0 = 0, s1 = 0, s2 = 0;
    // byte j of the span has channel j % 3; a 16-byte vector v starts at phase (16 v) % 3 = v % 3; inside a vector,
    // 32-bit word k starts at phase (v + 4k) % 3 = (v + k) % 3.
#pragma unroll 4
    for (int r = 0; r < F; ++r) {
        const uint4* rowp = reinterpret_cast<const uint4*>(base + static_cast<long long>(r) * pitch);
        uint4 v[VPR];
#pragma unroll
        for (int i = 0; i < VPR; ++i) v[i] = __ldg(rowp + i);
#pragma unroll
        for (int i = 0; i < VPR; ++i) {
            const unsigned wds[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int ph = (i + k) % 3;  // compile-time after unrolling
                s0 = __dp4a(wds[k], chan_mask(ph, 0), s0);
                s1 = __dp4a(wds[k], chan_mask(ph, 1), s1);
                s2 = __dp4a(wds[k], chan_mask(ph, 2), s2);
            }
        }
    }
    const float inv = __frcp_rn(static_cast<float>(F * F));   // cv2: scale = 1.f / area, correctly rounded (--use_fast_math must not touch it)
    uint8_t* o = dst + (static_cast<long long>(oy) * out_w + ox) * 3;
    o[0] = static_cast<uint8_t>(min(255, __float2int_rn(static_cast<float>(s0) * inv)));
    o[1] = static_cast<uint8_t>(min(255, __float2int_rn(static_cast<float>(s1) * inv)));
    o[2] = static_cast<uint8_t>(min(255, __float2int_rn(static_cast<float>(s2) * inv)));
}

// any integer factors (fx, fy): cv2's ResizeAreaFast sums the fx x fy block and multiplies by 1.f / (fx * fy)
__global__ void __launch_bounds__(128)
thumb_scalar_kernel(const uint8_t* __restrict__ src, long long pitch, int F, int FY, int out_w, int out_h, uint8_t* __restrict__ dst) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (ox >= out_w) return;
    const uint8_t* base = src + static_cast<long long>(oy) * FY * pitch + static_cast<long long>(ox) * 3 * F;
    unsigned s[3] = {0, 0, 0};
    for (int r = 0; r < FY; ++r) {
        const uint8_t* p = base + static_cast<long long>(r) * pitch;
        for (int i = 0; i < F; ++i) {
            s[0] += __ldg(p + 3 * i);
            s[1] += __ldg(p + 3 * i + 1);
            s[2] += __ldg(p + 3 * i + 2);
        }
    }
    const float inv = __frcp_rn(static_cast<float>(F * FY));
    uint8_t* o = dst + (static_cast<long long>(oy) * out_w + ox) * 3;
    for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>(min(255, __float2int_rn(static_cast<float>(s[c]) * inv)));
}

// General INTER_AREA down-scale (scale factors not both integral): OpenCV's computeResizeAreaTab + ResizeArea_<uchar, float>
// (resize.cpp), operation by operation: per source row a horizontal pass buf = sum_k S[si_k] * alpha_k (float32, separate multiply
// and add, table order), then sum += beta * buf over the rows of the cell, saturate_cast<uchar>(sum) = round half to even.
// oracle/thumbnail.py: area_resize_general restates the same arithmetic and is pinned against cv2 itself.
__global__ void __launch_bounds__(128)
thumb_area_general_kernel(const uint8_t* __restrict__ src, long long pitch, int out_w, int out_h, const int* __restrict__ xfirst,
                          const int* __restrict__ xsi, const float* __restrict__ xa, const int* __restrict__ yfirst,
                          const int* __restrict__ ysi, const float* __restrict__ ya, uint8_t* __restrict__ dst) {
    const int ox = blockIdx.x * blockDim.x + threadIdx.x;
    const int oy = blockIdx.y;
    if (ox >= out_w) return;
    const int k0 = xfirst[ox], k1 = xfirst[ox + 1];
    float sum[3] = {0.f, 0.f, 0.f};
    for (int j = yfirst[oy]; j < yfirst[oy + 1]; ++j) {
        const uint8_t* row = src + static_cast<long long>(ysi[j]) * pitch;
        const float beta = ya[j];
        float buf[3] = {0.f, 0.f, 0.f};
        for (int k = k0; k < k1; ++k) {
            const uint8_t* px = row + 3 * static_cast<long long>(xsi[k]);
            const float alpha = xa[k];
#pragma unroll
            for (int c = 0; c < 3; ++c) buf[c] = __fadd_rn(buf[c], __fmul_rn(static_cast<float>(__ldg(px + c)), alpha));
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) sum[c] = __fadd_rn(sum[c], __fmul_rn(beta, buf[c]));   // 0 + x == x for the first row
    }
    uint8_t* o = dst + (static_cast<long long>(oy) * out_w + ox) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = static_cast<uint8_t>(max(0, min(255, __float2int_rn(sum[c]))));
}

// computeResizeAreaTab (double arithmetic, float weights) for one axis: first[dsize + 1], si[], alpha[]
void build_area_tab(int ssize, int dsize, std::vector<int>& first, std::vector<int>& si, std::vector<float>& alpha) {
    const double scale = static_cast<double>(ssize) / dsize;
    first.assign(1, 0);
    si.clear();
    alpha.clear();
    for (int dx = 0; dx < dsize; ++dx) {
        const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
        const double cell = std::min(scale, ssize - fsx1);
        int sx1 = static_cast<int>(std::ceil(fsx1)), sx2 = static_cast<int>(std::floor(fsx2));
        sx2 = std::min(sx2, ssize - 1);
        sx1 = std::min(sx1, sx2);
        if (sx1 - fsx1 > 1e-3) { si.push_back(sx1 - 1); alpha.push_back(static_cast<float>((sx1 - fsx1) / cell)); }
        for (int sx = sx1; sx < sx2; ++sx) { si.push_back(sx); alpha.push_back(static_cast<float>(1.0 / cell)); }
        if (fsx2 - sx2 > 1e-3) { si.push_back(sx2); alpha.push_back(static_cast<float>(std::min(std::min(fsx2 - sx2, 1.), cell) / cell)); }
        first.push_back(static_cast<int>(si.size()));
    }
}

}  // namespace

extern "C" int ap_synth_render(ap_ctx* ctx, uint8_t* out_dev, int64_t pitch, int64_t W, int64_t H, uint32_t seed,
                               const int32_t* blobs_host, int n_blobs, const int32_t* holes_host, int n_holes, int64_t x0,
                               int64_t y0, int64_t w, int64_t h, void* stream) {
    if (!ctx) return AP_EINVAL;
    DeviceGuard guard(ctx);
    AP_REQUIRE(ctx, out_dev && w > 0 && h > 0 && pitch >= 3 * w, "synth_render: bad output geometry");
    AP_REQUIRE(ctx, n_blobs >= 0 && n_blobs <= MAX_BLOBS && n_holes >= 0 && n_holes <= MAX_HOLES, "synth_render: too many blobs/holes");
    AP_REQUIRE(ctx, h <= 0x7fffffffLL && (w + 1023) / 1024 <= 0x7fffffffLL, "synth_render: region too large");
    SynthParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.n_blobs = n_blobs; sp.n_holes = n_holes;
    if (n_blobs) memcpy(sp.blobs, blobs_host, sizeof(int) * 6 * n_blobs);
    if (n_holes) memcpy(sp.holes, holes_host, sizeof(int) * 3 * n_holes);
    // gridDim.y is limited to 65535: render in row bands
    for (int64_t yb = 0; yb < h; yb += 65535) {
        const int64_t hb = h - yb < 65535 ? h - yb : 65535;
        dim3 grid(static_cast<unsigned>((w + 1023) / 1024), static_cast<unsigned>(hb));
        synth_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(out_dev + yb * pitch, pitch, W, H, seed, sp, x0, y0 + yb, w, hb);
        AP_CHECK_LAUNCH(ctx, "synth_kernel");
    }
    return AP_OK;
}

extern "C" int ap_thumbnail_area(ap_ctx* ctx, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch, int factor,
                                 uint8_t* out_dev, void* stream) {
    if (!ctx) return AP_EINVAL;
    DeviceGuard guard(ctx);
    AP_REQUIRE(ctx, slide_dev && out_dev, "thumbnail_area: NULL pointer");
    AP_REQUIRE(ctx, factor >= 1 && W > 0 && H > 0 && W % factor == 0 && H % factor == 0,
               "thumbnail_area: integer factor %d must divide W=%lld and H=%lld (non-integer INTER_AREA is not implemented)", factor,
               (long long)W, (long long)H);
    AP_REQUIRE(ctx, pitch >= 3 * W, "thumbnail_area: pitch %lld < 3*W", (long long)pitch);
    const int out_w = static_cast<int>(W / factor), out_h = static_cast<int>(H / factor);
    AP_REQUIRE(ctx, out_h <= 65535, "thumbnail_area: output height %d > 65535", out_h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid((out_w + 127) / 128, out_h);
    ProfScope prof(ctx, st, AP_K_THUMBNAIL);
    const bool vec_ok = (pitch % 16 == 0) && ((reinterpret_cast<uintptr_t>(slide_dev) & 15) == 0);
    if (vec_ok && factor == 16) thumb_vec_kernel<16><<<grid, 128, 0, st>>>(slide_dev, pitch, out_w, out_h, out_dev);
    else if (vec_ok && factor == 32) thumb_vec_kernel<32><<<grid, 128, 0, st>>>(slide_dev, pitch, out_w, out_h, out_dev);
    else if (vec_ok && factor == 64) thumb_vec_kernel<64><<<grid, 128, 0, st>>>(slide_dev, pitch, out_w, out_h, out_dev);
    else thumb_scalar_kernel<<<grid, 128, 0, st>>>(slide_dev, pitch, factor, factor, out_w, out_h, out_dev);
    AP_CHECK_LAUNCH(ctx, "thumbnail kernel");
    return AP_OK;
}

// cv2.resize(level, (out_w, out_h), INTER_AREA) for any down-scale: what IWSI.get_thumbnail_at_power computes when the level
// size is not a multiple of the factor (out = round(W / ds) x round(H / ds), core/wsi/iwsi.py:302-321).
extern "C" int ap_thumbnail_resize(ap_ctx* ctx, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch, int out_w, int out_h,
                                   uint8_t* out_dev, void* stream) {
    if (!ctx) return AP_EINVAL;
    DeviceGuard guard(ctx);
    AP_REQUIRE(ctx, slide_dev && out_dev, "thumbnail_resize: NULL pointer");
    AP_REQUIRE(ctx, W > 0 && H > 0 && W <= 0x7fffffffLL && H <= 0x7fffffffLL && out_w >= 1 && out_h >= 1 && out_w <= W && out_h <= H,
               "thumbnail_resize: only down-scaling is built (%lldx%lld -> %dx%d)", (long long)W, (long long)H, out_w, out_h);
    AP_REQUIRE(ctx, pitch >= 3 * W, "thumbnail_resize: pitch %lld < 3*W", (long long)pitch);
    AP_REQUIRE(ctx, out_h <= 65535, "thumbnail_resize: output height %d > 65535", out_h);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // OpenCV's is_area_fast: both scale factors integral (resize.cpp)
    const double sx = static_cast<double>(W) / out_w, sy = static_cast<double>(H) / out_h;
    const int isx = static_cast<int>(sx), isy = static_cast<int>(sy);
    if (std::abs(sx - isx) < 2.220446049250313e-16 && std::abs(sy - isy) < 2.220446049250313e-16) {
        if (isx == isy) return ap_thumbnail_area(ctx, slide_dev, W, H, pitch, isx, out_dev, stream);
        ProfScope prof(ctx, st, AP_K_THUMBNAIL);
        thumb_scalar_kernel<<<dim3((out_w + 127) / 128, out_h), 128, 0, st>>>(slide_dev, pitch, isx, isy, out_w, out_h, out_dev);
        AP_CHECK_LAUNCH(ctx, "thumb_scalar_kernel");
        return AP_OK;
    }
    std::vector<int> xf, xs, yf, ys;
    std::vector<float> xa, ya;
    build_area_tab(static_cast<int>(W), out_w, xf, xs, xa);
    build_area_tab(static_cast<int>(H), out_h, yf, ys, ya);
    const size_t nb[6] = {xf.size() * 4, xs.size() * 4, xa.size() * 4, yf.size() * 4, ys.size() * 4, ya.size() * 4};
    const void* hp[6] = {xf.data(), xs.data(), xa.data(), yf.data(), ys.data(), ya.data()};
    size_t off[7] = {0};
    for (int i = 0; i < 6; ++i) off[i + 1] = off[i] + (nb[i] + 255) / 256 * 256;
    uint8_t* tab = nullptr;
    AP_CHECK_CUDA(ctx, cudaMallocAsync(reinterpret_cast<void**>(&tab), off[6], st));
    for (int i = 0; i < 6; ++i) AP_CHECK_CUDA(ctx, cudaMemcpyAsync(tab + off[i], hp[i], nb[i], cudaMemcpyHostToDevice, st));
    {
        ProfScope prof(ctx, st, AP_K_THUMBNAIL);
        thumb_area_general_kernel<<<dim3((out_w + 127) / 128, out_h), 128, 0, st>>>(
            slide_dev, pitch, out_w, out_h, reinterpret_cast<const int*>(tab + off[0]), reinterpret_cast<const int*>(tab + off[1]),
            reinterpret_cast<const float*>(tab + off[2]), reinterpret_cast<const int*>(tab + off[3]),
            reinterpret_cast<const int*>(tab + off[4]), reinterpret_cast<const float*>(tab + off[5]), out_dev);
        AP_CHECK_LAUNCH(ctx, "thumb_area_general_kernel");
    }
    AP_CHECK_CUDA(ctx, cudaStreamSynchronize(st));   // the pageable host tables must outlive the copies
    AP_CHECK_CUDA(ctx, cudaFreeAsync(tab, st));
    return AP_OK;
}
