// a11 + a12 for the DINOv2 encoders: patch read + transformers' BitImageProcessorFast, fused with the im2col of the
// patch-embedding convolution.
//
//   reference: atlas_patch/models/patch/dinov2.py:20-25,49 -> AutoImageProcessor(use_fast=True) of facebook/dinov2-*:
//   resize(shortest_edge 256, bicubic) -> center_crop(224) -> rescale(1/255) -> normalize(ImageNet).  The fast processor
//   resizes the uint8 tensor with torch.nn.functional.interpolate(mode="bicubic", antialias=True), i.e. ATen's separable
//   uint8 kernel (Pillow's ImagingResample): horizontal pass to uint8, then vertical pass to uint8, int16 weights with a
//   common precision p, out = clamp((sum w_j src_j + 2^(p-1)) >> p).  The tap tables are computed on the host at
//   ap_encoder_finalize (encoder.cu: build_resize_tables) with the same float64 arithmetic; oracle/resize_aa.py restates it
//   and is pinned bit-exactly against torch.
//   Rescale + normalise are folded into the conv weights, so the kernel's output is fp16((pixel - centre_c) / 256) like the
//   crop-only preprocess kernel.
//
// One CTA per (patch b, token row tr): the P cropped output rows of that token row need source rows [ymin, ymax) of the
// patch; they are staged in shared memory (zeros outside the slide, like IWSI.extract), resampled horizontally for the `image`
// cropped columns only, then vertically, and written as `g` im2col rows (k = c P^2 + ky P + kx).
#include <algorithm>
#include <cmath>
#include <vector>

#include "ap_internal.cuh"
#include "ptx.cuh"

namespace {

__global__ void __launch_bounds__(256)
preprocess_resize_kernel(const uint8_t* __restrict__ slide, int64_t W, int64_t H, int64_t pitch, const int32_t* __restrict__ coords,
                         int input_patch, int image, int P, const int32_t* __restrict__ tap_min, const int32_t* __restrict__ tap_cnt,
                         const int32_t* __restrict__ tap_w, int max_taps, int precision, __half* __restrict__ out,
                         int64_t out_row_stride, int3 centre) {
    extern __shared__ __align__(16) uint8_t smem[];
    ptx::pdl_wait();
    ptx::pdl_launch_dependents();
    const int g = image / P;
    const int b = blockIdx.x / g, tr = blockIdx.x % g;
    const int64_t x0 = coords[b * 5 + 0], y0 = coords[b * 5 + 1];
    const int oy0 = tr * P;
    const int ymin = tap_min[oy0];
    const int ymax = tap_min[oy0 + P - 1] + tap_cnt[oy0 + P - 1];
    const int nrows = ymax - ymin;
    const int src_stride = input_patch * 3, h_stride = image * 3;
    uint8_t* src = smem;                                           // [nrows][input_patch * 3]
    uint8_t* hbuf = smem + ((nrows * src_stride + 15) & ~15);      // [nrows][image * 3]
    const int tid = threadIdx.x;

    // ---- stage the source rows -------------------------------------------------------------------------------------
    for (int i = tid; i < nrows * src_stride; i += 256) {
        const int r = i / src_stride, u = i - r * src_stride;
        const int64_t gy = y0 + ymin + r, gxb = x0 * 3 + u;
        uint8_t v = 0;
        if (gy >= 0 && gy < H && gxb >= 0 && gxb < W * 3) v = __ldg(slide + gy * pitch + gxb);
        src[i] = v;
    }
    __syncthreads();

    // ---- horizontal pass (cropped columns only), uint8 result -------------------------------------------------------
    const int round_add = 1 << (precision - 1);
    for (int i = tid; i < nrows * h_stride; i += 256) {
        const int r = i / h_stride, rem = i - r * h_stride;
        const int ox = rem / 3, c = rem - ox * 3;
        const int xm = tap_min[ox], n = tap_cnt[ox];
        const int32_t* w = tap_w + ox * max_taps;
        const uint8_t* s = src + r * src_stride + xm * 3 + c;
        int acc = round_add;
        for (int k = 0; k < n; ++k) acc += w[k] * static_cast<int>(s[k * 3]);
        hbuf[i] = static_cast<uint8_t>(min(max(acc >> precision, 0), 255));
    }
    __syncthreads();

    // ---- vertical pass + im2col write: idx = ((tc * 3 + c) * P + ky) * P + kx ------------------------------------------
    const int per_token = 3 * P * P;
    for (int i = tid; i < g * per_token; i += 256) {
        const int tc = i / per_token, k = i - tc * per_token;
        const int c = k / (P * P), rem = k - c * P * P;
        const int ky = rem / P, kx = rem - ky * P;
        const int oy = oy0 + ky, ox = tc * P + kx;
        const int ym = tap_min[oy], n = tap_cnt[oy];
        const int32_t* w = tap_w + oy * max_taps;
        const uint8_t* s = hbuf + (ym - ymin) * h_stride + ox * 3 + c;
        int acc = round_add;
        for (int t = 0; t < n; ++t) acc += w[t] * static_cast<int>(s[t * h_stride]);
        const int pix = min(max(acc >> precision, 0), 255);
        const int cen = c == 0 ? centre.x : (c == 1 ? centre.y : centre.z);
        out[(static_cast<int64_t>(b) * g * g + tr * g + tc) * out_row_stride + k] = __float2half_rn(static_cast<float>(pix - cen) * (1.0f / 256.0f));
    }
}

double cubic_aa(double x) {  // Keys cubic, a = -0.5 (ATen HelperInterpCubic::aa_filter)
    const double a = -0.5;
    x = x < 0 ? -x : x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0;
    if (x < 2.0) return (((x - 5.0) * x + 8.0) * x - 4.0) * a;
    return 0.0;
}

}  // namespace

// Tap tables of the n_in -> n_out antialias resize for the `image` output indices that survive the centre crop.
//   kind 0: ATen's uint8 bicubic (a = -0.5, support 2): _compute_index_ranges_int16_weights, the precision is chosen over ALL
//           n_out outputs so that the largest weight fits int16 (transformers' BitImageProcessorFast, DINOv2);
//   kind 1: Pillow's own BILINEAR (triangle, support 1): precompute_coeffs + normalize_coeffs_8bpc, fixed 22-bit int32
//           coefficients (torchvision's ImageClassification preset on the PIL image the reference hands it; crop offset
//           int(round((n_out - image) / 2.0)) as torchvision's center_crop computes it);
//   kind 2: ATen's uint8 bilinear with antialias (triangle, support 1; int16 weights with the precision rule of kind 0): what
//           transformers' ViTImageProcessorFast runs for `resample = 2` checkpoints (atlas_patch/models/patch/phikon.py:15-21,46);
//   kind 3: Pillow's BICUBIC (Keys cubic a = -0.5, support 2) with Pillow's fixed 22-bit coefficients and torchvision's crop offset:
//           torchvision `Resize(256, BICUBIC) -> CenterCrop(224)` on the PIL patch (atlas_patch/models/patch/gigapath.py:17-26).
int ap_build_resize_tables(ap_ctx* ctx, int n_in, int n_out, int image, int kind, std::vector<int32_t>& tap_min, std::vector<int32_t>& tap_cnt,
                           std::vector<int32_t>& tap_w, int* max_taps, int* precision) {
    AP_REQUIRE(ctx, n_in > 0 && n_out >= image && image > 0, "resize tables: bad sizes %d -> %d crop %d", n_in, n_out, image);
    AP_REQUIRE(ctx, kind >= 0 && kind <= 3, "resize tables: unknown filter kind %d", kind);
    const bool cubic = kind == 0 || kind == 3, pillow = kind == 1 || kind == 3;
    const double scale = static_cast<double>(n_in) / n_out;
    const double fsup = cubic ? 2.0 : 1.0;
    const double support = scale >= 1.0 ? fsup * scale : fsup;
    const double invscale = scale >= 1.0 ? 1.0 / scale : 1.0;
    std::vector<std::vector<double>> ws(n_out);
    std::vector<int> mins(n_out);
    double wt_max = 0.0;
    int taps = 0;
    for (int i = 0; i < n_out; ++i) {
        const double center = scale * (i + 0.5);
        int xmin = static_cast<int>(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xend = static_cast<int>(center + support + 0.5);
        if (xend > n_in) xend = n_in;
        const int xsize = xend - xmin;
        double total = 0.0;
        ws[i].resize(xsize);
        for (int j = 0; j < xsize; ++j) {
            const double t = (j + xmin - center + 0.5) * invscale;
            ws[i][j] = cubic ? cubic_aa(t) : std::max(0.0, 1.0 - std::fabs(t));
            total += ws[i][j];
        }
        for (int j = 0; j < xsize; ++j) {
            if (total != 0.0) ws[i][j] /= total;
            if (ws[i][j] > wt_max) wt_max = ws[i][j];
        }
        mins[i] = xmin;
        if (xsize > taps) taps = xsize;
    }
    int prec = 22;                         // Pillow: PRECISION_BITS = 32 - 8 - 2
    if (!pillow)
        for (prec = 0; prec < 22; ++prec) {
            const int next = static_cast<int>(0.5 + wt_max * (1 << (prec + 1)));
            if (next >= (1 << 15)) break;
        }
    // transformers center_crop: top = (h - crop) // 2; torchvision center_crop: int(round((h - crop) / 2.0)) (half to even)
    const int off = !pillow ? (n_out - image) / 2 : static_cast<int>(std::nearbyint((n_out - image) / 2.0));
    tap_min.assign(image, 0);
    tap_cnt.assign(image, 0);
    tap_w.assign(static_cast<size_t>(image) * taps, 0);
    for (int o = 0; o < image; ++o) {
        const int i = o + off;
        tap_min[o] = mins[i];
        tap_cnt[o] = static_cast<int>(ws[i].size());
        for (size_t j = 0; j < ws[i].size(); ++j) {
            const double v = ws[i][j] * (1 << prec);
            tap_w[static_cast<size_t>(o) * taps + j] = static_cast<int32_t>(v + (ws[i][j] >= 0 ? 0.5 : -0.5));
        }
    }
    *max_taps = taps;
    *precision = prec;
    return AP_OK;
}

int ap_preprocess_resize_run(ap_ctx* ctx, const uint8_t* slide, int64_t W, int64_t H, int64_t pitch, const int32_t* coords, int64_t n,
                             int input_patch, int image, int patch, const int32_t* tap_min, const int32_t* tap_cnt, const int32_t* tap_w,
                             int max_taps, int precision, int max_src_rows, __half* out, int64_t out_row_stride, const int* centre,
                             cudaStream_t stream) {
    AP_REQUIRE(ctx, image % patch == 0 && patch <= 32, "preprocess(resize): bad geometry image %d patch %d", image, patch);
    if (n == 0) return AP_OK;
    const int g = image / patch;
    const size_t smem = ((static_cast<size_t>(max_src_rows) * input_patch * 3 + 15) & ~static_cast<size_t>(15)) +
                        static_cast<size_t>(max_src_rows) * image * 3;
    AP_REQUIRE(ctx, smem <= 200 * 1024, "preprocess(resize): input patch %d needs %zu bytes of shared memory (<= 200 KB)", input_patch, smem);
    static PerDeviceOnce attr;   // the attribute is per device: set it to the 200 KB ceiling checked above, once per device
    if (attr.need(ctx->device)) {
        AP_CHECK_CUDA(ctx, cudaFuncSetAttribute(preprocess_resize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr.done(ctx->device);
    }
    ProfScope prof(ctx, stream, AP_K_PREPROCESS);
    AP_CHECK_CUDA(ctx, ap_launch_pdl(preprocess_resize_kernel, dim3(static_cast<unsigned>(n * g)), dim3(256), smem, stream, 1, ctx->pdl != 0,
                                     slide, W, H, pitch, coords, input_patch, image, patch, tap_min, tap_cnt, tap_w, max_taps, precision, out,
                                     out_row_stride, make_int3(centre[0], centre[1], centre[2])));
    AP_CHECK_LAUNCH(ctx, "preprocess_resize_kernel");
    return AP_OK;
}

extern "C" int ap_resize_tap_tables(int filter, int n_in, int n_out, int image, int32_t* tap_min, int32_t* tap_cnt, int32_t* tap_w,
                                    int taps_capacity, int* max_taps, int* precision) {
    if (!tap_min || !tap_cnt || !tap_w || !max_taps || !precision) return AP_EINVAL;
    std::vector<int32_t> tmin, tcnt, tw;
    int rc = ap_build_resize_tables(nullptr, n_in, n_out, image, filter, tmin, tcnt, tw, max_taps, precision);
    if (rc) return rc;
    if (*max_taps > taps_capacity) return AP_EINVAL;
    for (int o = 0; o < image; ++o) {
        tap_min[o] = tmin[o];
        tap_cnt[o] = tcnt[o];
        for (int j = 0; j < taps_capacity; ++j) tap_w[static_cast<size_t>(o) * taps_capacity + j] = j < *max_taps ? tw[static_cast<size_t>(o) * *max_taps + j] : 0;
    }
    return AP_OK;
}
