// a4: SAM2 box-prompted tissue mask on the 1024 x 1024 thumbnail -- weights, constants and the forward schedule.
//
// Replaces _SAM2Predictor.predict_image's model calls (atlas_patch/services/segmentation.py:127-136):
//   SAM2ImagePredictor.set_image  = pixels/255 -> ImageNet normalise -> Hiera trunk + FPN neck -> conv_s0 / conv_s1 -> + no_mem_embed
//   SAM2ImagePredictor.predict(box = whole image, multimask_output = False)
//                                 = prompt encoder (2 corner points, labels 2/3, + 1 padding point) -> two-way transformer ->
//                                   upscaling + hyper-network (mask token 0) -> 256 x 256 logits -> bilinear x4
// Hyper-parameters follow atlas_patch/configs/sam2.1_hiera_t.yaml:4-28,87-118 (Hiera-T) and are parametric (ap_sam2_desc).
// Parameter names are those of transformers' Sam2Model, the only runnable restatement in this image (the reference's `sam2`
// package and checkpoint are not available offline).  All arithmetic is fp32.
#include <cmath>
#include <string>
#include <unordered_map>
#include <vector>

#include "sam2_internal.cuh"

struct ap_sam2 {
    ap_ctx* ctx = nullptr;
    ap_sam2_desc d{};
    bool finalized = false;
    std::unordered_map<std::string, std::vector<float>> host;
    std::unordered_map<std::string, float*> w;                          // device copies of the parameters
    std::unordered_map<std::string, std::pair<float*, size_t>> bufs;    // named activation buffers (also for ap_sam2_debug_copy)
    std::vector<void*> allocs;
    uint8_t* img_dev = nullptr;
};

namespace {

constexpr int IMG = 1024;

int dalloc(ap_sam2* s, float** p, size_t n) {
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(float));
    if (e != cudaSuccess) return ap_set_error(s->ctx, AP_ENOMEM, "sam2: cudaMalloc(%zu floats) failed: %s", n, cudaGetErrorString(e));
    s->allocs.push_back(*p);
    return AP_OK;
}

// named activation buffer, allocated on first use at the requested size (sizes are fixed by the architecture)
float* buf(ap_sam2* s, const std::string& name, size_t n) {
    auto it = s->bufs.find(name);
    if (it != s->bufs.end() && it->second.second >= n) return it->second.first;
    float* p = nullptr;
    if (dalloc(s, &p, n)) return nullptr;
    s->bufs[name] = {p, n};
    return p;
}

int upload(ap_sam2* s, const std::string& name, const std::vector<float>& v) {
    float* p = nullptr;
    int rc = dalloc(s, &p, v.size());
    if (rc) return rc;
    AP_CHECK_CUDA(s->ctx, cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    s->w[name] = p;
    return AP_OK;
}

const float* W(ap_sam2* s, const std::string& name) {
    auto it = s->w.find(name);
    if (it == s->w.end()) {
        ap_set_error(s->ctx, AP_ESTATE, "sam2: parameter '%s' missing", name.c_str());
        return nullptr;
    }
    return it->second;
}

const std::vector<float>* H(ap_sam2* s, const std::string& name, size_t numel) {
    auto it = s->host.find(name);
    if (it == s->host.end() || it->second.size() != numel) {
        ap_set_error(s->ctx, AP_ESTATE, "sam2: tensor '%s' missing or wrong size (%zu expected)", name.c_str(), numel);
        return nullptr;
    }
    return &it->second;
}

// torch F.interpolate(mode="bicubic", align_corners=False) coefficients (A = -0.75)
inline double cc1(double x, double A) { return ((A + 2) * x - (A + 3)) * x * x + 1; }
inline double cc2(double x, double A) { return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A; }

#define SAM_TRY(call)               \
    do {                            \
        int rc__ = (call);          \
        if (rc__) return rc__;      \
    } while (0)
#define SAM_PTR(var, expr)          \
    auto var = (expr);              \
    if (!var) return AP_ESTATE;

// Sam2Attention (modeling_sam2.py:859-927): out[Lq, 256] = o_proj(SDPA(q_proj(query), k_proj(key), v_proj(value)))
int sam_attn_module(ap_sam2* s, const std::string& p, const float* query, int Lq, const float* key, const float* value, int Lk, int hidden,
                    int internal, int heads, float* out, cudaStream_t st) {
    ap_ctx* ctx = s->ctx;
    float* q = buf(s, "dec_q", static_cast<size_t>(4096) * 256);
    float* k = buf(s, "dec_k", static_cast<size_t>(4096) * 256);
    float* v = buf(s, "dec_v", static_cast<size_t>(4096) * 256);
    float* a = buf(s, "dec_a", static_cast<size_t>(4096) * 256);
    if (!q || !k || !v || !a) return AP_ENOMEM;
    SAM_PTR(wq, W(s, p + "q_proj.weight")) SAM_PTR(bq, W(s, p + "q_proj.bias"))
    SAM_PTR(wk, W(s, p + "k_proj.weight")) SAM_PTR(bk, W(s, p + "k_proj.bias"))
    SAM_PTR(wv, W(s, p + "v_proj.weight")) SAM_PTR(bv, W(s, p + "v_proj.bias"))
    SAM_PTR(wo, W(s, p + "o_proj.weight")) SAM_PTR(bo, W(s, p + "o_proj.bias"))
    SAM_TRY(sam_linear(ctx, query, hidden, wq, bq, q, internal, Lq, internal, hidden, SAM_ACT_NONE, 0, st));
    SAM_TRY(sam_linear(ctx, key, hidden, wk, bk, k, internal, Lk, internal, hidden, SAM_ACT_NONE, 0, st));
    SAM_TRY(sam_linear(ctx, value, hidden, wv, bv, v, internal, Lk, internal, hidden, SAM_ACT_NONE, 0, st));
    const int hd = internal / heads;
    SAM_TRY(sam_attention(ctx, q, internal, k, v, internal, a, internal, 1, Lq, Lk, heads, hd, 1.0f / sqrtf(static_cast<float>(hd)), st));
    return sam_linear(ctx, a, internal, wo, bo, out, hidden, Lq, hidden, internal, SAM_ACT_NONE, 0, st);
}

int sam_ln_named(ap_sam2* s, const std::string& p, const float* x, float* y, int rows, int D, float eps, int act, cudaStream_t st) {
    SAM_PTR(g, W(s, p + "weight")) SAM_PTR(b, W(s, p + "bias"))
    return sam_layernorm(s->ctx, x, g, b, y, rows, D, eps, act, st);
}

}  // namespace

extern "C" int ap_sam2_create(ap_ctx* ctx, const ap_sam2_desc* desc, ap_sam2** out) {
    if (!ctx || !desc || !out) return AP_EINVAL;
    DeviceGuard guard(ctx);
    *out = nullptr;
    AP_REQUIRE(ctx, desc->embed_dim > 0 && desc->embed_dim % desc->heads_per_stage[0] == 0, "sam2: bad embed_dim / heads");
    for (int i = 0; i < 4; ++i) {
        const int dim = desc->embed_dim << i;
        AP_REQUIRE(ctx, desc->blocks_per_stage[i] >= 1 && desc->heads_per_stage[i] >= 1 && dim % desc->heads_per_stage[i] == 0 &&
                            dim / desc->heads_per_stage[i] <= 128 && desc->window_per_stage[i] >= 1,
                   "sam2: stage %d configuration unsupported", i);
    }
    AP_REQUIRE(ctx, desc->n_global >= 0 && desc->n_global <= 8, "sam2: at most 8 global-attention blocks");
    ap_sam2* s = new ap_sam2();
    s->ctx = ctx;
    s->d = *desc;
    *out = s;
    return AP_OK;
}

extern "C" int ap_sam2_destroy(ap_sam2* s) {
    if (!s) return AP_OK;
    DeviceGuard guard(s->ctx);
    cudaDeviceSynchronize();
    sam_state_free(s->ctx);   // the split-weight cache is keyed by this model's (about to be freed) weight pointers
    for (void* p : s->allocs) cudaFree(p);
    if (s->img_dev) cudaFree(s->img_dev);
    delete s;
    return AP_OK;
}

extern "C" int ap_sam2_set_tensor(ap_sam2* s, const char* name, const float* data_host, int64_t numel) {
    if (!s || !name || !data_host || numel <= 0) return AP_EINVAL;
    if (s->finalized) return ap_set_error(s->ctx, AP_ESTATE, "sam2: already finalized");
    s->host[name].assign(data_host, data_host + numel);
    return AP_OK;
}

extern "C" int ap_sam2_finalize(ap_sam2* s) {
    if (!s) return AP_EINVAL;
    DeviceGuard guard(s->ctx);
    if (s->finalized) return AP_OK;
    ap_ctx* ctx = s->ctx;
    const int C0 = s->d.embed_dim;
    // ---- derived constants (host) ------------------------------------------------------------------------------------
    {   // positional embedding of the trunk: bicubic(pos_embed 7x7 -> 256x256) + tiled window embedding (modeling_sam2.py:629-636)
        const int bg = 7, ws0 = s->d.window_per_stage[0], G = IMG / 4;
        SAM_PTR(pe, H(s, "vision_encoder.backbone.pos_embed", static_cast<size_t>(C0) * bg * bg))
        SAM_PTR(pw, H(s, "vision_encoder.backbone.pos_embed_window", static_cast<size_t>(C0) * ws0 * ws0))
        std::vector<float> pos(static_cast<size_t>(G) * G * C0);
        std::vector<int> ix(G * 4);
        std::vector<double> cw(G * 4);
        const double A = -0.75, scale = static_cast<double>(bg) / G;
        for (int o = 0; o < G; ++o) {
            const double real = scale * (o + 0.5) - 0.5;
            const double fl = std::floor(real);
            const double t = real - fl;
            const double c[4] = {cc2(t + 1, A), cc1(t, A), cc1(1 - t, A), cc2(2 - t, A)};
            for (int k = 0; k < 4; ++k) {
                int i = static_cast<int>(fl) - 1 + k;
                ix[o * 4 + k] = i < 0 ? 0 : (i > bg - 1 ? bg - 1 : i);
                cw[o * 4 + k] = c[k];
            }
        }
        for (int y = 0; y < G; ++y)
            for (int x = 0; x < G; ++x)
                for (int c = 0; c < C0; ++c) {
                    double acc = 0;
                    for (int ky = 0; ky < 4; ++ky) {
                        double row = 0;
                        for (int kx = 0; kx < 4; ++kx) row += cw[x * 4 + kx] * (*pe)[(static_cast<size_t>(c) * bg + ix[y * 4 + ky]) * bg + ix[x * 4 + kx]];
                        acc += cw[y * 4 + ky] * row;
                    }
                    pos[(static_cast<size_t>(y) * G + x) * C0 + c] =
                        static_cast<float>(acc) + (*pw)[(static_cast<size_t>(c) * ws0 + y % ws0) * ws0 + x % ws0];
                }
        SAM_TRY(upload(s, "__pos", pos));
    }
    {   // dense positional encoding of the 64x64 image embedding (get_image_wide_positional_embeddings)
        SAM_PTR(g, H(s, "shared_image_embedding.positional_embedding", 2 * 128))
        const int E = 64;
        std::vector<float> pe(static_cast<size_t>(E) * E * 256);
        for (int y = 0; y < E; ++y)
            for (int x = 0; x < E; ++x) {
                const float cx = 2.f * ((x + 0.5f) / E) - 1.f, cy = 2.f * ((y + 0.5f) / E) - 1.f;
                for (int j = 0; j < 128; ++j) {
                    const float v = 2.f * 3.14159265358979323846f * (cx * (*g)[j] + cy * (*g)[128 + j]);
                    pe[(static_cast<size_t>(y) * E + x) * 256 + j] = sinf(v);
                    pe[(static_cast<size_t>(y) * E + x) * 256 + 128 + j] = cosf(v);
                }
            }
        SAM_TRY(upload(s, "__image_pe", pe));
    }
    {   // decoder input tokens for the fixed prompt "box = whole image": [obj_score, iou, mask x4, corner(2), corner(3), pad]
        SAM_PTR(g, H(s, "prompt_encoder.shared_embedding.positional_embedding", 2 * 128))
        SAM_PTR(pt, H(s, "prompt_encoder.point_embed.weight", 4 * 256))
        SAM_PTR(nap, H(s, "prompt_encoder.not_a_point_embed.weight", 256))
        SAM_PTR(obj, H(s, "mask_decoder.obj_score_token.weight", 256))
        SAM_PTR(iou, H(s, "mask_decoder.iou_token.weight", 256))
        SAM_PTR(mt, H(s, "mask_decoder.mask_tokens.weight", 4 * 256))
        std::vector<float> tok(9 * 256);
        std::copy(obj->begin(), obj->end(), tok.begin());
        std::copy(iou->begin(), iou->end(), tok.begin() + 256);
        std::copy(mt->begin(), mt->end(), tok.begin() + 512);
        const float corner[2] = {0.f + 0.5f, static_cast<float>(IMG) + 0.5f};   // _embed_points: +0.5, /image size
        for (int c = 0; c < 2; ++c) {
            const float u = 2.f * (corner[c] / IMG) - 1.f;
            for (int j = 0; j < 128; ++j) {
                const float v = 2.f * 3.14159265358979323846f * (u * (*g)[j] + u * (*g)[128 + j]);
                tok[(6 + c) * 256 + j] = sinf(v) + (*pt)[(2 + c) * 256 + j];
                tok[(6 + c) * 256 + 128 + j] = cosf(v) + (*pt)[(2 + c) * 256 + 128 + j];
            }
        }
        std::copy(nap->begin(), nap->end(), tok.begin() + 8 * 256);
        SAM_TRY(upload(s, "__tokens0", tok));
        // constant added to the image embedding: no_memory_embedding + no_mask_embed (dense prompt)
        SAM_PTR(nm, H(s, "no_memory_embedding", 256))
        SAM_PTR(nk, H(s, "prompt_encoder.no_mask_embed.weight", 256))
        std::vector<float> dc(256);
        for (int j = 0; j < 256; ++j) dc[j] = (*nm)[j] + (*nk)[j];
        SAM_TRY(upload(s, "__dense_const", dc));
    }
    for (int u = 0; u < 2; ++u) {   // ConvTranspose2d(k=2, s=2) as a GEMM: Wt[(ky*2+kx)*Co + co][ci] = w[ci][co][ky][kx]
        const int Ci = u == 0 ? 256 : 64, Co = u == 0 ? 64 : 32;
        const std::string n = u == 0 ? "mask_decoder.upscale_conv1." : "mask_decoder.upscale_conv2.";
        SAM_PTR(w, H(s, n + "weight", static_cast<size_t>(Ci) * Co * 4))
        SAM_PTR(b, H(s, n + "bias", Co))
        std::vector<float> wt(static_cast<size_t>(4) * Co * Ci), bt(4 * Co);
        for (int ci = 0; ci < Ci; ++ci)
            for (int co = 0; co < Co; ++co)
                for (int k = 0; k < 4; ++k) wt[(static_cast<size_t>(k) * Co + co) * Ci + ci] = (*w)[(static_cast<size_t>(ci) * Co + co) * 4 + k];
        for (int k = 0; k < 4; ++k)
            for (int co = 0; co < Co; ++co) bt[k * Co + co] = (*b)[co];
        SAM_TRY(upload(s, n + "__wt", wt));
        SAM_TRY(upload(s, n + "__bt", bt));
    }
    {   // patch-embed weights [C0][3][7][7] as a [C0][148] matrix (147 taps + a zero column) for the im2col + GEMM form of the convolution
        SAM_PTR(w, H(s, "vision_encoder.backbone.patch_embed.projection.weight", static_cast<size_t>(C0) * 147))
        std::vector<float> w148(static_cast<size_t>(C0) * 148, 0.f);
        for (int c = 0; c < C0; ++c) std::copy(w->begin() + static_cast<size_t>(c) * 147, w->begin() + static_cast<size_t>(c + 1) * 147, w148.begin() + static_cast<size_t>(c) * 148);
        SAM_TRY(upload(s, "__pe_w148", w148));
    }
    for (auto& kv : s->host) SAM_TRY(upload(s, kv.first, kv.second));
    s->host.clear();
    AP_CHECK_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&s->img_dev), static_cast<size_t>(IMG) * IMG * 3));
    s->finalized = true;
    return AP_OK;
}

// image_dev: uint8 [1024, 1024, 3] on the device.  logits_dev: float [1024, 1024]; lowres_dev (optional): float [256, 256].
extern "C" int ap_sam2_forward(ap_sam2* s, const uint8_t* image_dev, float* logits_dev, float* lowres_dev, void* stream) {
    if (!s) return AP_EINVAL;
    DeviceGuard guard(s->ctx);
    ap_ctx* ctx = s->ctx;
    if (!s->finalized) return ap_set_error(ctx, AP_ESTATE, "sam2: ap_sam2_finalize has not been called");
    AP_REQUIRE(ctx, image_dev && logits_dev, "sam2_forward: NULL pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const ap_sam2_desc& d = s->d;
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    const float eps = 1e-6f;
    int Hc = IMG / 4, Wc = IMG / 4;
    const int C0 = d.embed_dim;

    // ---- trunk ----------------------------------------------------------------------------------------------------------
    float* cur = buf(s, "patch_embed", static_cast<size_t>(Hc) * Wc * C0);
    if (!cur) return AP_ENOMEM;
    {
        SAM_PTR(w, W(s, "vision_encoder.backbone.patch_embed.projection.weight"))
        SAM_PTR(b, W(s, "vision_encoder.backbone.patch_embed.projection.bias"))
        SAM_PTR(pos, W(s, "__pos"))
        if (ctx->sam_tensor_cores == 3) {   // im2col + the tcgen05 split GEMM, then + positional embedding
            SAM_PTR(w148, W(s, "__pe_w148"))
            float* cols = buf(s, "pe_cols", static_cast<size_t>(Hc) * Wc * 148);
            if (!cols) return AP_ENOMEM;
            SAM_TRY(sam_patch_im2col(ctx, image_dev, IMG, IMG, cols, mean, stdv, st));
            SAM_TRY(sam_linear(ctx, cols, 148, w148, b, cur, C0, Hc * Wc, C0, 148, SAM_ACT_NONE, 0, st));
            SAM_TRY(sam_add(ctx, cur, pos, cur, static_cast<int64_t>(Hc) * Wc * C0, C0, 0, st));
        } else {
            SAM_TRY(sam_patch_embed(ctx, image_dev, IMG, IMG, w, b, pos, cur, C0, mean, stdv, st));
        }
    }
    const float* stage_out[4] = {nullptr, nullptr, nullptr, nullptr};
    int stage_hw[4] = {0, 0, 0, 0};
    int blk = 0;
    for (int sidx = 0; sidx < 4; ++sidx) {
        for (int b = 0; b < d.blocks_per_stage[sidx]; ++b, ++blk) {
            const int dim_out = C0 << sidx;
            const int dim_in = (sidx > 0 && b == 0) ? (C0 << (sidx - 1)) : dim_out;
            int ws = (sidx > 0 && b == 0) ? d.window_per_stage[sidx - 1] : d.window_per_stage[sidx];
            for (int g = 0; g < d.n_global; ++g)
                if (d.global_blocks[g] == blk) ws = 0;
            const bool qpool = sidx > 0 && b == 0;   // num_query_pool_stages = 3: every stage transition
            const int heads = d.heads_per_stage[sidx], hd = dim_out / heads;
            const std::string p = "vision_encoder.backbone.blocks." + std::to_string(blk) + ".";
            const int T = Hc * Wc;
            AP_REQUIRE(ctx, !(qpool && ws == 0) && !(qpool && (ws % 2)), "sam2: block %d: q-pooling needs an even window", blk);

            float* ln = buf(s, "ln", static_cast<size_t>(T) * dim_in);
            if (!ln) return AP_ENOMEM;
            SAM_TRY(sam_ln_named(s, p + "layer_norm1.", cur, ln, T, dim_in, eps, SAM_ACT_NONE, st));
            const float* res = cur;
            if (dim_in != dim_out) {   // skip path: do_pool(proj(norm1(x)))
                float* tmp = buf(s, "proj_tmp", static_cast<size_t>(T) * dim_out);
                float* rp = buf(s, "res_pool", static_cast<size_t>(T / 4) * dim_out);
                if (!tmp || !rp) return AP_ENOMEM;
                SAM_PTR(w, W(s, p + "proj.weight")) SAM_PTR(bb, W(s, p + "proj.bias"))
                SAM_TRY(sam_linear(ctx, ln, dim_in, w, bb, tmp, dim_out, T, dim_out, dim_in, SAM_ACT_NONE, 0, st));
                SAM_TRY(sam_maxpool2(ctx, tmp, dim_out, rp, 1, Hc, Wc, dim_out, st));
                res = rp;
            }
            int nB = 1, Lk = T, nWx = 1, nWy = 1;
            const float* win = ln;
            const bool fuse_windows = ws > 0 && ctx->sam_tensor_cores == 3 && dim_in % 4 == 0;   // window_partition inside the operand staging
            if (ws > 0) {
                nWy = (Hc + ws - 1) / ws; nWx = (Wc + ws - 1) / ws;
                nB = nWy * nWx; Lk = ws * ws;
                if (!fuse_windows) {
                    float* wb = buf(s, "win", static_cast<size_t>(nB) * Lk * dim_in);
                    if (!wb) return AP_ENOMEM;
                    SAM_TRY(sam_window_gather(ctx, ln, wb, Hc, Wc, dim_in, ws, nWy, nWx, st));
                    win = wb;
                }
            }
            const int ntok = nB * Lk;
            float* qkv = buf(s, "qkv", static_cast<size_t>(ntok) * 3 * dim_out);
            if (!qkv) return AP_ENOMEM;
            {
                SAM_PTR(w, W(s, p + "attn.qkv.weight")) SAM_PTR(bb, W(s, p + "attn.qkv.bias"))
                if (fuse_windows) {
                    const SamGather g{Hc, Wc, ws, nWx};
                    SAM_TRY(sam_linear_windows(ctx, ln, dim_in, g, w, bb, qkv, 3 * dim_out, ntok, 3 * dim_out, dim_in, st));
                } else {
                    SAM_TRY(sam_linear(ctx, win, dim_in, w, bb, qkv, 3 * dim_out, ntok, 3 * dim_out, dim_in, SAM_ACT_NONE, 0, st));
                }
            }
            const float* q = qkv;
            int q_stride = 3 * dim_out, Lq = Lk, ws_out = ws;
            if (qpool) {   // 2x2 max pool of the queries inside every window (modeling_sam2.py:316-319)
                float* qp = buf(s, "q_pool", static_cast<size_t>(ntok / 4) * dim_out);
                if (!qp) return AP_ENOMEM;
                SAM_TRY(sam_maxpool2(ctx, qkv, 3 * dim_out, qp, nB, ws, ws, dim_out, st));
                q = qp; q_stride = dim_out; Lq = Lk / 4; ws_out = ws / 2;
                Hc /= 2; Wc /= 2;
            }
            float* att = buf(s, "att", static_cast<size_t>(nB) * Lq * dim_out);
            float* pr = buf(s, "att_proj", static_cast<size_t>(nB) * Lq * dim_out);
            if (!att || !pr) return AP_ENOMEM;
            SAM_TRY(sam_attention(ctx, q, q_stride, qkv + dim_out, qkv + 2 * dim_out, 3 * dim_out, att, dim_out, nB, Lq, Lk, heads, hd,
                                  1.0f / sqrtf(static_cast<float>(hd)), st));
            {
                SAM_PTR(w, W(s, p + "attn.proj.weight")) SAM_PTR(bb, W(s, p + "attn.proj.bias"))
                SAM_TRY(sam_linear(ctx, att, dim_out, w, bb, pr, dim_out, nB * Lq, dim_out, dim_out, SAM_ACT_NONE, 0, st));
            }
            const int T2 = Hc * Wc;
            float* xn = buf(s, "blk" + std::to_string(blk), static_cast<size_t>(T2) * dim_out);
            if (!xn) return AP_ENOMEM;
            if (ws > 0) SAM_TRY(sam_window_scatter_add(ctx, pr, res, xn, Hc, Wc, dim_out, ws_out, nWx, st));
            else SAM_TRY(sam_add(ctx, res, pr, xn, static_cast<int64_t>(T2) * dim_out, dim_out, 0, st));
            float* ln2 = buf(s, "ln2", static_cast<size_t>(T2) * dim_out);
            float* hb = buf(s, "mlp_h", static_cast<size_t>(T2) * 4 * dim_out);
            if (!ln2 || !hb) return AP_ENOMEM;
            SAM_TRY(sam_ln_named(s, p + "layer_norm2.", xn, ln2, T2, dim_out, eps, SAM_ACT_NONE, st));
            {
                SAM_PTR(w1, W(s, p + "mlp.proj_in.weight")) SAM_PTR(b1, W(s, p + "mlp.proj_in.bias"))
                SAM_PTR(w2, W(s, p + "mlp.proj_out.weight")) SAM_PTR(b2, W(s, p + "mlp.proj_out.bias"))
                SAM_TRY(sam_linear(ctx, ln2, dim_out, w1, b1, hb, 4 * dim_out, T2, 4 * dim_out, dim_out, SAM_ACT_GELU, 0, st));
                SAM_TRY(sam_linear(ctx, hb, 4 * dim_out, w2, b2, xn, dim_out, T2, dim_out, 4 * dim_out, SAM_ACT_NONE, 1, st));
            }
            cur = xn;
        }
        stage_out[sidx] = cur;
        stage_hw[sidx] = Hc;
    }

    // ---- FPN neck (modeling_sam2.py:199-246): lateral 1x1 convs, nearest x2 top-down on level 2 only, level 3 dropped -----------
    float* fpn[4];
    for (int i = 3; i >= 0; --i) {
        const int T = stage_hw[i] * stage_hw[i], Cin = C0 << i;
        fpn[i] = buf(s, "fpn" + std::to_string(i), static_cast<size_t>(T) * 256);
        if (!fpn[i]) return AP_ENOMEM;
        const std::string p = "vision_encoder.neck.convs." + std::to_string(3 - i) + ".";
        SAM_PTR(w, W(s, p + "weight")) SAM_PTR(bb, W(s, p + "bias"))
        SAM_TRY(sam_linear(ctx, stage_out[i], Cin, w, bb, fpn[i], 256, T, 256, Cin, SAM_ACT_NONE, 0, st));
        if (i == 2) SAM_TRY(sam_upsample2_add(ctx, fpn[3], fpn[2], stage_hw[3], stage_hw[3], 256, st));
    }
    const int E0 = stage_hw[0], E1 = stage_hw[1], E2 = stage_hw[2];   // 256, 128, 64
    float* feat_s0 = buf(s, "feat_s0", static_cast<size_t>(E0) * E0 * 32);
    float* feat_s1 = buf(s, "feat_s1", static_cast<size_t>(E1) * E1 * 64);
    float* keys = buf(s, "keys", static_cast<size_t>(E2) * E2 * 256);
    if (!feat_s0 || !feat_s1 || !keys) return AP_ENOMEM;
    {
        SAM_PTR(w0, W(s, "mask_decoder.conv_s0.weight")) SAM_PTR(b0, W(s, "mask_decoder.conv_s0.bias"))
        SAM_PTR(w1, W(s, "mask_decoder.conv_s1.weight")) SAM_PTR(b1, W(s, "mask_decoder.conv_s1.bias"))
        SAM_PTR(dc, W(s, "__dense_const"))
        SAM_TRY(sam_linear(ctx, fpn[0], 256, w0, b0, feat_s0, 32, E0 * E0, 32, 256, SAM_ACT_NONE, 0, st));
        SAM_TRY(sam_linear(ctx, fpn[1], 256, w1, b1, feat_s1, 64, E1 * E1, 64, 256, SAM_ACT_NONE, 0, st));
        // image embedding + no_mem_embed (services: directly_add_no_mem_embed) + dense "no mask" prompt embedding
        SAM_TRY(sam_add(ctx, fpn[2], dc, keys, static_cast<int64_t>(E2) * E2 * 256, 256, 1, st));
    }

    // ---- mask decoder: two-way transformer on 9 tokens x 4096 image tokens (modeling_sam2.py:930-1062) ---------------------
    const int NT = 9, NI = E2 * E2, HID = 256;
    SAM_PTR(tok0, W(s, "__tokens0"))
    SAM_PTR(ipe, W(s, "__image_pe"))
    float* queries = buf(s, "queries", NT * HID);
    float* qpe = buf(s, "q_pe", NT * HID);
    float* kpe = buf(s, "k_pe", static_cast<size_t>(NI) * HID);
    float* ao = buf(s, "attn_out", static_cast<size_t>(NI) * HID);
    float* mh = buf(s, "dec_mlp_h", NT * 2048);
    if (!queries || !qpe || !kpe || !ao || !mh) return AP_ENOMEM;
    AP_CHECK_CUDA(ctx, cudaMemcpyAsync(queries, tok0, NT * HID * sizeof(float), cudaMemcpyDeviceToDevice, st));
    const float dec_eps = 1e-5f;   // nn.LayerNorm default inside the two-way blocks
    for (int l = 0; l < 2; ++l) {
        const std::string p = "mask_decoder.transformer.layers." + std::to_string(l) + ".";
        if (l == 0) {   // skip_first_layer_pe: queries = self_attn(queries, queries, queries)
            SAM_TRY(sam_attn_module(s, p + "self_attn.", queries, NT, queries, queries, NT, HID, 256, 8, ao, st));
            AP_CHECK_CUDA(ctx, cudaMemcpyAsync(queries, ao, NT * HID * sizeof(float), cudaMemcpyDeviceToDevice, st));
        } else {
            SAM_TRY(sam_add(ctx, queries, tok0, qpe, NT * HID, HID, 0, st));
            SAM_TRY(sam_attn_module(s, p + "self_attn.", qpe, NT, qpe, queries, NT, HID, 256, 8, ao, st));
            SAM_TRY(sam_add(ctx, queries, ao, queries, NT * HID, HID, 0, st));
        }
        SAM_TRY(sam_ln_named(s, p + "layer_norm1.", queries, queries, NT, HID, dec_eps, SAM_ACT_NONE, st));
        // tokens -> image
        SAM_TRY(sam_add(ctx, queries, tok0, qpe, NT * HID, HID, 0, st));
        SAM_TRY(sam_add(ctx, keys, ipe, kpe, static_cast<int64_t>(NI) * HID, HID, 0, st));
        SAM_TRY(sam_attn_module(s, p + "cross_attn_token_to_image.", qpe, NT, kpe, keys, NI, HID, 128, 8, ao, st));
        SAM_TRY(sam_add(ctx, queries, ao, queries, NT * HID, HID, 0, st));
        SAM_TRY(sam_ln_named(s, p + "layer_norm2.", queries, queries, NT, HID, dec_eps, SAM_ACT_NONE, st));
        {   // MLP (ReLU)
            SAM_PTR(w1, W(s, p + "mlp.proj_in.weight")) SAM_PTR(b1, W(s, p + "mlp.proj_in.bias"))
            SAM_PTR(w2, W(s, p + "mlp.proj_out.weight")) SAM_PTR(b2, W(s, p + "mlp.proj_out.bias"))
            SAM_TRY(sam_linear(ctx, queries, HID, w1, b1, mh, 2048, NT, 2048, HID, SAM_ACT_RELU, 0, st));
            SAM_TRY(sam_linear(ctx, mh, 2048, w2, b2, queries, HID, NT, HID, 2048, SAM_ACT_NONE, 1, st));
        }
        SAM_TRY(sam_ln_named(s, p + "layer_norm3.", queries, queries, NT, HID, dec_eps, SAM_ACT_NONE, st));
        // image -> tokens
        SAM_TRY(sam_add(ctx, queries, tok0, qpe, NT * HID, HID, 0, st));
        SAM_TRY(sam_attn_module(s, p + "cross_attn_image_to_token.", kpe, NI, qpe, queries, NT, HID, 128, 8, ao, st));
        SAM_TRY(sam_add(ctx, keys, ao, keys, static_cast<int64_t>(NI) * HID, HID, 0, st));
        SAM_TRY(sam_ln_named(s, p + "layer_norm4.", keys, keys, NI, HID, dec_eps, SAM_ACT_NONE, st));
    }
    SAM_TRY(sam_add(ctx, queries, tok0, qpe, NT * HID, HID, 0, st));
    SAM_TRY(sam_add(ctx, keys, ipe, kpe, static_cast<int64_t>(NI) * HID, HID, 0, st));
    SAM_TRY(sam_attn_module(s, "mask_decoder.transformer.final_attn_token_to_image.", qpe, NT, kpe, keys, NI, HID, 128, 8, ao, st));
    SAM_TRY(sam_add(ctx, queries, ao, queries, NT * HID, HID, 0, st));
    SAM_TRY(sam_ln_named(s, "mask_decoder.transformer.layer_norm_final_attn.", queries, queries, NT, HID, dec_eps, SAM_ACT_NONE, st));

    // ---- upscaling + hyper-network for mask token 0 (modeling_sam2.py:1208-1232) --------------------------------------------
    float* lin1 = buf(s, "up_lin1", static_cast<size_t>(NI) * 256);
    float* u1 = buf(s, "up1", static_cast<size_t>(E1) * E1 * 64);
    float* lin2 = buf(s, "up_lin2", static_cast<size_t>(E1) * E1 * 128);
    float* u2 = buf(s, "upscaled", static_cast<size_t>(E0) * E0 * 32);
    float* hy = buf(s, "hyper", 2 * 256 + 32);
    float* low = buf(s, "low_res", static_cast<size_t>(E0) * E0);
    if (!lin1 || !u1 || !lin2 || !u2 || !hy || !low) return AP_ENOMEM;
    {
        SAM_PTR(wt1, W(s, "mask_decoder.upscale_conv1.__wt")) SAM_PTR(bt1, W(s, "mask_decoder.upscale_conv1.__bt"))
        SAM_PTR(wt2, W(s, "mask_decoder.upscale_conv2.__wt")) SAM_PTR(bt2, W(s, "mask_decoder.upscale_conv2.__bt"))
        SAM_TRY(sam_linear(ctx, keys, 256, wt1, bt1, lin1, 256, NI, 256, 256, SAM_ACT_NONE, 0, st));
        SAM_TRY(sam_pixel_shuffle_add(ctx, lin1, feat_s1, u1, E2, E2, 64, SAM_ACT_NONE, st));
        SAM_TRY(sam_ln_named(s, "mask_decoder.upscale_layer_norm.", u1, u1, E1 * E1, 64, 1e-6f, SAM_ACT_GELU, st));
        SAM_TRY(sam_linear(ctx, u1, 64, wt2, bt2, lin2, 128, E1 * E1, 128, 64, SAM_ACT_NONE, 0, st));
        SAM_TRY(sam_pixel_shuffle_add(ctx, lin2, feat_s0, u2, E1, E1, 32, SAM_ACT_GELU, st));
        const std::string p = "mask_decoder.output_hypernetworks_mlps.0.";
        SAM_PTR(w1, W(s, p + "proj_in.weight")) SAM_PTR(b1, W(s, p + "proj_in.bias"))
        SAM_PTR(wl, W(s, p + "layers.0.weight")) SAM_PTR(bl, W(s, p + "layers.0.bias"))
        SAM_PTR(w2, W(s, p + "proj_out.weight")) SAM_PTR(b2, W(s, p + "proj_out.bias"))
        SAM_TRY(sam_linear(ctx, queries + 2 * HID, HID, w1, b1, hy, 256, 1, 256, 256, SAM_ACT_RELU, 0, st));
        SAM_TRY(sam_linear(ctx, hy, 256, wl, bl, hy + 256, 256, 1, 256, 256, SAM_ACT_RELU, 0, st));
        SAM_TRY(sam_linear(ctx, hy + 256, 256, w2, b2, hy + 512, 32, 1, 32, 256, SAM_ACT_NONE, 0, st));
        // mask logits = hyper (32) . upscaled (32 per pixel)
        SAM_TRY(sam_linear(ctx, u2, 32, hy + 512, nullptr, low, 1, E0 * E0, 1, 32, SAM_ACT_NONE, 0, st));
    }
    if (lowres_dev) AP_CHECK_CUDA(ctx, cudaMemcpyAsync(lowres_dev, low, static_cast<size_t>(E0) * E0 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return sam_bilinear(ctx, low, E0, E0, logits_dev, IMG, IMG, st);
}

// Host-side convenience for the segmentation adapter / tests: uint8 image on the host in, logits on the host out.
extern "C" int ap_sam2_predict_host(ap_sam2* s, const uint8_t* image_host, float* logits_host, float* lowres_host) {
    if (!s || !image_host || !logits_host) return AP_EINVAL;
    DeviceGuard guard(s->ctx);
    ap_ctx* ctx = s->ctx;
    if (!s->finalized) return ap_set_error(ctx, AP_ESTATE, "sam2: ap_sam2_finalize has not been called");
    float* lg = buf(s, "logits", static_cast<size_t>(IMG) * IMG);
    float* lo = buf(s, "lowres_out", 256 * 256);
    if (!lg || !lo) return AP_ENOMEM;
    AP_CHECK_CUDA(ctx, cudaMemcpy(s->img_dev, image_host, static_cast<size_t>(IMG) * IMG * 3, cudaMemcpyHostToDevice));
    int rc = ap_sam2_forward(s, s->img_dev, lg, lo, nullptr);
    if (rc) return rc;
    AP_CHECK_CUDA(ctx, cudaMemcpy(logits_host, lg, static_cast<size_t>(IMG) * IMG * sizeof(float), cudaMemcpyDeviceToHost));
    if (lowres_host) AP_CHECK_CUDA(ctx, cudaMemcpy(lowres_host, lo, 256 * 256 * sizeof(float), cudaMemcpyDeviceToHost));
    return AP_OK;
}

// A batch of thumbnails (the reference's predict_batch: SAM2ImagePredictor.set_image_batch + predict_batch, one whole-image box per
// image, services/segmentation.py:142-180).  The model's activation buffers hold one image; the batch is pipelined instead: image
// i + 1 is uploaded on a copy stream while image i runs, its logits are downloaded while image i + 1 runs.  Outputs are binary-
// thresholded by the caller; logits_host: n x [1024, 1024] floats, lowres_host: n x [256, 256] floats or NULL.
extern "C" int ap_sam2_predict_batch_host(ap_sam2* s, const uint8_t* images_host, int n, float* logits_host, float* lowres_host) {
    if (!s || (n > 0 && (!images_host || !logits_host))) return AP_EINVAL;
    DeviceGuard guard(s->ctx);
    ap_ctx* ctx = s->ctx;
    if (!s->finalized) return ap_set_error(ctx, AP_ESTATE, "sam2: ap_sam2_finalize has not been called");
    if (n <= 0) return AP_OK;
    const size_t img_bytes = static_cast<size_t>(IMG) * IMG * 3, lg_n = static_cast<size_t>(IMG) * IMG, lo_n = 256 * 256;
    uint8_t* img[2] = {s->img_dev, nullptr};
    float* lg[2] = {buf(s, "logits", lg_n), buf(s, "logits_b", lg_n)};
    float* lo[2] = {buf(s, "lowres_out", lo_n), buf(s, "lowres_out_b", lo_n)};
    if (!lg[0] || !lg[1] || !lo[0] || !lo[1]) return AP_ENOMEM;
    AP_CHECK_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&img[1]), img_bytes));
    cudaStream_t s_copy = nullptr, s_run = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, done[2] = {nullptr, nullptr}, drained[2] = {nullptr, nullptr};
    int rc = AP_OK;
    auto ck = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && rc == AP_OK) rc = ap_set_error(ctx, AP_ECUDA, "%s failed: %s", what, cudaGetErrorString(e));
        return e == cudaSuccess;
    };
    ck(cudaStreamCreateWithFlags(&s_copy, cudaStreamNonBlocking), "cudaStreamCreate");
    ck(cudaStreamCreateWithFlags(&s_run, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int i = 0; i < 2; ++i) {
        ck(cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming), "cudaEventCreate");
        ck(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming), "cudaEventCreate");
        ck(cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming), "cudaEventCreate");
    }
    if (rc == AP_OK) {
        ck(cudaMemcpyAsync(img[0], images_host, img_bytes, cudaMemcpyHostToDevice, s_copy), "upload");
        ck(cudaEventRecord(up[0], s_copy), "record");
        for (int i = 0; i < n && rc == AP_OK; ++i) {
            const int b = i & 1;
            if (i + 1 < n) {   // next image in flight while this one runs (its buffer was last read by forward i - 1)
                if (i >= 1) ck(cudaStreamWaitEvent(s_copy, done[b ^ 1], 0), "wait");
                ck(cudaMemcpyAsync(img[b ^ 1], images_host + static_cast<size_t>(i + 1) * img_bytes, img_bytes, cudaMemcpyHostToDevice, s_copy), "upload");
                ck(cudaEventRecord(up[b ^ 1], s_copy), "record");
            }
            ck(cudaStreamWaitEvent(s_run, up[b], 0), "wait");
            if (i >= 2) ck(cudaStreamWaitEvent(s_run, drained[b], 0), "wait");   // logits buffer b has been downloaded
            if (rc != AP_OK) break;
            int frc = ap_sam2_forward(s, img[b], lg[b], lo[b], s_run);
            if (frc) { rc = frc; break; }
            ck(cudaEventRecord(done[b], s_run), "record");
            ck(cudaStreamWaitEvent(s_copy, done[b], 0), "wait");
            ck(cudaMemcpyAsync(logits_host + static_cast<size_t>(i) * lg_n, lg[b], lg_n * sizeof(float), cudaMemcpyDeviceToHost, s_copy), "download");
            if (lowres_host) ck(cudaMemcpyAsync(lowres_host + static_cast<size_t>(i) * lo_n, lo[b], lo_n * sizeof(float), cudaMemcpyDeviceToHost, s_copy), "download");
            ck(cudaEventRecord(drained[b], s_copy), "record");
        }
    }
    if (s_run) cudaStreamSynchronize(s_run);
    if (s_copy) cudaStreamSynchronize(s_copy);
    for (int i = 0; i < 2; ++i) {
        if (up[i]) cudaEventDestroy(up[i]);
        if (done[i]) cudaEventDestroy(done[i]);
        if (drained[i]) cudaEventDestroy(drained[i]);
    }
    if (s_copy) cudaStreamDestroy(s_copy);
    if (s_run) cudaStreamDestroy(s_run);
    cudaFree(img[1]);
    return rc;
}

extern "C" int ap_sam2_debug_copy(ap_sam2* s, const char* buffer_name, float* host_out, int64_t numel) {
    if (!s || !buffer_name || !host_out) return AP_EINVAL;
    DeviceGuard guard(s->ctx);
    auto it = s->bufs.find(buffer_name);
    if (it == s->bufs.end()) return ap_set_error(s->ctx, AP_EINVAL, "sam2: no buffer named '%s'", buffer_name);
    if (static_cast<size_t>(numel) > it->second.second) return ap_set_error(s->ctx, AP_EINVAL, "sam2: buffer '%s' has %zu floats", buffer_name, it->second.second);
    AP_CHECK_CUDA(s->ctx, cudaDeviceSynchronize());
    AP_CHECK_CUDA(s->ctx, cudaMemcpy(host_out, it->second.first, numel * sizeof(float), cudaMemcpyDeviceToHost));
    return AP_OK;
}
