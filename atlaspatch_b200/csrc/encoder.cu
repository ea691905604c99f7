// ViT encoder object: weight upload/packing, workspaces, and the forward schedule.
//
// Forward = torchvision VisionTransformer with heads -> Identity, as the reference builds it
// (atlas_patch/models/patch/vit.py:9-38, models/patch/base.py:148-180):
//   conv_proj (k = s = patch) -> [class_token ; tokens] + pos_embedding
//   -> layers x [ x += out_proj(MHA(ln_1(x))) ; x += mlp.3(GELU(mlp.0(ln_2(x)))) ] -> ln -> x[:, 0]
// Residual stream, LayerNorm and softmax are fp32; GEMM operands are fp16 with fp32 accumulation.
#include <map>
#include <string>
#include <cstdlib>
#include <thread>
#include <unordered_map>
#include <vector>

#include "ap_internal.cuh"

namespace {

struct LayerWeights {
    float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
    __half *w_qkv, *w_o, *w_1, *w_2;
    float *b_qkv, *b_o, *b_1, *b_2;
    GemmPlan p_qkv, p_o, p_1, p_2;
    GemmPlan pc_o, pc_1, pc_2;  // last layer only: class-token rows (M = images)
    int split = 0;  // bit mask of GEMMs whose weights are stored as [hi | lo] fp16 pairs (K doubled): 1 qkv, 2 out_proj, 4 mlp.0, 8 mlp.3
    int asplit = 0; // bit mask of GEMMs whose A operand is stored as [hi | lo] instead (1 qkv, 2 out_proj, 4 mlp.0): the LayerNorm /
                    // attention epilogue that produces it also writes lo = fp16(a - fp16(a)); weights stay single fp16
    bool fold1 = false, fold2 = false;  // LayerNorm 1 / 2 of this layer folded into in_proj / mlp.0 (not where that GEMM's A operand is split)
};

}  // namespace

struct ap_encoder {
    ap_ctx* ctx = nullptr;
    ap_vit_desc d{};
    int tokens = 0;   // patches per image (196)
    int lead = 1;     // rows in front of each image's patch tokens: class token + desc.registers register tokens
    int out_dim = 0;  // features per patch: hidden (class token) or 2 * hidden ([class || mean of patch tokens], desc.pool == 1)
    int kpe = 0;      // 3 * patch * patch
    int kpe_pad = 0;  // kpe rounded up to the GEMM's 64-wide K block (588 -> 640 for patch 14); pad columns stay zero
    // preprocess 1 (BitImageProcessorFast): tap tables of the input_patch -> resize_to antialias bicubic resize
    int32_t *tap_min = nullptr, *tap_cnt = nullptr, *tap_w = nullptr;
    int max_taps = 0, tap_precision = 0, max_src_rows = 0;
    int centre[3] = {0, 0, 0};  // integer pixel centre per channel = round(255 * mean_c)
    int max_batch = 0;
    int precise_layers = 1;
    bool finalized = false;
    std::unordered_map<std::string, std::vector<float>> host;  // staged fp32 tensors until finalize
    std::vector<void*> allocs;
    // packed weights
    __half* w_pe = nullptr;
    float *preln_g = nullptr, *preln_b = nullptr, *b_proj = nullptr;   // CLIP: LayerNorm after the embeddings; zero bias of the projection
    __half *w_proj = nullptr, *yc_proj = nullptr;                       // visual projection [proj_dim, 2 D] (hi | lo), its A operand [images, 2 D]
    GemmPlan p_proj;
    float *b_pe = nullptr, *cls = nullptr, *regs = nullptr, *pos = nullptr, *lnf_g = nullptr, *lnf_b = nullptr;
    GemmPlan p_pe;
    AttnPlan p_attn;
    bool attn_tc = false;
    std::vector<LayerWeights> layers;
    // workspaces
    __half *a_pe = nullptr, *y1 = nullptr, *y2 = nullptr, *qkv = nullptr, *hbuf = nullptr;
    __half *y1s = nullptr, *y2s = nullptr;   // [rows, 2 D]: LayerNorm / attention outputs as [hi | lo] pairs (A-split layers)
    float* x = nullptr;
    // class-token-only tail of the last layer
    __half *yc_attn = nullptr, *yc_ln = nullptr, *hc = nullptr;
    float* xc = nullptr;
    // host-patch path
    uint8_t *pin_in[2] = {nullptr, nullptr}, *dev_patches[2] = {nullptr, nullptr};
    float *pin_out[2] = {nullptr, nullptr}, *dev_feats[2] = {nullptr, nullptr};
    int32_t* dev_tall_coords = nullptr;
    // LayerNorm folding (see forward_chunk): per-row statistics partials written by the producers of x, unit gamma / zero beta for
    // the class-token tail; the fp16 copy of x lives in y1
    bool fold_ln = false;
    int ln_parts = 0;
    float2 *stats1 = nullptr, *stats2 = nullptr;
    float *ones = nullptr, *zeros = nullptr;
    // cv2.resize (INTER_LINEAR) tap tables per read size > input_patch, built on first use (device pointers)
    std::map<int, std::pair<int32_t*, int16_t*>> lin_tables;
    std::mutex lin_mu;
    cudaStream_t s_copy = nullptr, s_compute = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
};

namespace {

int dev_alloc(ap_encoder* e, void** p, size_t bytes) {
    cudaError_t err = cudaMalloc(p, bytes);
    if (err != cudaSuccess) return ap_set_error(e->ctx, AP_ENOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(err));
    e->allocs.push_back(*p);
    return AP_OK;
}

int upload_f32(ap_encoder* e, float** dst, const std::vector<float>& v) {
    int rc = dev_alloc(e, reinterpret_cast<void**>(dst), v.size() * sizeof(float));
    if (rc) return rc;
    AP_CHECK_CUDA(e->ctx, cudaMemcpy(*dst, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
    return AP_OK;
}

// [rows, K] fp32 -> [rows, 2K] fp16 = [hi | lo] with hi = fp16(w), lo = fp16(w - hi)
int upload_f16_split(ap_encoder* e, __half** dst, const float* src, size_t rows, size_t K) {
    std::vector<__half> h(rows * 2 * K);
    for (size_t r = 0; r < rows; ++r)
        for (size_t k = 0; k < K; ++k) {
            const float w = src[r * K + k];
            const __half hi = __float2half_rn(w);
            h[r * 2 * K + k] = hi;
            h[r * 2 * K + K + k] = __float2half_rn(w - __half2float(hi));
        }
    int rc = dev_alloc(e, reinterpret_cast<void**>(dst), h.size() * sizeof(__half));
    if (rc) return rc;
    AP_CHECK_CUDA(e->ctx, cudaMemcpy(*dst, h.data(), h.size() * sizeof(__half), cudaMemcpyHostToDevice));
    return AP_OK;
}

int upload_f16(ap_encoder* e, __half** dst, const float* src, size_t n) {
    std::vector<__half> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = __float2half_rn(src[i]);
    int rc = dev_alloc(e, reinterpret_cast<void**>(dst), n * sizeof(__half));
    if (rc) return rc;
    AP_CHECK_CUDA(e->ctx, cudaMemcpy(*dst, h.data(), n * sizeof(__half), cudaMemcpyHostToDevice));
    return AP_OK;
}

// LayerNorm(x) W^T + b = rstd (x - mean) (gamma o W)^T + (b + W beta).  Fold gamma / beta into the weights the LayerNorm feeds, and
// centre every row of gamma o W: with zero row sums, x W''^T = (x - mean 1) W'^T exactly, so the GEMM can read the raw x.
void fold_ln_affine(std::vector<float>& w, std::vector<float>& b, const std::vector<float>& gamma, const std::vector<float>& beta,
                    size_t rows, size_t K) {
    for (size_t r = 0; r < rows; ++r) {
        double acc = b[r], sum = 0.0;
        for (size_t k = 0; k < K; ++k) {
            acc += static_cast<double>(w[r * K + k]) * beta[k];
            sum += static_cast<double>(w[r * K + k]) * gamma[k];
        }
        const double mean = sum / static_cast<double>(K);
        for (size_t k = 0; k < K; ++k) w[r * K + k] = static_cast<float>(static_cast<double>(w[r * K + k]) * gamma[k] - mean);
        b[r] = static_cast<float>(acc);
    }
}

const std::vector<float>* find(ap_encoder* e, const std::string& name, size_t numel) {
    auto it = e->host.find(name);
    if (it == e->host.end()) {
        ap_set_error(e->ctx, AP_ESTATE, "encoder: tensor '%s' was never set", name.c_str());
        return nullptr;
    }
    if (it->second.size() != numel) {
        ap_set_error(e->ctx, AP_EINVAL, "encoder: tensor '%s' has %zu elements, expected %zu", name.c_str(), it->second.size(), numel);
        return nullptr;
    }
    return &it->second;
}

// Tables of the read_size -> input_patch resize (nullptrs for read_size == input_patch).
int get_linear_tables(ap_encoder* e, int read_size, const int32_t** taps, const int16_t** weights) {
    *taps = nullptr;
    *weights = nullptr;
    if (read_size == e->d.input_patch) return AP_OK;
    std::lock_guard<std::mutex> lk(e->lin_mu);
    auto it = e->lin_tables.find(read_size);
    if (it == e->lin_tables.end()) {
        std::vector<int32_t> t;
        std::vector<int16_t> w;
        int rc = ap_build_linear_tables(e->ctx, read_size, e->d.input_patch, t, w);
        if (rc) return rc;
        int32_t* dt = nullptr;
        int16_t* dw = nullptr;
        if ((rc = dev_alloc(e, (void**)&dt, t.size() * 4)) || (rc = dev_alloc(e, (void**)&dw, w.size() * 2))) return rc;
        AP_CHECK_CUDA(e->ctx, cudaMemcpy(dt, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
        AP_CHECK_CUDA(e->ctx, cudaMemcpy(dw, w.data(), w.size() * 2, cudaMemcpyHostToDevice));
        it = e->lin_tables.emplace(read_size, std::make_pair(dt, dw)).first;
    }
    *taps = it->second.first;
    *weights = it->second.second;
    return AP_OK;
}

// One forward chunk: nb images whose patches are at `coords` (device) inside `slide`.
int forward_chunk(ap_encoder* e, const uint8_t* slide, int64_t W, int64_t H, int64_t pitch, const int32_t* coords, int nb,
                  int read_size, float* out_feats, cudaStream_t st) {
    ap_ctx* ctx = e->ctx;
    const int D = e->d.hidden, T = e->tokens, T1 = T + e->lead;
    const int rows = nb * T1;
    int rc;
    const int32_t* lin_s = nullptr;
    const int16_t* lin_w = nullptr;
    if (e->d.preprocess >= 1) {
        AP_REQUIRE(ctx, read_size == e->d.input_patch, "encoder: reads larger than the patch are not implemented for the resizing preprocesses");
        if ((rc = ap_preprocess_resize_run(ctx, slide, W, H, pitch, coords, nb, e->d.input_patch, e->d.image_size, e->d.patch, e->tap_min,
                                           e->tap_cnt, e->tap_w, e->max_taps, e->tap_precision, e->max_src_rows, e->a_pe, e->kpe_pad,
                                           e->centre, st)))
            return rc;
    } else if ((rc = get_linear_tables(e, read_size, &lin_s, &lin_w)) ||
               (rc = ap_preprocess_run(ctx, slide, W, H, pitch, coords, nb, e->d.input_patch, e->d.image_size, e->d.patch, e->a_pe,
                                       e->kpe_pad, e->centre, 0, lin_s, lin_w, st)))
        return rc;
    // LayerNorm folding: x is normalised right before in_proj and mlp.0, and with W'' = row-centred (gamma o W), b' = b + W beta,
    //     LN(x) W^T + b = rstd (x W''^T) + b'
    // (zero row sums absorb the mean), so those two GEMMs read the RAW residual stream (its fp16 copy, written by whichever
    // epilogue produced x) and their epilogue only scales each row by 1 / sigma, computed from per-row (sum, sum of squares)
    // partials that the same producer epilogues emit.  24 of the 25 LayerNorm launches per chunk disappear together with their
    // 115 MB of traffic each.  Same number of operand roundings as before: the CPU simulation gives 5.5e-4 either way.
    const bool fold = e->fold_ln;
    GemmExtra prod1, prod2, cons;
    prod1.out_h = prod2.out_h = e->y1;
    prod1.stats_out = e->stats1;
    prod2.stats_out = e->stats2;
    prod1.ln_parts = prod2.ln_parts = cons.ln_parts = e->ln_parts;
    cons.ln_dim = D;
    cons.ln_eps = e->d.ln_eps;
    if ((rc = ap_cls_rows_run(ctx, e->x, e->cls, e->pos, e->regs, e->lead, nb, T1, D, fold ? e->y1 : nullptr, fold ? e->stats1 : nullptr, e->ln_parts, st)))
        return rc;
    {
        GemmPlan p = e->p_pe;
        p.M = nb * T;
        GemmExtra ex = fold ? prod1 : GemmExtra();
        ex.pos = e->pos;
        ex.tokens_per_image = T;
        ex.lead_tokens = e->lead;
        if ((rc = ap_gemm_run(ctx, &p, e->b_pe, nullptr, e->x, &ex, st))) return rc;
    }
    // CLIP (modeling_clip.py CLIPVisionTransformer: hidden = pre_layrnorm(embeddings)): the residual stream itself is normalised, in
    // place (each warp holds its row in registers); layer 0 then runs its own LayerNorm kernel (finalize: fold1 = false)
    if (e->d.pre_ln && (rc = ap_layernorm_run(ctx, e->x, D, e->preln_g, e->preln_b, e->d.ln_eps, nullptr, e->x, rows, D, st))) return rc;
    // final LayerNorm of the class rows (row stride xs) -> fp32 features, or -> [hi | lo] fp16 and through the bias-free visual projection
    // (CLIPModel.get_image_features: visual_projection(post_layernorm(hidden[:, 0])), three-product split GEMM = fp32-like precision)
    auto head = [&](const float* xsrc, int64_t xs) -> int {
        if (e->d.proj_dim <= 0) return ap_layernorm_run(ctx, xsrc, xs, e->lnf_g, e->lnf_b, e->d.ln_eps, nullptr, out_feats, nb, D, st);
        int r2 = ap_layernorm_run(ctx, xsrc, xs, e->lnf_g, e->lnf_b, e->d.ln_eps, e->yc_proj, nullptr, nb, D, st, 2 * D, 1);
        if (r2) return r2;
        GemmPlan pp = e->p_proj;
        pp.M = nb;
        return ap_gemm_run(ctx, &pp, e->b_proj, nullptr, out_feats, nullptr, st);
    };
    for (size_t li = 0; li < e->layers.size(); ++li) {
        auto& L = e->layers[li];
        GemmPlan p;
        if (!L.fold1 && (rc = (L.asplit & 1) ? ap_layernorm_run(ctx, e->x, D, L.ln1_g, L.ln1_b, e->d.ln_eps, e->y1s, nullptr, rows, D, st, 2 * D, 1)
                                          : ap_layernorm_run(ctx, e->x, D, L.ln1_g, L.ln1_b, e->d.ln_eps, e->y1, nullptr, rows, D, st)))
            return rc;
        p = L.p_qkv; p.M = rows;
        cons.stats_in = e->stats1;
        if ((rc = ap_gemm_run(ctx, &p, L.b_qkv, nullptr, e->qkv, L.fold1 ? &cons : nullptr, st))) return rc;
        if (li + 1 == e->layers.size() && ctx->cls_only_last_layer && e->d.pool == 0) {
            // Only x[:, 0] survives the final LayerNorm (models/patch/base.py:100 -> torchvision forward `x[:, 0]`): the last
            // layer's attention needs every token's K/V but only the class-token query, and out_proj / MLP only that row.
            if ((rc = ap_cls_attention_run(ctx, e->qkv, e->yc_attn, nb, T1, e->d.heads, 1, st))) return rc;
            if ((rc = ap_gather_rows_run(ctx, e->x, e->xc, nb, static_cast<int64_t>(T1) * D, D, st))) return rc;
            p = L.pc_o; p.M = nb;
            if ((rc = ap_gemm_run(ctx, &p, L.b_o, e->xc, e->xc, nullptr, st))) return rc;
            // with folding mlp.0 already carries gamma / beta: the tail's explicit LayerNorm only normalises
            if ((rc = ap_layernorm_run(ctx, e->xc, D, L.fold2 ? e->ones : L.ln2_g, L.fold2 ? e->zeros : L.ln2_b, e->d.ln_eps, e->yc_ln, nullptr, nb,
                                       D, st)))
                return rc;
            p = L.pc_1; p.M = nb;
            if ((rc = ap_gemm_run(ctx, &p, L.b_1, nullptr, e->hc, nullptr, st))) return rc;
            p = L.pc_2; p.M = nb;
            if ((rc = ap_gemm_run(ctx, &p, L.b_2, e->xc, e->xc, nullptr, st))) return rc;
            return head(e->xc, D);
        }
        if (L.asplit & 2) {   // out_proj reads [hi | lo] attention outputs (finalize only sets this bit with the tcgen05 kernel)
            if ((rc = ap_attention_tc_run(ctx, &e->p_attn, e->y2s, nb, T1, e->d.heads, st, 2 * D, 1))) return rc;
        } else if (e->attn_tc && ctx->attn_mode == 2) {
            if ((rc = ap_attention_tc_run(ctx, &e->p_attn, e->y2, nb, T1, e->d.heads, st))) return rc;
        } else if ((rc = ap_attention_run(ctx, e->qkv, e->y2, nb, T1, e->d.heads, st))) return rc;
        p = L.p_o; p.M = rows;
        if ((rc = ap_gemm_run(ctx, &p, L.b_o, e->x, e->x, fold ? &prod2 : nullptr, st))) return rc;
        if (!L.fold2 && (rc = (L.asplit & 4) ? ap_layernorm_run(ctx, e->x, D, L.ln2_g, L.ln2_b, e->d.ln_eps, e->y1s, nullptr, rows, D, st, 2 * D, 1)
                                          : ap_layernorm_run(ctx, e->x, D, L.ln2_g, L.ln2_b, e->d.ln_eps, e->y1, nullptr, rows, D, st)))
            return rc;
        p = L.p_1; p.M = rows;
        cons.stats_in = e->stats2;
        if ((rc = ap_gemm_run(ctx, &p, L.b_1, nullptr, e->hbuf, L.fold2 ? &cons : nullptr, st))) return rc;
        p = L.p_2; p.M = rows;
        if ((rc = ap_gemm_run(ctx, &p, L.b_2, e->x, e->x, fold ? &prod1 : nullptr, st))) return rc;
    }
    if (e->d.pool == 1)   // final LayerNorm on every token, then [class || mean of the patch tokens] (midnight.py:57-61)
        return ap_cls_mean_pool_run(ctx, e->x, nb, T1, e->lead, D, e->lnf_g, e->lnf_b, e->d.ln_eps, out_feats, st);
    // final LayerNorm on the class-token rows only -> fp32 features
    return head(e->x, static_cast<int64_t>(T1) * D);
}

}  // namespace

extern "C" int ap_encoder_create(ap_ctx* ctx, const ap_vit_desc* desc, ap_encoder** out_enc) {
    if (!ctx || !desc || !out_enc) return AP_EINVAL;
    DeviceGuard guard(ctx);
    *out_enc = nullptr;
    AP_REQUIRE(ctx, desc->preprocess >= 0 && desc->preprocess <= 4, "encoder: unknown preprocess %d", desc->preprocess);
    AP_REQUIRE(ctx, desc->registers >= 0 && desc->registers <= 8, "encoder: registers %d unsupported (0..8)", desc->registers);
    AP_REQUIRE(ctx, desc->pool == 0 || desc->pool == 1, "encoder: unknown pool %d (0 class token, 1 [class || mean of patch tokens])", desc->pool);
    AP_REQUIRE(ctx, desc->mlp_kind >= 0 && desc->mlp_kind <= 2, "encoder: unknown mlp_kind %d", desc->mlp_kind);
    AP_REQUIRE(ctx, desc->proj_dim >= 0 && desc->proj_dim % 128 == 0 && (desc->proj_dim == 0 || desc->pool == 0),
               "encoder: proj_dim %d unsupported (multiple of 128, class-token head only)", desc->proj_dim);
    AP_REQUIRE(ctx, desc->patch == 16 || desc->patch == 32 || (desc->preprocess >= 1 && desc->patch >= 4 && desc->patch <= 32),
               "encoder: conv patch %d unsupported (16 / 32 with the crop preprocess, 4..32 with the resizing preprocesses)", desc->patch);
    AP_REQUIRE(ctx, desc->preprocess == 0 || desc->resize_to >= desc->image_size, "encoder: resize_to %d < image_size %d", desc->resize_to,
               desc->image_size);
    AP_REQUIRE(ctx, desc->mlp_kind != 1 || (2 * desc->mlp) % 256 == 0, "encoder: SwiGLU needs 2 * mlp %% 256 == 0 (mlp %d)", desc->mlp);
    AP_REQUIRE(ctx, desc->image_size % desc->patch == 0, "encoder: image %d not a multiple of patch %d", desc->image_size, desc->patch);
    AP_REQUIRE(ctx, desc->hidden % desc->heads == 0 && desc->hidden / desc->heads == 64,
               "encoder: head_dim must be 64 (hidden %d, heads %d)", desc->hidden, desc->heads);
    AP_REQUIRE(ctx, desc->hidden % 128 == 0 && desc->mlp % 128 == 0, "encoder: hidden/mlp must be multiples of 128");
    AP_REQUIRE(ctx, desc->preprocess >= 1 || desc->input_patch >= desc->image_size,
               "encoder: input_patch %d < image_size %d (needs an up-sampling resize)", desc->input_patch, desc->image_size);
    AP_REQUIRE(ctx, desc->layers >= 1, "encoder: layers must be >= 1");
    ap_encoder* e = new ap_encoder();
    e->ctx = ctx;
    e->d = *desc;
    const int g = desc->image_size / desc->patch;
    e->tokens = g * g;
    e->out_dim = desc->proj_dim > 0 ? desc->proj_dim : desc->pool == 1 ? 2 * desc->hidden : desc->hidden;
    e->lead = 1 + desc->registers;
    e->kpe = 3 * desc->patch * desc->patch;
    e->kpe_pad = (e->kpe + 63) / 64 * 64;
    e->max_batch = desc->max_batch > 0 ? desc->max_batch : 127;
    // default: 1 layer with hi/lo split weights (DESIGN.md "precision").  Models deeper than 32 layers (DINOv2 giant: 40 layers of
    // fp16 roundings) get 8 leading layers in which the weights AND the A operands of qkv / out_proj / mlp.0 are split (three
    // products per term, ctx->precise_aw_layers): worst golden row 7.9e-4 (83 % black overhang patch), against 1.19e-3 with 20
    // weight-only layers and 9.1e-4 with 4 + 4 (gpurun_out/r02/giant_precision11.log -> profiles/r02_dinov2_giant_precision.log).
    // Every golden row has to be inside 1e-3 at the default.
    const int default_precise = desc->layers > 32 ? 8 : 1;
    e->precise_layers = desc->precise_layers < 0 ? default_precise : (desc->precise_layers > desc->layers ? desc->layers : desc->precise_layers);
    if ((e->tokens + e->lead) > 272) {
        delete e;
        return ap_set_error(ctx, AP_EINVAL, "encoder: sequence %d too long (<= 272)", g * g + 1 + desc->registers);
    }
    *out_enc = e;
    return AP_OK;
}

extern "C" int ap_encoder_destroy(ap_encoder* e) {
    if (!e) return AP_OK;
    DeviceGuard guard(e->ctx);
    cudaDeviceSynchronize();
    for (void* p : e->allocs) cudaFree(p);
    for (int i = 0; i < 2; ++i) {
        if (e->pin_in[i]) cudaFreeHost(e->pin_in[i]);
        if (e->pin_out[i]) cudaFreeHost(e->pin_out[i]);
        if (e->ev_h2d[i]) cudaEventDestroy(e->ev_h2d[i]);
        if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
    }
    if (e->s_copy) cudaStreamDestroy(e->s_copy);
    if (e->s_compute) cudaStreamDestroy(e->s_compute);
    delete e;
    return AP_OK;
}

extern "C" int ap_encoder_embedding_dim(const ap_encoder* e) { return e ? e->out_dim : 0; }

extern "C" int ap_encoder_set_tensor(ap_encoder* e, const char* name, const float* data_host, int64_t numel) {
    if (!e || !name || !data_host || numel <= 0) return AP_EINVAL;
    if (e->finalized) return ap_set_error(e->ctx, AP_ESTATE, "encoder: already finalized");
    e->host[name].assign(data_host, data_host + numel);
    return AP_OK;
}

extern "C" int ap_encoder_finalize(ap_encoder* e) {
    if (!e) return AP_EINVAL;
    DeviceGuard guard(e->ctx);
    ap_ctx* ctx = e->ctx;
    if (e->finalized) return AP_OK;
    const int D = e->d.hidden, M = e->d.mlp, P = e->d.patch, T = e->tokens, T1 = T + e->lead, K = e->kpe, Kp = e->kpe_pad;
    const int M1 = e->d.mlp_kind == 1 ? 2 * M : M;  // rows of mlp.0 (SwiGLU: gates and values)
    const int MB = e->max_batch;
    int rc;
#define AP_GET(var, name, n)                                   \
    const std::vector<float>* var = find(e, (name), (size_t)(n)); \
    if (!var) return AP_ESTATE;
    // ---- conv_proj with the preset's normalisation folded in -----------------------------------------
    // reference preprocess: x = (pixel/255 - mean_c) / std_c   ([tv]transforms/_presets.py:58-64)
    // kernel input:         a = (pixel - centre_c) / 256, centre_c = round(255 mean_c)   (exact in fp16)
    //   => W'[o,c,ky,kx] = W * 256 / (255 std_c)
    //      b'[o] = b[o] + sum_c (centre_c - 255 mean_c) / (255 std_c) * sum_{ky,kx} W[o,c,ky,kx]
    // conv_proj sees raw pixels, whose common mode is large against the signal, so fp16 rounding of W' alone costs
    // ~5.7e-4 of the 1e-3 feature budget (measured on the CPU simulation in DESIGN.md).  It is 0.7 % of the FLOPs,
    // so W' is kept as an fp16 hi/lo pair and the GEMM contracts over 2K = [W_hi | W_lo] while re-reading the same A
    // blocks (GemmPlan::Ka): ~22-bit weights.
    {
        AP_GET(w, "conv_proj.weight", (size_t)D * K)
        AP_GET(b, "conv_proj.bias", D)
        std::vector<float> wcat((size_t)D * 2 * Kp, 0.0f), bf(D);
        for (int c = 0; c < 3; ++c) e->centre[c] = static_cast<int>(lrint(255.0 * e->d.mean[c]));
        for (int o = 0; o < D; ++o) {
            double acc = (*b)[o];
            for (int c = 0; c < 3; ++c) {
                const double sc = 256.0 / (255.0 * e->d.std[c]);
                double s = 0.0;
                for (int i = 0; i < P * P; ++i) {
                    const float v = (*w)[(size_t)o * K + c * P * P + i];
                    const float wf = static_cast<float>(v * sc);
                    const float hi = __half2float(__float2half_rn(wf));
                    wcat[(size_t)o * 2 * Kp + c * P * P + i] = hi;
                    wcat[(size_t)o * 2 * Kp + Kp + c * P * P + i] = wf - hi;
                    s += v;
                }
                acc += s * (e->centre[c] - 255.0 * e->d.mean[c]) / (255.0 * e->d.std[c]);
            }
            bf[o] = static_cast<float>(acc);
        }
        if ((rc = upload_f16(e, &e->w_pe, wcat.data(), wcat.size()))) return rc;
        if ((rc = upload_f32(e, &e->b_pe, bf))) return rc;
    }
    {
        AP_GET(c, "class_token", D)
        AP_GET(p, "encoder.pos_embedding", (size_t)(T + 1) * D)   // class + patch positions; register tokens carry none
        AP_GET(g, "encoder.ln.weight", D)
        AP_GET(b, "encoder.ln.bias", D)
        if ((rc = upload_f32(e, &e->cls, *c))) return rc;
        if (e->d.pre_ln) {
            AP_GET(pg, "encoder.pre_ln.weight", D)
            AP_GET(pb, "encoder.pre_ln.bias", D)
            if ((rc = upload_f32(e, &e->preln_g, *pg)) || (rc = upload_f32(e, &e->preln_b, *pb))) return rc;
        }
        if (e->d.proj_dim > 0) {
            AP_GET(wp, "head.proj.weight", (size_t)e->d.proj_dim * D)
            std::vector<float> zb(e->d.proj_dim, 0.0f);
            if ((rc = upload_f16_split(e, &e->w_proj, wp->data(), e->d.proj_dim, D)) || (rc = upload_f32(e, &e->b_proj, zb))) return rc;
        }
        if (e->d.registers > 0) {
            AP_GET(r, "register_tokens", (size_t)e->d.registers * D)
            if ((rc = upload_f32(e, &e->regs, *r))) return rc;
        }
        if ((rc = upload_f32(e, &e->pos, *p))) return rc;
        if ((rc = upload_f32(e, &e->lnf_g, *g))) return rc;
        if ((rc = upload_f32(e, &e->lnf_b, *b))) return rc;
    }
    // LayerNorm folding is decided per layer and per LayerNorm: a GEMM whose A operand is split into [hi | lo] needs the LayerNorm kernel
    // that writes that pair, every other in_proj / mlp.0 reads the raw residual stream.  0 = off, 1 (default) = automatic, 2 = on.
    // Automatic (1) leaves encoders deeper than 32 layers unfolded as before: measured on the giant, folding its 32 plain layers trades
    // 27 ms of LayerNorm kernels for 59 ms of heavier producer / consumer epilogues per 1 016 patches (1 370 -> 1 375 patches/s: nothing).
    e->fold_ln = ctx->fold_ln == 2 || (ctx->fold_ln == 1 && e->d.layers <= 32);
    e->ln_parts = D / ((D % 256 == 0 ? 256 : 128) / 2);   // one partial per bn/2-column block of the N = D producer GEMMs
    e->layers.resize(e->d.layers);
    for (int i = 0; i < e->d.layers; ++i) {
        const std::string p = "encoder.layers.encoder_layer_" + std::to_string(i) + ".";
        LayerWeights& L = e->layers[i];
        AP_GET(ln1g, p + "ln_1.weight", D) AP_GET(ln1b, p + "ln_1.bias", D)
        AP_GET(ln2g, p + "ln_2.weight", D) AP_GET(ln2b, p + "ln_2.bias", D)
        AP_GET(wqkv, p + "self_attention.in_proj_weight", (size_t)3 * D * D) AP_GET(bqkv, p + "self_attention.in_proj_bias", 3 * D)
        AP_GET(wo, p + "self_attention.out_proj.weight", (size_t)D * D) AP_GET(bo, p + "self_attention.out_proj.bias", D)
        AP_GET(w1, p + "mlp.0.weight", (size_t)M1 * D) AP_GET(b1, p + "mlp.0.bias", M1)
        AP_GET(w2, p + "mlp.3.weight", (size_t)D * M) AP_GET(b2, p + "mlp.3.bias", D)
        // precise layers: W-split everywhere (kind 0), or -- where the activations' rounding dominates the error, i.e. the deep
        // DINOv2 giant (DESIGN.md "precision": per layer the A operands cost ~4x the weights' share of the squared error) -- the
        // A operands of in_proj / out_proj / mlp.0 as [hi | lo] pairs and only mlp.3's weights split (kind 1).  A-split needs the
        // LayerNorm kernels (not the folded form) and the tcgen05 attention epilogue; the class-token tail of the last layer keeps
        // single operands.
        const bool a_kind = ctx->precise_kind == 1;
        const bool a_ok = (T1 <= 257) && ctx->attn_mode == 2 && !(i + 1 == e->d.layers && ctx->cls_only_last_layer && e->d.pool == 0);
        const int pm = ctx->precise_mask & 15;
        if (a_kind) {            // A operands of qkv / out_proj / mlp.0 split, weights of mlp.3 split
            L.asplit = (i < e->precise_layers && a_ok) ? (pm & 7) : 0;
            L.split = i < e->precise_layers ? (pm & ~L.asplit) : 0;
        } else {                 // weights split; in the first precise_aw_layers layers the A operands as well (three products per term)
            L.split = i < e->precise_layers ? pm : 0;
            const int aw = ctx->precise_aw_layers >= 0 ? ctx->precise_aw_layers : (e->d.layers > 32 ? 8 : 0);
            L.asplit = (i < e->precise_layers && i < aw && a_ok) ? (pm & 7) : 0;
        }
        L.fold1 = e->fold_ln && !(L.asplit & 1) && !(i == 0 && e->d.pre_ln);   // after the in-place pre-LayerNorm the producers' fp16 copy / statistics of x are stale
        L.fold2 = e->fold_ln && !(L.asplit & 4);
        std::vector<float> wq(*wqkv), bq(*bqkv), wf(*w1), bf1(*b1);
        if (L.fold1) fold_ln_affine(wq, bq, *ln1g, *ln1b, (size_t)3 * D, D);
        if (L.fold2) fold_ln_affine(wf, bf1, *ln2g, *ln2b, (size_t)M1, D);
        if ((rc = upload_f32(e, &L.ln1_g, *ln1g)) || (rc = upload_f32(e, &L.ln1_b, *ln1b)) ||
            (rc = upload_f32(e, &L.ln2_g, *ln2g)) || (rc = upload_f32(e, &L.ln2_b, *ln2b)) ||
            (rc = upload_f32(e, &L.b_qkv, bq)) || (rc = upload_f32(e, &L.b_o, *bo)) ||
            (rc = upload_f32(e, &L.b_1, bf1)) || (rc = upload_f32(e, &L.b_2, *b2)))
            return rc;
        auto up = [&](__half** dst, const float* w, size_t n_elems, size_t rows_w, size_t k, int bit) {
            return (L.split & bit) ? upload_f16_split(e, dst, w, rows_w, k) : upload_f16(e, dst, w, n_elems);
        };
        if ((rc = up(&L.w_qkv, wq.data(), wq.size(), 3 * D, D, 1)) || (rc = up(&L.w_o, wo->data(), wo->size(), D, D, 2)) ||
            (rc = up(&L.w_1, wf.data(), wf.size(), M1, D, 4)) || (rc = up(&L.w_2, w2->data(), w2->size(), D, M, 8)))
            return rc;
    }
#undef AP_GET
    e->host.clear();

    // ---- workspaces (rows padded to the 128-row GEMM tile so TMA boxes never leave the allocation) ----
    const size_t rows = ((size_t)MB * T1 + 127) / 128 * 128;
    const size_t rows_pe = ((size_t)MB * T + 127) / 128 * 128;
    if ((rc = dev_alloc(e, (void**)&e->a_pe, rows_pe * Kp * 2)) || (rc = dev_alloc(e, (void**)&e->x, rows * D * 4)) ||
        (rc = dev_alloc(e, (void**)&e->y1, rows * D * 2)) || (rc = dev_alloc(e, (void**)&e->y2, rows * D * 2)) ||
        (rc = dev_alloc(e, (void**)&e->qkv, rows * 3 * D * 2)) || (rc = dev_alloc(e, (void**)&e->hbuf, rows * M * 2)))
        return rc;
    bool any_asplit = false;
    for (auto& L : e->layers) any_asplit |= L.asplit != 0;
    if (any_asplit) {
        if ((rc = dev_alloc(e, (void**)&e->y1s, rows * 2 * D * 2)) || (rc = dev_alloc(e, (void**)&e->y2s, rows * 2 * D * 2))) return rc;
        AP_CHECK_CUDA(ctx, cudaMemset(e->y1s, 0, rows * 2 * D * 2));
        AP_CHECK_CUDA(ctx, cudaMemset(e->y2s, 0, rows * 2 * D * 2));
    }
    {
        std::vector<float> ones(D, 1.0f), zeros(D, 0.0f);
        if ((rc = upload_f32(e, &e->ones, ones)) || (rc = upload_f32(e, &e->zeros, zeros)) ||
            (rc = dev_alloc(e, (void**)&e->stats1, rows * e->ln_parts * sizeof(float2))) ||
            (rc = dev_alloc(e, (void**)&e->stats2, rows * e->ln_parts * sizeof(float2))))
            return rc;
        AP_CHECK_CUDA(ctx, cudaMemset(e->stats1, 0, rows * e->ln_parts * sizeof(float2)));
        AP_CHECK_CUDA(ctx, cudaMemset(e->stats2, 0, rows * e->ln_parts * sizeof(float2)));
    }
    const size_t rows_c = ((size_t)MB + 127) / 128 * 128;
    if ((rc = dev_alloc(e, (void**)&e->yc_attn, rows_c * D * 2)) || (rc = dev_alloc(e, (void**)&e->yc_ln, rows_c * D * 2)) ||
        (rc = dev_alloc(e, (void**)&e->hc, rows_c * M * 2)) || (rc = dev_alloc(e, (void**)&e->xc, rows_c * D * 4)))
        return rc;
    if (e->d.proj_dim > 0) {
        if ((rc = dev_alloc(e, (void**)&e->yc_proj, rows_c * 2 * D * 2))) return rc;
        AP_CHECK_CUDA(ctx, cudaMemset(e->yc_proj, 0, rows_c * 2 * D * 2));
    }
    AP_CHECK_CUDA(ctx, cudaMemset(e->yc_attn, 0, rows_c * D * 2));
    AP_CHECK_CUDA(ctx, cudaMemset(e->yc_ln, 0, rows_c * D * 2));
    AP_CHECK_CUDA(ctx, cudaMemset(e->hc, 0, rows_c * M * 2));
    AP_CHECK_CUDA(ctx, cudaMemset(e->xc, 0, rows_c * D * 4));
    AP_CHECK_CUDA(ctx, cudaMemset(e->a_pe, 0, rows_pe * Kp * 2));
    AP_CHECK_CUDA(ctx, cudaMemset(e->y1, 0, rows * D * 2));
    AP_CHECK_CUDA(ctx, cudaMemset(e->y2, 0, rows * D * 2));
    AP_CHECK_CUDA(ctx, cudaMemset(e->hbuf, 0, rows * M * 2));
    AP_CHECK_CUDA(ctx, cudaMemset(e->qkv, 0, rows * 3 * D * 2));
    AP_CHECK_CUDA(ctx, cudaMemset(e->x, 0, rows * D * 4));

    // ---- GEMM plans (TMA descriptors over the fixed workspaces / weights) ------------------------------
    if ((rc = ap_gemm_plan(ctx, &e->p_pe, e->a_pe, e->w_pe, MB * T, D, 2 * Kp, AP_EPI_BIAS_F32, Kp))) return rc;
    if (e->d.proj_dim > 0 && (rc = ap_gemm_plan_split(ctx, &e->p_proj, e->yc_proj, e->w_proj, MB, e->d.proj_dim, D, AP_EPI_BIAS_F32, AP_SPLIT_AW))) return rc;
    const int epi1 = e->d.mlp_kind == 1 ? AP_EPI_BIAS_SWIGLU_F16 : e->d.mlp_kind == 2 ? AP_EPI_BIAS_QGELU_F16 : AP_EPI_BIAS_GELU_F16;
    for (auto& L : e->layers) {
        const int sq = (L.split & 1) ? 2 : 1, so = (L.split & 2) ? 2 : 1, s1 = (L.split & 4) ? 2 : 1, s2 = (L.split & 8) ? 2 : 1;
        auto mode = [&](int bit) { return ((L.asplit & bit) ? AP_SPLIT_A : 0) | ((L.split & bit) ? AP_SPLIT_W : 0); };
        if ((rc = (L.asplit & 1) ? ap_gemm_plan_split(ctx, &L.p_qkv, e->y1s, L.w_qkv, MB * T1, 3 * D, D, AP_EPI_BIAS_F16, mode(1))
                                 : ap_gemm_plan(ctx, &L.p_qkv, e->y1, L.w_qkv, MB * T1, 3 * D, sq * D, AP_EPI_BIAS_F16, D)) ||
            (rc = (L.asplit & 2) ? ap_gemm_plan_split(ctx, &L.p_o, e->y2s, L.w_o, MB * T1, D, D, AP_EPI_BIAS_RESID_F32, mode(2))
                                 : ap_gemm_plan(ctx, &L.p_o, e->y2, L.w_o, MB * T1, D, so * D, AP_EPI_BIAS_RESID_F32, D)) ||
            (rc = (L.asplit & 4) ? ap_gemm_plan_split(ctx, &L.p_1, e->y1s, L.w_1, MB * T1, M1, D, epi1, mode(4))
                                 : ap_gemm_plan(ctx, &L.p_1, e->y1, L.w_1, MB * T1, M1, s1 * D, epi1, D)) ||
            (rc = ap_gemm_plan(ctx, &L.p_2, e->hbuf, L.w_2, MB * T1, D, s2 * M, AP_EPI_BIAS_RESID_F32, M)))
            return rc;
        if (&L == &e->layers.back() &&
            ((rc = ap_gemm_plan(ctx, &L.pc_o, e->yc_attn, L.w_o, MB, D, so * D, AP_EPI_BIAS_RESID_F32, D)) ||
             (rc = ap_gemm_plan(ctx, &L.pc_1, e->yc_ln, L.w_1, MB, M1, s1 * D, epi1, D)) ||
             (rc = ap_gemm_plan(ctx, &L.pc_2, e->hc, L.w_2, MB, D, s2 * M, AP_EPI_BIAS_RESID_F32, M))))
            return rc;
    }

    {
        e->attn_tc = T1 <= 257;   // 257 (DINOv2 @ 224): 256 patch keys through the MMA + the class token as a rank-1 extra key
        if (e->attn_tc && (rc = ap_attention_tc_plan(ctx, &e->p_attn, e->qkv, (int)rows, T1, e->d.heads))) return rc;
    }

    // ---- tap tables of the resizing preprocess ----------------------------------------------------------------
    if (e->d.preprocess >= 1) {
        std::vector<int32_t> tmin, tcnt, tw;
        if ((rc = ap_build_resize_tables(ctx, e->d.input_patch, e->d.resize_to, e->d.image_size, e->d.preprocess - 1, tmin, tcnt, tw,
                                         &e->max_taps, &e->tap_precision)))
            return rc;
        for (int tr = 0; tr < e->d.image_size / P; ++tr) {
            const int last = tr * P + P - 1;
            const int nr = tmin[last] + tcnt[last] - tmin[tr * P];
            if (nr > e->max_src_rows) e->max_src_rows = nr;
        }
        if ((rc = dev_alloc(e, (void**)&e->tap_min, tmin.size() * 4)) || (rc = dev_alloc(e, (void**)&e->tap_cnt, tcnt.size() * 4)) ||
            (rc = dev_alloc(e, (void**)&e->tap_w, tw.size() * 4)))
            return rc;
        AP_CHECK_CUDA(ctx, cudaMemcpy(e->tap_min, tmin.data(), tmin.size() * 4, cudaMemcpyHostToDevice));
        AP_CHECK_CUDA(ctx, cudaMemcpy(e->tap_cnt, tcnt.data(), tcnt.size() * 4, cudaMemcpyHostToDevice));
        AP_CHECK_CUDA(ctx, cudaMemcpy(e->tap_w, tw.data(), tw.size() * 4, cudaMemcpyHostToDevice));
    }

    // ---- host-patch path: double-buffered pinned staging + its own streams -------------------------------
    const size_t IP = e->d.input_patch;
    const size_t patch_bytes = IP * IP * 3;
    for (int i = 0; i < 2; ++i) {
        AP_CHECK_CUDA(ctx, cudaMallocHost((void**)&e->pin_in[i], (size_t)MB * patch_bytes));
        AP_CHECK_CUDA(ctx, cudaMallocHost((void**)&e->pin_out[i], (size_t)MB * e->out_dim * 4));
        if ((rc = dev_alloc(e, (void**)&e->dev_patches[i], (size_t)MB * patch_bytes)) ||
            (rc = dev_alloc(e, (void**)&e->dev_feats[i], (size_t)MB * e->out_dim * 4)))
            return rc;
        AP_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&e->ev_h2d[i], cudaEventDisableTiming));
        AP_CHECK_CUDA(ctx, cudaEventCreateWithFlags(&e->ev_done[i], cudaEventDisableTiming));
    }
    AP_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&e->s_copy, cudaStreamNonBlocking));
    AP_CHECK_CUDA(ctx, cudaStreamCreateWithFlags(&e->s_compute, cudaStreamNonBlocking));
    {
        std::vector<int32_t> tall((size_t)MB * 5);
        for (int i = 0; i < MB; ++i) {
            tall[i * 5 + 0] = 0; tall[i * 5 + 1] = i * (int)IP; tall[i * 5 + 2] = (int)IP; tall[i * 5 + 3] = (int)IP; tall[i * 5 + 4] = 0;
        }
        if ((rc = dev_alloc(e, (void**)&e->dev_tall_coords, tall.size() * 4))) return rc;
        AP_CHECK_CUDA(ctx, cudaMemcpy(e->dev_tall_coords, tall.data(), tall.size() * 4, cudaMemcpyHostToDevice));
    }
    AP_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    e->finalized = true;
    return AP_OK;
}

extern "C" int ap_encoder_embed_coords(ap_encoder* e, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch,
                                       const int32_t* coords_dev, int64_t n, int read_size, float* out_features_dev, void* stream) {
    if (!e) return AP_EINVAL;
    DeviceGuard guard(e->ctx);
    ap_ctx* ctx = e->ctx;
    if (!e->finalized) return ap_set_error(ctx, AP_ESTATE, "encoder: ap_encoder_finalize has not been called");
    AP_REQUIRE(ctx, n >= 0, "embed_coords: n < 0");
    if (n == 0) return AP_OK;
    AP_REQUIRE(ctx, slide_dev && coords_dev && out_features_dev, "embed_coords: NULL pointer");
    AP_REQUIRE(ctx, pitch >= W * 3, "embed_coords: pitch %lld < 3*W", (long long)pitch);
    AP_REQUIRE(ctx, read_size >= e->d.input_patch, "embed_coords: read size %d is smaller than the patch size %d (up-sampling reads do not occur: "
               "the reference rejects target magnifications above the slide's)", read_size, e->d.input_patch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    for (int64_t s = 0; s < n; s += e->max_batch) {
        const int nb = static_cast<int>(n - s < e->max_batch ? n - s : e->max_batch);
        int rc = forward_chunk(e, slide_dev, W, H, pitch, coords_dev + s * 5, nb, read_size, out_features_dev + s * e->out_dim, st);
        if (rc) return rc;
    }
    return AP_OK;
}

// a12 alone (parity tests): the fp16 im2col rows the patch-embedding GEMM consumes, out[n * tokens, kpe_pad] (device).
extern "C" int ap_encoder_preprocess(ap_encoder* e, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch,
                                     const int32_t* coords_dev, int64_t n, int read_size, void* out_dev, int64_t* out_cols, void* stream) {
    if (!e) return AP_EINVAL;
    DeviceGuard guard(e->ctx);
    ap_ctx* ctx = e->ctx;
    if (!e->finalized) return ap_set_error(ctx, AP_ESTATE, "encoder: ap_encoder_finalize has not been called");
    if (out_cols) *out_cols = e->kpe_pad;
    AP_REQUIRE(ctx, n >= 0 && n <= e->max_batch, "encoder_preprocess: n=%lld must be in [0, max_batch=%d]", (long long)n, e->max_batch);
    if (n == 0) return AP_OK;
    AP_REQUIRE(ctx, slide_dev && coords_dev && out_dev, "encoder_preprocess: NULL pointer");
    AP_REQUIRE(ctx, read_size == e->d.input_patch || (e->d.preprocess == 0 && read_size > e->d.input_patch),
               "encoder_preprocess: read size %d unsupported for patch size %d", read_size, e->d.input_patch);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (e->d.preprocess >= 1)
        rc = ap_preprocess_resize_run(ctx, slide_dev, W, H, pitch, coords_dev, n, e->d.input_patch, e->d.image_size, e->d.patch, e->tap_min,
                                      e->tap_cnt, e->tap_w, e->max_taps, e->tap_precision, e->max_src_rows, e->a_pe, e->kpe_pad, e->centre, st);
    else {
        const int32_t* lin_s = nullptr;
        const int16_t* lin_w = nullptr;
        if ((rc = get_linear_tables(e, read_size, &lin_s, &lin_w))) return rc;
        rc = ap_preprocess_run(ctx, slide_dev, W, H, pitch, coords_dev, n, e->d.input_patch, e->d.image_size, e->d.patch, e->a_pe,
                               e->kpe_pad, e->centre, 0, lin_s, lin_w, st);
    }
    if (rc) return rc;
    AP_CHECK_CUDA(ctx, cudaMemcpyAsync(out_dev, e->a_pe, (size_t)n * e->tokens * e->kpe_pad * 2, cudaMemcpyDeviceToDevice, st));
    return AP_OK;
}

extern "C" int ap_encoder_embed_patches_host(ap_encoder* e, const uint8_t* const* patches_host, int64_t n,
                                             float* out_features_host) {
    if (!e) return AP_EINVAL;
    DeviceGuard guard(e->ctx);
    ap_ctx* ctx = e->ctx;
    if (!e->finalized) return ap_set_error(ctx, AP_ESTATE, "encoder: ap_encoder_finalize has not been called");
    AP_REQUIRE(ctx, n >= 0, "embed_patches_host: n < 0");
    if (n == 0) return AP_OK;
    AP_REQUIRE(ctx, patches_host && out_features_host, "embed_patches_host: NULL pointer");
    const int MB = e->max_batch, D = e->out_dim;
    const size_t IP = e->d.input_patch, patch_bytes = IP * IP * 3;
    // Chunk schedule: the first chunk's gather + H2D is not overlapped with anything, so with large workspaces the pipeline ramps up:
    // 127 patches, then max_batch - 127 (together one full chunk, so what follows stays aligned to max_batch), then full chunks.  With
    // 508-patch chunks the exposed prologue of a call falls from ~4.5 ms to ~1.5 ms.
    std::vector<int64_t> starts;
    std::vector<int> sizes;
    {
        int64_t s0 = 0;
        int idx = 0;
        while (s0 < n) {
            int nb = MB;
            if (MB >= 254 && idx == 0) nb = 127;
            else if (MB >= 254 && idx == 1) nb = MB - 127;
            if (n - s0 < nb) nb = static_cast<int>(n - s0);
            starts.push_back(s0);
            sizes.push_back(nb);
            s0 += nb;
            ++idx;
        }
    }
    const int64_t n_chunks = static_cast<int64_t>(starts.size());
    // Pipeline over chunks with two buffers: host gather (CPU threads) | H2D (copy stream) | forward (compute stream)
    // | D2H (copy stream).  Buffer i%2 is reused only after chunk i-2's features have been copied out.
    for (int64_t c = 0; c < n_chunks + 1; ++c) {
        if (c < n_chunks) {
            const int buf = static_cast<int>(c & 1);
            const int64_t s = starts[c];
            const int nb = sizes[c];
            if (c >= 2) AP_CHECK_CUDA(ctx, cudaEventSynchronize(e->ev_done[buf]));  // chunk c-2 fully drained
            if (c >= 2) memcpy(out_features_host + starts[c - 2] * D, e->pin_out[buf], (size_t)sizes[c - 2] * D * 4);
            {   // gather the (possibly scattered) host patches into pinned memory
                // 100 MB per 508-patch chunk: up to 8 threads, but never more than this process' share of the host cores (one process
                // per GPU: torchrun exports LOCAL_WORLD_SIZE)
                static const int max_threads = [] {
                    const char* lw = getenv("LOCAL_WORLD_SIZE");
                    const int ranks = lw && atoi(lw) > 0 ? atoi(lw) : 1;
                    const int hc = static_cast<int>(std::thread::hardware_concurrency());
                    const int share = hc > 0 ? hc / ranks : 4;
                    return share < 1 ? 1 : share > 8 ? 8 : share;
                }();
                const int nthreads = nb >= 256 ? max_threads : nb >= 32 ? (max_threads < 4 ? max_threads : 4) : 1;
                auto work = [&](int t) {
                    for (int i = t; i < nb; i += nthreads) memcpy(e->pin_in[buf] + (size_t)i * patch_bytes, patches_host[s + i], patch_bytes);
                };
                if (nthreads == 1) work(0);
                else {
                    std::vector<std::thread> th;
                    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, t);
                    for (auto& t : th) t.join();
                }
            }
            AP_CHECK_CUDA(ctx, cudaMemcpyAsync(e->dev_patches[buf], e->pin_in[buf], (size_t)nb * patch_bytes, cudaMemcpyHostToDevice, e->s_copy));
            AP_CHECK_CUDA(ctx, cudaEventRecord(e->ev_h2d[buf], e->s_copy));
            AP_CHECK_CUDA(ctx, cudaStreamWaitEvent(e->s_compute, e->ev_h2d[buf], 0));
            int rc = forward_chunk(e, e->dev_patches[buf], (int64_t)IP, (int64_t)IP * nb, (int64_t)IP * 3, e->dev_tall_coords, nb, (int)IP,
                                   e->dev_feats[buf], e->s_compute);
            if (rc) return rc;
            AP_CHECK_CUDA(ctx, cudaMemcpyAsync(e->pin_out[buf], e->dev_feats[buf], (size_t)nb * D * 4, cudaMemcpyDeviceToHost, e->s_compute));
            AP_CHECK_CUDA(ctx, cudaEventRecord(e->ev_done[buf], e->s_compute));
        }
    }
    // drain the last (up to) two chunks in order
    for (int64_t c = (n_chunks >= 2 ? n_chunks - 2 : 0); c < n_chunks; ++c) {
        const int buf = static_cast<int>(c & 1);
        AP_CHECK_CUDA(ctx, cudaEventSynchronize(e->ev_done[buf]));
        memcpy(out_features_host + starts[c] * D, e->pin_out[buf], (size_t)sizes[c] * D * 4);
    }
    return AP_OK;
}
