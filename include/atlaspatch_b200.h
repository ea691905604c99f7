/*
 * atlaspatch_b200 -- C ABI of the B200-native AtlasPatch hot path (libatlaspatch_b200.so).
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  Every entry point returns 0 on
 * success or a negative AP_E* code; the message is available from ap_last_error().  Device
 * pointers are ordinary CUDA device pointers of the ctx's device (e.g. torch `tensor.data_ptr()`),
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  The library owns
 * only encoder weights and workspaces; all inputs/outputs are caller-allocated.
 *
 * Each entry cites the reference interface it replaces (paths under the AtlasPatch repo).
 * The reference is pure Python and has no FFI of its own; INTEGRATION.md shows the ctypes
 * binding a maintainer adds at each of these seams.
 */
#ifndef ATLASPATCH_B200_H
#define ATLASPATCH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AP_OK 0
#define AP_EINVAL (-1)      /* bad argument / unsupported configuration */
#define AP_ECUDA (-2)       /* CUDA runtime or driver error */
#define AP_ENOMEM (-3)      /* allocation failed */
#define AP_ECAPACITY (-4)   /* caller-provided output buffer too small */
#define AP_ESTATE (-5)      /* object not ready (e.g. encoder weights missing) */

typedef struct ap_ctx ap_ctx;
typedef struct ap_encoder ap_encoder;

/* ---- context ------------------------------------------------------------------------- */
int ap_version(void);
/* sizeof("ap_vit_desc" | "ap_sam2_desc") as compiled into the library (-1: unknown name): lets a binding (ctypes, cgo ...)
 * assert that its own struct declaration matches before it passes one in. */
int ap_sizeof(const char* struct_name);
/* One ctx per GPU/process (SURVEY.md section 8b "Threading").  Fails (AP_ECUDA) when no sm_100 device. */
int ap_init(int device, ap_ctx** out_ctx);
int ap_destroy(ap_ctx* ctx);
/* Last error text for this ctx (ctx==NULL: last error of a failed ap_init).  Never NULL. */
const char* ap_last_error(const ap_ctx* ctx);
/* Number of kernels this library has launched on ctx since ap_init (bench.py "gpu_launches"). */
int64_t ap_launch_count(const ap_ctx* ctx);
int ap_sm_count(const ap_ctx* ctx);
/* Tuning knobs.  "gemm_cta_group": 2 (default; CTA-pair 256x256 tiles, tcgen05 cta_group::2) or 1 (128x256 tiles).
 * Affects GEMM plans created afterwards. */
int ap_set_option(ap_ctx* ctx, const char* key, int value);
/* Optional per-launch timing with CUDA events recorded on the launching stream (bench.py's live roofline
 * numbers).  Kernel classes: 0 gemm, 1 attention, 2 layernorm, 3 preprocess, 4 coords, 5 thumbnail, 6 other.
 * ap_profile_enable: on = 0 off, -1 all classes, else a bit mask (bit c = class c).
 * ap_profile_read sums and clears the records into total_ms[n_classes] / counts[n_classes] (n_classes >= 7). */
int ap_profile_enable(ap_ctx* ctx, int on);
int ap_profile_read(ap_ctx* ctx, double* total_ms, int64_t* counts, int n_classes);
/* The same for one class, split by launch tag (GEMM launches: (N << 32) | (K << 4) | epilogue); clears all records. */
int ap_profile_read_tagged(ap_ctx* ctx, int cls, int64_t* keys, double* total_ms, int64_t* counts, int cap, int* n_out);

/* ---- synthetic slide (benchmark input) ---------------------------------
 * Renders the region [x0,x0+w) x [y0,y0+h) of the synthetic slide (W x H, seed, blobs, holes)
 * into `out` (RGB uint8, HWC, row pitch `pitch` bytes).  blobs: n_blobs x 6 int32
 * (cx,cy,a,b,c,s), holes: n_holes x 3 int32 (cx,cy,r) -- HOST pointers.  Pixels outside the
 * slide are 0.  Same integer function as atlaspatch_b200/synthetic.py:render_region_host. */
int ap_synth_render(ap_ctx* ctx, uint8_t* out_dev, int64_t pitch, int64_t W, int64_t H, uint32_t seed,
                    const int32_t* blobs_host, int n_blobs, const int32_t* holes_host, int n_holes,
                    int64_t x0, int64_t y0, int64_t w, int64_t h, void* stream);

/* ---- a1: whole-level thumbnail -------------------------------------------------------------
 * Replaces IWSI.get_thumbnail_at_power's whole-level read + cv2.resize(INTER_AREA)
 * (atlas_patch/core/wsi/iwsi.py:293-321) for an integer factor f = base_mag/power with
 * W % f == H % f == 0: out[(H/f),(W/f),3] = round-half-even(mean of f x f block).  out is tightly
 * packed RGB.  One pass over W*H*3 bytes: the HBM-roofline kernel of the path. */
int ap_thumbnail_area(ap_ctx* ctx, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch,
                      int factor, uint8_t* out_dev, void* stream);
/* The general case of the same reference step: the reference's output size is round(W/ds) x round(H/ds)
 * (core/wsi/iwsi.py:302-303); when the level size is not a multiple of ds, cv2.resize(INTER_AREA) leaves its
 * integer-factor path and weights partial source cells in float32 (computeResizeAreaTab / ResizeArea_).  This entry
 * reproduces cv2.resize(level, (out_w, out_h), INTER_AREA) bit for bit for any down-scale and dispatches to the
 * integer-factor kernels when both scale factors are integral. */
int ap_thumbnail_resize(ap_ctx* ctx, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch,
                        int out_w, int out_h, uint8_t* out_dev, void* stream);


/* ---- a9: patch-coordinate extraction -------------------------------------------------------
 * Replaces PatchExtractionService._iter_patch_entries (fast mode) + _in_tissue +
 * FourPointContainment (atlas_patch/services/extraction.py:67-103, utils/contours.py:22-38).
 * Inputs are the level-0 contours produced by _prepare_contours (extraction.py:30-42), flattened:
 *   contour_xy      int32 (sum K_c) x 2         vertices of all tissue contours, in order
 *   contour_offsets int32 n_contours+1          vertex ranges
 *   hole_xy         int32 (sum K_h) x 2         vertices of all hole contours
 *   hole_offsets    int32 n_holes+1             vertex ranges
 *   hole_first      int32 n_contours+1          holes of contour c are [hole_first[c], hole_first[c+1])
 * (all HOST pointers; they are a few KB).  patch_src/step_src/read_w/read_h/level come from
 * _prepare_geometry (extraction.py:44-64).
 * Output rows (x, y, read_w, read_h, level) int32, in the reference's order (contour-major, y, x),
 * are written to out_rows_dev (device, capacity rows) and/or out_rows_host (host, capacity rows);
 * either may be NULL.  *out_count receives N.  Synchronous on `stream`. */
int64_t ap_coords_capacity(const int32_t* contour_xy, const int32_t* contour_offsets, int n_contours,
                           int step_src);
int ap_extract_coords(ap_ctx* ctx, const int32_t* contour_xy, const int32_t* contour_offsets, int n_contours,
                      const int32_t* hole_xy, const int32_t* hole_offsets, const int32_t* hole_first,
                      int patch_src, int step_src, int read_w, int read_h, int level,
                      int32_t* out_rows_dev, int32_t* out_rows_host, int64_t capacity,
                      int64_t* out_count, void* stream);

/* ---- a9b: --no-fast-mode content filter -------------------------------------------------------
 * Replaces the per-candidate read + cv2.resize + is_black_patch / is_white_patch of
 * PatchExtractionService._iter_patch_entries with fast_mode off (atlas_patch/services/extraction.py:105-119,
 * atlas_patch/utils/image.py:7-41) for a slide resident in device memory.
 *   rows_dev      int32 n x 5 candidates (x, y, read_w, read_h, level) from ap_extract_coords (device)
 *   read_size     level-0 pixels read per candidate (>= patch_size); a larger read is cv2.resize()d to patch_size first
 *                 (OpenCV's 8-bit INTER_LINEAR restated bit for bit, any ratio)
 *   black_thresh  ExtractionConfig.black_threshold (gray < t), white_thresh = ExtractionConfig.white_threshold (s < t and
 *                 v >= 200), min_fraction = 0.7 in the reference (utils/image.py defaults)
 * Kept rows are written in input order to out_rows_dev and/or out_rows_host (capacity n rows; either may be NULL);
 * counts_dev (optional, int32 n x 2) receives the per-candidate (black, white) pixel counts.  Pixels outside the slide
 * read as 0, like IWSI.extract.  Synchronous on `stream`. */
int ap_filter_patches(ap_ctx* ctx, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch,
                      const int32_t* rows_dev, int64_t n, int read_size, int patch_size, int black_thresh,
                      int white_thresh, double min_fraction, int32_t* out_rows_dev, int32_t* out_rows_host,
                      int64_t* out_count, int32_t* counts_dev, void* stream);

/* ---- a11-a13: patch read -> preprocess -> encoder forward ------------------------------------
 * Replaces PatchFeatureEmbeddingService._iter_patch_entries_coords + PatchDataset/preprocess +
 * PatchFeatureExtractor.extract_batch + the torchvision VisionTransformer forward
 * (atlas_patch/services/feature_embedding.py:81-96, models/patch/base.py:32-107,
 *  models/patch/vit.py:9-38).  fp16 operands, fp32 accumulation / residual / LayerNorm / softmax. */
typedef struct ap_vit_desc {
    int image_size;   /* 224 */
    int patch;        /* 16  */
    int layers;       /* 12  */
    int heads;        /* 12  */
    int hidden;       /* 768 */
    int mlp;          /* 3072 */
    int input_patch;  /* edge of the RGB patches fed in (256): centre crop to image_size */
    int max_batch;    /* patches per forward chunk (workspace size); 0 = default 127 */
    int precise_layers; /* leading encoder layers whose GEMM weights are kept as fp16 hi/lo pairs (2 MMAs per weight):
                           keeps worst-case feature error under 1e-3 (DESIGN.md "precision"); -1 = default (1; 20 when layers > 32) */
    float ln_eps;     /* 1e-6 */
    float mean[3];    /* ImageNet mean / std of the torchvision preset */
    float std[3];
    int preprocess;   /* 0: centre crop of the input patch to image_size (torchvision preset whose resize == input_patch);
                         1: transformers BitImageProcessorFast as configured by the DINOv2 checkpoints
                            (atlas_patch/models/patch/dinov2.py:20-25,49): uint8 bicubic-antialias resize of the
                            input_patch^2 patch to resize_to^2, centre crop to image_size, rescale + normalise;
                         2: torchvision's ImageClassification preset for an input patch whose size differs from resize_to
                            (atlas_patch/models/patch/base.py:42-45,170: PIL image -> Image.resize(BILINEAR) = Pillow's antialiased
                            triangle filter in 22-bit fixed point -> centre crop -> /255 -> normalise); input_patch == resize_to needs
                            no resize and uses preprocess 0;
                         3: transformers ViTImageProcessorFast with resample = 2 (atlas_patch/models/patch/phikon.py:15-21,46 for
                            owkin/phikon): uint8 bilinear-antialias resize (ATen's separable uint8 kernel, triangle filter) of the
                            input_patch^2 patch to resize_to^2, centre crop to image_size (none when equal), rescale + normalise;
                         4: as 2 with Pillow's BICUBIC filter: torchvision Resize(resize_to, BICUBIC) -> CenterCrop(image_size) on the PIL
                            patch (atlas_patch/models/patch/gigapath.py:17-26) */
    int resize_to;    /* 256 (preprocess 1, 2); 224 (preprocess 3) */
    int mlp_kind;     /* 0: Linear - GELU - Linear, mlp = hidden features;
                         1: SwiGLU (Dinov2SwiGLUFFN): "mlp.0" = weights_in with 2 * mlp rows interleaved as AP_EPI_BIAS_SWIGLU_F16
                            expects, "mlp.3" = weights_out [hidden, mlp];
                         2: Linear - QuickGELU (x sigmoid(1.702 x)) - Linear: OpenAI CLIP checkpoints (atlas_patch/models/patch/plip.py:34,
                            quilt.py:56 via transformers CLIPModel) */
    int pool;         /* 0: feature = final LayerNorm of the class token (torchvision heads -> Identity, base.py:100; transformers
                            last_hidden_state[:, 0], dinov2.py:60-62), hidden floats per patch;
                         1: [class || mean of the patch tokens] of the final-LayerNorm'd sequence, 2 * hidden floats per patch
                            (atlas_patch/models/patch/midnight.py:57-61, virchow.py:57-61); register tokens are left out of the
                            mean (virchow.py:110-114, hoptimus.py:157-161) */
    int registers;    /* register tokens between the class token and the patch tokens ("register_tokens" [registers, hidden], no
                         position embedding: transformers Dinov2WithRegistersEmbeddings; hibou.py, openmidnight.py:49, the reg4 timm
                         models of hoptimus.py): sequence = 1 + registers + (image_size / patch)^2 <= 272.  0 = none */
    int pre_ln;       /* 1: LayerNorm ("encoder.pre_ln.weight / bias") on the embedded sequence before the first layer (CLIP's
                         pre_layrnorm, transformers modeling_clip.py CLIPVisionTransformer) */
    int proj_dim;     /* > 0: feature = "head.proj.weight" [proj_dim, hidden] x final LayerNorm of the class token, no bias
                         (CLIPModel.get_image_features = visual_projection(pooler_output): plip.py:56, quilt.py:60); multiple of 128.
                         0: no projection */
} ap_vit_desc;

int ap_encoder_create(ap_ctx* ctx, const ap_vit_desc* desc, ap_encoder** out_enc);
int ap_encoder_destroy(ap_encoder* enc);
/* Upload one fp32 tensor by its torchvision state_dict name (HOST pointer), e.g.
 * "conv_proj.weight", "encoder.layers.encoder_layer_3.self_attention.in_proj_weight". */
int ap_encoder_set_tensor(ap_encoder* enc, const char* name, const float* data_host, int64_t numel);
/* Pack weights (fp16, normalisation folded into conv_proj), build TMA descriptors.
 * AP_ESTATE if a tensor is missing. */
int ap_encoder_finalize(ap_encoder* enc);
int ap_encoder_embedding_dim(const ap_encoder* enc);
/* Device-resident fast path: patches are cut straight out of the level-0 slide in HBM at
 * coords_dev rows (x, y, read_w, read_h, level) [int32, n x 5]; features (n x D fp32) to
 * out_features_dev.  Asynchronous on `stream`.  read_size = the rows' read_w = read_h >= input_patch.  A larger read is
 * resized like the reference's cv2.resize to the patch size (feature_embedding.py:93-95): OpenCV's 8-bit INTER_LINEAR
 * (two taps per axis, 11-bit coefficients, fixed-point passes), bit for bit, for any ratio; the DINOv2 preprocess
 * (ap_vit_desc.preprocess = 1) only takes read_size == input_patch. */
int ap_encoder_embed_coords(ap_encoder* enc, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch,
                            const int32_t* coords_dev, int64_t n, int read_size, float* out_features_dev, void* stream);
/* Host-only (no device, no context): the integer tap tables the resizing preprocesses run from, for the `image` output indices that
 * survive the centre crop of an n_in -> n_out resize.  filter = ap_vit_desc.preprocess - 1 (0 ATen uint8 bicubic-antialias, 1 Pillow
 * BILINEAR, 2 ATen uint8 bilinear-antialias, 3 Pillow BICUBIC).  tap_min / tap_cnt: [image]; tap_w: [image * taps_capacity], row o
 * holds tap_cnt[o] weights; *max_taps = widest row, *precision = the shift of both passes.  AP_EINVAL when taps_capacity is too
 * small (*max_taps then says how much is needed).  What CPU tests compare against the reference libraries' own coefficients. */
int ap_resize_tap_tables(int filter, int n_in, int n_out, int image, int32_t* tap_min, int32_t* tap_cnt, int32_t* tap_w,
                         int taps_capacity, int* max_taps, int* precision);
/* Host-only: OpenCV's 8-bit INTER_LINEAR taps of an n_src -> n_dst down-scale (the cv2.resize of reads larger than the patch,
 * feature_embedding.py:93-95), as the crop preprocess and the content filter use them: taps[2 d], taps[2 d + 1] = the two source
 * indices of output d, weights[2 d], weights[2 d + 1] their coefficients in units of 1 / 2048.  n_src > n_dst. */
int ap_linear_tap_tables(int n_src, int n_dst, int32_t* taps, int16_t* weights);
/* a12 alone (used by the parity tests): run only the patch read + preprocess of n <= max_batch coordinates and copy the fp16
 * im2col rows the patch-embedding GEMM consumes to out_dev [n * tokens, *out_cols]: value = (pixel - round(255 mean_c)) / 256 at
 * column c * patch^2 + ky * patch + kx (exact in fp16), columns >= 3 * patch^2 are zero padding. */
int ap_encoder_preprocess(ap_encoder* enc, const uint8_t* slide_dev, int64_t W, int64_t H, int64_t pitch,
                          const int32_t* coords_dev, int64_t n, int read_size, void* out_dev, int64_t* out_cols, void* stream);
/* extract_batch-compatible path: n HOST patches (each input_patch x input_patch x 3 uint8,
 * contiguous) given by pointer; features (n x D fp32) to HOST.  Includes H2D / D2H; synchronous. */
int ap_encoder_embed_patches_host(ap_encoder* enc, const uint8_t* const* patches_host, int64_t n,
                                  float* out_features_host);

/* ---- a4: SAM2 box-prompted tissue mask on the 1024 x 1024 thumbnail ---------------------------------------------------
 * Replaces the model calls of _SAM2Predictor.predict_image (atlas_patch/services/segmentation.py:127-136):
 * SAM2ImagePredictor.set_image + predict(box = whole image, multimask_output = False) -> mask logits.  The caller keeps the
 * reference's host steps around it (PIL resize to 1024^2, `> mask_threshold`, NEAREST back to the thumbnail).
 * Hyper-parameters: atlas_patch/configs/sam2.1_hiera_t.yaml (Hiera-T: embed_dim 96, blocks {1,2,7,2}, heads {1,2,4,8},
 * windows {8,4,14,7}, global attention in blocks {5,7,9}).  Parameter names: transformers' Sam2Model state_dict.  fp32. */
typedef struct ap_sam2 ap_sam2;
typedef struct ap_sam2_desc {
    int embed_dim;
    int blocks_per_stage[4];
    int heads_per_stage[4];
    int window_per_stage[4];
    int n_global;
    int global_blocks[8];
} ap_sam2_desc;
int ap_sam2_create(ap_ctx* ctx, const ap_sam2_desc* desc, ap_sam2** out);
int ap_sam2_destroy(ap_sam2* s);
int ap_sam2_set_tensor(ap_sam2* s, const char* name, const float* data_host, int64_t numel);
int ap_sam2_finalize(ap_sam2* s);
/* image_dev: uint8 RGB [1024,1024,3] (device); logits_dev: float [1024,1024]; lowres_dev: float [256,256] or NULL. */
int ap_sam2_forward(ap_sam2* s, const uint8_t* image_dev, float* logits_dev, float* lowres_dev, void* stream);
/* Same with host buffers (synchronous). */
int ap_sam2_predict_host(ap_sam2* s, const uint8_t* image_host, float* logits_host, float* lowres_host);
/* A batch of n thumbnails, replacing predict_batch (atlas_patch/services/segmentation.py:142-180: set_image_batch + predict_batch
 * with one whole-image box each): images_host n x [1024,1024,3] uint8, logits_host n x [1024,1024], lowres_host n x [256,256] or
 * NULL.  Uploads / downloads of neighbouring images overlap the forward passes (two streams, double-buffered). */
int ap_sam2_predict_batch_host(ap_sam2* s, const uint8_t* images_host, int n, float* logits_host, float* lowres_host);
/* Parity tests: copy a named intermediate activation ("patch_embed", "blk0".."blk11", "fpn0".."fpn2", "feat_s0", "feat_s1",
 * "keys", "queries", "upscaled", "low_res") to the host. */
int ap_sam2_debug_copy(ap_sam2* s, const char* buffer_name, float* host_out, int64_t numel);

/* ---- building-block ops (kernel-level parity tests; also what the encoder is made of) -------- */
#define AP_EPI_BIAS_F16 0        /* out fp16 = acc + bias                          */
#define AP_EPI_BIAS_GELU_F16 1   /* out fp16 = gelu_erf(acc + bias)                */
#define AP_EPI_BIAS_RESID_F32 2  /* out fp32 = resid + acc + bias (resid may == out) */
#define AP_EPI_BIAS_F32 3        /* out fp32 = acc + bias                          */
#define AP_EPI_BIAS_SWIGLU_F16 4 /* out fp16 [M, N/2] = silu(g) * v, (g, v) = acc + bias in 16-column blocks: columns
                                    32b..32b+15 hold the gates and 32b+16..32b+31 the values of outputs 16b..16b+15
                                    (Dinov2SwiGLUFFN with weights_in rows interleaved); N % 256 == 0 */
#define AP_EPI_BIAS_QGELU_F16 5  /* out fp16 = v sigmoid(1.702 v), v = acc + bias (CLIP's QuickGELU: transformers activations.QuickGELUActivation) */
/* out[M,N] = epilogue(A[M,K] (fp16, row-major) x W[N,K]^T (fp16, row-major)); tcgen05/TMEM/TMA.
 * K % 64 == 0, N % 128 == 0. */
int ap_gemm_f16(ap_ctx* ctx, const void* A_dev, const void* W_dev, const float* bias_dev,
                const float* resid_dev, void* out_dev, int M, int N, int K, int epilogue, void* stream);
/* The same contraction with fp16 hi/lo SPLIT operands, fp32-like precision on the tensor cores (what the "precise" encoder layers
 * and conv_proj use): split 1: W_dev is [N, 2K] = [W_hi | W_lo]; 2: A_dev is [M, 2K] = [A_hi | A_lo]; 3: both (A_hi W_hi +
 * A_lo W_hi + A_hi W_lo); 0: plain.  x = hi + lo with hi = fp16(x), lo = fp16(x - hi). */
int ap_gemm_f16_split(ap_ctx* ctx, const void* A_dev, const void* W_dev, const float* bias_dev, const float* resid_dev,
                      void* out_dev, int M, int N, int K, int epilogue, int split, void* stream);

/* y fp16 [rows, D] = LayerNorm(x fp32 [rows, D]; gamma, beta, eps), x row stride in elements. */
int ap_layernorm_f16(ap_ctx* ctx, const float* x_dev, int64_t x_row_stride, const float* gamma_dev,
                     const float* beta_dev, float eps, void* y_dev, int rows, int D, void* stream);
/* Multi-head self-attention on packed QKV fp16 [B*S, 3*D] -> out fp16 [B*S, D]; head_dim 64. */
int ap_attention_f16(ap_ctx* ctx, const void* qkv_dev, void* out_dev, int B, int S, int heads, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ATLASPATCH_B200_H */
