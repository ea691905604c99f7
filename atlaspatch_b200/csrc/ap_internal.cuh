// Internal declarations shared by the translation units of libatlaspatch_b200.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <vector>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "atlaspatch_b200.h"

struct ap_ctx {
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    char err[1024] = {0};
    std::atomic<int64_t> launches{0};
    // cuTensorMapEncodeTiled, fetched through cudaGetDriverEntryPoint (no link-time libcuda dependency)
    void* encode_tiled = nullptr;
    int fold_ln = 1;              // LayerNorm folded into the GEMMs around it: 0 off, 1 automatic (<= 32 layers), 2 on (read at
                                  // ap_encoder_finalize; "fold_ln"; encoder.cu)
    int sam_tensor_cores = 3;     // SAM2 linears: 3 tcgen05 GEMM with split-fp16 operands (A_hi W_hi + A_lo W_hi + A_hi W_lo), 1 the same
                                  // three products on mma.sync, 2 plain fp16 on mma.sync (1 MMA, not pinnable), 0 fp32 SIMT
    void* sam_state = nullptr;    // sam2_kernels.cu: split-weight cache + scratch of the tcgen05 linears (sam_state_free)
    int precise_mask = 15;        // which GEMMs of the precise layers get hi/lo split operands: 1 qkv, 2 out_proj, 4 mlp.0, 8 mlp.3
    int precise_kind = 0;         // what is split there: 0 the weights, 1 the A operands of qkv / out_proj / mlp.0 (+ mlp.3's weights)
                                  // ("precise_kind"; encoder.cu: ap_encoder_finalize; measured: profiles/r02_dinov2_giant_precision.log)
    int precise_aw_layers = -1;   // with kind 0: leading layers whose qkv / out_proj / mlp.0 ALSO get split A operands (3 products per term);
                                  // -1 = automatic (8 for encoders deeper than 32 layers, else 0)
    int pdl = 1;                  // programmatic dependent launch for the encoder kernel chain (ap_set_option "pdl")
    int cls_only_last_layer = 1;  // last layer: attention / out_proj / MLP only for the class-token row (ap_set_option)
    int attn_mode = 2;       // 2: tcgen05 attention when 16 <= S_pad <= 256, 1: warp-MMA (mma.sync) kernel
    int attn_variant = 0;    // diagnostics
    int attn_emu = 0;        // attention (<= 208 keys): exponential PAIRS per 16 evaluated on the FMA pipe instead of MUFU (0, 4, 6, 8)
    int gemm_debug = 0;      // diagnostics: see EpiParams::debug
    int gemm_cta_group = 2;  // default GEMM flavour (ap_set_option "gemm_cta_group"; env AP_GEMM_CTA_GROUP)
    // optional per-launch CUDA-event timing (ap_profile_*): bench.py's live roofline measurement
    unsigned profiling = 0;  // bit mask of ApKernelClass values to time (0 = off)
    int prof_stride = 1;     // time every prof_stride-th launch of a class ("profile_stride"): sampling keeps the event records from
    unsigned prof_seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // serialising the programmatic-dependent-launch chain around every kernel
    std::mutex prof_mu;
    struct ProfRec { cudaEvent_t start, stop; int cls; int64_t tag; };
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
};

enum ApKernelClass { AP_K_GEMM = 0, AP_K_ATTENTION = 1, AP_K_LAYERNORM = 2, AP_K_PREPROCESS = 3, AP_K_COORDS = 4,
                     AP_K_THUMBNAIL = 5, AP_K_OTHER = 6, AP_K_NUM = 7 };

// Records a CUDA-event pair around the launches issued in its scope when profiling is on (same stream as the kernel).
struct ProfScope {
    ap_ctx* ctx; cudaStream_t st; cudaEvent_t stop = nullptr;
    ProfScope(ap_ctx* c, cudaStream_t s, int cls, int64_t tag = 0);  // tag: e.g. the GEMM shape, see ap_profile_read_tagged
    ~ProfScope();
};

int ap_set_error(ap_ctx* ctx, int code, const char* fmt, ...);
void sam_state_free(ap_ctx* ctx);   // sam2_kernels.cu

// Every C-ABI entry runs on ITS context's device whatever the calling thread's current device is (another thread, or torch having
// switched devices since ap_init), and leaves the caller's current device as it found it.
struct DeviceGuard {
    int prev = -1, want = -1;
    explicit DeviceGuard(const ap_ctx* ctx) {
        if (!ctx) return;
        want = ctx->device;
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != want) cudaSetDevice(want);
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != want) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

#define AP_CHECK_CUDA(ctx, call)                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return ap_set_error((ctx), AP_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                __FILE__, __LINE__);                                               \
    } while (0)

#define AP_CHECK_LAUNCH(ctx, what)                                                                 \
    do {                                                                                           \
        (ctx)->launches.fetch_add(1, std::memory_order_relaxed);                                   \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess)                                                                    \
            return ap_set_error((ctx), AP_ECUDA, "launch of %s failed: %s (%s:%d)", what,          \
                                cudaGetErrorString(e__), __FILE__, __LINE__);                      \
    } while (0)

#define AP_REQUIRE(ctx, cond, ...)                                                                 \
    do {                                                                                           \
        if (!(cond)) return ap_set_error((ctx), AP_EINVAL, __VA_ARGS__);                           \
    } while (0)

// Launch with the programmatic-stream-serialization attribute (PDL): every kernel of the encoder chain starts with
// griddepcontrol.wait, so its prologue (barrier init, TMEM allocation, descriptor prefetch, index math) overlaps the tail
// of its predecessor.  cluster_x > 1 adds the thread-block-cluster dimension.
template <typename... KArgs, typename... Args>
inline cudaError_t ap_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster_x,
                                 bool pdl, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attrs[2];
    int n = 0;
    if (cluster_x > 1) {
        attrs[n].id = cudaLaunchAttributeClusterDimension;
        attrs[n].val.clusterDim.x = cluster_x;
        attrs[n].val.clusterDim.y = 1;
        attrs[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl) {
        attrs[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attrs[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = attrs;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncSetAttribute is per device: remember per device (bit = device index) that an instantiation has been configured
struct PerDeviceOnce {
    std::atomic<uint64_t> mask{0};
    bool need(int dev) const { return ((mask.load(std::memory_order_acquire) >> (dev & 63)) & 1ull) == 0; }
    void done(int dev) { mask.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

// ---- TMA descriptor helper (host) --------------------------------------------------------------
// 2-D row-major fp16 matrix [rows, cols] (cols contiguous), box = box_cols x box_rows, 128B swizzle.
int ap_make_tmap_f16_2d(ap_ctx* ctx, CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols);

// ---- kernels' host launchers (internal) --------------------------------------------------------
struct GemmPlan {
    CUtensorMap map_a;
    CUtensorMap map_w;       // box = bn rows of W (single-CTA tiles)
    CUtensorMap map_w_half;  // box = bn/2 rows of W (each CTA of a cta_group::2 pair loads half)
    int M, N, K, epilogue;
    int Ka;         // width of A in memory
    // Split operands (DESIGN.md "precision"): the contraction runs over `segs` segments of Kseg columns; segment s multiplies
    // A columns [a_seg[s] Kseg, +Kseg) with W columns [w_seg[s] Kseg, +Kseg).  W-split (W = [W_hi | W_lo], A stored once):
    // a_seg = {0, 0}, w_seg = {0, 1}.  A-split (A = [A_hi | A_lo], W stored once): a_seg = {0, 1}, w_seg = {0, 0}.
    int segs, Kseg;
    int a_seg[3], w_seg[3];
    int bn;         // N tile (256 or 128)
    int cta_group;  // 2: CTA pair owns a 256 x 256 tile; 1: one CTA owns a 128 x bn tile
};
int ap_gemm_plan(ap_ctx* ctx, GemmPlan* plan, const void* A, const void* W, int M, int N, int K, int epilogue, int Ka = 0);
// split: AP_SPLIT_NONE, AP_SPLIT_W (W holds [hi | lo], 2 Kbase wide), AP_SPLIT_A (A holds [hi | lo]), AP_SPLIT_AW (both: 3 segments)
enum { AP_SPLIT_NONE = 0, AP_SPLIT_W = 1, AP_SPLIT_A = 2, AP_SPLIT_AW = 3 };
int ap_gemm_plan_split(ap_ctx* ctx, GemmPlan* plan, const void* A, const void* W, int M, int N, int Kbase, int epilogue, int split);
// Patch-embed epilogue parameters: output row remap (b*T + t -> b*(T+1) + 1 + t) and +pos[1+t].
struct GemmExtra {
    const float* pos = nullptr;  // [(T+1), N] fp32 or null
    int tokens_per_image = 0;    // T (0 = no remap)
    int lead_tokens = 1;         // rows in front of each image's T patch rows: class token + register tokens (b*T + t -> b*(T+lead) + lead + t)
    float alpha = 1.0f;          // scale applied to the accumulator before bias
    // ---- LayerNorm folded into the GEMMs around it (encoder.cu "LN folding") -----------------------------------------------
    // producer side (fp32-output epilogues): also write the output as fp16 (the next GEMM's A operand) and, per output row,
    // the partial (sum, sum of squares) of each block of bn/2 columns: stats_out[out_row * (N / (bn/2)) + column block]
    __half* out_h = nullptr;
    float2* stats_out = nullptr;
    // consumer side (fp16-output epilogues): A holds raw x, W has gamma folded in and its rows centred (zero row sums absorb the
    // mean: x W''^T = (x - mean) W'^T); the epilogue only scales by this row's 1 / sigma from the ln_parts partials of the row:
    //   out = rstd[m] * acc + bias'[n]
    const float2* stats_in = nullptr;
    int ln_parts = 0;
    int ln_dim = 0;                  // length of the normalised rows (hidden size)
    float ln_eps = 0.f;
    // AP_EPI_BIAS_F32 only: destination row stride in floats (0 = N), real columns (0 = N; the GEMM's N may be padded to 128),
    // activation on the fp32 result (0 none, 1 GELU erf, 2 ReLU)
    int out_ld = 0, n_valid = 0, act = 0;
};
int ap_gemm_run(ap_ctx* ctx, const GemmPlan* plan, const float* bias, const float* resid, void* out,
                const GemmExtra* extra, cudaStream_t stream);

// y_f16 rows are y_ld halfs apart (0 = D); with split_lo the row holds [hi | lo] (y_ld >= 2 D): lo = fp16(y - hi)
int ap_layernorm_run(ap_ctx* ctx, const float* x, int64_t x_row_stride, const float* gamma, const float* beta,
                     float eps, __half* y_f16, float* y_f32, int rows, int D, cudaStream_t stream, int y_ld = 0, int split_lo = 0);
int ap_attention_run(ap_ctx* ctx, const __half* qkv, __half* out, int B, int S, int heads, cudaStream_t stream);
// tcgen05 attention (S <= 257): TMA descriptors over the packed QKV buffer [rows, 3 * heads * 64]
struct AttnPlan {
    CUtensorMap map_q;   // box 128 rows x 64
    CUtensorMap map_q16; // box 16 rows x 64: the third query tile of a 257-token sequence holds one row (attention_units.cu)
    CUtensorMap map_kv;  // box S_pad rows x 64
    const __half* qkv;   // the packed QKV buffer the maps describe
    int S_pad;           // MMA keys rounded up to 16
    int q0, nq;          // query token window per image
    int k0, nk;          // MMA key token window per image
    int xkey;            // extra key token folded in as a rank-1 update (257-token sequences: the class token), else -1
};
int ap_attention_tc_plan(ap_ctx* ctx, AttnPlan* plan, const __half* qkv, int rows, int S, int heads);
// out rows are out_ld halfs apart (0 = heads * 64); with split_lo a row holds [hi | lo] (out_ld >= 2 * heads * 64)
int ap_attention_tc_run(ap_ctx* ctx, const AttnPlan* plan, __half* out, int B, int S, int heads, cudaStream_t stream, int out_ld = 0,
                        int split_lo = 0);
int ap_preprocess_run(ap_ctx* ctx, const uint8_t* slide, int64_t W, int64_t H, int64_t pitch, const int32_t* coords,
                      int64_t n, int input_patch, int image, int patch, __half* out, int64_t out_row_stride,
                      const int* centre, int dup, const int32_t* lin_s, const int16_t* lin_w, cudaStream_t stream);
int ap_build_linear_tables(ap_ctx* ctx, int n_src, int n_dst, std::vector<int32_t>& taps, std::vector<int16_t>& weights);
// DINOv2 preprocess (transformers BitImageProcessorFast): antialias bicubic resize + centre crop + im2col (preprocess_resize.cu)
int ap_build_resize_tables(ap_ctx* ctx, int n_in, int n_out, int image, int kind, std::vector<int32_t>& tap_min,
                           std::vector<int32_t>& tap_cnt, std::vector<int32_t>& tap_w, int* max_taps, int* precision);
int ap_preprocess_resize_run(ap_ctx* ctx, const uint8_t* slide, int64_t W, int64_t H, int64_t pitch, const int32_t* coords, int64_t n,
                             int input_patch, int image, int patch, const int32_t* tap_min, const int32_t* tap_cnt, const int32_t* tap_w,
                             int max_taps, int precision, int max_src_rows, __half* out, int64_t out_row_stride, const int* centre,
                             cudaStream_t stream);
// class-token query only; image b's output row is out[b * out_row_stride] (1: compact rows, S: row 0 of every image's block)
int ap_cls_attention_run(ap_ctx* ctx, const __half* qkv, __half* out, int B, int S, int heads, int out_row_stride, cudaStream_t stream);
// [CLS || mean(patch tokens)] of the final-LayerNorm'd sequence: x [n_images, tokens1, D] fp32 -> out [n_images, 2 D] fp32
int ap_cls_mean_pool_run(ap_ctx* ctx, const float* x, int n_images, int tokens1, int lead, int D, const float* gamma, const float* beta, float eps,
                         float* out, cudaStream_t stream);
int ap_gather_rows_run(ap_ctx* ctx, const float* src, float* dst, int n_rows, int64_t src_row_stride, int D, cudaStream_t stream);
int ap_cls_rows_run(ap_ctx* ctx, float* x, const float* cls, const float* pos, const float* regs, int lead, int n_images, int tokens, int D, __half* xh,
                    float2* stats, int parts, cudaStream_t stream);
