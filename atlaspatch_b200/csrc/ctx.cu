// Context, error reporting and the TMA descriptor helper.
#include <cstdlib>

#include "ap_internal.cuh"

static thread_local char g_init_error[1024] = "no error";

int ap_set_error(ap_ctx* ctx, int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    char* dst = ctx ? ctx->err : g_init_error;
    vsnprintf(dst, 1024, fmt, ap);
    va_end(ap);
    return code;
}

extern "C" int ap_version(void) { return 2; }

// sizeof of the structs that cross the ABI, so that a binding can check its own declaration against the library it loaded
extern "C" int ap_sizeof(const char* struct_name) {
    if (!struct_name) return -1;
    if (!strcmp(struct_name, "ap_vit_desc")) return static_cast<int>(sizeof(ap_vit_desc));
    if (!strcmp(struct_name, "ap_sam2_desc")) return static_cast<int>(sizeof(ap_sam2_desc));
    return -1;
}

extern "C" const char* ap_last_error(const ap_ctx* ctx) {
    if (!ctx) return g_init_error;
    return ctx->err[0] ? ctx->err : "no error";
}

extern "C" int64_t ap_launch_count(const ap_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
extern "C" int ap_sm_count(const ap_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

extern "C" int ap_init(int device, ap_ctx** out_ctx) {
    if (!out_ctx) return ap_set_error(nullptr, AP_EINVAL, "ap_init: out_ctx is NULL");
    *out_ctx = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return ap_set_error(nullptr, AP_ECUDA, "ap_init: no CUDA device (%s); this library has no CPU fallback",
                            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return ap_set_error(nullptr, AP_EINVAL, "ap_init: device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return ap_set_error(nullptr, AP_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return ap_set_error(nullptr, AP_ECUDA, "ap_init: device %d is sm_%d%d; this library is built for sm_100a (B200) only",
                            device, prop.major, prop.minor);
    int prev_device = -1;
    if (cudaGetDevice(&prev_device) != cudaSuccess) prev_device = -1;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return ap_set_error(nullptr, AP_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    ap_ctx* ctx = new ap_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = prop.sharedMemPerBlockOptin;
    cudaDriverEntryPointQueryResult qres;
    e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ctx->encode_tiled, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ctx->encode_tiled) {
        delete ctx;
        return ap_set_error(nullptr, AP_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    }
    if (const char* env = getenv("AP_GEMM_CTA_GROUP")) ctx->gemm_cta_group = atoi(env) == 1 ? 1 : 2;
    if (prev_device >= 0 && prev_device != device) cudaSetDevice(prev_device);   // the caller's current device is left as it was;
    *out_ctx = ctx;                                                               // every later entry selects ctx->device itself
    return AP_OK;
}

extern "C" int ap_set_option(ap_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return AP_EINVAL;
    if (!strcmp(key, "gemm_cta_group")) {
        AP_REQUIRE(ctx, value == 1 || value == 2, "gemm_cta_group must be 1 or 2");
        ctx->gemm_cta_group = value;
        return AP_OK;
    }
    if (!strcmp(key, "attn_mode")) {
        AP_REQUIRE(ctx, value == 1 || value == 2, "attn_mode must be 1 or 2");
        ctx->attn_mode = value;
        return AP_OK;
    }
    if (!strcmp(key, "precise_kind")) {
        AP_REQUIRE(ctx, value >= 0 && value <= 1, "precise_kind must be 0 or 1");
        ctx->precise_kind = value;
        return AP_OK;
    }
    if (!strcmp(key, "precise_aw_layers")) {
        AP_REQUIRE(ctx, value >= -1, "precise_aw_layers must be >= -1 (-1 = automatic)");
        ctx->precise_aw_layers = value;
        return AP_OK;
    }
    if (!strcmp(key, "pdl")) {
        ctx->pdl = value != 0;
        return AP_OK;
    }
    if (!strcmp(key, "cls_only_last_layer")) {
        ctx->cls_only_last_layer = value != 0;
        return AP_OK;
    }
    if (!strcmp(key, "profile_stride")) {
        AP_REQUIRE(ctx, value >= 1, "profile_stride must be >= 1");
        ctx->prof_stride = value;
        for (auto& v : ctx->prof_seen) v = 0;
        return AP_OK;
    }
    if (!strcmp(key, "fold_ln")) {   // read at ap_encoder_finalize
        AP_REQUIRE(ctx, value >= 0 && value <= 2, "fold_ln must be 0 (off), 1 (automatic) or 2 (on)");
        ctx->fold_ln = value;
        return AP_OK;
    }
    if (!strcmp(key, "sam_tensor_cores")) {
        AP_REQUIRE(ctx, value >= 0 && value <= 3, "sam_tensor_cores must be 0..3");
        ctx->sam_tensor_cores = value;   // 3: tcgen05 split GEMM, 1: mma.sync split-fp16, 2: mma.sync plain fp16, 0: fp32 SIMT
        return AP_OK;
    }
    if (!strcmp(key, "precise_mask")) {   // read at ap_encoder_finalize
        AP_REQUIRE(ctx, value >= 0 && value <= 15, "precise_mask must be in [0, 15]");
        ctx->precise_mask = value;
        return AP_OK;
    }
    if (!strcmp(key, "attn_emu")) {
        AP_REQUIRE(ctx, value >= 0 && value <= 8, "attn_emu must be in [0, 8]");
        ctx->attn_emu = value;
        return AP_OK;
    }
    if (!strcmp(key, "attn_variant")) {
        ctx->attn_variant = value;
        return AP_OK;
    }
    if (!strcmp(key, "gemm_debug")) {
        ctx->gemm_debug = value;
        return AP_OK;
    }
    return ap_set_error(ctx, AP_EINVAL, "unknown option '%s'", key);
}

extern "C" int ap_destroy(ap_ctx* ctx) {
    if (ctx) {
        sam_state_free(ctx);
        for (auto& r : ctx->prof_recs) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
        for (auto e : ctx->prof_pool) cudaEventDestroy(e);
    }
    delete ctx;
    return AP_OK;
}

// ---- per-launch event timing -----------------------------------------------------------------------
static cudaEvent_t prof_get_event(ap_ctx* ctx) {
    if (!ctx->prof_pool.empty()) {
        cudaEvent_t e = ctx->prof_pool.back();
        ctx->prof_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(ap_ctx* c, cudaStream_t s, int cls, int64_t tag) : ctx(c), st(s) {
    if (!ctx || !((ctx->profiling >> cls) & 1u)) return;
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    if (ctx->prof_stride > 1 && (ctx->prof_seen[cls & 7]++ % static_cast<unsigned>(ctx->prof_stride)) != 0) return;
    cudaEvent_t start = prof_get_event(ctx);
    stop = prof_get_event(ctx);
    cudaEventRecord(start, st);
    ctx->prof_recs.push_back({start, stop, cls, tag});
}
ProfScope::~ProfScope() {
    if (stop) cudaEventRecord(stop, st);
}

extern "C" int ap_profile_enable(ap_ctx* ctx, int on) {
    if (!ctx) return AP_EINVAL;
    ctx->profiling = on < 0 ? 0xFFFFFFFFu : static_cast<unsigned>(on);  // -1 (or 1 | 2 | 4 ...): bit c = time kernel class c
    return AP_OK;
}

// Sums (and clears) the recorded event pairs: total_ms[AP_K_NUM], counts[AP_K_NUM].  Synchronises the device.
extern "C" int ap_profile_read(ap_ctx* ctx, double* total_ms, int64_t* counts, int n_classes) {
    if (!ctx || !total_ms || !counts || n_classes < AP_K_NUM) return AP_EINVAL;
    DeviceGuard guard(ctx);
    AP_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    for (int i = 0; i < n_classes; ++i) { total_ms[i] = 0.0; counts[i] = 0; }
    for (auto& r : ctx->prof_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) { total_ms[r.cls] += ms; counts[r.cls] += 1; }
        ctx->prof_pool.push_back(r.start);
        ctx->prof_pool.push_back(r.stop);
    }
    ctx->prof_recs.clear();
    return AP_OK;
}

// Like ap_profile_read for ONE kernel class, but split by the launches' tags (GEMM: (N << 32) | (K << 4) | epilogue).
// Clears all records.  keys / total_ms / counts have room for `cap` distinct tags; *n_out receives how many were seen.
extern "C" int ap_profile_read_tagged(ap_ctx* ctx, int cls, int64_t* keys, double* total_ms, int64_t* counts, int cap, int* n_out) {
    if (!ctx || !keys || !total_ms || !counts || !n_out || cap <= 0) return AP_EINVAL;
    DeviceGuard guard(ctx);
    AP_CHECK_CUDA(ctx, cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(ctx->prof_mu);
    int n = 0;
    for (auto& r : ctx->prof_recs) {
        float ms = 0.f;
        if (r.cls == cls && cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
            int i = 0;
            while (i < n && keys[i] != r.tag) ++i;
            if (i == n && n < cap) { keys[n] = r.tag; total_ms[n] = 0.0; counts[n] = 0; ++n; }
            if (i < n) { total_ms[i] += ms; counts[i] += 1; }
        }
        ctx->prof_pool.push_back(r.start);
        ctx->prof_pool.push_back(r.stop);
    }
    ctx->prof_recs.clear();
    *n_out = n;
    return AP_OK;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int ap_make_tmap_f16_2d(ap_ctx* ctx, CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols) {
    AP_REQUIRE(ctx, (reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base %p not 16-byte aligned", base);
    AP_REQUIRE(ctx, (row_stride_elems * 2) % 16 == 0, "TMA row stride %llu B not a multiple of 16",
               (unsigned long long)(row_stride_elems * 2));
    AP_REQUIRE(ctx, box_cols * 2 == 128 && box_rows <= 256, "TMA box %ux%u unsupported", box_rows, box_cols);
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = reinterpret_cast<EncodeTiledFn>(ctx->encode_tiled)(
        map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return ap_set_error(ctx, AP_ECUDA, "cuTensorMapEncodeTiled failed: CUresult %d", (int)r);
    return AP_OK;
}
