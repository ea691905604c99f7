// Context, error reporting and the TMA descriptor helper.
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

extern "C" int ap_version(void) { return 1; }

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
    *out_ctx = ctx;
    return AP_OK;
}

extern "C" int ap_destroy(ap_ctx* ctx) {
    delete ctx;
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
