// Shared between the attention kernels (attention_tcgen05.cu: whole-tile pipelines; attention_units.cu: key-block units): launch
// arguments and the packed-fp32 / exponential helpers of the softmax.
#pragma once
#include "ap_internal.cuh"
#include "ptx.cuh"

// launch arguments (a named type: it crosses translation units)
struct AttnArgs {
    int B, S, heads;     // images, tokens per image (row pitch of the QKV buffer per image), heads
    int q0, nq;          // query token window
    int k0, nk;          // MMA key token window
    int xkey;            // extra key token (XK kernels), else -1
    int S_pad;           // nk rounded up to 16
    int variant;         // diagnostics
    int out_ld;          // halfs between output rows (heads * 64, or 2 x that when the row holds [hi | lo])
    int split_lo;        // also write lo = fp16(o - fp16(o)) at column offset heads * 64 (A-operand split of out_proj)
};

namespace {

constexpr int Q_TILE_BYTES = 128 * 128;  // 128 rows x 64 fp16


__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float dot8_h(const uint4& a, const uint4& b, float acc) {
    const __half2* pa = reinterpret_cast<const __half2*>(&a);
    const __half2* pb = reinterpret_cast<const __half2*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 x = __half22float2(pa[i]), y = __half22float2(pb[i]);
        acc = fmaf(x.x, y.x, acc);
        acc = fmaf(x.y, y.y, acc);
    }
    return acc;
}

__device__ __forceinline__ uint64_t pk2f(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void upk2f(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2f(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2f(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t add2f_rm(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ex2_mufu(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// 2^x for two values x <= 0 on the FMA pipe.  t = x + 1.5 * 2^23 rounded DOWN keeps floor(x) in the low mantissa bits; f = x - floor(x)
// in [0, 1); p(f) = 2^f by a degree-3 minimax polynomial (max rel. error 7.5e-5, a sixth of the fp16 rounding P gets anyway); the
// exponent is spliced in by adding floor(x) << 23 to the bits of p.  x is clamped to >= -32 (P < 2^-24 rounds to zero in fp16 anyway).
__device__ __forceinline__ void ex2_emu2(uint64_t x, float& p0, float& p1) {
    float x0, x1;
    upk2f(x, x0, x1);
    x = pk2f(fmaxf(x0, -32.f), fmaxf(x1, -32.f));
    const float magic = 12582912.f;   // 1.5 * 2^23
    const uint64_t t = add2f_rm(x, pk2f(magic, magic));
    const uint64_t nfl = fma2f(t, pk2f(-1.f, -1.f), pk2f(magic, magic));   // -(floor x), exact
    const uint64_t f = add2f(x, nfl);
    const float c3 = 0.0780240297f, c2 = 0.2260670662f, c1 = 0.6958339810f, c0 = 0.9999251366f;   // max rel. error 7.5e-5 on [0, 1]
    uint64_t p = fma2f(f, pk2f(c3, c3), pk2f(c2, c2));
    p = fma2f(p, f, pk2f(c1, c1));
    p = fma2f(p, f, pk2f(c0, c0));
    float q0, q1, t0, t1;
    upk2f(p, q0, q1);
    upk2f(t, t0, t1);
    p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(t0) << 23));
    p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(t1) << 23));
}
template <int W>
__device__ __forceinline__ void tmem_ld_w(uint32_t taddr, uint32_t (&r)[W]) {
    if constexpr (W == 64) ptx::tmem_ld_32x64(taddr, r);
    else if constexpr (W == 32) ptx::tmem_ld_32x32(taddr, r);
    else ptx::tmem_ld_32x16(taddr, r);
}

}  // namespace

// attention_units.cu: the key-block-unit pipeline (three rotating 128-column score buffers, one O tile per softmax warpgroup)
int ap_attention_units_run(ap_ctx* ctx, const AttnPlan* plan, __half* out, const AttnArgs& a, int grid, cudaStream_t stream);
