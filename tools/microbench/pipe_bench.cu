// Issue-rate microbenchmark for the instructions of the attention softmax (sm_100a): clocks per warp instruction on ONE scheduler
// (SM sub-partition) for MUFU.EX2, F2FP (fp32x2 -> fp16x2), FFMA2, FADD2, FMNMX3 and the softmax mix, with 1, 2 and 4 warps on that
// scheduler (warps w, w + 4, w + 8, ... share a scheduler).  Standalone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/pipe_bench tools/microbench/pipe_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ float ex2(float x) { float r; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ uint32_t f2fp(float a, float b) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t add2rm(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmax3(float a, float b, float c) { float r; asm volatile("max.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float ffma(float a, float b, float c) { float r; asm volatile("fma.rn.ftz.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ uint32_t shladd(uint32_t a, uint32_t b) { uint32_t r; asm volatile("{.reg .b32 t; shl.b32 t, %1, 23; add.s32 %0, t, %2;}" : "=r"(r) : "r"(a), "r"(b)); return r; }

constexpr int UNROLL = 16, ITERS = 256;

template <int KIND>
__global__ void bench(long long* out, float seed, int warps_per_sched) {
    const int warp = threadIdx.x >> 5;
    float v[UNROLL];
    uint64_t w[UNROLL];
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) { v[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; w[i] = (uint64_t)__float_as_uint(v[i]) << 32 | __float_as_uint(v[i]); }
    const uint64_t c2 = w[0];
    __syncthreads();
    const long long t0 = clock64();
    if ((warp & 3) == 0 && (warp >> 2) < warps_per_sched) {      // only scheduler 0 works
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int i = 0; i < UNROLL; ++i) {
                if (KIND == 0) v[i] = ex2(v[i]);
                if (KIND == 1) v[i] = __uint_as_float(f2fp(v[i], v[(i + 1) % UNROLL]));
                if (KIND == 2) w[i] = fma2(w[i], c2, c2);
                if (KIND == 3) w[i] = add2(w[i], c2);
                if (KIND == 4) v[i] = fmax3(v[i], v[(i + 1) % UNROLL], seed);
                if (KIND == 5) v[i] = ffma(v[i], seed, seed);
                if (KIND == 6) w[i] = add2rm(w[i], c2);
                if (KIND == 7) v[i] = __uint_as_float(shladd(__float_as_uint(v[i]), __float_as_uint(v[(i + 1) % UNROLL])));
                if (KIND == 8) {   // softmax mix per PAIR of elements: 1 FFMA2, 2 MUFU, 1 FADD2, 1 F2FP  (5 instructions)
                    uint64_t x = fma2(w[i], c2, c2);
                    float a = ex2(__uint_as_float((uint32_t)x)), b = ex2(__uint_as_float((uint32_t)(x >> 32)));
                    w[i] = add2(w[i], (uint64_t)__float_as_uint(b) << 32 | __float_as_uint(a));
                    v[i] = __uint_as_float(f2fp(a, b));
                }
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) acc += v[i] + __uint_as_float((uint32_t)w[i]);
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 12345.678f) out[1] = 1;
}

template <int KIND>
int run(const char* name, int per_iter, long long* d) {
    for (int wps : {1, 2, 4}) {
        bench<KIND><<<1, 32 * 4 * 4>>>(d, 0.5f, wps);
        CK(cudaDeviceSynchronize());
        long long h = 0;
        CK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
        const double n = (double)ITERS * UNROLL * per_iter * wps;
        printf("%-44s warps/scheduler %d : %8lld clk, %.2f clk per warp instruction on the scheduler\n", name, wps, h, h / n);
    }
    return 0;
}

int main() {
    long long* d;
    CK(cudaMalloc(&d, 64));
    run<0>("MUFU.EX2 (ex2.approx.ftz.f32)", 1, d);
    run<1>("F2FP.F16.F32.PACK_AB (cvt.rn.f16x2.f32)", 1, d);
    run<2>("FFMA2 (fma.rn.f32x2)", 1, d);
    run<3>("FADD2 (add.rn.f32x2)", 1, d);
    run<6>("FADD2.RM (add.rm.f32x2)", 1, d);
    run<4>("FMNMX3 (max.f32 a, b, c)", 1, d);
    run<5>("FFMA", 1, d);
    run<7>("SHL + IADD (exponent splice)", 1, d);
    run<8>("softmax mix: FFMA2 + 2 MUFU + FADD2 + F2FP", 5, d);
    return 0;
}
