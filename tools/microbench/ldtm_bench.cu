// tcgen05.ld (LDTM) throughput microbenchmark: how many bytes per clock one SM can move TMEM -> registers, by instruction
// shape and by number of issuing warps.  The attention softmax and every GEMM epilogue are paced by this unit
// (DESIGN.md section 4.1b).  Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/ldtm_bench ldtm_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int X> __device__ __forceinline__ void ld32x32(uint32_t taddr, uint32_t& sink);
template <> __device__ __forceinline__ void ld32x32<16>(uint32_t taddr, uint32_t& sink) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) sink ^= r[i];
}
template <> __device__ __forceinline__ void ld32x32<32>(uint32_t taddr, uint32_t& sink) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
}
// two x32 loads in flight before one wait (what the pipelined epilogues do)
__device__ __forceinline__ void ld32x32_2x32(uint32_t taddr, uint32_t& sink) {
    uint32_t r[64];
#pragma unroll
    for (int h = 0; h < 2; ++h)
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                     "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                     : "=r"(r[32 * h + 0]), "=r"(r[32 * h + 1]), "=r"(r[32 * h + 2]), "=r"(r[32 * h + 3]), "=r"(r[32 * h + 4]), "=r"(r[32 * h + 5]),
                       "=r"(r[32 * h + 6]), "=r"(r[32 * h + 7]), "=r"(r[32 * h + 8]), "=r"(r[32 * h + 9]), "=r"(r[32 * h + 10]), "=r"(r[32 * h + 11]),
                       "=r"(r[32 * h + 12]), "=r"(r[32 * h + 13]), "=r"(r[32 * h + 14]), "=r"(r[32 * h + 15]), "=r"(r[32 * h + 16]),
                       "=r"(r[32 * h + 17]), "=r"(r[32 * h + 18]), "=r"(r[32 * h + 19]), "=r"(r[32 * h + 20]), "=r"(r[32 * h + 21]),
                       "=r"(r[32 * h + 22]), "=r"(r[32 * h + 23]), "=r"(r[32 * h + 24]), "=r"(r[32 * h + 25]), "=r"(r[32 * h + 26]),
                       "=r"(r[32 * h + 27]), "=r"(r[32 * h + 28]), "=r"(r[32 * h + 29]), "=r"(r[32 * h + 30]), "=r"(r[32 * h + 31])
                     : "r"(taddr + 32 * h) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) sink ^= r[i];
}
// 16 lanes x 256 bits: every thread gets 4 registers per repetition; .x8 = 32 registers (16 lanes x 64 columns)
__device__ __forceinline__ void ld16x256_x8(uint32_t taddr, uint32_t& sink) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
}
// packed 16-bit: two adjacent columns' low halves per register (x32 registers = 64 columns)
__device__ __forceinline__ void ld32x32_pack16_x32(uint32_t taddr, uint32_t& sink) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.pack::16b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
                   "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
                   "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) sink ^= r[i];
}

// mode: 0 = 32x32b.x16, 1 = 32x32b.x32, 2 = 2 x (32x32b.x32) per wait, 3 = 16x256b.x8, 4 = 32x32b.pack::16b.x32
__global__ void __launch_bounds__(512, 1) ldtm_kernel(int mode, int n_warps, int iters, long long* cycles, uint32_t* sink_out) {
    __shared__ uint32_t tmem_ptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_ptr)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tmem_ptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t sink = 0;
    long long t0 = 0, t1 = 0;
    __syncthreads();
    if (warp < n_warps) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t a = base + ((i * 64) & 255) + (warp >= 4 ? 256 : 0) * 0;   // stay inside the 512 allocated columns
            switch (mode) {
                case 0: ld32x32<16>(a, sink); break;
                case 1: ld32x32<32>(a, sink); break;
                case 2: ld32x32_2x32(a, sink); break;
                case 3: ld16x256_x8(a, sink); break;
                case 4: ld32x32_pack16_x32(a, sink); break;
            }
        }
        t1 = clock64();
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && warp < n_warps) cycles[blockIdx.x * 16 + warp] = t1 - t0;
    if (sink == 0x12345678u) sink_out[0] = sink;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_ptr), "r"(512) : "memory");
}

int main() {
    long long* cyc;
    uint32_t* sink;
    CK(cudaMalloc(&cyc, 148 * 16 * sizeof(long long)));
    CK(cudaMalloc(&sink, 4));
    const char* names[5] = {"32x32b.x16", "32x32b.x32", "2 x 32x32b.x32 per wait", "16x256b.x8", "32x32b.pack::16b.x32"};
    const int bytes_per_instr[5] = {32 * 16 * 4, 32 * 32 * 4, 2 * 32 * 32 * 4, 16 * 64 * 4, 32 * 64 * 2};   // TMEM bytes delivered to registers
    const int iters = 2000;
    for (int grid : {1, 148})
        for (int mode = 0; mode < 5; ++mode)
            for (int nw : {1, 2, 4, 8, 16}) {
                ldtm_kernel<<<grid, 512>>>(mode, nw, iters, cyc, sink);
                CK(cudaDeviceSynchronize());
                long long h[16];
                CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
                long long mx = 0;
                for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
                const double bpc = (double)bytes_per_instr[mode] * iters * nw / (double)mx;
                printf("grid %3d  %-26s warps %2d : %8lld clk for %d instr/warp -> %.1f clk/instr/warp, %.1f B/clk/SM (register bytes)\n", grid,
                       names[mode], nw, mx, iters, (double)mx / iters, bpc);
            }
    return 0;
}
