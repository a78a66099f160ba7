// microbench_lsu.cu -- what limits the LSU/MIO path that SHFL and shared-memory loads share?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_lsu microbench_lsu.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 256
template <int OP>
__global__ void bench(uint32_t* out, int active_mask, long long* cycles) {
    extern __shared__ __align__(16) uint32_t sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * 2654435761u;
    __syncthreads();
    uint32_t x[8];
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 7 + i;
    const bool on = (active_mask >> (warp & 3)) & 1;     // which schedulers (warp % 4) take part
    long long t0 = clock64();
    if (on) {
#pragma unroll 1
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (OP == 0) x[k] = __shfl_xor_sync(0xffffffffu, x[k], 1);
                else if (OP == 1) { uint2 v = *reinterpret_cast<uint2*>(&sm[((x[k] >> 3) & 15) * 2 + (k * 64 & 1023)]); x[k] ^= v.x + v.y; }           // LDS.64, 16 distinct entries (broadcast)
                else if (OP == 2) { uint4 v = *reinterpret_cast<uint4*>(&sm[((x[k] >> 3) & 15) * 4 + (k * 128 & 2047)]); x[k] ^= v.x + v.w; }          // LDS.128, 16 distinct entries
                else if (OP == 3) { uint4 v = *reinterpret_cast<uint4*>(&sm[lane * 4 + (k * 128 & 2047)]); x[k] ^= v.x + v.w; }                       // LDS.128, 32 distinct (512 B)
                else if (OP == 4) { *reinterpret_cast<uint4*>(&sm[lane * 4 + warp * 128 + (k & 1) * 2048]) = make_uint4(x[k], x[k], x[k], x[k]); x[k] += 1; } // STS.128
                else if (OP == 5) { x[k] = __shfl_xor_sync(0xffffffffu, x[k], 1); uint2 v = *reinterpret_cast<uint2*>(&sm[((x[(k + 1) & 7] >> 3) & 15) * 2]); x[(k + 4) & 7] ^= v.x; } // SHFL + LDS.64
                asm volatile("" : "+r"(x[k]));
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= x[i];
    out[threadIdx.x] = s;
    if (on && lane == 0) atomicMax((unsigned long long*)cycles, (unsigned long long)(t1 - t0));
}
template <int OP>
void run(const char* name, int per) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    printf("%-34s", name);
    for (int cfg = 0; cfg < 5; cfg++) {
        // (warps per block, scheduler mask): one warp on one scheduler; 4 warps on one scheduler; 1 warp on each; 4 on each
        const int nw[5] = {4, 16, 4, 16, 32}, mask[5] = {1, 1, 15, 15, 15};
        cudaMemset(cyc, 0, 8);
        bench<OP><<<1, nw[cfg] * 32, 16384>>>(out, mask[cfg], cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        int active = 0;
        for (int w = 0; w < nw[cfg]; w++) active += (mask[cfg] >> (w & 3)) & 1;
        printf("  %dw/%s: %5.2f", active, mask[cfg] == 1 ? "1sched" : "4sched", (double)active * ITERS * 8 * per / c);
    }
    printf("   (LSU instr/clk/SM)\n");
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("SHFL.BFLY", 1);
    run<1>("LDS.64 (16 distinct 8B entries)", 1);
    run<2>("LDS.128 (16 distinct 16B entries)", 1);
    run<3>("LDS.128 (32 distinct, 512 B)", 1);
    run<4>("STS.128 (512 B)", 1);
    run<5>("SHFL + LDS.64 pair", 2);
    return 0;
}
