// microbench_lat.cu -- dependent-chain latencies of the ACS loop's instructions, one warp on one SM (clocks per link).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_lat microbench_lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 2048
template <int OP>
__global__ void lat(uint32_t* out, uint32_t a, uint32_t one, long long* cycles) {
    uint32_t v = threadIdx.x * 7 + a, w = threadIdx.x ^ a, y = a * 3 + threadIdx.x, z = a * 5 ^ threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v) : "r"(one), "r"(a));                       // IMAD -> IMAD
            else if (OP == 1) asm volatile("max.s16x2 %0, %0, %1;" : "+r"(v) : "r"(w));                                 // VIMNMX -> VIMNMX (no predicates)
            else if (OP == 2) asm volatile("{.reg .b32 t; mad.lo.u32 t, %0, %1, %2; max.s16x2 %0, t, %3;}" : "+r"(v) : "r"(one), "r"(a), "r"(w));   // IMAD -> VIMNMX
            else if (OP == 3) asm volatile("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3; .reg .b32 t, m;\n\t"
                                           "mad.lo.u32 t, %0, %2, %3; max.s16x2 m, t, %4; mov.b32 {r0, r1}, m; mov.b32 {r2, r3}, t; setp.eq.s16 pv, r0, r2; setp.eq.s16 pu, r1, r3;\n\t"
                                           "selp.b32 %1, %1, %4, pv; mov.b32 %0, m;}" : "+r"(v), "+r"(y) : "r"(one), "r"(a), "r"(w));              // IMAD -> VIMNMX.P (value chain) with a SEL hanging off
            else if (OP == 4) asm volatile("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3; .reg .b32 m;\n\t"
                                           "max.s16x2 m, %0, %2; mov.b32 {r0, r1}, m; mov.b32 {r2, r3}, %0; setp.eq.s16 pv, r0, r2; setp.eq.s16 pu, r1, r3;\n\t"
                                           "selp.b32 %0, %1, %2, pv;}" : "+r"(v) : "r"(y), "r"(w));                                                 // VIMNMX.P -> SEL -> VIMNMX.P (predicate on the chain)
            else if (OP == 5) v = __shfl_xor_sync(0xffffffffu, v, 1);                                                    // SHFL -> SHFL
            else if (OP == 6) { v = __shfl_xor_sync(0xffffffffu, v, 1); asm volatile("{.reg .b32 t; mad.lo.u32 t, %0, %1, %2; max.s16x2 %0, t, %3;}" : "+r"(v) : "r"(one), "r"(a), "r"(w)); }  // SHFL -> IMAD -> VIMNMX
            else if (OP == 7) { extern __shared__ uint32_t sm[]; v = sm[(v & 31) ^ 1]; }                                // LDS -> address -> LDS
            else if (OP == 8) asm volatile("prmt.b32 %0, %0, %1, 0x1032;" : "+r"(v) : "r"(w));                           // PRMT -> PRMT
            else if (OP == 9) asm volatile("{.reg .pred p; setp.ne.u32 p, %0, %1; selp.b32 %0, %2, %3, p;}" : "+r"(v) : "r"(a), "r"(y), "r"(z));  // ISETP -> SEL
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = v ^ y;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}
template <int OP> void run(const char* name, int links) {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    lat<OP><<<1, 32, 1024>>>(out, 3, 1, cyc); lat<OP><<<1, 32, 1024>>>(out, 3, 1, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-52s %6.1f clocks per iteration of the chain (%d dependent instructions)\n", name, (double)c / (ITERS * 8), links);
}
int main() {
    run<0>("IMAD -> IMAD", 1);
    run<1>("VIMNMX.S16x2 -> VIMNMX.S16x2", 1);
    run<2>("IMAD -> VIMNMX.S16x2", 2);
    run<3>("IMAD -> VIMNMX.S16x2 (+P,P; SEL off the chain)", 2);
    run<4>("VIMNMX.S16x2 P -> SEL -> (chain through predicate)", 2);
    run<5>("SHFL.BFLY -> SHFL.BFLY", 1);
    run<6>("SHFL.BFLY -> IMAD -> VIMNMX.S16x2", 3);
    run<7>("LDS -> LDS (address dependent)", 1);
    run<8>("PRMT -> PRMT", 1);
    run<9>("ISETP -> SEL", 2);
    return 0;
}
