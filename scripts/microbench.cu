// microbench.cu -- issue throughput of the instructions the ACS loop is made of, on one SM.
// Each kernel runs NW warps (1 block) of 8 independent dependency chains; reports warp-instructions
// per clock per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define ITERS 512
#define CHAINS 8

template <int OP>
__device__ __forceinline__ void op(uint32_t (&x)[CHAINS], uint32_t (&y)[CHAINS], uint32_t (&z)[CHAINS], uint32_t a, uint32_t b, bool p0, bool p1) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        uint32_t& v = x[i];
        uint32_t w = x[(i + 1) % CHAINS];
        if (OP == 0) v = v + w;                                   // IADD3 / IMAD.IADD (ptxas picks)
        else if (OP == 1) v = v * a + b;                          // IMAD
        else if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b));
        else if (OP == 3) v = ((i & 1) ? p0 : p1) ? w : v;        // SEL
        else if (OP == 4) v = __vmaxs2(v, w);                     // VIMNMX.S16x2
        else if (OP == 5) v = __vadd2(v, a);                      // VIADD.16x2
        else if (OP == 6) v = __byte_perm(v, a, b);               // PRMT
        else if (OP == 7) v = __shfl_xor_sync(0xffffffffu, v, 1); // SHFL.BFLY
        else if (OP == 8) { __half2 h = *(__half2*)&v, k = *(__half2*)&a; h = __hadd2(h, k); v = *(uint32_t*)&h; }
        else if (OP == 9) { __half2 h = *(__half2*)&v, k = *(__half2*)&a; h = __hmax2(h, k); v = *(uint32_t*)&h; }
        else if (OP == 10) { __half2 h = *(__half2*)&v, k = *(__half2*)&a; v = __hlt2_mask(h, k) ^ v; }   // HSET2 + LOP
        else if (OP == 11) { bool ph, pl; v = __vibmax_s16x2(v, a, &ph, &pl); }                            // VIMNMX w/ preds unused
        else if (OP == 12) { bool ph, pl; v = __vibmax_s16x2(v, w, &ph, &pl); y[i] = ph ? y[(i + 1) % CHAINS] : y[i]; z[i] = pl ? z[(i + 1) % CHAINS] : z[i]; asm volatile("" : "+r"(y[i]), "+r"(z[i])); } // VIMNMX + 2 SEL
        else if (OP == 13) { v = v + a; v = v * a + b; }          // add + imad pair
        else if (OP == 14) { v = v + a; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b)); } // add + lop3
        else if (OP == 15) { v = v * a + b; asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v) : "r"(a), "r"(b)); } // imad + lop3
        else if (OP == 16) { v = __vmaxs2(v, a); v = v * a + b; } // vimnmx + imad
        else if (OP == 17) { v = __vmaxs2(v, a); v = p0 ? w : v; } // vimnmx + sel
        else if (OP == 18) { __half2 h = *(__half2*)&v, k = *(__half2*)&a; h = __hadd2(h, k); v = *(uint32_t*)&h; v = __vmaxs2(v, b); } // hadd2 + vimnmx
        else if (OP == 20) { uint32_t c1 = v + a; asm volatile("" : "+r"(c1)); uint32_t c2 = w + b; asm volatile("" : "+r"(c2)); bool ph, pl; v = __vibmax_s16x2(c1, c2, &ph, &pl); y[i] = ph ? y[(i + 1) % CHAINS] : y[i]; z[i] = pl ? z[(i + 1) % CHAINS] : z[i]; asm volatile("" : "+r"(y[i]), "+r"(z[i])); } // ACS unit: 2 add + VIMNMX + 2 SEL
        else if (OP == 19) { __half2 h = *(__half2*)&v, k = *(__half2*)&a; h = __hfma2(h, k, k); v = *(uint32_t*)&h; } // HFMA2
        asm volatile("" : "+r"(v));
    }
}

template <int OP>
__global__ void bench(uint32_t* out, uint32_t a, uint32_t b, long long* cycles) {
    uint32_t x[CHAINS], y[CHAINS], z[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { x[i] = threadIdx.x * 7 + i; y[i] = a * i + threadIdx.x; z[i] = b * i ^ threadIdx.x; }
    bool p0 = (a & 1), p1 = (b & 1);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        op<OP>(x, y, z, a, b, p0, p1); op<OP>(x, y, z, a, b, p0, p1); op<OP>(x, y, z, a, b, p0, p1); op<OP>(x, y, z, a, b, p0, p1);
    }
    long long t1 = clock64();
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s ^= x[i] ^ y[i] ^ z[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_unit) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    printf("%-28s", name);
    for (int nw : {4, 8, 16, 32}) {
        bench<OP><<<1, nw * 32>>>(out, 3, 5, cyc);
        bench<OP><<<1, nw * 32>>>(out, 3, 5, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double winst = (double)nw * ITERS * 4 * CHAINS * ops_per_unit;
        printf("  nw=%2d: %6.3f wi/clk/SM", nw, winst / c);
    }
    printf("\n");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("add (IADD3|IMAD.IADD)", 1);
    run<1>("IMAD", 1);
    run<2>("LOP3", 1);
    run<3>("SEL", 1);
    run<4>("VIMNMX.S16x2", 1);
    run<5>("VIADD.16x2", 1);
    run<6>("PRMT", 1);
    run<7>("SHFL.BFLY", 1);
    run<8>("HADD2", 1);
    run<9>("HMNMX2", 1);
    run<10>("HSET2+LOP3", 2);
    run<11>("VIMNMX.S16x2 (preds dead)", 1);
    run<12>("VIMNMX+P,P + 2xSEL", 3);
    run<13>("add + IMAD", 2);
    run<14>("add + LOP3", 2);
    run<15>("IMAD + LOP3", 2);
    run<16>("VIMNMX + IMAD", 2);
    run<17>("VIMNMX + SEL", 2);
    run<18>("HADD2 + VIMNMX", 2);
    run<19>("HFMA2", 1);
    run<20>("ACS unit 2add+VIMNMX+2SEL", 5);
    return 0;
}
