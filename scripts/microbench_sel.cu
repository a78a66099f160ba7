// microbench_sel.cu -- which pipes can carry the register-exchange survivor selects?
// Each kernel runs NW warps (1 block) of 8 independent chains on one SM and reports warp-instructions
// per clock per SM.  Select flavours: SEL (ALU pipe), predicated IMAD move (FMA-heavy pipe), FSEL
// (selp.f32: which pipe?), and the full ACS unit (2 adds + VIMNMX.S16x2 with both predicates + 2 selects)
// with the selects spread over them.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_sel microbench_sel.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512
#define CHAINS 8

// ACS unit: c1 = v*one + a ; c2 = w*one + b ; v = max.s16x2(c1, c2) with predicates ; two selects of flavour S0, S1
// flavours: 0 SEL, 1 @p IMAD, 2 FSEL, 3 @p MOV
template <int S>
__device__ __forceinline__ void sel(uint32_t& dst, uint32_t src, uint32_t one, int which);

template <int OP>
__device__ __forceinline__ void op(uint32_t (&x)[CHAINS], uint32_t (&y)[CHAINS], uint32_t (&z)[CHAINS], uint32_t a, uint32_t b, uint32_t one) {
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        uint32_t& v = x[i];
        uint32_t w = x[(i + 1) % CHAINS];
        uint32_t ys = y[(i + 1) % CHAINS], zs = z[(i + 1) % CHAINS];
        if (OP == 0) {   // FSEL alone (predicate from a loop-invariant compare)
            asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(*(float*)&v) : "f"(*(float*)&w), "r"(a));
        } else if (OP == 1) {   // SEL alone
            asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %1, %0, p;}" : "+r"(v) : "r"(w), "r"(a));
        } else if (OP == 2) {   // FSEL + SEL
            asm volatile("{.reg .pred p; setp.ne.u32 p, %4, 0; selp.f32 %0, %2, %0, p; selp.b32 %1, %3, %1, p;}" : "+f"(*(float*)&v), "+r"(y[i]) : "f"(*(float*)&w), "r"(ys), "r"(a));
        } else if (OP == 3) {   // FSEL + IMAD
            asm volatile("{.reg .pred p; setp.ne.u32 p, %4, 0; selp.f32 %0, %2, %0, p; mad.lo.u32 %1, %1, %3, %4;}" : "+f"(*(float*)&v), "+r"(y[i]) : "f"(*(float*)&w), "r"(one), "r"(a));
        } else if (OP == 4) {   // FSEL + SEL + IMAD
            asm volatile("{.reg .pred p; setp.ne.u32 p, %6, 0; selp.f32 %0, %3, %0, p; selp.b32 %1, %4, %1, p; mad.lo.u32 %2, %2, %5, %6;}" : "+f"(*(float*)&v), "+r"(y[i]), "+r"(z[i]) : "f"(*(float*)&w), "r"(ys), "r"(one), "r"(a));
        } else if (OP >= 10 && OP < 20) {
            // ACS unit, selects: 10: SEL,SEL  11: SEL,@IMAD  12: FSEL,@IMAD  13: FSEL,FSEL  14: FSEL,SEL  15: @IMAD,@IMAD
            uint32_t r;
            if (OP == 10)
                asm volatile("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3; .reg .b32 c1, c2;\n\t"
                    "mad.lo.u32 c1, %0, %7, %5; mad.lo.u32 c2, %4, %7, %6;\n\t"
                    "max.s16x2 %0, c1, c2; mov.b32 {r0, r1}, %0; mov.b32 {r2, r3}, c1; setp.eq.s16 pv, r0, r2; setp.eq.s16 pu, r1, r3;\n\t"
                    "selp.b32 %1, %8, %1, pv; selp.b32 %2, %9, %2, pu;}"
                    : "+r"(v), "+r"(y[i]), "+r"(z[i]), "=r"(r) : "r"(w), "r"(a), "r"(b), "r"(one), "r"(ys), "r"(zs));
            else if (OP == 11)
                asm volatile("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3; .reg .b32 c1, c2;\n\t"
                    "mad.lo.u32 c1, %0, %7, %5; mad.lo.u32 c2, %4, %7, %6;\n\t"
                    "max.s16x2 %0, c1, c2; mov.b32 {r0, r1}, %0; mov.b32 {r2, r3}, c1; setp.eq.s16 pv, r0, r2; setp.eq.s16 pu, r1, r3;\n\t"
                    "selp.b32 %1, %8, %1, pv; @pu mad.lo.u32 %2, %9, %7, 0;}"
                    : "+r"(v), "+r"(y[i]), "+r"(z[i]), "=r"(r) : "r"(w), "r"(a), "r"(b), "r"(one), "r"(ys), "r"(zs));
            else if (OP == 12)
                asm volatile("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3; .reg .b32 c1, c2; .reg .f32 f1, f2;\n\t"
                    "mad.lo.u32 c1, %0, %7, %5; mad.lo.u32 c2, %4, %7, %6;\n\t"
                    "max.s16x2 %0, c1, c2; mov.b32 {r0, r1}, %0; mov.b32 {r2, r3}, c1; setp.eq.s16 pv, r0, r2; setp.eq.s16 pu, r1, r3;\n\t"
                    "mov.b32 f1, %1; mov.b32 f2, %8; selp.f32 f1, f2, f1, pv; mov.b32 %1, f1; @pu mad.lo.u32 %2, %9, %7, 0;}"
                    : "+r"(v), "+r"(y[i]), "+r"(z[i]), "=r"(r) : "r"(w), "r"(a), "r"(b), "r"(one), "r"(ys), "r"(zs));
            else if (OP == 13)
                asm volatile("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3; .reg .b32 c1, c2; .reg .f32 f1, f2, f3, f4;\n\t"
                    "mad.lo.u32 c1, %0, %7, %5; mad.lo.u32 c2, %4, %7, %6;\n\t"
                    "max.s16x2 %0, c1, c2; mov.b32 {r0, r1}, %0; mov.b32 {r2, r3}, c1; setp.eq.s16 pv, r0, r2; setp.eq.s16 pu, r1, r3;\n\t"
                    "mov.b32 f1, %1; mov.b32 f2, %8; selp.f32 f1, f2, f1, pv; mov.b32 %1, f1; mov.b32 f3, %2; mov.b32 f4, %9; selp.f32 f3, f4, f3, pu; mov.b32 %2, f3;}"
                    : "+r"(v), "+r"(y[i]), "+r"(z[i]), "=r"(r) : "r"(w), "r"(a), "r"(b), "r"(one), "r"(ys), "r"(zs));
            else if (OP == 14)
                asm volatile("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3; .reg .b32 c1, c2; .reg .f32 f1, f2;\n\t"
                    "mad.lo.u32 c1, %0, %7, %5; mad.lo.u32 c2, %4, %7, %6;\n\t"
                    "max.s16x2 %0, c1, c2; mov.b32 {r0, r1}, %0; mov.b32 {r2, r3}, c1; setp.eq.s16 pv, r0, r2; setp.eq.s16 pu, r1, r3;\n\t"
                    "mov.b32 f1, %1; mov.b32 f2, %8; selp.f32 f1, f2, f1, pv; mov.b32 %1, f1; selp.b32 %2, %9, %2, pu;}"
                    : "+r"(v), "+r"(y[i]), "+r"(z[i]), "=r"(r) : "r"(w), "r"(a), "r"(b), "r"(one), "r"(ys), "r"(zs));
            else if (OP == 15)
                asm volatile("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3; .reg .b32 c1, c2;\n\t"
                    "mad.lo.u32 c1, %0, %7, %5; mad.lo.u32 c2, %4, %7, %6;\n\t"
                    "max.s16x2 %0, c1, c2; mov.b32 {r0, r1}, %0; mov.b32 {r2, r3}, c1; setp.eq.s16 pv, r0, r2; setp.eq.s16 pu, r1, r3;\n\t"
                    "@pv mad.lo.u32 %1, %8, %7, 0; @pu mad.lo.u32 %2, %9, %7, 0;}"
                    : "+r"(v), "+r"(y[i]), "+r"(z[i]), "=r"(r) : "r"(w), "r"(a), "r"(b), "r"(one), "r"(ys), "r"(zs));
            else if (OP == 16)   // adds on the ALU (VIADD.16x2-style add.s16x2?) : plain add -> ptxas picks IADD3 or IMAD.IADD
                asm volatile("{.reg .pred pu, pv; .reg .s16 r0, r1, r2, r3; .reg .b32 c1, c2; .reg .f32 f1, f2;\n\t"
                    "add.u32 c1, %0, %5; mad.lo.u32 c2, %4, %7, %6;\n\t"
                    "max.s16x2 %0, c1, c2; mov.b32 {r0, r1}, %0; mov.b32 {r2, r3}, c1; setp.eq.s16 pv, r0, r2; setp.eq.s16 pu, r1, r3;\n\t"
                    "mov.b32 f1, %1; mov.b32 f2, %8; selp.f32 f1, f2, f1, pv; mov.b32 %1, f1; @pu mad.lo.u32 %2, %9, %7, 0;}"
                    : "+r"(v), "+r"(y[i]), "+r"(z[i]), "=r"(r) : "r"(w), "r"(a), "r"(b), "r"(one), "r"(ys), "r"(zs));
        }
        asm volatile("" : "+r"(v));
    }
}

template <int OP>
__global__ void bench(uint32_t* out, uint32_t a, uint32_t b, uint32_t one, long long* cycles) {
    uint32_t x[CHAINS], y[CHAINS], z[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { x[i] = threadIdx.x * 7 + i; y[i] = a * i + threadIdx.x; z[i] = b * i ^ threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        op<OP>(x, y, z, a, b, one); op<OP>(x, y, z, a, b, one); op<OP>(x, y, z, a, b, one); op<OP>(x, y, z, a, b, one);
    }
    long long t1 = clock64();
    __syncthreads();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s ^= x[i] ^ y[i] ^ z[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cycles = t1 - t0;
}

// FSEL must move all 32 bits unchanged (NaN payloads, denormals, -0): checked on every bit pattern class
__global__ void fsel_bits(const uint32_t* in, uint32_t* out, int n, int pick) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t a = in[i], b = ~in[i], r;
    asm volatile("{.reg .pred p; .reg .f32 fa, fb; setp.ne.s32 p, %3, 0; mov.b32 fa, %1; mov.b32 fb, %2; selp.f32 fa, fa, fb, p; mov.b32 %0, fa;}" : "=r"(r) : "r"(a), "r"(b), "r"(pick));
    out[i] = r;
}

template <int OP>
void run(const char* name, int ops_per_unit) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    printf("%-34s", name);
    for (int nw : {4, 8, 12, 16, 32}) {
        bench<OP><<<1, nw * 32>>>(out, 3, 5, 1, cyc);
        bench<OP><<<1, nw * 32>>>(out, 3, 5, 1, cyc);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double winst = (double)nw * ITERS * 4 * CHAINS * ops_per_unit;
        printf("  nw=%2d: %6.3f", nw, winst / c);
    }
    printf("   wi/clk/SM\n");
    cudaFree(out); cudaFree(cyc);
}

int main() {
    // bit-exactness of FSEL as a move
    {
        const int n = 1 << 20;
        uint32_t* h = (uint32_t*)malloc(n * 4), *g = (uint32_t*)malloc(n * 4);
        uint32_t s = 12345;
        for (int i = 0; i < n; i++) {
            s = s * 1664525u + 1013904223u;
            uint32_t v = s;
            if ((i & 7) == 1) v = 0x7f800000u | (s & 0x7fffffu) | (s & 0x80000000u);   // NaN / Inf payloads
            if ((i & 7) == 2) v = (s & 0x807fffffu);                                     // denormals, +-0
            h[i] = v;
        }
        uint32_t *din, *dout; cudaMalloc(&din, n * 4); cudaMalloc(&dout, n * 4);
        cudaMemcpy(din, h, n * 4, cudaMemcpyHostToDevice);
        long bad = 0;
        for (int pick = 0; pick < 2; pick++) {
            fsel_bits<<<n / 256, 256>>>(din, dout, n, pick);
            cudaMemcpy(g, dout, n * 4, cudaMemcpyDeviceToHost);
            for (int i = 0; i < n; i++) bad += g[i] != (pick ? h[i] : ~h[i]);
        }
        printf("FSEL bit-exact move over %d patterns (NaN payloads, denormals, -0): %ld mismatches\n", 2 * n, bad);
    }
    run<0>("FSEL", 1);
    run<1>("SEL", 1);
    run<2>("FSEL + SEL", 2);
    run<3>("FSEL + IMAD", 2);
    run<4>("FSEL + SEL + IMAD", 3);
    run<10>("ACS: 2IMAD VIMNMX SEL SEL", 5);
    run<11>("ACS: 2IMAD VIMNMX SEL @IMAD", 5);
    run<12>("ACS: 2IMAD VIMNMX FSEL @IMAD", 5);
    run<13>("ACS: 2IMAD VIMNMX FSEL FSEL", 5);
    run<14>("ACS: 2IMAD VIMNMX FSEL SEL", 5);
    run<15>("ACS: 2IMAD VIMNMX @IMAD @IMAD", 5);
    run<16>("ACS: add IMAD VIMNMX FSEL @IMAD", 5);
    return 0;
}
