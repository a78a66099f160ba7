// stripped_bound.cu -- measured upper bounds for the register-exchange design ("next round" item 4): the product
// kernel source compiled with parts stubbed out (-DVIT_STRIP_BUILD: operand table built once and reused; -DVIT_STRIP_TRACEBACK:
// ring flush only, no traceback / output store; -DVIT_STRIP_SHFL: no half exchanges) and timed on the headline workload shape
// (one 32 Mbit s4 stream, int16x2 core, 1600 one-warp blocks).  The stubbed builds decode garbage -- only their time means
// anything: it bounds what removing that part entirely could buy.  -DVIT_STAGGER_NS=<ns> delays warps by (id % 3) * ns.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr [-D...] -o stripped scripts/stripped_bound.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../gpu-accelerated-viterbi-decoder_b200/csrc/vit_kernel.cuh"

#ifndef SB_MET
#define SB_MET vitk::MET_B16
#endif
#ifndef SB_IN
#define SB_IN vitk::IN_S4
#endif
#ifndef SB_TBL
#define SB_TBL 96
#endif
#ifndef SB_STREAMS
#define SB_STREAMS 1
#endif

int main() {
    const size_t n_bits = 32000000, N = 2 * n_bits;
    const size_t M = (n_bits - 64) / 32 * 32;
    const size_t in_bytes = SB_IN == vitk::IN_HARD ? N / 8 : SB_IN == vitk::IN_S4 ? N / 2 : SB_IN == vitk::IN_S8 ? N : SB_IN == vitk::IN_S16 ? 2 * N : 4 * N;
    const size_t in_stride = (in_bytes + 255) / 256 * 256, out_stride = (M / 8 + 255) / 256 * 256;
    const int nbuf = SB_STREAMS > 1 ? 2 : 8;
    uint8_t *in_d, *out_d;
    cudaMalloc(&in_d, in_stride * SB_STREAMS * nbuf + 256);
    cudaMalloc(&out_d, out_stride * SB_STREAMS + 256);
    std::vector<uint8_t> h(in_stride * SB_STREAMS * nbuf);
    unsigned x = 12345;
    for (auto& b : h) { x = x * 1664525u + 1013904223u; b = (uint8_t)(x >> 24); }
    cudaMemcpy(in_d, h.data(), h.size(), cudaMemcpyHostToDevice);
    vitk::KParams kp{};
    kp.out = out_d; kp.in_stride = in_stride; kp.out_stride = out_stride; kp.in_bytes = in_bytes; kp.packs = M / 32;
    kp.segments = 6400; kp.seg_first = 0; kp.seg_limit = 6400; kp.nstreams = SB_STREAMS; kp.one = 1; kp.gate_n = 0;
    using namespace vitk;
    auto kern = l8::vit_decode_kernel<SB_MET, SB_IN, 32, SB_TBL>;
    const int smem = l8::Smem<SB_IN, SB_TBL>::TOTAL;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid(1600, SB_STREAMS);
    float best = 1e9f, sum = 0.f;
    const int reps = 30;
    for (int r = -5; r < reps; r++) {
        kp.in = in_d + (size_t)((r + 5) % nbuf) * in_stride * SB_STREAMS;
        cudaEventRecord(e0);
        kern<<<grid, 32, smem>>>(kp);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r >= 0) { sum += ms; if (ms < best) best = ms; }
    }
    cudaError_t e = cudaGetLastError();
    printf("%-52s mean %.4f ms  best %.4f ms  -> %.1f Gb/s (mean)  %s\n", SB_NAME, sum / reps, best, (double)M * SB_STREAMS / (sum / reps) / 1e6, cudaGetErrorString(e));
    return 0;
}
