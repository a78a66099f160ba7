// l1_bench.cu -- measurement harness of the one-lane-per-segment geometry (csrc/vit_kernel_l1.inc), which the library does not
// instantiate yet.  Result of its one run on a B200: profiles/r2_l1_bench.txt (119.2 against 108.8 Gb/s, outputs identical).
// Decodes L1_STREAMS random streams of L1_BITS message bits (s8 input, int16x2 core, 32-bit packs: the shape of BASELINE.json
// configs[4]) in ONE launch with the product's 8-lane kernel (TBL=32 build, as the library picks for such launches) and with
// the l1 kernel, compares the two outputs word for word and prints both times.  Random bytes are valid s8 symbols; both
// kernels must agree on them exactly (ties included).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -o l1_bench scripts/l1_bench.cu && ./l1_bench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../gpu-accelerated-viterbi-decoder_b200/csrc/vit_kernel.cuh"

#ifndef L1_STREAMS
#define L1_STREAMS 16
#endif
#ifndef L1_BITS
#define L1_BITS 32000000
#endif
#ifndef L1_IN
#define L1_IN vitk::IN_S8
#endif

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { printf("%s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main() {
    using namespace vitk;
    const size_t n_bits = L1_BITS, N = 2 * n_bits, M = (n_bits - 64) / 32 * 32;
    const size_t in_bytes = L1_IN == IN_HARD ? N / 8 : L1_IN == IN_S4 ? N / 2 : L1_IN == IN_S8 ? N : L1_IN == IN_S16 ? 2 * N : 4 * N;
    const size_t in_stride = (in_bytes + 255) / 256 * 256, out_stride = (M / 8 + 255) / 256 * 256;
    uint8_t *in_d, *out8_d, *out1_d;
    CK(cudaMalloc(&in_d, in_stride * L1_STREAMS + 256));
    CK(cudaMalloc(&out8_d, out_stride * L1_STREAMS + 256));
    CK(cudaMalloc(&out1_d, out_stride * L1_STREAMS + 256));
    {
        // one stream of random bytes from the host; the others are the same bytes rotated by 4096 * s (distinct data per stream
        // without minutes of host random numbers)
        std::vector<uint32_t> h(in_stride / 4);
        unsigned x = 12345;
        for (auto& w : h) { x = x * 1664525u + 1013904223u; w = x ^ (x >> 15); }
        CK(cudaMemcpy(in_d, h.data(), in_stride, cudaMemcpyHostToDevice));
        for (size_t s = 1; s < L1_STREAMS; s++) {
            const size_t rot = (4096 * s) % in_stride;
            CK(cudaMemcpy(in_d + s * in_stride, in_d + rot, in_stride - rot, cudaMemcpyDeviceToDevice));
            CK(cudaMemcpy(in_d + s * in_stride + (in_stride - rot), in_d, rot, cudaMemcpyDeviceToDevice));
        }
    }
    CK(cudaMemset(out8_d, 0xEE, out_stride * L1_STREAMS));
    CK(cudaMemset(out1_d, 0xEE, out_stride * L1_STREAMS));
    KParams kp{};
    kp.in = in_d; kp.in_stride = in_stride; kp.out_stride = out_stride; kp.in_bytes = in_bytes; kp.packs = M / 32;
    kp.segments = 6400; kp.seg_first = 0; kp.seg_limit = 6400; kp.nstreams = L1_STREAMS; kp.one = 1; kp.gate_n = 0;
    auto k8 = l8::vit_decode_kernel<MET_B16, L1_IN, 32, 32>;
    auto k1 = l1::vit_decode_kernel_l1<MET_B16, L1_IN, 32>;
    const int smem8 = l8::Smem<L1_IN, 32>::TOTAL, smem1 = l1::SmemL1<L1_IN>::TOTAL;
    CK(cudaFuncSetAttribute(k8, cudaFuncAttributeMaxDynamicSharedMemorySize, smem8));
    CK(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    int occ8 = 0, occ1 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ8, k8, 32, smem8);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k1, 32, smem1);
    setvbuf(stdout, nullptr, _IONBF, 0);
    printf("%d streams x %zu bits; resident warps per SM: 8-lane kernel %d (%d B smem), l1 kernel %d (%d B smem)\n", L1_STREAMS, n_bits, occ8, smem8, occ1, smem1);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const dim3 grid8(1600, L1_STREAMS), grid1(200, L1_STREAMS);
    float t8 = 0.f, t1 = 0.f;
    const int reps = 3;
    for (int r = -1; r < reps; r++) {
        float ms;
        kp.out = out8_d;
        CK(cudaEventRecord(e0)); k8<<<grid8, 32, smem8>>>(kp); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if (r >= 0) t8 += ms;
        kp.out = out1_d;
        CK(cudaEventRecord(e0)); k1<<<grid1, 32, smem1>>>(kp); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if (r >= 0) t1 += ms;
    }
    CK(cudaGetLastError());
    std::vector<uint8_t> a(out_stride * L1_STREAMS), b(out_stride * L1_STREAMS);
    CK(cudaMemcpy(a.data(), out8_d, a.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(b.data(), out1_d, b.size(), cudaMemcpyDeviceToHost));
    size_t diff = 0;
    for (size_t i = 0; i < a.size(); i++) diff += a[i] != b[i];
    const double bits = (double)M * L1_STREAMS;
    printf("8-lane kernel (TBL=32): %.3f ms = %.1f Gb/s\nl1 kernel:              %.3f ms = %.1f Gb/s  (x%.2f)\noutput bytes that differ: %zu of %zu%s\n",
           t8 / reps, bits / (t8 / reps) / 1e6, t1 / reps, bits / (t1 / reps) / 1e6, t8 / t1, diff, a.size(), diff ? "  <-- MISMATCH" : "  (identical)");
    return diff ? 2 : 0;
}
