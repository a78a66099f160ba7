// upload_pattern_probe.cu -- how fast can vit_run's time-sliced upload go?  Replays the copy pattern of the headline
// configuration (32 Mbit s4: 6400 segments, 1598 of them 157 packs = 5024 B, the rest 156 packs = 4992 B) in variants:
// slices per segment, one or two copy streams (one per segment group), gate flags by 4-byte memcpy or by stream write-value.
// nvcc -O3 -o upload_pattern_probe upload_pattern_probe.cu
#include <cstdio>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*WriteValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

int main() {
    const size_t W = 6400, q = 156, r = 1598, pack = 32;
    const size_t pitch1 = (q + 1) * pack, pitch2 = q * pack, grp2 = r * pitch1, total = grp2 + (W - r) * pitch2 + 2048;
    char *h, *d; unsigned *flag_h, *flag_d;
    cudaHostAlloc(&h, total + 4096, cudaHostAllocDefault);
    cudaHostAlloc(&flag_h, 64, cudaHostAllocDefault); *flag_h = 1;
    cudaMalloc(&d, total + 4096); cudaMalloc(&flag_d, 256);
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
    WriteValue32Fn wv = nullptr;
    cudaDriverEntryPointQueryResult qres;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess) wv = (WriteValue32Fn)fn;
    printf("cuStreamWriteValue32 %s\n", wv ? "available" : "NOT available");
    struct Plan { const char* name; std::vector<double> cuts; };
    std::vector<Plan> plans = {
        {"1 slice (contiguous 1D copy)", {}},
        {"4 slices 1/2 1/4 1/8 1/8", {0.5, 0.75, 0.875}},
        {"3 slices 1/2 3/8 1/8", {0.5, 0.875}},
        {"3 slices 5/8 1/4 1/8", {0.625, 0.875}},
        {"2 slices 3/4 1/4", {0.75}},
        {"5 slices 1/2 1/4 1/8 1/16 1/16", {0.5, 0.75, 0.875, 0.9375}},
    };
    for (auto& p : plans) {
        for (int two = 0; two < 2; two++) for (int usewv = 0; usewv < 2; usewv++) {
            if (usewv && !wv) continue;
            if (p.cuts.empty() && (two || usewv)) continue;
            float best = 1e9;
            for (int rep = 0; rep < 6; rep++) {
                cudaStream_t sa = s1, sb = two ? s2 : s1;
                cudaEventRecord(e0, s1);
                if (two) cudaStreamWaitEvent(s2, e0, 0);
                if (p.cuts.empty()) cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, s1);
                else {
                    size_t nsl = p.cuts.size() + 1;
                    for (size_t g = 0; g < nsl; g++) {
                        size_t lo = g == 0 ? 0 : (size_t)(p.cuts[g - 1] * pitch2) / 96 * 96 - 16;
                        size_t hi = g + 1 == nsl ? (size_t)-1 : (size_t)(p.cuts[g] * pitch2) / 96 * 96 + 48;
                        size_t w1 = hi < pitch1 ? hi : pitch1, w2 = hi < pitch2 ? hi : pitch2;
                        cudaMemcpy2DAsync(d + lo, pitch1, h + lo, pitch1, w1 - lo, r, cudaMemcpyHostToDevice, sa);
                        if (usewv) wv((CUstream)sa, (CUdeviceptr)(flag_d + g), 1, 0); else cudaMemcpyAsync(flag_d + g, flag_h, 4, cudaMemcpyHostToDevice, sa);
                        cudaMemcpy2DAsync(d + grp2 + lo, pitch2, h + grp2 + lo, pitch2, w2 - lo, W - r, cudaMemcpyHostToDevice, sb);
                        if (two) { if (usewv) wv((CUstream)sb, (CUdeviceptr)(flag_d + 16 + g), 1, 0); else cudaMemcpyAsync(flag_d + 16 + g, flag_h, 4, cudaMemcpyHostToDevice, sb); }
                    }
                }
                if (two) { cudaEventRecord(e2, s2); cudaStreamWaitEvent(s1, e2, 0); }
                cudaEventRecord(e1, s1);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("%-34s %s  flags by %-10s : %.3f ms = %.1f GB/s\n", p.name, two ? "2 copy streams" : "1 copy stream ", usewv ? "writeValue" : "memcpy", best, total / best / 1e6);
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
