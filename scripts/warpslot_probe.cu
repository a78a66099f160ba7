// warpslot_probe.cu -- where do the two warps of 64-thread blocks land?  Records (%smid, %warpid) of both warps of
// every block of a 1600-block launch with the warp-specialised kernel's footprint (13.4 KB smem, 80 regs).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warpslot_probe warpslot_probe.cu
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(64) probe(unsigned* out, long long spin) {
    extern __shared__ unsigned char sm[];
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    long long t0 = clock64();
    while (clock64() - t0 < spin) { sm[threadIdx.x] = (unsigned char)wid; }
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * 2 + (threadIdx.x >> 5)) * 2] = smid; out[(blockIdx.x * 2 + (threadIdx.x >> 5)) * 2 + 1] = wid; }
}
int main() {
    const int nb = 1600;
    unsigned* d; cudaMalloc(&d, nb * 4 * 4);
    probe<<<nb, 64, 13400>>>(d, 2000000);
    cudaDeviceSynchronize();
    std::vector<unsigned> h(nb * 4);
    cudaMemcpy(h.data(), d, nb * 16, cudaMemcpyDeviceToHost);
    std::map<int, int> delta, sched0, sched1, rule;
    std::map<unsigned, std::vector<int>> per_sm;
    for (int b = 0; b < nb; b++) {
        unsigned s0 = h[b * 4], w0 = h[b * 4 + 1], s1 = h[b * 4 + 2], w1 = h[b * 4 + 3];
        delta[(int)w1 - (int)w0]++;
        sched0[w0 & 3]++; sched1[w1 & 3]++;
        unsigned decoder = (w0 >> 2) & 1;
        unsigned wd = decoder ? w1 : w0;
        rule[wd & 3]++;
        per_sm[s0].push_back(wd & 3);
        if (s0 != s1) printf("block %d on two SMs?!\n", b);
    }
    printf("warpid(warp1) - warpid(warp0):"); for (auto& kv : delta) printf("  %d: %d blocks", kv.first, kv.second); printf("\n");
    printf("warp0 slot%%4:"); for (auto& kv : sched0) printf("  %d: %d", kv.first, kv.second); printf("\n");
    printf("warp1 slot%%4:"); for (auto& kv : sched1) printf("  %d: %d", kv.first, kv.second); printf("\n");
    printf("decoder slot%%4 with rule (warpid0>>2)&1:"); for (auto& kv : rule) printf("  %d: %d", kv.first, kv.second); printf("\n");
    std::map<int, int> worst;
    for (auto& kv : per_sm) { int c[4] = {0, 0, 0, 0}; for (int v : kv.second) c[v]++; int m = 0; for (int i = 0; i < 4; i++) m = c[i] > m ? c[i] : m; worst[m]++; }
    printf("max decoders on one scheduler, per SM:"); for (auto& kv : worst) printf("  %d: %d SMs", kv.first, kv.second); printf("\n");
    // print the slot list of SM 0
    printf("SM %u blocks (warpid0,warpid1):", per_sm.begin()->first);
    for (int b = 0; b < nb; b++) if (h[b * 4] == per_sm.begin()->first) printf(" (%u,%u)", h[b * 4 + 1], h[b * 4 + 3]);
    printf("\n");
    return 0;
}
