// memcpy2d_probe.cu -- pinned host -> device bandwidth of strided (2D) copies: rows of `width` bytes every `pitch` bytes.
// Decides whether a time-sliced upload (a column block of every stream segment per copy) can run at PCIe speed.
// nvcc -O3 -o memcpy2d_probe memcpy2d_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    const size_t pitch = 5000, rows = 6400, total = pitch * rows;
    char *h, *d;
    cudaHostAlloc(&h, total + 4096, cudaHostAllocDefault);
    cudaMalloc(&d, total + 4096);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t width : {64, 128, 256, 512, 625, 1000, 1250, 2500, 5000}) {
        const int nslice = (int)(pitch / width);
        float best = 1e9;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0, st);
            for (int s = 0; s < nslice; s++)
                cudaMemcpy2DAsync(d + s * width, pitch, h + s * width, pitch, width, rows, cudaMemcpyHostToDevice, st);
            cudaEventRecord(e1, st);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("width %5zu B x %zu rows x %2d slices: %.3f ms = %.1f GB/s\n", width, rows, nslice, best, (double)width * rows * nslice / best / 1e6);
    }
    float best = 1e9;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0, st);
        cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, st);
        cudaEventRecord(e1, st); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("1D %zu B: %.3f ms = %.1f GB/s\n", total, best, total / best / 1e6);
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
