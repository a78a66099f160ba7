// vit_synth.cu -- device-side synthetic channel source: the twin, on the GPU, of the
// reference harness's host chain  RandBitGen | ConvolutionalEncoder | AddNoise | SoftDecisionPacker
// (reference src/viterbiDF.h:20-167) so that the 256 Mbit / 4 Gbit / many-stream configurations do not need a
// host pipeline (minutes of mt19937 and ~100 GB of std::vector in the reference, main.cpp:131-142).
//
// Everything is counter based and integer only, so the CPU twin (oracle.make_channel_det(bits_source="hash"),
// test infrastructure) reproduces it bit for bit:
//   message bit i       = top bit of splitmix64(i + seed * 0xD1B54A32D192ED03)
//   coded symbols 2i,2i+1 = parities of the 7-bit buffer (bit 6 = newest) with 0171 / 0133   (viterbiDF.h:48-60)
//   symbol value (Q8)   = +-(amp << 8) + ((u * sigma_q16) >> 16),  u = sum of the four 16-bit lanes of
//                         splitmix64(j + seed * 0x100000001B3) - 2*65535   (~Gaussian, sd = 0.577 * sigma_q16 / 256)
//   quantise            = floor(value / 256), then the reference's saturating quantiser and MSB-first packing
//                         (viterbiDF.h:105-166); FP32: value / 16 as float
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/vit_b200.h"

namespace {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

struct SynthParams {
    unsigned long long n_bits, seed;
    long long amp_q8, sigma_q16;
    int input_type, zero;
};

__device__ inline int msg_bit(const SynthParams& p, long long i) {
    if (i < 0) return 0;                                   // encoder starts from the all-zero state
    return (int)(splitmix64((uint64_t)i + p.seed * 0xD1B54A32D192ED03ull) >> 63);
}

__device__ inline long long symbol_value(const SynthParams& p, unsigned long long j, int coded) {
    if (p.zero) return 0;
    long long v = (coded ? 1 : -1) * p.amp_q8;
    if (p.sigma_q16) {
        uint64_t h = splitmix64(j + p.seed * 0x100000001B3ull);
        long long u = (long long)((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48)) - 2 * 65535;
        v += (u * p.sigma_q16) >> 16;
    }
    return v >> 8;
}

// one thread = one 32-bit pack (HARD 16 message bits, SOFT4 4, SOFT8 2, SOFT16 1) or one symbol pair (FP32)
__global__ void synth_kernel(SynthParams p, uint32_t* __restrict__ packed, uint8_t* __restrict__ bits_out,
                             unsigned long long n_words) {
    const unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const int bits_per_word = p.input_type == 0 ? 16 : p.input_type == 1 ? 4 : p.input_type == 2 ? 2 : 1;
    const long long i0 = (long long)w * bits_per_word;
    unsigned sr = 0;                                        // encoder buffer, bit 6 = newest
    for (int k = 6; k >= 1; k--) sr = (sr >> 1) | ((unsigned)msg_bit(p, i0 - k) << 6);
    uint32_t word = 0;
    float f[2] = {0.f, 0.f};
    for (int b = 0; b < bits_per_word; b++) {
        const long long i = i0 + b;
        const int u = (unsigned long long)i < p.n_bits ? msg_bit(p, i) : 0;
        if (bits_out && (unsigned long long)i < p.n_bits) bits_out[i] = (uint8_t)u;
        sr = (sr >> 1) | ((unsigned)u << 6);
        const int c[2] = {__popc(sr & 0171) & 1, __popc(sr & 0133) & 1};
        for (int k = 0; k < 2; k++) {
            long long v = (unsigned long long)i < p.n_bits ? symbol_value(p, 2ull * i + k, c[k]) : 0;
            switch (p.input_type) {
                case 0: word = (word << 1) | (v > 0 ? 1u : 0u); break;
                case 1: v = v < -8 ? -8 : v > 7 ? 7 : v; word = (word << 4) | ((uint32_t)v & 0xFu); break;
                case 2: v = v < -128 ? -128 : v > 127 ? 127 : v; word = (word << 8) | ((uint32_t)v & 0xFFu); break;
                case 3: v = v < -32768 ? -32768 : v > 32767 ? 32767 : v; word = (word << 16) | ((uint32_t)v & 0xFFFFu); break;
                default: f[k] = (float)v / 16.0f; break;
            }
        }
    }
    if (p.input_type == 4) {
        reinterpret_cast<float2*>(packed)[w] = make_float2(f[0], f[1]);
    } else {
        packed[w] = word;
    }
}

// decoded bit j (MSB-first inside its pack) vs message bit j + 26 (reference main.cpp:153-169)
template <int BPP>
__global__ void count_errors_kernel(const void* __restrict__ out, const uint8_t* __restrict__ bits, unsigned long long n_packs,
                                    unsigned long long* __restrict__ total) {
    unsigned long long errs = 0;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < n_packs;
         w += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t word = BPP == 16 ? static_cast<const uint16_t*>(out)[w] : static_cast<const uint32_t*>(out)[w];
        uint32_t gen = 0;
        const uint8_t* b = bits + w * BPP + 26;
#pragma unroll 8
        for (int k = 0; k < BPP; k++) gen = (gen << 1) | (b[k] & 1u);
        errs += __popc(word ^ gen);
    }
    for (int d = 16; d > 0; d >>= 1) errs += __shfl_down_sync(0xffffffffu, errs, d);
    if ((threadIdx.x & 31) == 0 && errs) atomicAdd(total, errs);
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

// Fill packed_d (vit_input_size(options, 2*n_bits) bytes, rounded up to whole 32-bit packs) with a synthetic
// received stream for n_bits message bits; bits_d (optional, n_bits bytes) receives the message bits for BER.
// amp: symbol amplitude in quantiser units (0 = default per input type); sigma: noise sd relative to amp.
int vit_synth_device(int input_type, size_t n_bits, unsigned seed, int amp, double sigma, int zero,
                     void* packed_d, void* bits_d, void* cuda_stream) {
    if (input_type < 0 || input_type > 4 || !packed_d) return VIT_ERR_ARG;
    static const int def_amp[5] = {64, 3, 40, 9000, 48};
    if (amp <= 0) amp = def_amp[input_type];
    SynthParams p;
    p.n_bits = n_bits; p.seed = seed; p.input_type = input_type; p.zero = zero;
    p.amp_q8 = (long long)amp << 8;
    p.sigma_q16 = (long long)(sigma * amp * 256.0 / 0.57735 + 0.5);
    const int bits_per_word = input_type == 0 ? 16 : input_type == 1 ? 4 : input_type == 2 ? 2 : 1;
    const unsigned long long n_words = (n_bits + bits_per_word - 1) / bits_per_word;
    if (n_words == 0) return VIT_OK;
    const unsigned threads = 256;
    const unsigned long long blocks = (n_words + threads - 1) / threads;
    if (blocks > 0x7fffffffull) return VIT_ERR_ARG;
    synth_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        p, static_cast<uint32_t*>(packed_d), static_cast<uint8_t*>(bits_d), n_words);
    return cudaGetLastError() == cudaSuccess ? VIT_OK : VIT_ERR_CUDA;
}

// Bit errors of a decoded stream against the message bits (one byte per bit, e.g. from vit_synth_device),
// counted on the device: errors = #{ j < messageLen : out bit j != bits[j + 26] }.  Synchronous.
int vit_count_errors_device(int options, const void* out_d, const void* bits_d, size_t messageLen,
                            unsigned long long* errors, void* cuda_stream) {
    if (!out_d || !bits_d || !errors) return VIT_ERR_ARG;
    const int bpp = ((options >> 8) & 0xf) == 1 ? 16 : 32;
    unsigned long long* acc = nullptr;
    if (cudaMalloc(&acc, sizeof *acc) != cudaSuccess) return VIT_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    cudaMemsetAsync(acc, 0, sizeof *acc, st);
    const unsigned long long n_packs = messageLen / bpp;
    if (n_packs) {
        const unsigned blocks = (unsigned)((n_packs + 255) / 256 < 148 * 16 ? (n_packs + 255) / 256 : 148 * 16);
        if (bpp == 16) count_errors_kernel<16><<<blocks, 256, 0, st>>>(out_d, static_cast<const uint8_t*>(bits_d), n_packs, acc);
        else count_errors_kernel<32><<<blocks, 256, 0, st>>>(out_d, static_cast<const uint8_t*>(bits_d), n_packs, acc);
    }
    cudaError_t e = cudaMemcpyAsync(errors, acc, sizeof *acc, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(acc);
    return e == cudaSuccess ? VIT_OK : VIT_ERR_CUDA;
}

// plain device-memory helpers so that host-only callers (no CUDA headers) can use the device-resident entry points
int vit_dev_alloc(void** ptr, size_t bytes) { return cudaMalloc(ptr, bytes ? bytes : 1) == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
void vit_dev_free(void* ptr) { if (ptr) cudaFree(ptr); }
int vit_dev_sync(void) { return cudaDeviceSynchronize() == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
int vit_dev_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
// page-locked host memory for callers without CUDA headers: vit_run takes its time-sliced upload path from such buffers
int vit_host_alloc(void** ptr, size_t bytes) { return cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
void vit_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

#pragma GCC visibility pop
}
