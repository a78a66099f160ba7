// vit_synth.cu -- device-side synthetic channel source: the twin, on the GPU, of the
// reference harness's host chain  RandBitGen | ConvolutionalEncoder | AddNoise | SoftDecisionPacker
// (reference src/viterbiDF.h:20-167) so that the 256 Mbit / 4 Gbit / many-stream configurations do not need a
// host pipeline (minutes of mt19937 and ~100 GB of std::vector in the reference, main.cpp:131-142).
//
// Everything is counter based and integer only, so the CPU twin (oracle.make_channel_det(bits_source="hash"),
// test infrastructure) reproduces it bit for bit:
//   message bit i       = top bit of splitmix64(i + seed * 0xD1B54A32D192ED03)   (source 0), or
//                         bit i of the PRBS-31 sequence x^31 + x^28 + 1 started from state `seed` (source 1: the
//                         bench's message source; CPU twin vo_prbs31)
//   coded symbols 2i,2i+1 = parities of the 7-bit buffer (bit 6 = newest) with the generator polynomials
//                         (vit_code.h; 0171 / 0133 as in the reference, viterbiDF.h:48-60)
//   symbol value (Q8)   = +-(amp << 8) + ((u * sigma_q16) >> 16),  u = sum of the four 16-bit lanes of
//                         splitmix64(j + seed * 0x100000001B3) - 2*65535   (~Gaussian, sd = 0.577 * sigma_q16 / 256)
//   quantise            = floor(value / 256), then the reference's saturating quantiser and MSB-first packing
//                         (viterbiDF.h:105-166); FP32: value / 16 as float
#include <cuda_runtime.h>
#include <cstdint>

#include "../../include/vit_b200.h"
#include "vit_code.h"

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
    int source;                  // 0: counter hash, 1: PRBS-31
};

// ---- PRBS-31 (x^31 + x^28 + 1), the bench's message source (oracle twin: vo_prbs31) ----------------
// state s (31 bits), one step: nb = s[30] ^ s[27]; s = (s << 1) | nb; output nb.  Bit i of the sequence is the output
// of step i from s_0 = seed.  To start anywhere, the state is advanced by i steps with the precomputed powers
// T^(2^j) of the step matrix (columns = images of the unit vectors).
constexpr int PRBS_POWERS = 40;
__constant__ uint32_t c_prbs_pow[PRBS_POWERS][31];

inline uint32_t prbs_step_host(uint32_t s) {
    const uint32_t nb = ((s >> 30) ^ (s >> 27)) & 1u;
    return ((s << 1) | nb) & 0x7fffffffu;
}
inline uint32_t matvec_host(const uint32_t* col, uint32_t v) {
    uint32_t r = 0;
    for (int k = 0; k < 31; k++) if ((v >> k) & 1u) r ^= col[k];
    return r;
}
cudaError_t upload_prbs_tables() {
    static uint32_t tab[PRBS_POWERS][31];
    static bool done[64] = {};                                       // per device (constant memory is per device)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && done[dev]) return cudaSuccess;
    for (int k = 0; k < 31; k++) tab[0][k] = prbs_step_host(1u << k);
    for (int j = 1; j < PRBS_POWERS; j++)
        for (int k = 0; k < 31; k++) tab[j][k] = matvec_host(tab[j - 1], tab[j - 1][k]);
    const cudaError_t e = cudaMemcpyToSymbol(c_prbs_pow, tab, sizeof tab);
    if (e == cudaSuccess && dev >= 0 && dev < 64) done[dev] = true;
    return e;
}
__device__ inline uint32_t prbs_jump(uint32_t s, unsigned long long steps) {
    for (int j = 0; steps; j++, steps >>= 1) {
        if (!(steps & 1ull)) continue;
        uint32_t r = 0;
#pragma unroll
        for (int k = 0; k < 31; k++) r ^= ((s >> k) & 1u) ? c_prbs_pow[j][k] : 0u;
        s = r;
    }
    return s;
}
__host__ __device__ inline uint32_t prbs_seed_state(unsigned long long seed) {
    uint32_t s = (uint32_t)seed & 0x7fffffffu;
    return s ? s : 0x7fffffffu;
}

// sequential reader of the message bits from index `i` on (i may be negative: the encoder starts from zeros)
struct BitReader {
    const SynthParams& p;
    long long i;
    uint32_t s;
    __device__ BitReader(const SynthParams& p_, long long i0) : p(p_), i(i0), s(0) {
        if (p.source == 1) s = prbs_jump(prbs_seed_state(p.seed), i0 > 0 ? (unsigned long long)i0 : 0ull);
    }
    __device__ int next() {
        int b;
        if (i < 0) b = 0;
        else if (p.source == 1) {
            b = (int)(((s >> 30) ^ (s >> 27)) & 1u);
            s = ((s << 1) | (uint32_t)b) & 0x7fffffffu;
        } else b = (int)(splitmix64((uint64_t)i + p.seed * 0xD1B54A32D192ED03ull) >> 63);
        i++;
        return b;
    }
};

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

// one thread = WPT consecutive 32-bit packs (HARD 16 message bits each, SOFT4 4, SOFT8 2, SOFT16 1) or symbol pairs (FP32)
constexpr int WPT = 8;
__global__ void synth_kernel(SynthParams p, uint32_t* __restrict__ packed, uint8_t* __restrict__ bits_out,
                             unsigned long long n_words) {
    const unsigned long long w0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * WPT;
    if (w0 >= n_words) return;
    const int bits_per_word = p.input_type == 0 ? 16 : p.input_type == 1 ? 4 : p.input_type == 2 ? 2 : 1;
    const long long i0 = (long long)w0 * bits_per_word;
    BitReader rd(p, i0 - 6);
    unsigned sr = 0;                                        // encoder buffer, bit 6 = newest
    for (int k = 0; k < 6; k++) sr = (sr >> 1) | ((unsigned)rd.next() << 6);
    for (int ww = 0; ww < WPT && w0 + ww < n_words; ww++) {
        const unsigned long long w = w0 + ww;
        uint32_t word = 0;
        float f[2] = {0.f, 0.f};
        for (int b = 0; b < bits_per_word; b++) {
            const long long i = (long long)w * bits_per_word + b;
            const bool live = (unsigned long long)i < p.n_bits;
            int u = rd.next();
            if (!live) u = 0;
            if (bits_out && live) bits_out[i] = (uint8_t)u;
            sr = (sr >> 1) | ((unsigned)u << 6);
            const int c[2] = {__popc(sr & VIT_POLY1) & 1, __popc(sr & VIT_POLY2) & 1};
            for (int k = 0; k < 2; k++) {
                long long v = live ? symbol_value(p, 2ull * i + k, c[k]) : 0;
                switch (p.input_type) {
                    case 0: word = (word << 1) | (v > 0 ? 1u : 0u); break;
                    case 1: v = v < -8 ? -8 : v > 7 ? 7 : v; word = (word << 4) | ((uint32_t)v & 0xFu); break;
                    case 2: v = v < -128 ? -128 : v > 127 ? 127 : v; word = (word << 8) | ((uint32_t)v & 0xFFu); break;
                    case 3: v = v < -32768 ? -32768 : v > 32767 ? 32767 : v; word = (word << 16) | ((uint32_t)v & 0xFFFFu); break;
                    default: f[k] = (float)v / 16.0f; break;
                }
            }
        }
        if (p.input_type == 4) reinterpret_cast<float2*>(packed)[w] = make_float2(f[0], f[1]);
        else packed[w] = word;
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

// the same count with the message bits regenerated from the source (no bits buffer): one thread = one pack
template <int BPP>
__global__ void count_errors_synth_kernel(SynthParams p, const void* __restrict__ out, unsigned long long n_packs,
                                          unsigned long long* __restrict__ total) {
    unsigned long long errs = 0;
    const unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n_packs) {
        const uint32_t word = BPP == 16 ? static_cast<const uint16_t*>(out)[w] : static_cast<const uint32_t*>(out)[w];
        BitReader rd(p, (long long)(w * BPP + 26));
        uint32_t gen = 0;
        for (int k = 0; k < BPP; k++) gen = (gen << 1) | (uint32_t)rd.next();
        errs = __popc(word ^ gen);
    }
    for (int d = 16; d > 0; d >>= 1) errs += __shfl_down_sync(0xffffffffu, errs, d);
    if ((threadIdx.x & 31) == 0 && errs) atomicAdd(total, errs);
}

// ---- depuncturing (the part that needs no new trellis): a punctured rate-k/n stream of soft symbols is
// expanded to the rate-1/2 stream the decoder takes, with zero-valued symbols (erasures: they contribute 0 to every branch
// metric) where the transmitter dropped one.  One thread = one 32-bit pack (or one symbol pair for FP32) of the OUTPUT.
struct PunctParams {
    unsigned period;          // stages per puncturing period
    unsigned keep[2];         // bit t of keep[w]: symbol w (0: polynomial 0171, 1: 0133) of stage t of the period is transmitted
    unsigned kept_per_period;
    unsigned prefix[2][32];   // transmitted symbols of the period that precede (stage t, symbol w), in transmission order
    int input_type;
    unsigned long long n_out_syms, n_in_syms;
};

__device__ inline int32_t load_symbol(const void* in, int input_type, unsigned long long idx) {
    switch (input_type) {
        case 1: { uint32_t w = static_cast<const uint32_t*>(in)[idx >> 3]; int v = (int)((w >> (28 - 4 * (idx & 7))) & 0xF); return (v ^ 8) - 8; }
        case 2: { uint32_t w = static_cast<const uint32_t*>(in)[idx >> 2]; return (int8_t)(w >> (24 - 8 * (idx & 3))); }
        case 3: { uint32_t w = static_cast<const uint32_t*>(in)[idx >> 1]; return (int16_t)(w >> (16 - 16 * (idx & 1))); }
        default: return 0;
    }
}

__global__ void depuncture_kernel(PunctParams p, const void* __restrict__ in, uint32_t* __restrict__ out, unsigned long long n_words) {
    const unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_words) return;
    const int per_word = p.input_type == 1 ? 8 : p.input_type == 2 ? 4 : 2;   // FP32: one symbol pair per thread
    uint32_t word = 0;
    float f[2] = {0.f, 0.f};
    for (int j = 0; j < per_word; j++) {
        const unsigned long long s = w * per_word + j;          // output symbol index
        const unsigned long long stage = s >> 1;
        const unsigned which = (unsigned)(s & 1), t = (unsigned)(stage % p.period);
        bool have = s < p.n_out_syms && ((p.keep[which] >> t) & 1u);
        const unsigned long long src = (stage / p.period) * p.kept_per_period + p.prefix[which][t];
        have = have && src < p.n_in_syms;
        if (p.input_type == 4) f[j] = have ? static_cast<const float*>(in)[src] : 0.f;
        else {
            const int32_t v = have ? load_symbol(in, p.input_type, src) : 0;
            const int width = p.input_type == 1 ? 4 : p.input_type == 2 ? 8 : 16;
            word = (word << width) | ((uint32_t)v & ((1u << width) - 1u));
        }
    }
    if (p.input_type == 4) reinterpret_cast<float2*>(out)[w] = make_float2(f[0], f[1]);
    else out[w] = word;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

// Fill packed_d (vit_input_size(options, 2*n_bits) bytes, rounded up to whole 32-bit packs) with a synthetic
// received stream for n_bits message bits; bits_d (optional, n_bits bytes) receives the message bits for BER.
// amp: symbol amplitude in quantiser units (0 = default per input type); sigma: noise sd relative to amp.
int vit_synth_device_ex(int input_type, size_t n_bits, unsigned seed, int amp, double sigma, int zero, int source,
                        void* packed_d, void* bits_d, void* cuda_stream) {
    if (input_type < 0 || input_type > 4 || !packed_d || source < 0 || source > 1) return VIT_ERR_ARG;
    static const int def_amp[5] = {64, 3, 40, 9000, 48};
    if (amp <= 0) amp = def_amp[input_type];
    SynthParams p;
    p.n_bits = n_bits; p.seed = seed; p.input_type = input_type; p.zero = zero; p.source = source;
    if (source == 1 && upload_prbs_tables() != cudaSuccess) return VIT_ERR_CUDA;
    p.amp_q8 = (long long)amp << 8;
    p.sigma_q16 = (long long)(sigma * amp * 256.0 / 0.57735 + 0.5);
    const int bits_per_word = input_type == 0 ? 16 : input_type == 1 ? 4 : input_type == 2 ? 2 : 1;
    const unsigned long long n_words = (n_bits + bits_per_word - 1) / bits_per_word;
    if (n_words == 0) return VIT_OK;
    const unsigned threads = 256;
    const unsigned long long blocks = ((n_words + WPT - 1) / WPT + threads - 1) / threads;
    if (blocks > 0x7fffffffull) return VIT_ERR_ARG;
    synth_kernel<<<(unsigned)blocks, threads, 0, static_cast<cudaStream_t>(cuda_stream)>>>(
        p, static_cast<uint32_t*>(packed_d), static_cast<uint8_t*>(bits_d), n_words);
    return cudaGetLastError() == cudaSuccess ? VIT_OK : VIT_ERR_CUDA;
}

int vit_synth_device(int input_type, size_t n_bits, unsigned seed, int amp, double sigma, int zero,
                     void* packed_d, void* bits_d, void* cuda_stream) {
    return vit_synth_device_ex(input_type, n_bits, seed, amp, sigma, zero, 0, packed_d, bits_d, cuda_stream);
}

// Bit errors against the message bits of the synthetic source itself (regenerated on the fly: no bits buffer, which
// for multi-Gbit streams would be larger than the decoded output by 8x).  Synchronous.
int vit_count_errors_synth_device(int options, const void* out_d, size_t messageLen, unsigned seed, int source,
                                  unsigned long long* errors, void* cuda_stream) {
    if (!out_d || !errors || source < 0 || source > 1) return VIT_ERR_ARG;
    const int bpp = ((options >> 8) & 0xf) == 1 ? 16 : 32;
    SynthParams p{};
    p.seed = seed; p.source = source; p.n_bits = ~0ull;
    if (source == 1 && upload_prbs_tables() != cudaSuccess) return VIT_ERR_CUDA;
    unsigned long long* acc = nullptr;
    if (cudaMalloc(&acc, sizeof *acc) != cudaSuccess) return VIT_ERR_CUDA;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    cudaMemsetAsync(acc, 0, sizeof *acc, st);
    const unsigned long long n_packs = messageLen / bpp;
    if (n_packs) {
        const unsigned long long blocks = (n_packs + 255) / 256;
        if (blocks > 0x7fffffffull) { cudaFree(acc); return VIT_ERR_ARG; }
        if (bpp == 16) count_errors_synth_kernel<16><<<(unsigned)blocks, 256, 0, st>>>(p, out_d, n_packs, acc);
        else count_errors_synth_kernel<32><<<(unsigned)blocks, 256, 0, st>>>(p, out_d, n_packs, acc);
    }
    cudaError_t e = cudaMemcpyAsync(errors, acc, sizeof *acc, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(acc);
    return e == cudaSuccess ? VIT_OK : VIT_ERR_CUDA;
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

// Expand a punctured stream of soft symbols to the rate-1/2 stream the decoder takes.  keep0 / keep1: bit t set = the 0171 /
// 0133 symbol of stage t of the period is transmitted (stage order, 0171 before 0133 inside a stage); e.g. rate 3/4:
// period 3, keep0 = 0b101, keep1 = 0b011.  in_d: n_in_syms transmitted symbols packed like the decoder's input type (SOFT4,
// SOFT8, SOFT16 or FP32 -- hard decisions have no erasure value); out_d receives n_out_stages * 2 symbols in the same
// packing, zero where nothing was transmitted.
int vit_depuncture_device(int input_type, const void* in_d, size_t n_in_syms, unsigned period, unsigned keep0, unsigned keep1,
                          void* out_d, size_t n_out_stages, void* cuda_stream) {
    if (input_type < 1 || input_type > 4 || !in_d || !out_d || period < 1 || period > 32) return VIT_ERR_ARG;
    PunctParams p{};
    p.period = period; p.keep[0] = keep0; p.keep[1] = keep1; p.input_type = input_type;
    unsigned n = 0;
    for (unsigned t = 0; t < period; t++)
        for (unsigned w = 0; w < 2; w++) {
            p.prefix[w][t] = n;
            if ((p.keep[w] >> t) & 1u) n++;
        }
    if (n == 0) return VIT_ERR_ARG;
    p.kept_per_period = n;
    p.n_out_syms = 2ull * n_out_stages; p.n_in_syms = n_in_syms;
    const unsigned per_word = input_type == 1 ? 8 : input_type == 2 ? 4 : 2;
    const unsigned long long n_words = (p.n_out_syms + per_word - 1) / per_word;
    if (n_words == 0) return VIT_OK;
    const unsigned long long blocks = (n_words + 255) / 256;
    if (blocks > 0x7fffffffull) return VIT_ERR_ARG;
    depuncture_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(p, in_d, static_cast<uint32_t*>(out_d), n_words);
    return cudaGetLastError() == cudaSuccess ? VIT_OK : VIT_ERR_CUDA;
}

// plain device-memory helpers so that host-only callers (no CUDA headers) can use the device-resident entry points
int vit_dev_alloc(void** ptr, size_t bytes) { return cudaMalloc(ptr, bytes ? bytes : 1) == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
void vit_dev_free(void* ptr) { if (ptr) cudaFree(ptr); }
int vit_dev_copy_to_host(void* dst_h, const void* src_d, size_t bytes) { return cudaMemcpy(dst_h, src_d, bytes, cudaMemcpyDeviceToHost) == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
int vit_dev_copy_from_host(void* dst_d, const void* src_h, size_t bytes) { return cudaMemcpy(dst_d, src_h, bytes, cudaMemcpyHostToDevice) == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
int vit_dev_set(int device) { return cudaSetDevice(device) == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
int vit_dev_sync(void) { return cudaDeviceSynchronize() == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
int vit_dev_count(void) { int n = 0; return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0; }
// page-locked host memory for callers without CUDA headers: vit_run takes its time-sliced upload path from such buffers
int vit_host_alloc(void** ptr, size_t bytes) { return cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault) == cudaSuccess ? VIT_OK : VIT_ERR_CUDA; }
void vit_host_free(void* ptr) { if (ptr) cudaFreeHost(ptr); }

#pragma GCC visibility pop
}
