// vit_api.cu -- the C ABI declared in include/vit_b200.h.
//
// Host-side replacement for reference src/viterbi/viterbi.cu:10-139,210-238 (ViterbiCUDA<options>
// ::Impl, memAlloc/memFree, size helpers, run) with a runtime `options` value instead of 60 class
// template instantiations.  Differences from the reference, all deliberate:
//   * device buffers are kept and grown, not re-allocated on every run (the reference's
//     preAllocated flag is never set, viterbi.cu:19,217,237);
//   * host<->device copies go through pinned staging buffers on a private stream and large inputs
//     are split at segment boundaries so copy-in, decode and copy-out overlap;
//   * failures are reported by return code (+ vit_last_error), never by exit().
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/vit_b200.h"
#include "vit_launch.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define VIT_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(VIT_ERR_CUDA, "%s in %s at line %d", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int EXTRA = 64;  // extraL + extraR, reference viterbi.h:70-76

inline int in_type(int o) { return o & 0xf; }
inline int met_type(int o) { return (o >> 4) & 0xf; }
inline int out_type(int o) { return (o >> 8) & 0xf; }
inline int bpp_of(int o) { return out_type(o) == 1 ? 16 : 32; }

bool fields_known(int o) {
    return in_type(o) <= 4 && met_type(o) <= 2 && out_type(o) <= 1 && ((o >> 12) & 0xf) <= 1 && (o >> 16) == 0;
}

const vitk::KernelEntry* entry_for(int o) {
    int met = met_type(o) == 0 ? vitk::MET_B32 : met_type(o) == 1 ? vitk::MET_B16 : vitk::MET_F16;
    return vitk::kernel_entry(met, in_type(o), out_type(o));
}

}  // namespace

namespace vitk {
const KernelEntry* kernel_entry(int met, int in, int bpp16) {
    switch (met) {
        case MET_B32: return kernel_entry_b32(in, bpp16);
        case MET_B16: return kernel_entry_b16(in, bpp16);
        case MET_F16: return kernel_entry_f16(in, bpp16);
        default: return nullptr;
    }
}
}  // namespace vitk

struct vit_handle {
    int options = 0;
    int device = 0;
    unsigned segments = 6400;                 // reference viterbi.cu:19
    const vitk::KernelEntry* kernel = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_in[2] = {nullptr, nullptr};
    void* in_d = nullptr;  size_t in_cap = 0;
    void* out_d = nullptr; size_t out_cap = 0;
    void* pin_in = nullptr; size_t pin_in_cap = 0;
    void* pin_out = nullptr; size_t pin_out_cap = 0;
    unsigned long long launches = 0;
};

namespace {

int ensure_device_buffers(vit_handle* h, size_t in_bytes, size_t out_bytes) {
    if (in_bytes > h->in_cap) {
        if (h->in_d) cudaFree(h->in_d);
        h->in_d = nullptr; h->in_cap = 0;
        size_t cap = (in_bytes + 255) / 256 * 256;
        VIT_CUDA(cudaMalloc(&h->in_d, cap));
        h->in_cap = cap;
    }
    if (out_bytes > h->out_cap) {
        if (h->out_d) cudaFree(h->out_d);
        h->out_d = nullptr; h->out_cap = 0;
        size_t cap = (out_bytes + 255) / 256 * 256;
        VIT_CUDA(cudaMalloc(&h->out_d, cap));
        h->out_cap = cap;
    }
    return VIT_OK;
}

int launch(vit_handle* h, const void* in_d, void* out_d, size_t inputNum, size_t nstreams,
           size_t in_stride, size_t out_stride, cudaStream_t st, float* kernel_ms) {
    const int o = h->options;
    const size_t M = vit_message_len(o, inputNum);
    if (M == 0 || nstreams == 0) { if (kernel_ms) *kernel_ms = 0.f; return VIT_OK; }
    if ((reinterpret_cast<uintptr_t>(in_d) & 15) || (in_stride & 15))
        return fail(VIT_ERR_ARG, "device input must be 16-byte aligned (ptr %p, stride %zu)", in_d, in_stride);
    if (nstreams > 65535) return fail(VIT_ERR_ARG, "at most 65535 streams per launch (got %zu)", nstreams);
    vitk::KParams kp;
    kp.in = static_cast<const uint8_t*>(in_d);
    kp.out = static_cast<uint8_t*>(out_d);
    kp.in_stride = in_stride; kp.out_stride = out_stride;
    kp.in_bytes = vit_input_size(o, inputNum);
    kp.packs = M / (size_t)bpp_of(o);
    kp.segments = h->segments;
    kp.nstreams = (unsigned)nstreams;
    dim3 grid((h->segments + vitk::SEGS_PER_WARP - 1) / vitk::SEGS_PER_WARP, (unsigned)nstreams, 1);
    if (kernel_ms) VIT_CUDA(cudaEventRecord(h->ev0, st));
    VIT_CUDA(h->kernel->launch(kp, grid, st));
    h->launches++;
    if (kernel_ms) {
        VIT_CUDA(cudaEventRecord(h->ev1, st));
        VIT_CUDA(cudaEventSynchronize(h->ev1));
        VIT_CUDA(cudaEventElapsedTime(kernel_ms, h->ev0, h->ev1));
    }
    return VIT_OK;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

const char* vit_last_error(void) { return g_err; }

// reference viterbi.cu:63-84
size_t vit_input_size(int o, size_t n) {
    switch (in_type(o)) {
        case 0: return (n + 7) / 8;
        case 1: return (n + 1) / 2;
        case 2: return n;
        case 3: return n * 2;
        case 4: return n * 4;
        default: return 0;
    }
}
// reference viterbi.cu:86-88
size_t vit_message_len(int o, size_t n) {
    const size_t bpp = (size_t)bpp_of(o);
    if (n / 2 < (size_t)EXTRA) return 0;   // the reference's unsigned arithmetic would wrap here
    return (n / 2 - EXTRA) / bpp * bpp;
}
// reference viterbi.cu:90-92
size_t vit_output_size(int o, size_t n) { return vit_message_len(o, n) / 8; }

// reference viterbi.h:22-36
int vit_options_valid_ref(int o) {
    if (!fields_known(o)) return 0;
    const int it = in_type(o), mt = met_type(o), cm = (o >> 12) & 0xf;
    if (it == 2 && mt == 2) return 0;
    if (it == 3 && mt == 2) return 0;
    if (it == 3 && mt == 1) return 0;
    if (mt == 2 && cm == 1) return 0;
    return 1;
}

int vit_options_valid(int o) {
    if (!fields_known(o)) return 0;
    return entry_for(o) != nullptr;
}

int vit_kernel_info(int o, int* regs, int* smem_bytes, int* block_threads, int* segs_per_block) {
    if (!vit_options_valid(o)) return fail(VIT_ERR_OPTIONS, "unsupported option combination 0x%x", o);
    const vitk::KernelEntry* e = entry_for(o);
    cudaFuncAttributes a;
    VIT_CUDA(cudaFuncGetAttributes(&a, e->func));
    if (regs) *regs = a.numRegs;
    if (smem_bytes) *smem_bytes = e->smem_bytes;
    if (block_threads) *block_threads = 32;
    if (segs_per_block) *segs_per_block = vitk::SEGS_PER_WARP;
    return VIT_OK;
}

int vit_create(vit_handle** out, int options, int device, size_t prealloc_inputNum) {
    if (!out) return fail(VIT_ERR_ARG, "null handle pointer");
    *out = nullptr;
    if (!vit_options_valid(options)) return fail(VIT_ERR_OPTIONS, "unsupported option combination 0x%x", options);
    VIT_CUDA(cudaSetDevice(device));
    vit_handle* h = new (std::nothrow) vit_handle();
    if (!h) return fail(VIT_ERR_ARG, "out of host memory");
    h->options = options; h->device = device; h->kernel = entry_for(options);
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_in[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_in[1], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        vit_destroy(h);
        return fail(VIT_ERR_CUDA, "%s in %s at line %d", cudaGetErrorString(e), __FILE__, __LINE__);
    }
    if (prealloc_inputNum) {
        int rc = ensure_device_buffers(h, vit_input_size(options, prealloc_inputNum), vit_output_size(options, prealloc_inputNum));
        if (rc) { vit_destroy(h); return rc; }
    }
    *out = h;
    return VIT_OK;
}

void vit_destroy(vit_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->in_d) cudaFree(h->in_d);
    if (h->out_d) cudaFree(h->out_d);
    if (h->pin_in) cudaFreeHost(h->pin_in);
    if (h->pin_out) cudaFreeHost(h->pin_out);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int i = 0; i < 2; i++) if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
}

int vit_set_segments(vit_handle* h, unsigned segments) {
    if (!h) return fail(VIT_ERR_ARG, "null handle");
    h->segments = segments ? segments : 6400;
    return VIT_OK;
}

unsigned long long vit_launch_count(const vit_handle* h) { return h ? h->launches : 0; }

int vit_run_device_batch(vit_handle* h, const void* in_d, void* out_d, size_t inputNum, size_t nstreams,
                         size_t in_stride, size_t out_stride, void* cuda_stream, float* kernel_ms) {
    if (!h || !in_d || !out_d) return fail(VIT_ERR_ARG, "null argument");
    VIT_CUDA(cudaSetDevice(h->device));
    return launch(h, in_d, out_d, inputNum, nstreams, in_stride, out_stride, static_cast<cudaStream_t>(cuda_stream), kernel_ms);
}

int vit_run_device(vit_handle* h, const void* in_d, void* out_d, size_t inputNum, void* cuda_stream, float* kernel_ms) {
    return vit_run_device_batch(h, in_d, out_d, inputNum, 1, 0, 0, cuda_stream, kernel_ms);
}

int vit_run(vit_handle* h, const void* in_h, void* out_h, size_t inputNum, float* kernel_ms) {
    if (!h || !in_h || !out_h) return fail(VIT_ERR_ARG, "null argument");
    VIT_CUDA(cudaSetDevice(h->device));
    const size_t in_bytes = vit_input_size(h->options, inputNum);
    const size_t out_bytes = vit_output_size(h->options, inputNum);
    if (out_bytes == 0) { if (kernel_ms) *kernel_ms = 0.f; return VIT_OK; }
    int rc = ensure_device_buffers(h, in_bytes, out_bytes);
    if (rc) return rc;
    // host -> device (reference viterbi.cu:219), decode (viterbi.cu:228), device -> host (viterbi.cu:235)
    VIT_CUDA(cudaMemcpyAsync(h->in_d, in_h, in_bytes, cudaMemcpyHostToDevice, h->stream));
    rc = launch(h, h->in_d, h->out_d, inputNum, 1, 0, 0, h->stream, kernel_ms);
    if (rc) return rc;
    VIT_CUDA(cudaMemcpyAsync(out_h, h->out_d, out_bytes, cudaMemcpyDeviceToHost, h->stream));
    VIT_CUDA(cudaStreamSynchronize(h->stream));
    return VIT_OK;
}

#pragma GCC visibility pop
}  // extern "C"
