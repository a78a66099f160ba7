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
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "../../include/vit_b200.h"
#include "vit_code.h"
#include "vit_internal.h"
#include "vit_launch.h"
#include "vit_stage_pool.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
}  // namespace
namespace {
struct RemoteRange { const char* base; size_t bytes; int owner; };
std::mutex g_remote_mu;
std::vector<RemoteRange> g_remote;
}  // namespace
void vit_note_remote_range(const void* base, size_t bytes, int owner_device) {
    std::lock_guard<std::mutex> lk(g_remote_mu);
    for (const auto& r : g_remote)
        if (r.base == base) return;                      // several threads of one process register the same buffer
    g_remote.push_back({static_cast<const char*>(base), bytes, owner_device});
}
void vit_forget_remote_range(const void* base) {
    std::lock_guard<std::mutex> lk(g_remote_mu);
    for (size_t i = 0; i < g_remote.size(); i++)
        if (g_remote[i].base == base) { g_remote.erase(g_remote.begin() + (long)i); break; }
}
bool vit_in_remote_range(const void* p, int device) {
    std::lock_guard<std::mutex> lk(g_remote_mu);
    for (const auto& r : g_remote)
        if (static_cast<const char*>(p) >= r.base && static_cast<const char*>(p) < r.base + r.bytes) return r.owner != device;
    return false;
}
// shared with the other translation units of the library (vit_mg.cu); not exported
int vit_set_error(int code, const char* msg) {
    snprintf(g_err, sizeof g_err, "%s", msg);
    return code;
}
namespace {
#define VIT_CUDA(call)                                                                             \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(VIT_ERR_CUDA, "%s in %s at line %d", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int EXTRA = 64;  // extraL + extraR, reference viterbi.h:70-76

// The API calls run on the handle's device but leave the caller's current device as they found it (a multi-GPU
// process that holds one handle per device must not have its thread's device changed under it).
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    cudaError_t enter(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev == device) return cudaSuccess;
        switched = true;
        return cudaSetDevice(device);
    }
    ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
};
#define VIT_ON_DEVICE(h) DeviceGuard guard_; VIT_CUDA(guard_.enter((h)->device))

using vit_host::StagePool;
using vit_host::stage_copy;
using vit_host::stage_fence;

inline int in_type(int o) { return o & 0xf; }
inline int met_type(int o) { return (o >> 4) & 0xf; }
inline int out_type(int o) { return (o >> 8) & 0xf; }
inline int bpp_of(int o) { return out_type(o) == 1 ? 16 : 32; }

inline int comp_mode(int o) { return (o >> 12) & 0xf; }

// CompMode: 0 REG, 1 DPX (the reference's flag; selects the same core as REG there and here), 2 DPX_TIES (extension: the
// tie rule the reference's DPX code would have if it were instantiated -- differs from REG for the int32 core only)
bool fields_known(int o) {
    return in_type(o) <= 4 && met_type(o) <= 2 && out_type(o) <= 1 && comp_mode(o) <= 2 && (o >> 16) == 0;
}

const vitk::KernelEntry* entry_for(int o) {
    int met = met_type(o) == 0 ? (comp_mode(o) == 2 ? vitk::MET_B32D : vitk::MET_B32) : met_type(o) == 1 ? vitk::MET_B16 : vitk::MET_F16;
    if (comp_mode(o) == 2 && met_type(o) == 2) return nullptr;      // the reference has no half2 DPX code (viterbi.h:30-31)
    return vitk::kernel_entry(met, in_type(o), out_type(o));
}

}  // namespace

namespace vitk {
const KernelEntry* kernel_entry(int met, int in, int bpp16) {
    switch (met) {
        case MET_B32: return kernel_entry_b32(in, bpp16);
        case MET_B32D: return kernel_entry_b32d(in, bpp16);
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
    static constexpr int MAX_CHUNKS = 8;
    cudaStream_t stream = nullptr;        // decode kernels
    cudaStream_t chunk_stream[MAX_CHUNKS] = {};   // one per pipeline chunk so that chunk kernels co-reside
    cudaStream_t copy_stream = nullptr;   // host -> device
    cudaStream_t out_stream = nullptr;    // device -> host
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_k0[MAX_CHUNKS] = {}, ev_k1[MAX_CHUNKS] = {};
    void* in_d = nullptr;  size_t in_cap = 0;
    void* out_d = nullptr; size_t out_cap = 0;
    void* pin_in = nullptr; size_t pin_in_cap = 0;
    void* pin_out = nullptr; size_t pin_out_cap = 0;
    unsigned long long launches = 0;
    // upload gates of the time-sliced host-buffer path (vit_run)
    unsigned* gate_d = nullptr;          // [8] slice-arrived words + [1] error word
    unsigned* epoch_h = nullptr;         // pinned: [0] the value the copy stream writes into a gate, [16] error word
    unsigned* gate_err_d = nullptr;      // device address of epoch_h[16] (mapped): set by a warp that gave up waiting
    unsigned epoch = 0;
    cudaEvent_t ev_kdone = nullptr;
    // The time-sliced upload needs copies to make progress while a kernel that waits for them is running.  That is not
    // the case when a tool injected into the CUDA driver serialises kernels and copies (ncu, nsys, compute-sanitizer),
    // when every stream shares one hardware queue (CUDA_DEVICE_MAX_CONNECTIONS=1) or when launches block the host
    // (CUDA_LAUNCH_BLOCKING=1): detected up front, vit_run then takes the segment-range chunk pipeline.
    bool gates_disabled = env_forbids_gates();
    // chunked decode of an endless stream (vit_stream_push): the symbols of the previous window that could not be
    // decoded yet (its last 64..64+bitsPerPack-1 stages) wait here and are prepended to the next chunk
    void* carry_d = nullptr; size_t carry_cap = 0; size_t carry_syms = 0;
    unsigned long long stream_bits = 0;   // bits emitted since vit_stream_reset
    unsigned last_stage_out = 0;                                     // KParams::stage_out of the last launch
    bool force_stage_out = getenv("VIT_STAGE_OUT") != nullptr;       // measurement hook: staged output stores for local buffers too
    int upload_mode = VIT_UPLOAD_AUTO;    // vit_set_upload_mode
    int geometry = env_geometry();        // vit_set_geometry: lane geometry of gate-free launches
    int last_geometry = VIT_GEOMETRY_L8;
    static int env_geometry() {
        const char* e = getenv("VIT_GEOMETRY");
        return e && (strcmp(e, "l1") == 0 || strcmp(e, "L1") == 0 || strcmp(e, "1") == 0) ? VIT_GEOMETRY_L1 : VIT_GEOMETRY_L8;
    }
    unsigned long long gate_timeout_ns = 2000000000ull;
    StagePool* pool = nullptr;            // created on the first vit_run with pageable buffers
    static bool env_forbids_gates() {
        auto is = [](const char* name, const char* v) { const char* e = getenv(name); return e && strcmp(e, v) == 0; };
        return getenv("CUDA_INJECTION64_PATH") != nullptr || getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr ||
               is("CUDA_DEVICE_MAX_CONNECTIONS", "1") || is("CUDA_LAUNCH_BLOCKING", "1");
    }
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

// decode segments [seg_first, seg_limit) of every stream; timing (optional) between caller-supplied events
struct GatePlan { unsigned n = 0; unsigned super[8] = {}; };

int launch_range(vit_handle* h, const void* in_d, void* out_d, size_t inputNum, size_t nstreams,
                 size_t in_stride, size_t out_stride, cudaStream_t st, unsigned seg_first, unsigned seg_limit,
                 cudaEvent_t e0, cudaEvent_t e1, const GatePlan* gp = nullptr) {
    const int o = h->options;
    vitk::KParams kp;
    kp.in = static_cast<const uint8_t*>(in_d);
    kp.out = static_cast<uint8_t*>(out_d);
    kp.in_stride = in_stride; kp.out_stride = out_stride;
    kp.in_bytes = vit_input_size(o, inputNum);
    kp.packs = vit_message_len(o, inputNum) / (size_t)bpp_of(o);
    kp.segments = h->segments;
    kp.seg_first = seg_first; kp.seg_limit = seg_limit;
    kp.nstreams = (unsigned)nstreams;
    kp.one = 1u;
    kp.gate = h->gate_d; kp.gate_err = h->gate_err_d; kp.gate_epoch = h->epoch; kp.gate_n = 0;
    kp.gate_timeout_ns = h->gate_timeout_ns;
    // output in another GPU's memory (VIT_GATHER_DIRECT): the kernel stages 8 slides per store.  Mappings handed out by
    // vit_comm_shared_alloc are registered (an IPC mapping looks like local memory to cudaPointerGetAttributes); any other
    // peer pointer is recognised by its attributes.
    kp.stage_out = 0;
    if (out_d != h->out_d) {
        bool remote = vit_in_remote_range(out_d, h->device);
        if (!remote) {
            cudaPointerAttributes pa;
            remote = cudaPointerGetAttributes(&pa, out_d) == cudaSuccess && pa.type == cudaMemoryTypeDevice && pa.device != h->device;
            cudaGetLastError();
        }
        kp.stage_out = (remote || h->force_stage_out) ? 1u : 0u;
    }
    h->last_stage_out = kp.stage_out;
    for (int i = 0; i < 8; i++) kp.gate_super[i] = 0;
    if (gp && h->gate_d) {
        kp.gate_n = gp->n;
        for (unsigned i = 0; i < gp->n; i++) kp.gate_super[i] = gp->super[i];
    }
    // the experimental one-lane-per-segment geometry: only where the launch needs neither upload gates nor staged stores
    const bool l1 = h->geometry == VIT_GEOMETRY_L1 && h->kernel->launch_l1 && kp.gate_n == 0 && kp.stage_out == 0;
    h->last_geometry = l1 ? VIT_GEOMETRY_L1 : VIT_GEOMETRY_L8;
    if (e0) VIT_CUDA(cudaEventRecord(e0, st));
    VIT_CUDA(l1 ? h->kernel->launch_l1(kp, st) : h->kernel->launch(kp, st));
    h->launches++;
    if (e1) VIT_CUDA(cudaEventRecord(e1, st));
    return VIT_OK;
}

int launch(vit_handle* h, const void* in_d, void* out_d, size_t inputNum, size_t nstreams,
           size_t in_stride, size_t out_stride, cudaStream_t st, float* kernel_ms) {
    const int o = h->options;
    const size_t M = vit_message_len(o, inputNum);
    if (M == 0 || nstreams == 0) { if (kernel_ms) *kernel_ms = 0.f; return VIT_OK; }
    if ((reinterpret_cast<uintptr_t>(in_d) & 15) || (in_stride & 15))
        return fail(VIT_ERR_ARG, "device input must be 16-byte aligned (ptr %p, stride %zu)", in_d, in_stride);
    const size_t pack_align = (size_t)bpp_of(o) / 8;
    if ((reinterpret_cast<uintptr_t>(out_d) & (pack_align - 1)) || (out_stride & (pack_align - 1)))
        return fail(VIT_ERR_ARG, "device output must be aligned to its %zu-byte packs (ptr %p, stride %zu)", pack_align, out_d, out_stride);
    if (nstreams > 65535) return fail(VIT_ERR_ARG, "at most 65535 streams per launch (got %zu)", nstreams);
    int rc = launch_range(h, in_d, out_d, inputNum, nstreams, in_stride, out_stride, st, 0, h->segments,
                          kernel_ms ? h->ev0 : nullptr, kernel_ms ? h->ev1 : nullptr);
    if (rc) return rc;
    if (kernel_ms) {
        VIT_CUDA(cudaEventSynchronize(h->ev1));
        VIT_CUDA(cudaEventElapsedTime(kernel_ms, h->ev0, h->ev1));
    }
    return VIT_OK;
}

}  // namespace

namespace {

// geometry of one host-buffer decode (reference viterbi.cu:156-165 for the segment partition)
struct HostRun {
    const char* in_h; char* out_h;
    size_t inputNum, in_bytes, out_bytes;
    size_t bpp, W;            // bits per decoded pack, stream segments
    size_t P, q, r;           // decoded packs; packs per segment; segments that are one pack longer
    size_t b96;               // channel bytes per 96 trellis stages
    size_t pack_bytes;        // channel bytes per decoded pack
    size_t nsuper;            // 96-stage super-steps of the longest segment (64-stage tail included)
    int nch;                  // chunks of the segment-range pipeline
    size_t start_pack(size_t w) const { return q * w + std::min(w, r); }
};

HostRun host_run_geometry(const vit_handle* h, const void* in_h, void* out_h, size_t inputNum) {
    const int o = h->options;
    HostRun g;
    g.in_h = static_cast<const char*>(in_h); g.out_h = static_cast<char*>(out_h);
    g.inputNum = inputNum; g.in_bytes = vit_input_size(o, inputNum); g.out_bytes = vit_output_size(o, inputNum);
    g.bpp = (size_t)bpp_of(o); g.W = h->segments;
    g.P = vit_message_len(o, inputNum) / g.bpp; g.q = g.P / g.W; g.r = g.P % g.W;
    g.b96 = in_type(o) == 0 ? 24 : in_type(o) == 1 ? 96 : in_type(o) == 2 ? 192 : in_type(o) == 3 ? 384 : 768;
    g.pack_bytes = g.bpp * g.b96 / 96;
    const size_t Lmax = (g.q + (g.r ? 1 : 0)) * g.bpp, Tmax = 64 + 32 * ((Lmax + 31) / 32);
    g.nsuper = (Tmax + 95) / 96;
    g.nch = (int)std::min<size_t>(vit_handle::MAX_CHUNKS, g.in_bytes / (4u << 20));
    return g;
}

// host -> device (reference viterbi.cu:219), decode (viterbi.cu:228), device -> host (viterbi.cu:235);
// kernel_ms = device time of the single decode launch, exactly the reference's measurement (viterbi.cu:224-232)
int run_sequential(vit_handle* h, const HostRun& g, float* kernel_ms) {
    VIT_CUDA(cudaMemcpyAsync(h->in_d, g.in_h, g.in_bytes, cudaMemcpyHostToDevice, h->stream));
    int rc = launch(h, h->in_d, h->out_d, g.inputNum, 1, 0, 0, h->stream, kernel_ms);
    if (rc) return rc;
    VIT_CUDA(cudaMemcpyAsync(g.out_h, h->out_d, g.out_bytes, cudaMemcpyDeviceToHost, h->stream));
    VIT_CUDA(cudaStreamSynchronize(h->stream));
    return VIT_OK;
}

bool is_pinned(const void* p) {
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, p) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return pinned;
}

bool gated_upload_applies(const vit_handle* h, const HostRun& g) {
    if (h->upload_mode == VIT_UPLOAD_GATED) return h->gate_d && h->gate_err_d && g.q >= 1 && g.nsuper >= 8;
    if (h->gates_disabled || !h->gate_d || !h->gate_err_d) return false;
    if (g.q < 1 || g.nsuper < 16 || g.in_bytes < (2u << 20)) return false;      // too short to be worth slicing
    return true;
}

int ensure_staging(vit_handle* h, size_t in_bytes, size_t out_bytes) {
    if (in_bytes > h->pin_in_cap) {
        if (h->pin_in) cudaFreeHost(h->pin_in);
        h->pin_in = nullptr; h->pin_in_cap = 0;
        VIT_CUDA(cudaHostAlloc(&h->pin_in, in_bytes + 256, cudaHostAllocDefault));
        h->pin_in_cap = in_bytes;
    }
    if (out_bytes > h->pin_out_cap) {
        if (h->pin_out) cudaFreeHost(h->pin_out);
        h->pin_out = nullptr; h->pin_out_cap = 0;
        VIT_CUDA(cudaHostAlloc(&h->pin_out, out_bytes + 256, cudaHostAllocDefault));
        h->pin_out_cap = out_bytes;
    }
    if (!h->pool) {
        const char* e = getenv("VIT_STAGE_THREADS");
        // default: up to 16 threads, and this process's share of the cores when several ranks share the host (torchrun and
        // MPI launchers export the number of local ranks): the workers spin between blocks, oversubscription is ruinous
        unsigned share = 1;
        for (const char* name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE"})
            if (const char* v = getenv(name)) { share = (unsigned)std::max(1, atoi(v)); break; }
        int n = e ? atoi(e) : (int)std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency() / share));
        n = std::max(1, std::min(n, 64));
        h->pool = new (std::nothrow) StagePool(n - 1);
        if (!h->pool) return fail(VIT_ERR_ARG, "out of host memory");
    }
    return VIT_OK;
}

// Time-sliced upload: ONE decode launch starts at once and every warp waits at "upload gates" for the next column block
// of its segments; the copy stream uploads block b of EVERY segment (two strided copies: the first P%W segments are one
// pack longer) and then opens gate b.  The decode therefore finishes one short block after the last byte has landed,
// instead of a whole per-segment chain (0.25-0.4 ms) after it as with segment-range chunks.
//   pinned caller buffers: the copies read the caller's input in place;
//   pageable (stage_in / stage_out): the worker threads first copy block b of every segment into the pinned staging buffer (same
//                          layout), which overlaps with the upload of block b-1; the decoded packs come back through
//                          the pinned output staging buffer.
// *gave_up is set when a warp stopped waiting for a gate (the output is then incomplete).
int run_gated(vit_handle* h, const HostRun& g, bool stage_in, bool stage_out, bool* gave_up) {
    const bool staged = stage_in || stage_out;
    GatePlan gp;
    if (stage_in) {
        // equal column blocks: the staging copy of block b+1 hides behind the upload of block b
        static const int env_blocks = [] { const char* e = getenv("VIT_STAGE_BLOCKS"); return e ? atoi(e) : 0; }();   // measurement hook
        gp.n = (unsigned)std::min<size_t>(env_blocks > 0 ? (size_t)std::min(env_blocks, 8) : 4, std::max<size_t>(2, g.nsuper / 4));
        for (unsigned b = 0; b < gp.n; b++) gp.super[b] = (unsigned)(g.nsuper * b / gp.n);
    } else {
        // column blocks of 1/2, 3/8 and 1/8 of a segment (profiles/r1_upload_pattern_probe.txt): fewer, wider strided
        // copies upload faster (0.66 ms for 32 MB against 0.69 with four blocks), a short last block keeps the tail
        // short.  That is for inputs whose upload takes longer than their decode (>= 1 byte per decoded bit at ~50 GB/s
        // against 12-16 ps per bit).  Hard-decision input (0.25 byte per bit) is decode bound: a small first block lets
        // the kernel start early and the rest arrives long before it is needed.
        gp.n = 3;
        gp.super[0] = 0;
        if (in_type(h->options) == 0) { gp.super[1] = (unsigned)(g.nsuper / 8); gp.super[2] = (unsigned)(g.nsuper / 2); }
        else { gp.super[1] = (unsigned)(g.nsuper / 2); gp.super[2] = (unsigned)(g.nsuper * 7 / 8); }
    }
    VIT_CUDA(cudaStreamSynchronize(h->stream));           // a kernel abandoned by a failed earlier call may still own the error word
    h->epoch_h[16] = 0;
    h->epoch++;
    if (h->epoch == 0) h->epoch = 1;
    *h->epoch_h = h->epoch;
    static const bool dbg = getenv("VIT_RUN_DEBUG") != nullptr;
    auto now = [] { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e3 + t.tv_nsec * 1e-6; };
    const double t_begin = dbg ? now() : 0.0;
    // a warp gives up on a gate after the time the whole upload would take at 1 GB/s plus 0.2 s (a healthy upload runs
    // at ~50 GB/s): long enough for any working configuration, short enough not to look like a hang
    h->gate_timeout_ns = 200000000ull + (unsigned long long)g.in_bytes;
    int rc = launch_range(h, h->in_d, h->out_d, g.inputNum, 1, 0, 0, h->stream, 0, (unsigned)g.W, nullptr, h->ev_kdone, &gp);
    if (rc) return rc;
    const char* user = g.in_h;
    char* stage = static_cast<char*>(h->pin_in);
    const char* src = stage_in ? stage : user;
    char* dst = static_cast<char*>(h->in_d);
    const size_t pitch1 = (g.q + 1) * g.pack_bytes, pitch2 = g.q * g.pack_bytes;   // segment strides in bytes
    const size_t grp2 = g.r * pitch1;                                                // first byte of segment r
    const size_t body_end = g.P * g.pack_bytes;                                      // first byte after the last segment's own packs
    if (staged) h->pool->begin_call();
    // the last 64 stages of the stream (warm-up tail of the last segment) go with the first block
    if (g.in_bytes > body_end) {
        if (stage_in) memcpy(stage + body_end, user + body_end, g.in_bytes - body_end);
        VIT_CUDA(cudaMemcpyAsync(dst + body_end, src + body_end, g.in_bytes - body_end, cudaMemcpyHostToDevice, h->copy_stream));
    }
    static const bool lose_gate = getenv("VIT_TEST_LOSE_GATE") != nullptr;           // test hook: never open the last gate
    double t_stage = 0.0;
    for (unsigned b = 0; b < gp.n; b++) {
        // columns [lo, hi) of every segment row: super-steps [super[b], super[b+1]) plus the read-ahead the kernel's
        // 16-byte staging pieces take (<= 15 bytes before, <= 27 bytes after)
        const size_t lo = b == 0 ? 0 : (size_t)gp.super[b] * g.b96 - 16;
        const size_t hi = b + 1 == gp.n ? (size_t)-1 : (size_t)gp.super[b + 1] * g.b96 + 48;
        const size_t w1 = std::min(hi, pitch1), w2 = std::min(hi, pitch2);
        if (stage_in) {
            // rows are dealt to the workers in groups of 32 (a group of short rows is still tens of kilobytes)
            const size_t groups = (g.W + 31) / 32;
            const std::function<void(size_t)> job = [&](size_t gi) {
                for (size_t w = gi * 32; w < std::min(g.W, gi * 32 + 32); w++) {
                    const size_t row = w < g.r ? w * pitch1 : grp2 + (w - g.r) * pitch2;
                    const size_t wd = w < g.r ? w1 : w2;
                    if (wd > lo) stage_copy(stage + row + lo, user + row + lo, wd - lo);
                }
                stage_fence();
            };
            const double ts = dbg ? now() : 0.0;
            h->pool->parallel_for(groups, job);
            if (dbg) t_stage += now() - ts;
        }
        if (g.r && w1 > lo)
            VIT_CUDA(cudaMemcpy2DAsync(dst + lo, pitch1, src + lo, pitch1, w1 - lo, g.r, cudaMemcpyHostToDevice, h->copy_stream));
        if (w2 > lo)
            VIT_CUDA(cudaMemcpy2DAsync(dst + grp2 + lo, pitch2, src + grp2 + lo, pitch2, w2 - lo, g.W - g.r, cudaMemcpyHostToDevice, h->copy_stream));
        if (lose_gate && b + 1 == gp.n) continue;
        VIT_CUDA(cudaMemcpyAsync(h->gate_d + b, h->epoch_h, sizeof(unsigned), cudaMemcpyHostToDevice, h->copy_stream));
    }
    // The download is queued only now, after every upload: a device-to-host copy into PAGEABLE memory blocks the host
    // until it has run, i.e. until the kernel is done -- queued earlier it would keep the uploads the kernel waits for
    // from ever being issued.  (Storing the decoded packs straight into pinned host memory from the kernel was measured:
    // 4-byte stores over PCIe double the time of the PCIe-bound s4 case.)
    char* out_target = stage_out ? static_cast<char*>(h->pin_out) : g.out_h;
    VIT_CUDA(cudaMemcpyAsync(out_target, h->out_d, g.out_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (dbg) {
        const double t_issued = now();
        cudaStreamSynchronize(h->copy_stream);
        const double t_copy = now();
        cudaEventSynchronize(h->ev_kdone);
        const double t_kernel = now();
        cudaStreamSynchronize(h->stream);
        fprintf(stderr, "[vit_run gated%s] %u blocks, issue %.3f ms (staging copies %.3f), upload done +%.3f, kernel done +%.3f, download done +%.3f\n",
                staged ? " staged" : "", gp.n, t_issued - t_begin, t_stage, t_copy - t_begin, t_kernel - t_begin, now() - t_begin);
    }
    cudaError_t e = cudaStreamSynchronize(h->copy_stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        if (staged) h->pool->end_call();
        return fail(VIT_ERR_CUDA, "%s in %s at line %d", cudaGetErrorString(e), __FILE__, __LINE__);
    }
    *gave_up = *static_cast<volatile unsigned*>(h->epoch_h + 16) != 0;
    h->epoch_h[16] = 0;
    if (staged) {
        if (stage_out && !*gave_up) {
            const size_t piece = 256u << 10, n = (g.out_bytes + piece - 1) / piece;
            const std::function<void(size_t)> job = [&](size_t i) {
                memcpy(g.out_h + i * piece, out_target + i * piece, std::min(piece, g.out_bytes - i * piece));
            };
            h->pool->parallel_for(n, job);
        }
        h->pool->end_call();
    }
    return VIT_OK;
}

// Segment-range chunk pipeline (pinned host input when gates are unavailable, e.g. under a profiler): the stream is cut at segment boundaries
// into nch chunks (segments are independent: chunk i needs the input bytes up to the end of its last segment's tail
// and produces a contiguous range of output packs).  Copies queue in order on one stream; each chunk's kernel runs on
// its own stream as soon as its bytes have landed, so the kernels co-reside (all 1600 warps fit on the device at once)
// instead of serialising.
int run_chunked(vit_handle* h, const HostRun& g) {
    size_t in_lo = 0;
    for (int i = 0; i < g.nch; i++) {
        const unsigned a = (unsigned)(g.W * i / g.nch / 8 * 8), b = (i + 1 == g.nch) ? (unsigned)g.W : (unsigned)(g.W * (i + 1) / g.nch / 8 * 8);
        size_t in_hi = g.in_bytes;
        if (i + 1 < g.nch) {
            const size_t last_bits = (g.q + ((size_t)(b - 1) < g.r ? 1 : 0)) * g.bpp;
            const size_t end_stage = g.start_pack(b - 1) * g.bpp + 64 + 32 * ((last_bits + 31) / 32) + 32;
            in_hi = std::min(g.in_bytes, (end_stage * g.b96 + 95) / 96 + 64);
            in_hi = std::max(in_hi, in_lo);
        }
        if (in_hi > in_lo)
            VIT_CUDA(cudaMemcpyAsync(static_cast<char*>(h->in_d) + in_lo, g.in_h + in_lo, in_hi - in_lo, cudaMemcpyHostToDevice, h->copy_stream));
        in_lo = in_hi;
        VIT_CUDA(cudaEventRecord(h->ev_in[i], h->copy_stream));
        VIT_CUDA(cudaStreamWaitEvent(h->chunk_stream[i], h->ev_in[i], 0));
        int rc = launch_range(h, h->in_d, h->out_d, g.inputNum, 1, 0, 0, h->chunk_stream[i], a, b, nullptr, h->ev_k1[i]);
        if (rc) return rc;
        const size_t out_lo = g.start_pack(a) * g.bpp / 8, out_hi = (b >= g.W) ? g.out_bytes : g.start_pack(b) * g.bpp / 8;
        VIT_CUDA(cudaStreamWaitEvent(h->out_stream, h->ev_k1[i], 0));
        if (out_hi > out_lo)
            VIT_CUDA(cudaMemcpyAsync(g.out_h + out_lo, static_cast<const char*>(h->out_d) + out_lo, out_hi - out_lo, cudaMemcpyDeviceToHost, h->out_stream));
    }
    VIT_CUDA(cudaStreamSynchronize(h->out_stream));
    for (int i = 0; i < g.nch; i++) VIT_CUDA(cudaStreamSynchronize(h->chunk_stream[i]));
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

// reference viterbi.h:61-63 (constLen, polyn1, polyn2): the values this build was compiled for (vit_code.h)
void vit_code_parameters(int* constLen, int* polyn1, int* polyn2) {
    if (constLen) *constLen = VIT_CONST_LEN;
    if (polyn1) *polyn1 = VIT_POLY1;
    if (polyn2) *polyn2 = VIT_POLY2;
}

// reference viterbi.h:22-36
int vit_options_valid_ref(int o) {
    if (!fields_known(o)) return 0;
    const int it = in_type(o), mt = met_type(o), cm = (o >> 12) & 0xf;
    if (cm > 1) return 0;
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
    if (segs_per_block) *segs_per_block = vitk::l8::SEGS_PER_WARP;     // single-stream build; multi-stream launches: l4, 8
    return VIT_OK;
}

int vit_create(vit_handle** out, int options, int device, size_t prealloc_inputNum) {
    if (!out) return fail(VIT_ERR_ARG, "null handle pointer");
    *out = nullptr;
    if (!vit_options_valid(options)) return fail(VIT_ERR_OPTIONS, "unsupported option combination 0x%x", options);
    DeviceGuard guard_;
    VIT_CUDA(guard_.enter(device));
    vit_handle* h = new (std::nothrow) vit_handle();
    if (!h) return fail(VIT_ERR_ARG, "out of host memory");
    h->options = options; h->device = device; h->kernel = entry_for(options);
    if (const char* e = getenv("VIT_RUN_MODE")) {            // same values as vit_set_upload_mode
        const int m = atoi(e);
        if (m >= VIT_UPLOAD_AUTO && m <= VIT_UPLOAD_GATED) h->upload_mode = m;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_kdone, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&h->gate_d), 16 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemset(h->gate_d, 0, 16 * sizeof(unsigned));
    if (e == cudaSuccess) e = cudaHostAlloc(reinterpret_cast<void**>(&h->epoch_h), 128, cudaHostAllocMapped);
    if (e == cudaSuccess) { memset(h->epoch_h, 0, 128); e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->gate_err_d), h->epoch_h + 16, 0); }
    for (int i = 0; i < vit_handle::MAX_CHUNKS && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&h->chunk_stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreate(&h->ev_k0[i]);
        if (e == cudaSuccess) e = cudaEventCreate(&h->ev_k1[i]);
    }
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
    DeviceGuard guard_;
    guard_.enter(h->device);
    if (h->in_d) cudaFree(h->in_d);
    if (h->out_d) cudaFree(h->out_d);
    delete h->pool;
    if (h->carry_d) cudaFree(h->carry_d);
    if (h->pin_in) cudaFreeHost(h->pin_in);
    if (h->pin_out) cudaFreeHost(h->pin_out);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_kdone) cudaEventDestroy(h->ev_kdone);
    if (h->gate_d) cudaFree(h->gate_d);
    if (h->epoch_h) cudaFreeHost(h->epoch_h);
    for (int i = 0; i < vit_handle::MAX_CHUNKS; i++) {
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_k0[i]) cudaEventDestroy(h->ev_k0[i]);
        if (h->ev_k1[i]) cudaEventDestroy(h->ev_k1[i]);
        if (h->chunk_stream[i]) cudaStreamDestroy(h->chunk_stream[i]);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    delete h;
}

int vit_set_segments(vit_handle* h, unsigned segments) {
    if (!h) return fail(VIT_ERR_ARG, "null handle");
    h->segments = segments ? segments : 6400;
    return VIT_OK;
}

unsigned long long vit_launch_count(const vit_handle* h) { return h ? h->launches : 0; }
int vit_last_launch_staged_output(const vit_handle* h) { return h ? (int)h->last_stage_out : 0; }

int vit_run_device_batch(vit_handle* h, const void* in_d, void* out_d, size_t inputNum, size_t nstreams,
                         size_t in_stride, size_t out_stride, void* cuda_stream, float* kernel_ms) {
    if (!h || !in_d || !out_d) return fail(VIT_ERR_ARG, "null argument");
    VIT_ON_DEVICE(h);
    return launch(h, in_d, out_d, inputNum, nstreams, in_stride, out_stride, static_cast<cudaStream_t>(cuda_stream), kernel_ms);
}

int vit_run_device(vit_handle* h, const void* in_d, void* out_d, size_t inputNum, void* cuda_stream, float* kernel_ms) {
    return vit_run_device_batch(h, in_d, out_d, inputNum, 1, 0, 0, cuda_stream, kernel_ms);
}

int vit_run(vit_handle* h, const void* in_h, void* out_h, size_t inputNum, float* kernel_ms) {
    if (!h || !in_h || !out_h) return fail(VIT_ERR_ARG, "null argument");
    VIT_ON_DEVICE(h);
    const size_t in_bytes = vit_input_size(h->options, inputNum);
    const size_t out_bytes = vit_output_size(h->options, inputNum);
    if (out_bytes == 0) { if (kernel_ms) *kernel_ms = 0.f; return VIT_OK; }
    int rc = ensure_device_buffers(h, in_bytes, out_bytes);
    if (rc) return rc;
    const HostRun hr = host_run_geometry(h, in_h, out_h, inputNum);
    // A kernel-time request keeps the reference's exact copy -> launch -> copy sequence (viterbi.cu:219-235): the time
    // between the events is then the decode alone, as the reference reports it.
    if (kernel_ms || hr.W < 64 || h->upload_mode == VIT_UPLOAD_SEQUENTIAL) return run_sequential(h, hr, kernel_ms);
    const bool pin_in = is_pinned(in_h), pin_out = is_pinned(out_h);
    if (h->upload_mode != VIT_UPLOAD_CHUNKED && gated_upload_applies(h, hr)) {
        // pageable buffers (the reference's calling convention: std::vector storage, viterbiDF.h:188-193) go through the
        // pinned staging buffers, copied by the worker threads block by block
        if (!pin_in || !pin_out) {
            rc = ensure_staging(h, pin_in ? 0 : in_bytes, pin_out ? 0 : out_bytes);
            if (rc) return rc;
        }
        bool gave_up = false;
        rc = run_gated(h, hr, !pin_in, !pin_out, &gave_up);
        if (rc || !gave_up) return rc;
        // A warp gave up waiting for its upload gate: something kept the copies from running beside the kernel.
        // Nothing was lost: decode again below and stop using gates on this handle.
        h->gates_disabled = true;
        if (h->upload_mode == VIT_UPLOAD_GATED) return fail(VIT_ERR_CUDA, "upload gate timed out (copies do not overlap kernels in this environment)");
    }
    // Without gates: pinned buffers take the segment-range chunk pipeline; pageable memory is staged by the driver,
    // synchronously and piece by piece, so one large copy is the fastest way through.
    if (!pin_in || !pin_out || hr.nch < 2) return run_sequential(h, hr, nullptr);
    return run_chunked(h, hr);
}

// ---- chunked decode of an endless stream ---------------------------------------------------------
namespace {
size_t symbols_per_word(int o) { return in_type(o) == 0 ? 32 : in_type(o) == 1 ? 8 : in_type(o) == 2 ? 4 : in_type(o) == 3 ? 2 : 1; }

// One window = carried symbols ++ this chunk.  `chunk` is a host pointer (chunk_on_device == false) or a device pointer.
int stream_push(vit_handle* h, const void* chunk, bool chunk_on_device, size_t inputNum, void* out, bool out_on_device,
                size_t out_cap, size_t* out_bytes, cudaStream_t st) {
    const int o = h->options;
    if (inputNum % symbols_per_word(o))
        return fail(VIT_ERR_ARG, "a stream chunk must be whole 32-bit channel packs (%zu symbols per pack, got %zu symbols)", symbols_per_word(o), inputNum);
    const size_t n = h->carry_syms + inputNum;
    const size_t carry_bytes = vit_input_size(o, h->carry_syms), chunk_bytes = vit_input_size(o, inputNum), win_bytes = carry_bytes + chunk_bytes;
    const size_t M = vit_message_len(o, n), need_out = M / 8;
    if (out_bytes) *out_bytes = 0;
    if (need_out > out_cap) return fail(VIT_ERR_ARG, "this chunk completes %zu bytes of decoded packs, the output buffer holds %zu", need_out, out_cap);
    int rc = ensure_device_buffers(h, win_bytes + 16, need_out);
    if (rc) return rc;
    char* win = static_cast<char*>(h->in_d);
    if (carry_bytes) VIT_CUDA(cudaMemcpyAsync(win, h->carry_d, carry_bytes, cudaMemcpyDeviceToDevice, st));
    if (chunk_bytes) VIT_CUDA(cudaMemcpyAsync(win + carry_bytes, chunk, chunk_bytes, chunk_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (M) {
        void* out_d = out_on_device ? out : h->out_d;
        rc = launch(h, win, out_d, n, 1, 0, 0, st, nullptr);
        if (rc) return rc;
        if (!out_on_device) VIT_CUDA(cudaMemcpyAsync(out, h->out_d, need_out, cudaMemcpyDeviceToHost, st));
    }
    // the next window starts at stage M of this one: its first decoded bit is then message bit M + 26 of this window,
    // the bit after the last one emitted now
    const size_t tail_off = vit_input_size(o, 2 * M), tail_bytes = win_bytes - tail_off;
    if (tail_bytes > h->carry_cap) {
        void* p = nullptr;
        VIT_CUDA(cudaMalloc(&p, tail_bytes + 1024));
        VIT_CUDA(cudaStreamSynchronize(st));
        if (h->carry_d) cudaFree(h->carry_d);
        h->carry_d = p; h->carry_cap = tail_bytes + 1024;
    }
    if (tail_bytes) VIT_CUDA(cudaMemcpyAsync(h->carry_d, win + tail_off, tail_bytes, cudaMemcpyDeviceToDevice, st));
    h->carry_syms = n - 2 * M;
    h->stream_bits += M;
    if (out_bytes) *out_bytes = need_out;
    if (!out_on_device || !chunk_on_device) VIT_CUDA(cudaStreamSynchronize(st));
    return VIT_OK;
}
}  // namespace

int vit_stream_reset(vit_handle* h) {
    if (!h) return fail(VIT_ERR_ARG, "null handle");
    h->carry_syms = 0; h->stream_bits = 0;
    return VIT_OK;
}

int vit_stream_push(vit_handle* h, const void* in_h, size_t inputNum, void* out_h, size_t out_cap, size_t* out_bytes) {
    if (!h || (!in_h && inputNum) || (!out_h && out_cap)) return fail(VIT_ERR_ARG, "null argument");
    VIT_ON_DEVICE(h);
    return stream_push(h, in_h, false, inputNum, out_h, false, out_cap, out_bytes, h->stream);
}

int vit_stream_push_device(vit_handle* h, const void* in_d, size_t inputNum, void* out_d, size_t out_cap, size_t* out_bytes,
                           void* cuda_stream) {
    if (!h || (!in_d && inputNum) || (!out_d && out_cap)) return fail(VIT_ERR_ARG, "null argument");
    VIT_ON_DEVICE(h);
    return stream_push(h, in_d, true, inputNum, out_d, true, out_cap, out_bytes, static_cast<cudaStream_t>(cuda_stream));
}

size_t vit_stream_pending(const vit_handle* h) { return h ? h->carry_syms : 0; }
unsigned long long vit_stream_bits(const vit_handle* h) { return h ? h->stream_bits : 0; }

int vit_set_upload_mode(vit_handle* h, int mode) {
    if (!h) return fail(VIT_ERR_ARG, "null handle");
    if (mode < VIT_UPLOAD_AUTO || mode > VIT_UPLOAD_GATED) return fail(VIT_ERR_ARG, "unknown upload mode %d", mode);
    h->upload_mode = mode;
    return VIT_OK;
}

int vit_set_geometry(vit_handle* h, int geometry) {
    if (!h) return fail(VIT_ERR_ARG, "null handle");
    if (geometry != VIT_GEOMETRY_L8 && geometry != VIT_GEOMETRY_L1) return fail(VIT_ERR_ARG, "unknown geometry %d", geometry);
    h->geometry = geometry;
    return VIT_OK;
}

int vit_last_launch_geometry(const vit_handle* h) { return h ? h->last_geometry : VIT_GEOMETRY_L8; }

int vit_upload_mode_in_effect(const vit_handle* h) {
    if (!h) return -1;
    if (h->upload_mode != VIT_UPLOAD_AUTO) return h->upload_mode;
    return h->gates_disabled ? VIT_UPLOAD_CHUNKED : VIT_UPLOAD_GATED;
}

#pragma GCC visibility pop
}  // extern "C"
