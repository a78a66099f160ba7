// sim_runtime.cpp -- TEST SCAFFOLDING ONLY: the stand-in CUDA runtime declared in tests/sim/cuda_runtime.h, and the kernel
// table that hands the launches of csrc/vit_api.cu (compiled for the host) to the emulator of the kernel source (tests/emu).
//
// Scheduling model (deliberately adversarial for the host code's upload protocols):
//   * every stream is an in-order queue; the queues are pumped after every enqueue, so an operation runs as soon as it is at
//     the head of its stream and its event dependencies are met -- in particular a kernel of the segment-range chunk pipeline
//     runs as soon as ITS chunk's bytes have been copied, before any later upload is even enqueued;
//   * "device" memory is poisoned on allocation, so bytes that have not been uploaded yet are garbage;
//   * a kernel launched with upload gates (vit_run's time-sliced upload) waits at the head of its stream and is RE-RUN from
//     scratch at every gate opening with the memory as it is at that moment: all warps advance to the first closed gate and
//     give up.  What such an attempt emits must already be final: every attempt's emitted words are compared with the
//     completed run's (sim_violations counts differences).  The attempt at a synchronisation point is the last one: if a
//     gate is still closed then, the kernel reports the time-out exactly as the device code does.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <vector>

#include "../../gpu-accelerated-viterbi-decoder_b200/csrc/vit_launch.h"

void vit_emu_run_kparams(const vitk::KParams& kp, int met, int in, int bpp, int tbl, int lanes);   // tests/emu/vit_emu.cpp

namespace {

struct Op {
    enum Kind { COPY, COPY2D, LAUNCH, RECORD, WAIT } kind;
    void* dst = nullptr; const void* src = nullptr; size_t n = 0, dpitch = 0, spitch = 0, width = 0, height = 0;
    SimEvent* ev = nullptr; unsigned long long ticket = 0;
    vitk::KParams kp{}; int met = 0, in = 0, bpp = 0, lanes = 8;
    std::vector<std::vector<uint8_t>> attempts;      // output snapshots of the gate-limited attempts
};
}  // namespace
struct SimStream { std::deque<Op> q; };
struct SimEvent { unsigned long long enqueued = 0, executed = 0; };

namespace {
std::vector<SimStream*> g_streams;
SimStream g_default_stream;
std::map<const char*, size_t> g_device, g_pinned;      // base -> bytes
unsigned long long g_violations = 0, g_kernel_runs = 0, g_gated_attempts = 0, g_copies = 0;
int g_tbl = 96;
bool g_in_pump = false;

SimStream* S(cudaStream_t s) { return s ? s : &g_default_stream; }

size_t out_bytes_of(const Op& op) { return (size_t)op.kp.packs * (size_t)(op.bpp / 8); }

bool in_range(const std::map<const char*, size_t>& m, const void* p) {
    auto it = m.upper_bound(static_cast<const char*>(p));
    if (it == m.begin()) return false;
    --it;
    return static_cast<const char*>(p) < it->first + it->second;
}

// run a launch; returns true if it completed (no warp gave up at a closed gate)
bool run_launch(Op& op, bool final_attempt) {
    const bool gated = op.kp.gate_n > 0;
    if (!gated) {
        vit_emu_run_kparams(op.kp, op.met, op.in, op.bpp, g_tbl, op.lanes);
        g_kernel_runs++;
        return true;
    }
    // one stream per gated launch in vit_run: poison the output, run, look at what was emitted
    const size_t nb = out_bytes_of(op);
    std::vector<uint8_t> before(op.kp.out, op.kp.out + nb);
    memset(op.kp.out, 0xDE, nb);
    const unsigned err_before = *op.kp.gate_err;
    *op.kp.gate_err = 0;
    vit_emu_run_kparams(op.kp, op.met, op.in, op.bpp, g_tbl, op.lanes);
    g_kernel_runs++; g_gated_attempts++;
    const bool gave_up = *op.kp.gate_err != 0;
    if (gave_up && !final_attempt) {
        op.attempts.emplace_back(op.kp.out, op.kp.out + nb);
        memcpy(op.kp.out, before.data(), nb);
        *op.kp.gate_err = err_before;
        return false;
    }
    if (!gave_up) {
        for (const auto& snap : op.attempts)
            for (size_t i = 0; i + 1 < nb; i += 2)            // 16-bit granules: the smallest pack
                if (!(snap[i] == 0xDE && snap[i + 1] == 0xDE) && (snap[i] != op.kp.out[i] || snap[i + 1] != op.kp.out[i + 1])) g_violations++;
    }
    return true;                                               // completed, or gave up for good (the host sees gate_err)
}

bool flag_write_of(const Op& copy, const Op& launch) {
    const char* d = static_cast<const char*>(copy.dst);
    const char* g = reinterpret_cast<const char*>(launch.kp.gate);
    return launch.kp.gate_n > 0 && d >= g && d < g + 8 * sizeof(unsigned);
}

// execute whatever can run; `drain` = a synchronisation point (parked gated kernels get their last attempt)
void pump(bool drain) {
    if (g_in_pump) return;
    g_in_pump = true;
    bool progress = true;
    while (progress) {
        progress = false;
        for (size_t si = 0; si <= g_streams.size(); si++) {
            SimStream* st = si < g_streams.size() ? g_streams[si] : &g_default_stream;
            while (!st->q.empty()) {
                Op& op = st->q.front();
                if (op.kind == Op::WAIT) { if (op.ev->executed < op.ticket) break; }
                else if (op.kind == Op::RECORD) op.ev->executed = op.ticket;
                else if (op.kind == Op::COPY) {
                    memmove(op.dst, op.src, op.n); g_copies++;
                    // a gate flag has been written: every parked kernel waiting on that gate array gets an attempt
                    for (SimStream* other : g_streams)
                        if (!other->q.empty() && other->q.front().kind == Op::LAUNCH && flag_write_of(op, other->q.front()))
                            if (run_launch(other->q.front(), false)) { other->q.pop_front(); }
                } else if (op.kind == Op::COPY2D) {
                    for (size_t r = 0; r < op.height; r++)
                        memmove(static_cast<char*>(op.dst) + r * op.dpitch, static_cast<const char*>(op.src) + r * op.spitch, op.width);
                    g_copies++;
                } else {   // LAUNCH
                    if (op.kp.gate_n > 0 && !drain) break;                        // parked until a gate opens or a sync
                    if (!run_launch(op, drain)) break;
                }
                st->q.pop_front();
                progress = true;
            }
        }
    }
    g_in_pump = false;
}

void enqueue(cudaStream_t s, Op&& op) {
    S(s)->q.push_back(std::move(op));
    pump(false);
}
}  // namespace

// ---- test hooks ------------------------------------------------------------------------------------------------------
extern "C" {
unsigned long long sim_violations(void) { return g_violations; }
unsigned long long sim_kernel_runs(void) { return g_kernel_runs; }
unsigned long long sim_gated_attempts(void) { return g_gated_attempts; }
void sim_reset_counters(void) { g_violations = g_kernel_runs = g_gated_attempts = g_copies = 0; }
void sim_set_table(int tbl) { g_tbl = tbl == 32 ? 32 : 96; }
void* sim_pinned_alloc(size_t n) { void* p = nullptr; cudaHostAlloc(&p, n, 0); return p; }
void sim_pinned_free(void* p) { cudaFreeHost(p); }
void* sim_device_alloc(size_t n) { void* p = nullptr; cudaMalloc(&p, n); return p; }
void sim_device_free(void* p) { cudaFree(p); }
}

// ---- the runtime -----------------------------------------------------------------------------------------------------
const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "simulated CUDA error"; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
// Every allocation sits between two 256-byte canaries (0xC5) that are checked when it is freed (and by sim_check_canaries):
// a write outside a device or pinned buffer -- by the host code's copies or by an emulated kernel -- counts as a violation.
static char* guarded_alloc(std::map<const char*, size_t>& reg, size_t n, int fill) {
    char* raw = static_cast<char*>(aligned_alloc(256, (n + 768 + 255) / 256 * 256));
    if (!raw) return nullptr;
    memset(raw, 0xC5, 256);
    memset(raw + 256, fill, n);
    memset(raw + 256 + n, 0xC5, 512);                         // (n + 768 bytes in all)
    reg[raw + 256] = n;
    return raw + 256;
}
static void check_canaries(const char* user, size_t n) {
    const unsigned char* lo = reinterpret_cast<const unsigned char*>(user) - 256;
    const unsigned char* hi = reinterpret_cast<const unsigned char*>(user) + n;
    for (int i = 0; i < 256; i++) if (lo[i] != 0xC5 || hi[i] != 0xC5) { g_violations++; return; }
}
static void guarded_free(std::map<const char*, size_t>& reg, void* p) {
    auto it = reg.find(static_cast<const char*>(p));
    if (it == reg.end()) { g_violations++; return; }          // not ours, or freed twice
    check_canaries(it->first, it->second);
    reg.erase(it);
    free(static_cast<char*>(p) - 256);
}
extern "C" void sim_check_canaries(void) {
    for (const auto& kv : g_device) check_canaries(kv.first, kv.second);
    for (const auto& kv : g_pinned) check_canaries(kv.first, kv.second);
}
cudaError_t cudaMalloc(void** p, size_t n) {
    char* m = guarded_alloc(g_device, n, 0xCD);              // not-yet-uploaded bytes are garbage
    if (!m) return cudaErrorMemoryAllocation;
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) { if (p) { pump(true); guarded_free(g_device, p); } return cudaSuccess; }
cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
    char* m = guarded_alloc(g_pinned, n, 0xAB);
    if (!m) return cudaErrorMemoryAllocation;
    *p = m;
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void* p) { if (p) { pump(true); guarded_free(g_pinned, p); } return cudaSuccess; }
cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
cudaError_t cudaMemset(void* p, int v, size_t n) { pump(true); memset(p, v, n); return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    a->device = 0; a->devicePointer = const_cast<void*>(p); a->hostPointer = const_cast<void*>(p);
    a->type = in_range(g_pinned, p) ? cudaMemoryTypeHost : in_range(g_device, p) ? cudaMemoryTypeDevice : cudaMemoryTypeUnregistered;
    return cudaSuccess;
}
cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, const void*) { a->numRegs = 0; a->sharedSizeBytes = 0; return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new SimStream(); g_streams.push_back(*s); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
    pump(true);
    for (size_t i = 0; i < g_streams.size(); i++) if (g_streams[i] == s) { g_streams.erase(g_streams.begin() + (long)i); break; }
    delete s;
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { pump(true); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    Op op; op.kind = Op::WAIT; op.ev = e; op.ticket = e->enqueued;
    enqueue(s, std::move(op));
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new SimEvent(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { pump(true); delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    Op op; op.kind = Op::RECORD; op.ev = e; op.ticket = ++e->enqueued;
    enqueue(s, std::move(op));
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { pump(true); return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t s) {
    Op op; op.kind = Op::COPY; op.dst = dst; op.src = src; op.n = n;
    enqueue(s, std::move(op));
    return cudaSuccess;
}
cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t s) {
    Op op; op.kind = Op::COPY2D; op.dst = dst; op.src = src; op.dpitch = dpitch; op.spitch = spitch; op.width = width; op.height = height;
    enqueue(s, std::move(op));
    return cudaSuccess;
}

// ---- kernel table: every (core, input type, pack width) "kernel" is the emulator --------------------------------------
namespace vitk {
namespace {
cudaError_t enqueue_launch(const KParams& kp, cudaStream_t st, int met, int in, int bpp, int lanes = 8) {
    Op op; op.kind = Op::LAUNCH; op.kp = kp; op.met = met; op.in = in; op.bpp = bpp; op.lanes = lanes;
    // a gate-waiting kernel is launched BEFORE its input is uploaded: whatever an earlier call left in the device buffer
    // must not be able to stand in for bytes that have not arrived yet
    if (kp.gate_n > 0) memset(const_cast<uint8_t*>(kp.in), 0xCD, (size_t)kp.in_bytes);
    enqueue(st, std::move(op));
    return cudaSuccess;
}
template <int MET, int IN, int BPP> cudaError_t sim_launch(const KParams& kp, cudaStream_t st) { return enqueue_launch(kp, st, MET, IN, BPP); }
template <int MET, int IN, int BPP> cudaError_t sim_launch_l1(const KParams& kp, cudaStream_t st) { return enqueue_launch(kp, st, MET, IN, BPP, 1); }
template <int MET> const KernelEntry* entry(int in, int bpp16) {
    // as in the library: the one-lane-per-segment geometry exists for the packed cores only
#define E(IN, BPP) {&sim_launch<MET, IN, BPP>, (const void*)&sim_launch<MET, IN, BPP>, 0, (MET == MET_B16 || MET == MET_F16) ? &sim_launch_l1<MET, IN, BPP> : nullptr}
    static const KernelEntry table[5][2] = {{E(0, 32), E(0, 16)}, {E(1, 32), E(1, 16)}, {E(2, 32), E(2, 16)}, {E(3, 32), E(3, 16)}, {E(4, 32), E(4, 16)}};
#undef E
    if (in < 0 || in > 4) return nullptr;
    if (MET == MET_B16 && in == IN_S16) return nullptr;               // as the library: rejected like the reference
    return &table[in][bpp16 ? 1 : 0];
}
}  // namespace
const KernelEntry* kernel_entry_b32(int in, int bpp16) { return entry<MET_B32>(in, bpp16); }
const KernelEntry* kernel_entry_b32d(int in, int bpp16) { return entry<MET_B32D>(in, bpp16); }
const KernelEntry* kernel_entry_b16(int in, int bpp16) { return entry<MET_B16>(in, bpp16); }
const KernelEntry* kernel_entry_f16(int in, int bpp16) { return entry<MET_F16>(in, bpp16); }
}  // namespace vitk
