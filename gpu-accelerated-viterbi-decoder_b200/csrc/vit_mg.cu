// vit_mg.cu -- multi-GPU part of the C ABI (include/vit_b200.h): independent codeword streams sharded over the GPUs
// of one box, one decoder per GPU, no stream ever split, and ONLY the packed output bits cross GPUs: they are gathered
// into one stream-major buffer on a root GPU.
//
// The reference has nothing here (cudaSetDevice(0), reference src/viterbi/viterbi.cu:134); SURVEY.md 8e / BASELINE.json
// configs[4] define the job: 1024 independent 256-Mbit s8 streams over 1/2/4/8 B200.
//
// Gather mechanisms (vit_comm_gatherv `mode`), all landing in the same root buffer:
//   VIT_GATHER_NCCL    grouped ncclSend / ncclRecv to the root on a side stream, per finished wave of streams, so that
//                      it overlaps the decode of the next wave (NCCL kernels: a few CTAs beside the decode kernel);
//   VIT_GATHER_COPY    one device-to-device copy per finished wave into the root's buffer (mapped into every rank
//                      through CUDA IPC or peer access): copy engines over NVLink, no SM at all;
//   VIT_GATHER_DIRECT  no gather step: the decode kernel stores its packs straight into the mapped root buffer over
//                      NVLink (4-byte stores, ~14 GB/s per GPU against 900 GB/s per link direction);
//   VIT_GATHER_NONE    outputs stay on the decoding GPU (the "without gather" figure).
// NCCL is loaded at run time (dlopen "libnccl.so.2": in a torch process that is torch's own copy), so the library itself
// has no NCCL link dependency and loads on a box without it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/vit_b200.h"
#include "vit_internal.h"

namespace {

int mg_fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return vit_set_error(code, buf);
}
#define MG_CUDA(call)                                                                                 \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return mg_fail(VIT_ERR_CUDA, "%s in %s at line %d", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ---- NCCL entry points, resolved at run time ---------------------------------------------------------------------
struct NcclApi {
    void* so = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

const NcclApi* nccl() {
    static NcclApi api;
    static const bool ok = [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.so) break;
        }
        if (!api.so) return false;
#define VIT_SYM(field, name) *(void**)(&api.field) = dlsym(api.so, name); if (!api.field) return false;
        VIT_SYM(GetUniqueId, "ncclGetUniqueId")
        VIT_SYM(CommInitRank, "ncclCommInitRank")
        VIT_SYM(CommDestroy, "ncclCommDestroy")
        VIT_SYM(Send, "ncclSend")
        VIT_SYM(Recv, "ncclRecv")
        VIT_SYM(AllReduce, "ncclAllReduce")
        VIT_SYM(Broadcast, "ncclBroadcast")
        VIT_SYM(GroupStart, "ncclGroupStart")
        VIT_SYM(GroupEnd, "ncclGroupEnd")
        VIT_SYM(GetVersion, "ncclGetVersion")
        VIT_SYM(GetErrorString, "ncclGetErrorString")
#undef VIT_SYM
        *(void**)(&api.CommInitRankConfig) = dlsym(api.so, "ncclCommInitRankConfig");   // optional
        return true;
    }();
    return ok ? &api : nullptr;
}
#define MG_NCCL(call)                                                                                 \
    do {                                                                                              \
        ncclResult_t r_ = (call);                                                                     \
        if (r_ != ncclSuccess)                                                                        \
            return mg_fail(VIT_ERR_NCCL, "NCCL: %s in %s at line %d", nccl()->GetErrorString(r_), __FILE__, __LINE__); \
    } while (0)

struct DevGuard {
    int prev = -1;
    explicit DevGuard(int device) { cudaGetDevice(&prev); if (prev != device) cudaSetDevice(device); else prev = -1; }
    ~DevGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

struct vit_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
    bool single_process = false;         // created by vit_comm_init_all: peers are threads of this process
    int peer_devices[64] = {};           // single_process: device ordinal of every rank
    cudaStream_t gstream = nullptr;      // gathers run here, beside the decode stream
    cudaEvent_t ev_prod = nullptr, ev_done = nullptr;
    cudaEvent_t ev_mark[8] = {};         // vit_comm_mark / vit_comm_stream_wait_mark
    bool mark_set[8] = {};
    int* scratch_d = nullptr;            // barrier word + IPC handle exchange
    struct Shared { void* ptr; bool mapped; };
    std::vector<Shared> shared;          // vit_comm_shared_alloc results (root: owned, peers: IPC mappings)
};

namespace {

int comm_finish_init(vit_comm* c) {
    DevGuard g(c->device);
    int lo = 0, hi = 0;
    MG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    // the gathers are short and latency matters (they release output slots): highest priority, so their blocks are
    // placed as soon as a decode block leaves an SM
    MG_CUDA(cudaStreamCreateWithPriority(&c->gstream, cudaStreamNonBlocking, hi));
    MG_CUDA(cudaEventCreateWithFlags(&c->ev_prod, cudaEventDisableTiming));
    MG_CUDA(cudaEventCreateWithFlags(&c->ev_done, cudaEventDisableTiming));
    for (auto& e : c->ev_mark) MG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    MG_CUDA(cudaMalloc(reinterpret_cast<void**>(&c->scratch_d), 256));
    MG_CUDA(cudaMemset(c->scratch_d, 0, 256));
    return VIT_OK;
}

ncclConfig_t comm_config() {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    // NCCL's kernels share the SMs with an issue-bound decode kernel: VIT_NCCL_MAX_CTAS bounds how many they take
    if (const char* e = getenv("VIT_NCCL_MAX_CTAS")) {
        const int n = atoi(e);
        if (n > 0) { cfg.maxCTAs = n; cfg.minCTAs = 1; }
    }
    return cfg;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int vit_comm_available(void) { return nccl() != nullptr; }

int vit_comm_nccl_version(void) {
    int v = 0;
    if (!nccl() || nccl()->GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

int vit_comm_get_unique_id(void* id128) {
    if (!id128) return mg_fail(VIT_ERR_ARG, "null id buffer");
    if (!nccl()) return mg_fail(VIT_ERR_NCCL, "libnccl.so.2 not found");
    ncclUniqueId id;
    MG_NCCL(nccl()->GetUniqueId(&id));
    static_assert(sizeof id == VIT_COMM_ID_BYTES, "NCCL unique id size");
    memcpy(id128, &id, sizeof id);
    return VIT_OK;
}

int vit_comm_init_rank(vit_comm** out, int nranks, int rank, const void* id128, int device) {
    if (!out || !id128 || nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) return mg_fail(VIT_ERR_ARG, "bad communicator arguments");
    *out = nullptr;
    if (!nccl()) return mg_fail(VIT_ERR_NCCL, "libnccl.so.2 not found");
    vit_comm* c = new (std::nothrow) vit_comm();
    if (!c) return mg_fail(VIT_ERR_ARG, "out of host memory");
    c->rank = rank; c->nranks = nranks; c->device = device;
    DevGuard g(device);
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclConfig_t cfg = comm_config();
    ncclResult_t r = nccl()->CommInitRankConfig ? nccl()->CommInitRankConfig(&c->comm, nranks, id, rank, &cfg)
                                                : nccl()->CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { delete c; return mg_fail(VIT_ERR_NCCL, "NCCL: %s (communicator init)", nccl()->GetErrorString(r)); }
    int rc = comm_finish_init(c);
    if (rc) { vit_comm_destroy(c); return rc; }
    *out = c;
    return VIT_OK;
}

int vit_comm_init_all(vit_comm** out, int ndev, const int* devices) {
    if (!out || ndev < 1 || ndev > 64) return mg_fail(VIT_ERR_ARG, "bad communicator arguments");
    for (int i = 0; i < ndev; i++) out[i] = nullptr;
    if (!nccl()) return mg_fail(VIT_ERR_NCCL, "libnccl.so.2 not found");
    ncclUniqueId id;
    MG_NCCL(nccl()->GetUniqueId(&id));
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<vit_comm*> cs(ndev, nullptr);
    for (int i = 0; i < ndev; i++) {
        cs[i] = new (std::nothrow) vit_comm();
        if (!cs[i]) return mg_fail(VIT_ERR_ARG, "out of host memory");
        cs[i]->rank = i; cs[i]->nranks = ndev; cs[i]->device = devices ? devices[i] : i; cs[i]->single_process = true;
    }
    for (int i = 0; i < ndev; i++)
        for (int j = 0; j < ndev; j++) cs[i]->peer_devices[j] = cs[j]->device;
    ncclConfig_t cfg = comm_config();
    ncclResult_t r = nccl()->GroupStart();
    for (int i = 0; i < ndev && r == ncclSuccess; i++) {
        cudaSetDevice(cs[i]->device);
        r = nccl()->CommInitRankConfig ? nccl()->CommInitRankConfig(&cs[i]->comm, ndev, id, i, &cfg)
                                       : nccl()->CommInitRank(&cs[i]->comm, ndev, id, i);
    }
    ncclResult_t r2 = nccl()->GroupEnd();
    if (r == ncclSuccess) r = r2;
    int rc = VIT_OK;
    if (r != ncclSuccess) rc = mg_fail(VIT_ERR_NCCL, "NCCL: %s (communicator init)", nccl()->GetErrorString(r));
    for (int i = 0; i < ndev && rc == VIT_OK; i++) rc = comm_finish_init(cs[i]);
    // copies and kernel stores into the root's buffer need peer access in the single-process case
    for (int i = 0; i < ndev && rc == VIT_OK; i++) {
        cudaSetDevice(cs[i]->device);
        for (int j = 0; j < ndev; j++) {
            if (i == j) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, cs[i]->device, cs[j]->device);
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(cs[j]->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) rc = mg_fail(VIT_ERR_CUDA, "%s (peer access %d -> %d)", cudaGetErrorString(e), cs[i]->device, cs[j]->device);
                cudaGetLastError();
            }
        }
    }
    cudaSetDevice(prev);
    if (rc != VIT_OK) {
        for (int i = 0; i < ndev; i++) vit_comm_destroy(cs[i]);
        return rc;
    }
    for (int i = 0; i < ndev; i++) out[i] = cs[i];
    return VIT_OK;
}

void vit_comm_destroy(vit_comm* c) {
    if (!c) return;
    DevGuard g(c->device);
    for (auto& s : c->shared) {
        if (!s.ptr) continue;
        if (s.mapped) { vit_forget_remote_range(s.ptr); cudaIpcCloseMemHandle(s.ptr); }
        else { vit_forget_remote_range(s.ptr); cudaFree(s.ptr); }
    }
    if (c->gstream) cudaStreamDestroy(c->gstream);
    if (c->ev_prod) cudaEventDestroy(c->ev_prod);
    if (c->ev_done) cudaEventDestroy(c->ev_done);
    for (auto e : c->ev_mark) if (e) cudaEventDestroy(e);
    if (c->scratch_d) cudaFree(c->scratch_d);
    if (c->comm && nccl()) nccl()->CommDestroy(c->comm);
    delete c;
}

int vit_comm_rank(const vit_comm* c) { return c ? c->rank : 0; }
int vit_comm_size(const vit_comm* c) { return c ? c->nranks : 1; }

// contiguous block partition: rank r owns streams [first, first + count); blocks differ by at most one stream
void vit_shard_range(size_t nstreams, int nranks, int rank, size_t* first, size_t* count) {
    if (nranks < 1) nranks = 1;
    const size_t q = nstreams / (size_t)nranks, r = nstreams % (size_t)nranks;
    const size_t k = (size_t)rank;
    if (first) *first = q * k + std::min(k, r);
    if (count) *count = q + (k < r ? 1 : 0);
}

int vit_shard_owner(size_t nstreams, int nranks, size_t stream) {
    if (nranks < 1) nranks = 1;
    const size_t q = nstreams / (size_t)nranks, r = nstreams % (size_t)nranks, edge = (q + 1) * r;
    if (stream < edge) return (int)(stream / (q + 1));
    return (int)(r + (q ? (stream - edge) / q : 0));
}

// all ranks: wait (on the host) for every gather this rank issued, then meet the other ranks.  After it returns the
// root's buffer holds every block gathered before the call.
int vit_comm_barrier(vit_comm* c) {
    if (!c) return VIT_OK;
    DevGuard g(c->device);
    MG_CUDA(cudaStreamSynchronize(c->gstream));
    if (c->nranks > 1) {
        MG_NCCL(nccl()->AllReduce(c->scratch_d, c->scratch_d, 1, ncclInt32, ncclSum, c->comm, c->gstream));
        MG_CUDA(cudaStreamSynchronize(c->gstream));
    }
    return VIT_OK;
}

// make `stream` wait for every gather issued so far on this rank (e.g. before a decode overwrites a gathered slot)
int vit_comm_stream_wait(vit_comm* c, void* stream) {
    if (!c) return VIT_OK;
    DevGuard g(c->device);
    MG_CUDA(cudaEventRecord(c->ev_done, c->gstream));
    MG_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->ev_done, 0));
    return VIT_OK;
}

// Markers: vit_comm_mark(k) remembers the point the gather stream has reached (after the gathers issued so far);
// vit_comm_stream_wait_mark(k) makes `stream` wait for that point only -- unlike vit_comm_stream_wait it does not wait
// for gathers issued after the mark (e.g. the one that is still reading the slot decoded a moment ago).
int vit_comm_mark(vit_comm* c, int k) {
    if (!c) return VIT_OK;
    if (k < 0 || k >= 8) return mg_fail(VIT_ERR_ARG, "marker index out of range");
    DevGuard g(c->device);
    MG_CUDA(cudaEventRecord(c->ev_mark[k], c->gstream));
    c->mark_set[k] = true;
    return VIT_OK;
}
int vit_comm_stream_wait_mark(vit_comm* c, int k, void* stream) {
    if (!c) return VIT_OK;
    if (k < 0 || k >= 8) return mg_fail(VIT_ERR_ARG, "marker index out of range");
    if (!c->mark_set[k]) return VIT_OK;
    DevGuard g(c->device);
    MG_CUDA(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), c->ev_mark[k], 0));
    return VIT_OK;
}

void* vit_comm_stream(vit_comm* c) { return c ? c->gstream : nullptr; }

// Collective: a device buffer of `bytes` on the root's GPU that every rank can address.  *ptr is the root's own
// allocation on the root and a mapping of it on the other ranks (CUDA IPC between processes, peer access between the
// threads of one process).  Freed by vit_comm_destroy.
int vit_comm_shared_alloc(vit_comm* c, void** ptr, size_t bytes, int root) {
    if (!c || !ptr || root < 0 || root >= c->nranks) return mg_fail(VIT_ERR_ARG, "bad shared allocation arguments");
    *ptr = nullptr;
    DevGuard g(c->device);
    void* mine = nullptr;
    if (c->rank == root) {
        MG_CUDA(cudaMalloc(&mine, bytes ? bytes : 256));
        c->shared.push_back({mine, false});
    }
    if (c->nranks == 1) { *ptr = mine; return VIT_OK; }
    // ship 64 bytes of IPC handle + the raw pointer (for threads of the same process) through a broadcast
    struct Msg { cudaIpcMemHandle_t h; unsigned long long raw; } msg;
    memset(&msg, 0, sizeof msg);
    static_assert(sizeof(Msg) <= 128, "scratch size");
    if (c->rank == root) {
        msg.raw = (unsigned long long)(uintptr_t)mine;
        if (!c->single_process) MG_CUDA(cudaIpcGetMemHandle(&msg.h, mine));
        MG_CUDA(cudaMemcpyAsync(c->scratch_d + 16, &msg, sizeof msg, cudaMemcpyHostToDevice, c->gstream));
    }
    MG_NCCL(nccl()->Broadcast(c->scratch_d + 16, c->scratch_d + 16, sizeof msg, ncclUint8, root, c->comm, c->gstream));
    MG_CUDA(cudaMemcpyAsync(&msg, c->scratch_d + 16, sizeof msg, cudaMemcpyDeviceToHost, c->gstream));
    MG_CUDA(cudaStreamSynchronize(c->gstream));
    if (c->rank == root) { *ptr = mine; return VIT_OK; }
    if (c->single_process) {
        *ptr = (void*)(uintptr_t)msg.raw;                  // same address space; peer access was enabled at init
        vit_note_remote_range(*ptr, bytes, c->peer_devices[root]);
        return VIT_OK;
    }
    void* mapped = nullptr;
    MG_CUDA(cudaIpcOpenMemHandle(&mapped, msg.h, cudaIpcMemLazyEnablePeerAccess));
    c->shared.push_back({mapped, true});
    vit_note_remote_range(mapped, bytes, -1);
    *ptr = mapped;
    return VIT_OK;
}

// Gather blocks of packed output bits to the root.  Every rank passes the same offsets[] / sizes[] (bytes, one entry per
// rank): rank p's block (sizes[p] bytes at send_d on rank p) lands at recv_base + offsets[p] on the root.  recv_base is
// the root's buffer -- for VIT_GATHER_COPY it must come from vit_comm_shared_alloc (the sender writes through its
// mapping).  The operation is ordered after the work already queued on producer_stream and runs on the communicator's
// own stream; completion: vit_comm_barrier (host) or vit_comm_stream_wait.
int vit_comm_gatherv(vit_comm* c, int mode, const void* send_d, void* recv_base, const size_t* offsets,
                     const size_t* sizes, int root, void* producer_stream) {
    if (!c || !offsets || !sizes || root < 0 || root >= c->nranks) return mg_fail(VIT_ERR_ARG, "bad gather arguments");
    if (mode == VIT_GATHER_NONE || mode == VIT_GATHER_DIRECT) return VIT_OK;
    if (mode != VIT_GATHER_NCCL && mode != VIT_GATHER_COPY) return mg_fail(VIT_ERR_ARG, "unknown gather mode %d", mode);
    DevGuard g(c->device);
    MG_CUDA(cudaEventRecord(c->ev_prod, static_cast<cudaStream_t>(producer_stream)));
    MG_CUDA(cudaStreamWaitEvent(c->gstream, c->ev_prod, 0));
    char* base = static_cast<char*>(recv_base);
    const size_t mine = sizes[c->rank];
    if (c->rank == root) {
        // the root's own block: already in place when it decoded straight into the buffer
        if (mine && send_d != base + offsets[root])
            MG_CUDA(cudaMemcpyAsync(base + offsets[root], send_d, mine, cudaMemcpyDeviceToDevice, c->gstream));
    }
    if (c->nranks == 1) return VIT_OK;
    if (mode == VIT_GATHER_COPY) {
        if (c->rank != root && mine)
            MG_CUDA(cudaMemcpyAsync(base + offsets[c->rank], send_d, mine, cudaMemcpyDefault, c->gstream));
        return VIT_OK;
    }
    MG_NCCL(nccl()->GroupStart());
    ncclResult_t r = ncclSuccess;
    if (c->rank != root) {
        if (mine) r = nccl()->Send(send_d, mine, ncclUint8, root, c->comm, c->gstream);
    } else {
        for (int p = 0; p < c->nranks && r == ncclSuccess; p++)
            if (p != root && sizes[p]) r = nccl()->Recv(base + offsets[p], sizes[p], ncclUint8, p, c->comm, c->gstream);
    }
    ncclResult_t r2 = nccl()->GroupEnd();
    if (r == ncclSuccess) r = r2;
    if (r != ncclSuccess) return mg_fail(VIT_ERR_NCCL, "NCCL: %s (gather)", nccl()->GetErrorString(r));
    return VIT_OK;
}

#pragma GCC visibility pop
}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// The sharded stream job (BASELINE.json configs[4]): nstreams independent streams of n_bits message bits, generated on
// the device per batch, decoded `wave` streams per launch, outputs gathered per finished wave.
// ---------------------------------------------------------------------------------------------------------------------
struct vit_job {
    vit_job_config cfg;
    vit_comm* comm = nullptr;
    int device = 0, rank = 0, nranks = 1;
    vit_handle* dec = nullptr;
    size_t N = 0, M = 0, in_bytes = 0, out_bytes = 0, in_stride = 0, out_stride = 0;
    size_t first = 0, count = 0;         // this rank's streams
    size_t max_count = 0;                // largest block of any rank (all ranks run that many batch rounds)
    void* in_d = nullptr;                // batch x in_stride
    void* out_d = nullptr;               // batch x out_stride (ranks that do not decode straight into the root buffer)
    void* gathered = nullptr;            // root buffer: nstreams x out_stride (own allocation on the root, mapping elsewhere)
    bool direct = false;                 // this rank decodes straight into `gathered`
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_c = nullptr, ev_s0 = nullptr, ev_s1 = nullptr;
    std::vector<unsigned long long> errs;
};

namespace {

char* job_out_ptr(const vit_job* j, size_t local_idx, size_t idx_in_batch) {
    if (j->direct) return static_cast<char*>(j->gathered) + (j->first + local_idx) * j->out_stride;
    return static_cast<char*>(j->out_d) + idx_in_batch * j->out_stride;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int vit_job_create(vit_job** out, vit_comm* comm, int device, const vit_job_config* cfg) {
    if (!out || !cfg) return mg_fail(VIT_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->nstreams == 0 || cfg->n_bits < 64 + 32 || cfg->wave == 0 || cfg->batch == 0)
        return mg_fail(VIT_ERR_ARG, "bad job configuration");
    if (cfg->gather < VIT_GATHER_NONE || cfg->gather > VIT_GATHER_DIRECT) return mg_fail(VIT_ERR_ARG, "unknown gather mode %d", cfg->gather);
    vit_job* j = new (std::nothrow) vit_job();
    if (!j) return mg_fail(VIT_ERR_ARG, "out of host memory");
    j->cfg = *cfg; j->comm = comm; j->device = comm ? comm->device : device;
    if (!comm) j->cfg.root = 0;
    if (j->cfg.root < 0 || (comm && j->cfg.root >= comm->nranks)) { delete j; return mg_fail(VIT_ERR_ARG, "bad root rank"); }
    j->rank = comm ? comm->rank : 0; j->nranks = comm ? comm->nranks : 1;
    if (j->cfg.batch < j->cfg.wave) j->cfg.batch = j->cfg.wave;
    j->cfg.batch = j->cfg.batch / j->cfg.wave * j->cfg.wave;
    DevGuard g(j->device);
    int rc = vit_create(&j->dec, cfg->options, j->device, 0);
    if (rc) { delete j; return rc; }
    j->N = 2 * cfg->n_bits;
    j->M = vit_message_len(cfg->options, j->N);
    j->in_bytes = vit_input_size(cfg->options, j->N);
    j->out_bytes = vit_output_size(cfg->options, j->N);
    j->in_stride = (j->in_bytes + 4 + 255) / 256 * 256;       // the source writes whole 32-bit packs
    j->out_stride = (j->out_bytes + 255) / 256 * 256;
    vit_shard_range(cfg->nstreams, j->nranks, j->rank, &j->first, &j->count);
    size_t f0 = 0;
    vit_shard_range(cfg->nstreams, j->nranks, 0, &f0, &j->max_count);
    const int root = j->cfg.root;
    const bool gathers = cfg->gather != VIT_GATHER_NONE;
    auto bail = [&](int code) { vit_job_destroy(j); return code; };
    cudaError_t e = cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking);
    for (cudaEvent_t* ev : {&j->ev_a, &j->ev_b, &j->ev_c, &j->ev_s0, &j->ev_s1})
        if (e == cudaSuccess) e = cudaEventCreate(ev);
    if (e == cudaSuccess) e = cudaMalloc(&j->in_d, (size_t)j->cfg.batch * j->in_stride);
    if (e != cudaSuccess) return bail(mg_fail(VIT_ERR_CUDA, "%s in %s at line %d", cudaGetErrorString(e), __FILE__, __LINE__));
    if (gathers) {
        // one stream-major buffer on the root for the whole job
        const size_t total = (size_t)cfg->nstreams * j->out_stride;
        if (comm) rc = vit_comm_shared_alloc(comm, &j->gathered, total, root);
        else rc = cudaMalloc(&j->gathered, total) == cudaSuccess ? VIT_OK : mg_fail(VIT_ERR_CUDA, "cudaMalloc of the gathered buffer (%zu bytes) failed", total);
        if (rc) return bail(rc);
    }
    j->direct = gathers && (j->rank == root || cfg->gather == VIT_GATHER_DIRECT);
    if (!j->direct) {
        e = cudaMalloc(&j->out_d, (size_t)j->cfg.batch * j->out_stride);
        if (e != cudaSuccess) return bail(mg_fail(VIT_ERR_CUDA, "%s in %s at line %d", cudaGetErrorString(e), __FILE__, __LINE__));
    }
    j->errs.assign(j->count, 0);
    *out = j;
    return VIT_OK;
}

void vit_job_destroy(vit_job* j) {
    if (!j) return;
    DevGuard g(j->device);
    if (j->dec) vit_destroy(j->dec);
    if (j->in_d) cudaFree(j->in_d);
    if (j->out_d) cudaFree(j->out_d);
    if (j->gathered && !j->comm) cudaFree(j->gathered);        // with a communicator: owned by it (shared allocation)
    for (cudaEvent_t ev : {j->ev_a, j->ev_b, j->ev_c, j->ev_s0, j->ev_s1})
        if (ev) cudaEventDestroy(ev);
    if (j->stream) cudaStreamDestroy(j->stream);
    delete j;
}

const void* vit_job_gathered(const vit_job* j, size_t* out_stride) {
    if (!j) return nullptr;
    if (out_stride) *out_stride = j->out_stride;
    return j->gathered;
}

int vit_job_stream_range(const vit_job* j, size_t* first, size_t* count) {
    if (!j) return mg_fail(VIT_ERR_ARG, "null job");
    if (first) *first = j->first;
    if (count) *count = j->count;
    return VIT_OK;
}

int vit_job_stream_errors(const vit_job* j, unsigned long long* errs, size_t cap) {
    if (!j || !errs) return mg_fail(VIT_ERR_ARG, "null argument");
    for (size_t i = 0; i < std::min(cap, j->errs.size()); i++) errs[i] = j->errs[i];
    return VIT_OK;
}

// One pass over the whole job.  Per batch round: generate this rank's next `batch` streams (untimed), meet the other
// ranks, then -- timed with events -- decode them `wave` streams per launch and gather every finished wave on the
// communicator's stream while the next wave decodes; the bit errors of every stream are counted after the timed part.
int vit_job_run(vit_job* j, vit_job_result* res) {
    if (!j || !res) return mg_fail(VIT_ERR_ARG, "null argument");
    memset(res, 0, sizeof *res);
    DevGuard g(j->device);
    const vit_job_config& c = j->cfg;
    const int it = c.options & 0xf;
    const size_t rounds = (j->max_count + c.batch - 1) / c.batch;
    std::vector<size_t> offsets(j->nranks), sizes(j->nranks);
    for (size_t round = 0; round < rounds; round++) {
        const size_t b0 = round * c.batch;                                         // local index of the batch's first stream
        const size_t nb = b0 < j->count ? std::min<size_t>(c.batch, j->count - b0) : 0;
        // ---- generation (untimed) ----
        MG_CUDA(cudaEventRecord(j->ev_s0, j->stream));
        for (size_t k = 0; k < nb; k++) {
            const unsigned seed = c.seed + (unsigned)(j->first + b0 + k);
            int rc = vit_synth_device_ex(it, c.n_bits, seed, c.amp, c.sigma, 0, c.source,
                                         static_cast<char*>(j->in_d) + k * j->in_stride, nullptr, j->stream);
            if (rc) return mg_fail(rc, "synthetic source failed for stream %zu", j->first + b0 + k);
        }
        MG_CUDA(cudaEventRecord(j->ev_s1, j->stream));
        MG_CUDA(cudaStreamSynchronize(j->stream));
        float ms = 0.f;
        MG_CUDA(cudaEventElapsedTime(&ms, j->ev_s0, j->ev_s1));
        res->synth_ms += ms;
        if (j->comm) { int rc = vit_comm_barrier(j->comm); if (rc) return rc; }     // ranks enter the timed part together
        // ---- decode + gather (timed) ----
        MG_CUDA(cudaEventRecord(j->ev_a, j->stream));
        const size_t max_nb = b0 < j->max_count ? std::min<size_t>(c.batch, j->max_count - b0) : 0;
        for (size_t w0 = 0; w0 < max_nb; w0 += c.wave) {
            const size_t nw = w0 < nb ? std::min<size_t>(c.wave, nb - w0) : 0;
            if (nw) {
                int rc = vit_run_device_batch(j->dec, static_cast<char*>(j->in_d) + w0 * j->in_stride, job_out_ptr(j, b0 + w0, w0),
                                              j->N, nw, j->in_stride, j->out_stride, j->stream, nullptr);
                if (rc) return rc;
                res->launches++;
            }
            if (j->comm && (c.gather == VIT_GATHER_NCCL || c.gather == VIT_GATHER_COPY)) {
                // every rank derives every rank's block of this wave from the partition alone
                for (int p = 0; p < j->nranks; p++) {
                    size_t pf, pc;
                    vit_shard_range(c.nstreams, j->nranks, p, &pf, &pc);
                    const size_t pnb = b0 < pc ? std::min<size_t>(c.batch, pc - b0) : 0;
                    const size_t pnw = w0 < pnb ? std::min<size_t>(c.wave, pnb - w0) : 0;
                    offsets[p] = (pf + b0 + w0) * j->out_stride;
                    sizes[p] = pnw * j->out_stride;
                }
                int rc = vit_comm_gatherv(j->comm, c.gather, nw ? job_out_ptr(j, b0 + w0, w0) : nullptr, j->gathered,
                                          offsets.data(), sizes.data(), c.root, j->stream);
                if (rc) return rc;
            }
        }
        MG_CUDA(cudaEventRecord(j->ev_b, j->stream));
        cudaStream_t last = j->stream;
        if (j->comm && (c.gather == VIT_GATHER_NCCL || c.gather == VIT_GATHER_COPY)) last = j->comm->gstream;
        if (last != j->stream) MG_CUDA(cudaStreamWaitEvent(last, j->ev_b, 0));
        MG_CUDA(cudaEventRecord(j->ev_c, last));
        MG_CUDA(cudaEventSynchronize(j->ev_c));
        MG_CUDA(cudaEventElapsedTime(&ms, j->ev_a, j->ev_b));
        res->decode_ms += ms;
        MG_CUDA(cudaEventElapsedTime(&ms, j->ev_a, j->ev_c));
        res->job_ms += ms;
        // ---- bit errors per stream (untimed; the message bits are regenerated on the device) ----
        for (size_t k = 0; k < nb; k++) {
            unsigned long long e = 0;
            int rc = vit_count_errors_synth_device(c.options, job_out_ptr(j, b0 + k, k), j->M, c.seed + (unsigned)(j->first + b0 + k),
                                                   c.source, &e, j->stream);
            if (rc) return mg_fail(rc, "error count failed for stream %zu", j->first + b0 + k);
            j->errs[b0 + k] = e;
            res->bit_errors += e;
            if (e > res->max_stream_errors) res->max_stream_errors = e;
        }
        // a sender's output slots are reused by the next round: its gathers have completed (ev_c), and the root's
        // buffer is complete once every rank has passed the next barrier
    }
    if (j->comm) { int rc = vit_comm_barrier(j->comm); if (rc) return rc; }
    res->streams = (unsigned)j->count;
    res->decoded_bits = (unsigned long long)j->count * j->M;
    return VIT_OK;
}

#pragma GCC visibility pop
}  // extern "C"
