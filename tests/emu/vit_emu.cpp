// vit_emu.cpp -- host SIMT emulator for the product kernel source (TEST SCAFFOLDING ONLY).
//
// Compiles gpu-accelerated-viterbi-decoder_b200/csrc/vit_kernel.cuh with a plain host compiler and
// runs each warp as 32 ucontext fibers in lockstep (every shuffle / syncwarp is a round-robin yield),
// so the kernel's trellis mapping, operand tables, survivor-field insertion and traceback can be
// checked against the oracle in the GPU-less authoring container.  It is never part of the product
// library: the product path has no CPU fallback.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ucontext.h>

#include "../../gpu-accelerated-viterbi-decoder_b200/csrc/vit_kernel.cuh"

namespace {
constexpr int NL = 32;
ucontext_t g_ctx[NL], g_main;
int g_cur = 0;
uint32_t g_slot[2][NL];
unsigned g_shfl_k[NL];
char* g_stacks = nullptr;
constexpr size_t STACK = 512 * 1024;

void yield_next() {
    int me = g_cur;
    int nx = (me + 1) % NL;
    g_cur = nx;
    swapcontext(&g_ctx[me], &g_ctx[nx]);
}
}  // namespace

namespace vitk {
uint32_t emu_shfl_xor(uint32_t v, int m) {
    int me = g_cur;
    unsigned b = g_shfl_k[me]++ & 1;
    g_slot[b][me] = v;
    yield_next();
    return g_slot[b][me ^ m];
}
uint32_t emu_shfl_idx(uint32_t v, int src) {
    int me = g_cur;
    unsigned b = g_shfl_k[me]++ & 1;
    g_slot[b][me] = v;
    yield_next();
    return g_slot[b][src & 31];
}
void emu_syncwarp() { yield_next(); }
}  // namespace vitk

// ---- upload gates and the record of requested input bytes (tests/test_emu_kernel.py: the kernel side of vit_run's time-sliced
// upload: a warp must not request channel words of a super-step whose gate is still closed) ----------------------------------
namespace {
unsigned g_gate_flags[8] = {};
unsigned g_gate_err = 0;
unsigned g_gate_n = 0, g_gate_super[8] = {};
const uint8_t* g_in_base = nullptr;
unsigned g_cur_warp = 0, g_cur_lanes = 8;
struct ReadRec { unsigned seg; unsigned long long off; unsigned n; };
ReadRec* g_reads = nullptr;
size_t g_nreads = 0, g_reads_cap = 0;
}  // namespace
namespace vitk {
void emu_note_global_read(const void* src, unsigned n) {
    if (!g_reads || g_nreads >= g_reads_cap) return;
    const unsigned seg = g_cur_warp * (32 / g_cur_lanes) + (unsigned)g_cur / g_cur_lanes;
    g_reads[g_nreads++] = ReadRec{seg, (unsigned long long)((const uint8_t*)src - g_in_base), n};
}
}  // namespace vitk
extern "C" void vit_emu_set_gates(unsigned n, const unsigned* super, unsigned open_count) {
    g_gate_n = n > 8 ? 8 : n;
    for (unsigned i = 0; i < 8; i++) { g_gate_super[i] = i < g_gate_n ? super[i] : 0; g_gate_flags[i] = i < open_count ? 1u : 0u; }
    g_gate_err = 0;
}
extern "C" unsigned vit_emu_gate_error() { return g_gate_err; }
// start (cap_records > 0) or stop (0) recording the input reads of the following decodes; fetch returns them as
// (segment, byte offset in the stream, bytes) triples of uint64
extern "C" void vit_emu_record_reads(size_t cap_records) {
    static ReadRec* store = nullptr;
    static size_t store_cap = 0;
    g_reads = nullptr; g_nreads = 0; g_reads_cap = 0;
    if (!cap_records) return;
    if (store_cap < cap_records) { free(store); store = (ReadRec*)malloc(sizeof(ReadRec) * cap_records); store_cap = cap_records; }
    g_reads = store; g_reads_cap = cap_records;
}
extern "C" size_t vit_emu_fetch_reads(unsigned long long* buf, size_t cap_records) {
    const size_t n = g_nreads < cap_records ? g_nreads : cap_records;
    for (size_t i = 0; i < n; i++) { buf[3 * i] = g_reads[i].seg; buf[3 * i + 1] = g_reads[i].off; buf[3 * i + 2] = g_reads[i].n; }
    return g_nreads;
}

namespace {
struct Job {
    vitk::KParams kp;
    unsigned warp, stream;
    uint8_t* smem;
    int met, in, bpp;
    int tbl = 96;      // operand-table variant of the kernel: 96 (per super-step) or 32 (per slide)
    int lanes = 8;     // lane geometry: 8 lanes per segment (4 segments per warp) or 4 (8 segments per warp)
};
Job g_job;

template <int MET, int IN, int BPP>
void run_lane(int lane) {
#if !defined(VIT_EMU_L8_ONLY) || defined(VIT_EMU_WITH_L1)
    if (g_job.lanes == 1) {
        vitk::l1::warp_body_l1<MET, IN, BPP>(g_job.kp, g_job.warp, g_job.stream, lane, g_job.smem);
    } else
#endif
#if !defined(VIT_EMU_L8_ONLY)   // (builds for other code parameters instantiate the product geometry only: a third of the compile time)
    if (g_job.lanes == 16) {
        if (g_job.tbl == 32) vitk::l16::warp_body<MET, IN, BPP, 32>(g_job.kp, g_job.warp, g_job.stream, lane, g_job.smem);
        else vitk::l16::warp_body<MET, IN, BPP, 96>(g_job.kp, g_job.warp, g_job.stream, lane, g_job.smem);
    } else if (g_job.lanes == 4) {
        if (g_job.tbl == 32) vitk::l4::warp_body<MET, IN, BPP, 32>(g_job.kp, g_job.warp, g_job.stream, lane, g_job.smem);
        else vitk::l4::warp_body<MET, IN, BPP, 96>(g_job.kp, g_job.warp, g_job.stream, lane, g_job.smem);
    } else
#endif
    {
        if (g_job.tbl == 32) vitk::l8::warp_body<MET, IN, BPP, 32>(g_job.kp, g_job.warp, g_job.stream, lane, g_job.smem);
        else vitk::l8::warp_body<MET, IN, BPP, 96>(g_job.kp, g_job.warp, g_job.stream, lane, g_job.smem);
    }
}

template <int MET, int IN>
void run_lane_bpp(int lane) {
    if (g_job.bpp == 16) run_lane<MET, IN, 16>(lane); else run_lane<MET, IN, 32>(lane);
}
template <int MET>
void run_lane_in(int lane) {
    switch (g_job.in) {
        case 0: run_lane_bpp<MET, 0>(lane); break;
        case 1: run_lane_bpp<MET, 1>(lane); break;
        case 2: run_lane_bpp<MET, 2>(lane); break;
        case 3: if constexpr (MET != vitk::MET_B16) run_lane_bpp<MET, 3>(lane); break;
        default: run_lane_bpp<MET, 4>(lane); break;
    }
}
void fiber_main(int lane) {
    switch (g_job.met) {
        case vitk::MET_B32: run_lane_in<vitk::MET_B32>(lane); break;
        case vitk::MET_B32D: run_lane_in<vitk::MET_B32D>(lane); break;
        case vitk::MET_B16: run_lane_in<vitk::MET_B16>(lane); break;
        default: run_lane_in<vitk::MET_F16>(lane); break;
    }
    g_cur = (lane + 1) % NL;   // falls through uc_link to the next lane (or main after lane 31)
}

void run_warp() {
    if (!g_stacks) g_stacks = (char*)malloc(STACK * NL);
    for (int i = 0; i < NL; i++) {
        getcontext(&g_ctx[i]);
        g_ctx[i].uc_stack.ss_sp = g_stacks + STACK * i;
        g_ctx[i].uc_stack.ss_size = STACK;
        g_ctx[i].uc_link = (i == NL - 1) ? &g_main : &g_ctx[i + 1];
        makecontext(&g_ctx[i], (void (*)())fiber_main, 1, i);
        g_shfl_k[i] = 0;
    }
    g_cur = 0;
    swapcontext(&g_main, &g_ctx[0]);
}
}  // namespace

// the generator polynomials this emulator build was compiled for (-DVIT_POLY1= / -DVIT_POLY2=, csrc/vit_code.h)
extern "C" void vit_emu_polynomials(int* p1, int* p2) { *p1 = vitk::POLY1; *p2 = vitk::POLY2; }

static void run_grid();
static unsigned g_stage_out = 0;
static unsigned g_seg_first = 0, g_seg_limit = 0;   // 0, 0: the whole stream
// decode only segments [first, limit) in the following calls (the segment-range launches of vit_run's chunk pipeline)
extern "C" void vit_emu_set_segment_range(unsigned first, unsigned limit) { g_seg_first = first; g_seg_limit = limit; }
extern "C" void vit_emu_set_stage_out(int on) { g_stage_out = on ? 1u : 0u; }
extern "C" void vit_emu_set_table(int tbl) { g_job.tbl = tbl == 32 ? 32 : 96; }
#if defined(VIT_EMU_L8_ONLY) && defined(VIT_EMU_WITH_L1)
extern "C" void vit_emu_set_lanes(int lanes) { g_job.lanes = lanes == 1 ? 1 : 8; }
#elif defined(VIT_EMU_L8_ONLY)
extern "C" void vit_emu_set_lanes(int) { g_job.lanes = 8; }
#else
extern "C" void vit_emu_set_lanes(int lanes) { g_job.lanes = lanes == 4 ? 4 : lanes == 16 ? 16 : lanes == 1 ? 1 : 8; }
#endif

extern "C" int vit_emu_decode(int options, const void* in, void* out, size_t inputNum, unsigned segments,
                              unsigned nstreams, size_t in_stride, size_t out_stride) {
    int it = options & 0xf, mt = (options >> 4) & 0xf, bpp = ((options >> 8) & 0xf) ? 16 : 32;
    if (it > 4 || mt > 2) return -1;
    if (mt == 1 && it == 3) return -1;
    size_t in_bytes = it == 0 ? (inputNum + 7) / 8 : it == 1 ? (inputNum + 1) / 2 : it == 2 ? inputNum : it == 3 ? inputNum * 2 : inputNum * 4;
    if (inputNum / 2 < 64) return 0;
    size_t M = (inputNum / 2 - 64) / bpp * bpp;
    g_job.kp.in = (const uint8_t*)in; g_job.kp.out = (uint8_t*)out;
    g_job.kp.in_stride = in_stride; g_job.kp.out_stride = out_stride;
    g_job.kp.in_bytes = in_bytes; g_job.kp.packs = M / bpp;
    g_job.kp.gate = g_gate_flags; g_job.kp.gate_err = &g_gate_err; g_job.kp.gate_epoch = 1u; g_job.kp.gate_n = g_gate_n; g_job.kp.gate_timeout_ns = 0;
    for (int i = 0; i < 8; i++) g_job.kp.gate_super[i] = g_gate_super[i];
    g_in_base = (const uint8_t*)in; g_cur_lanes = (unsigned)g_job.lanes;
    g_job.kp.segments = segments; g_job.kp.seg_first = g_seg_limit ? g_seg_first : 0; g_job.kp.seg_limit = g_seg_limit ? g_seg_limit : segments; g_job.kp.nstreams = nstreams; g_job.kp.one = 1u; g_job.kp.stage_out = g_stage_out;
    g_job.met = mt == 0 ? (((options >> 12) & 0xf) == 2 ? vitk::MET_B32D : vitk::MET_B32) : mt == 1 ? vitk::MET_B16 : vitk::MET_F16;
    g_job.in = it; g_job.bpp = bpp;
    run_grid();
    return 0;
}

// every warp of the launch g_job describes (grid = segment groups x streams), one after the other
static void run_grid() {
    g_job.smem = (uint8_t*)aligned_alloc(128, 64 * 1024);
    const unsigned spw = 32 / g_job.lanes;
    const unsigned nwarps = (g_job.kp.seg_limit - g_job.kp.seg_first + spw - 1) / spw;
    for (unsigned s = 0; s < g_job.kp.nstreams; s++)
        for (unsigned w = 0; w < nwarps; w++) {
            g_job.warp = w; g_job.stream = s; g_cur_warp = w;
            memset(g_job.smem, 0xA5, 64 * 1024);   // catch reads of unwritten shared memory
            run_warp();
        }
    free(g_job.smem);
}

// a launch exactly as the library's host code describes it (tests/sim: csrc/vit_api.cu compiled for the host hands its
// KParams to the emulator instead of a GPU); product lane geometry, operand-table build `tbl`
void vit_emu_run_kparams(const vitk::KParams& kp, int met, int in, int bpp, int tbl, int lanes) {
    g_job.kp = kp;
    g_job.met = met; g_job.in = in; g_job.bpp = bpp;
    g_job.lanes = lanes == 1 ? 1 : 8; g_job.tbl = tbl == 32 ? 32 : 96;
    g_in_base = kp.in; g_cur_lanes = (unsigned)g_job.lanes;
    run_grid();
}
