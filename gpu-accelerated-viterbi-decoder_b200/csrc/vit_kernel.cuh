// vit_kernel.cuh -- fused, persistent-per-segment Viterbi decode kernel for sm_100a (B200).
//
// Replaces the reference's viterbi_core kernel and its three device headers
// (reference src/viterbi/viterbi.cu:144-207, viterbiBM.cuh, viterbiACS.cuh, viterbiTB.cuh) with a
// different mapping of the same algorithm (K=7, 64 states, polys 0171/0133 by default, 6400 stream segments,
// 64-stage warm-up, register-exchange survivor words, 96-stage ring, traceback from state 0):
//
//   * 8 lanes per stream segment, 4 segments per warp (the reference: 32 lanes per segment).
//     Each lane owns 8 trellis states: 4 packed int16x2 / half2 registers, or 8 int32 registers.
//   * The map (lane, register, half) -> trellis state is GF(2)-linear and changes with the stage (period 6) in
//     such a way that EVERY butterfly joins two registers, or the two halves of one register, of the same
//     lane.  Two stages of six need no data movement at all; before each of the other four every lane
//     swaps half of its registers (2 metric + 4 survivor words) with lane ^ 1, 2, 4, 7 ("half exchange"):
//     24 shuffles per 6 stages, where an in-place map that exchanges whole register sets on the three
//     lane-bit stages needs 36.  See "trellis state <-> (lane, register, half) map" below.
//   * A 96-stage super-step (lcm of the 6 phases and the 32-stage slide) is three straight-line
//     slides, each a 5-iteration loop over the 6-stage phase period plus a 2-stage tail with the
//     ring flush/traceback: shuffle masks, operand choices and row offsets are immediates, there is
//     one loop branch per 6 stages, and the code (~25 KB) stays in the instruction cache.
//   * Branch metrics: the channel words of the NEXT 96-stage super-step arrive by 16-byte cp.async
//     into a raw staging area while the current super-step computes; they are unpacked into a
//     shared-memory table of ready-to-add packed operands (one 8-byte entry per stage x lane-class;
//     class 3-c holds the negated operands of class c), so each ACS stage costs two LDS.64 per lane.
//     The table covers 96 stages (TBL=96, built once per super-step) or 32 (TBL=32, rebuilt before every
//     slide: a third of the shared memory, twice the resident warps -- used by multi-stream launches).
//   * Survivors: 32-bit register-exchange words moved with predicated selects (VIMNMX.S16x2 yields
//     both decision predicates; the int32 core derives the decision from a fused VIADDMNMX).  The
//     decision bits themselves are never shifted in one by one: the last 6 message bits of a survivor
//     are its state index, so every 6 stages each word is shifted left by 6 and its register's index field is added
//     (one IMAD per word, the same code in every loop iteration).
//   * The one-pointer ring (3 x 64 words per segment) lives in shared memory, not global.
//   * Optional upload gates let one launch start before its input has arrived (vit_run's time-sliced copy-in).
//
// The same source compiles for the host (VIT_HOST_EMU) where 32 fibers run the warp in lockstep;
// tests/emu uses that to check the kernel logic against the oracle without a GPU.  The host build
// is test scaffolding only and is never linked into the product library.
#pragma once

#include <stddef.h>
#include <stdint.h>
#include <utility>

#include "vit_code.h"

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define VIT_HD __host__ __device__ __forceinline__
#define VIT_D __device__ __forceinline__
#else
#define VIT_HD inline
#define VIT_D inline
#endif

namespace vitk {

enum { IN_HARD = 0, IN_S4 = 1, IN_S8 = 2, IN_S16 = 3, IN_F32 = 4 };
enum { MET_B32 = 0, MET_B16 = 1, MET_F16 = 2,
       // int32 core with the tie rule of the reference's never-instantiated DPX variant (viterbiACS.cuh:123-134: the partner
       // wins ties in every phase; the REG variant lets the odd predecessor win at phase 0).  Same instructions as MET_B32.
       MET_B32D = 3 };
constexpr bool is_b32(int met) { return met == MET_B32 || met == MET_B32D; }

struct KParams {
    const uint8_t* in;              // stream 0 channel words
    uint8_t* out;                   // stream 0 decoded packs
    unsigned long long in_stride;   // bytes between consecutive streams
    unsigned long long out_stride;
    unsigned long long in_bytes;    // valid channel bytes per stream (reads beyond are zero-filled)
    unsigned long long packs;       // decoded packs per stream (messageLen / bitsPerPack)
    unsigned segments;              // stream segments (reference: 6400, viterbi.cu:19)
    unsigned seg_first, seg_limit;  // this launch decodes segments [seg_first, seg_limit) (chunked host pipeline)
    unsigned nstreams;
    unsigned one;                   // == 1, opaque to the compiler: `x*one + y` is emitted as IMAD so that
                                    // the metric adds run on the FMA pipe while VIMNMX/SEL own the ALU pipe
    // Upload gates (vit_run with host buffers, time-sliced copy-in): the channel words arrive while the kernel
    // runs, every segment's super-steps [gate_super[g], gate_super[g+1]) with copy slice g.  A warp may read
    // super-step x only after gate[g] == gate_epoch for every g with gate_super[g] <= x.  gate_n == 0: no gating.
    const unsigned* gate;
    unsigned* gate_err;             // set to 1 by a warp that gave up waiting (upload never arrived)
    unsigned gate_epoch;
    unsigned gate_n;
    unsigned gate_super[8];
    unsigned long long gate_timeout_ns;   // a warp gives up on a gate after this long (vit_api.cu derives it from the copy size)
    unsigned stage_out;                   // != 0: store decoded packs 8 slides (32 bytes per segment) at a time (output in a peer GPU's memory)
};

// code parameters of this build (vit_code.h): K = 7 is fixed, the generator polynomials default to the reference's
constexpr int CONST_LEN = VIT_CONST_LEN;
constexpr int POLY1 = VIT_POLY1, POLY2 = VIT_POLY2;

// ------------------------------------------------------------------------------------------------
// compile-time trellis geometry
// ------------------------------------------------------------------------------------------------
constexpr int mod6(int x) { return ((x % 6) + 6) % 6; }
constexpr int SUPER = 96;       // unroll period
constexpr int par3(int v) { return ((v >> 0) ^ (v >> 1) ^ (v >> 2) ^ (v >> 3)) & 1; }   // parity of up to 4 bits
constexpr int par6(int v) { return (v ^ (v >> 1) ^ (v >> 2) ^ (v >> 3) ^ (v >> 4) ^ (v >> 5)) & 1; }   // ... of up to 6 (compile-time uses)

template <int IN> struct InTraits;
template <> struct InTraits<IN_HARD> { static constexpr int B96 = 24; };
template <> struct InTraits<IN_S4> { static constexpr int B96 = 96; };
template <> struct InTraits<IN_S8> { static constexpr int B96 = 192; };
template <> struct InTraits<IN_S16> { static constexpr int B96 = 384; };
template <> struct InTraits<IN_F32> { static constexpr int B96 = 768; };
template <int IN> constexpr int raw_pieces() { return (InTraits<IN>::B96 + 12 + 15) / 16; }

// offset added to int16 branch metrics so every packed operand is a non-negative 16-bit value
// (lets plain 32-bit adds act as two independent 16-bit adds); == max |symbol sum| / 1
template <int IN> constexpr int b16_offset() { return IN == IN_HARD ? 1 : IN == IN_S8 ? 256 : 16; }
// the same for the half2 core (s8/s16 symbols are pre-scaled to 5 bits there): operands and metrics stay non-negative,
// so their IEEE bit patterns order like integers and VIMNMX.S16x2 can do the compare (see Core<MET_F16>)
template <int IN> constexpr int f16_offset() { return IN == IN_HARD ? 1 : (IN == IN_S8 || IN == IN_S16) ? 32 : 16; }

// ------------------------------------------------------------------------------------------------
// SIMT primitives: device intrinsics, or the host emulator's lockstep fibers
// ------------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
VIT_D uint32_t shfl_xor(uint32_t v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
VIT_D uint32_t shfl_idx(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
VIT_D void syncwarp() { __syncwarp(); }
VIT_D void cp_async16(void* smem_dst, const void* gsrc, unsigned src_bytes) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
VIT_D void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> VIT_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
VIT_D uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
VIT_D uint32_t brev32(uint32_t v) { return __brev(v); }
// wait until *flag == epoch (written by the host's copy stream after the slice's bytes); gives up after timeout_ns.
// The verdict is warp-uniform (every lane polls, the warp votes): no lane may leave while others go on to shuffles.
VIT_D bool gate_spin(const unsigned* flag, unsigned epoch, unsigned long long timeout_ns) {
    unsigned long long t0 = 0;
    for (unsigned it = 1;; it++) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (__all_sync(0xffffffffu, v == epoch)) return true;
        __nanosleep(it < 8 ? 200 : 1000);
        bool expired = false;
        if ((it & 63u) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else expired = now - t0 > timeout_ns;
        }
        if (__any_sync(0xffffffffu, expired)) return false;
    }
}
#else
uint32_t emu_shfl_xor(uint32_t v, int m);
uint32_t emu_shfl_idx(uint32_t v, int src);
void emu_syncwarp();
inline uint32_t shfl_xor(uint32_t v, int m) { return emu_shfl_xor(v, m); }
inline uint32_t shfl_idx(uint32_t v, int src) { return emu_shfl_idx(v, src); }
inline void syncwarp() { emu_syncwarp(); }
void emu_note_global_read(const void* src, unsigned n);       // the emulator records which input bytes a lane requests
inline void cp_async16(void* dst, const void* src, unsigned n) {
    uint8_t* d = (uint8_t*)dst;
    const uint8_t* s = (const uint8_t*)src;
    if (n) emu_note_global_read(src, n);
    for (unsigned i = 0; i < 16; i++) d[i] = i < n ? s[i] : 0;
}
inline void cp_async_commit() {}
template <int N> inline void cp_async_wait() {}
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t sel = (s >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
inline uint32_t brev32(uint32_t v) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((v >> i) & 1u) << (31 - i);
    return r;
}
inline bool gate_spin(const unsigned* flag, unsigned epoch, unsigned long long) { return *flag == epoch; }
#endif

// ------------------------------------------------------------------------------------------------
// metric cores: candidate = metric +/- operand, max with "partner chosen" predicates
// ------------------------------------------------------------------------------------------------
#if !defined(__CUDA_ARCH__)
// host model of IEEE half for small exact integers (|v| < 2048): we only ever hold integers, so a
// float pair stands in for half2 on the host; the device uses real __half2 arithmetic.
struct EmuH2 { float lo, hi; };
inline uint32_t emu_h2_pack(EmuH2 v) {
    // store as two int16 two's-complement integers (host emulation of exact-integer halves)
    return ((uint32_t)(uint16_t)(int16_t)v.lo) | ((uint32_t)(uint16_t)(int16_t)v.hi << 16);
}
inline EmuH2 emu_h2_unpack(uint32_t w) { return EmuH2{(float)(int16_t)(w & 0xffff), (float)(int16_t)(w >> 16)}; }
#endif

template <int MET, int IN> struct Core;

// ---- int16x2: non-negative 15-bit metrics, operands carry a +offset --------------------------
template <int IN> struct Core<MET_B16, IN> {
    static constexpr bool PACKED = true;
    static constexpr uint32_t K2 = (uint32_t)(2 * b16_offset<IN>()) * 0x10001u;
    // all metric adds are `x*one + y` = IMAD on the FMA pipe; the ALU pipe is left to VIMNMX/SEL
    static VIT_HD uint32_t plus(uint32_t pm, uint32_t w, uint32_t one) { return pm * one + w; }
    // Add-compare-select of one packed register = two states.  part/own are the two candidates; the
    // partner wins ties (reference int16 core, viterbiACS.cuh:112-119,215-220: __vibmax_s16x2(partner - bm,
    // own + bm) -> pred = (a >= b)).  PTX max.s16x2 + setp.eq on the halves is what __vibmax_s16x2 expands to and
    // ptxas fuses it into ONE VIMNMX.S16x2 Rd, P0, P1.  The survivor words follow the decision either through
    // SEL (ALU pipe, acs_sel) or through a predicated IMAD "move" (FMA pipe, acs_mov: `x*one + 0` with the opaque
    // multiplier is not folded back into SEL), so the selects can be split between the two pipes.
    static VIT_HD uint32_t acs_sel(uint32_t part, uint32_t own, uint32_t keep_lo, uint32_t src_lo, uint32_t keep_hi,
                                   uint32_t src_hi, uint32_t& out_lo, uint32_t& out_hi, bool /*own_wins_tie*/) {
#if defined(__CUDA_ARCH__)
        uint32_t v;
        asm("{.reg .pred pu, pv; \n\t"
            ".reg .s16 rs0, rs1, rs2, rs3; \n\t"
            "max.s16x2 %0, %3, %4; \n\t"
            "mov.b32 {rs0, rs1}, %0; \n\t"
            "mov.b32 {rs2, rs3}, %3; \n\t"
            "setp.eq.s16 pv, rs0, rs2; \n\t"
            "setp.eq.s16 pu, rs1, rs3; \n\t"
            "selp.b32 %1, %6, %5, pv; \n\t"
            "selp.b32 %2, %8, %7, pu;} \n\t"
            : "=r"(v), "=r"(out_lo), "=r"(out_hi)
            : "r"(part), "r"(own), "r"(keep_lo), "r"(src_lo), "r"(keep_hi), "r"(src_hi));
        return v;
#else
        bool pl, ph;
        uint32_t v = emu_max(part, own, pl, ph);
        out_lo = pl ? src_lo : keep_lo; out_hi = ph ? src_hi : keep_hi;
        return v;
#endif
    }
    static VIT_HD uint32_t acs_mov(uint32_t part, uint32_t own, uint32_t& io_lo, uint32_t src_lo, uint32_t& io_hi,
                                   uint32_t src_hi, uint32_t one, bool /*own_wins_tie*/) {
#if defined(__CUDA_ARCH__)
        uint32_t v;
        asm("{.reg .pred pu, pv; \n\t"
            ".reg .s16 rs0, rs1, rs2, rs3; \n\t"
            "max.s16x2 %0, %3, %4; \n\t"
            "mov.b32 {rs0, rs1}, %0; \n\t"
            "mov.b32 {rs2, rs3}, %3; \n\t"
            "setp.eq.s16 pv, rs0, rs2; \n\t"
            "setp.eq.s16 pu, rs1, rs3; \n\t"
            "@pv mad.lo.u32 %1, %5, %7, 0; \n\t"
            "@pu mad.lo.u32 %2, %6, %7, 0;} \n\t"
            : "=r"(v), "+r"(io_lo), "+r"(io_hi)
            : "r"(part), "r"(own), "r"(src_lo), "r"(src_hi), "r"(one));
        return v;
#else
        bool pl, ph;
        uint32_t v = emu_max(part, own, pl, ph);
        if (pl) io_lo = src_lo;
        if (ph) io_hi = src_hi;
        (void)one;
        return v;
#endif
    }
#if !defined(__CUDA_ARCH__)
    static uint32_t emu_max(uint32_t part, uint32_t own, bool& p_lo, bool& p_hi) {
        int16_t al = (int16_t)(part & 0xffff), ah = (int16_t)(part >> 16);
        int16_t bl = (int16_t)(own & 0xffff), bh = (int16_t)(own >> 16);
        p_lo = al >= bl; p_hi = ah >= bh;
        return (uint32_t)(uint16_t)(p_lo ? al : bl) | ((uint32_t)(uint16_t)(p_hi ? ah : bh) << 16);
    }
#endif
    static VIT_HD uint32_t vmin(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
        return __vmins2(a, b);
#else
        int16_t al = (int16_t)(a & 0xffff), ah = (int16_t)(a >> 16), bl = (int16_t)(b & 0xffff), bh = (int16_t)(b >> 16);
        return (uint32_t)(uint16_t)(al < bl ? al : bl) | ((uint32_t)(uint16_t)(ah < bh ? ah : bh) << 16);
#endif
    }
    static VIT_HD uint32_t sub(uint32_t a, uint32_t m) { return a - m; }   // per-half, no borrow: a >= m
    static VIT_HD uint32_t enc(int v) { return (uint32_t)(v + b16_offset<IN>()) & 0xffffu; }
    static VIT_HD uint32_t zero() { return 0; }
};

// ---- half2: exact small integers on the FP16 pipe ---------------------------------------------
template <int IN> struct Core<MET_F16, IN> {
    static constexpr bool PACKED = true;
#if defined(__CUDA_ARCH__)
    static VIT_D __half2 h2(uint32_t w) { return *reinterpret_cast<__half2*>(&w); }
    static VIT_D uint32_t u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
    static VIT_D uint32_t plus(uint32_t pm, uint32_t w, uint32_t) { return u32(__hadd2(h2(pm), h2(w))); }
    // Metric adds are HADD2 on the FP16 pipe; values are exact integers in [0, 2048).  Operands carry a +offset, so
    // every candidate is a non-negative half and compares like its bit pattern: the compare/select is the same single
    // VIMNMX.S16x2 Rd, P0, P1 as in the int16x2 core (HMNMX2 + HSETP2 would be two instructions).
    // own wins ties (reference half2 core, viterbiACS.cuh:146-157,249-256: __hlt2_mask(own, partner)): the maximum is
    // taken with the own candidate first and "own chosen" = (max == own).
    static VIT_D uint32_t acs_sel(uint32_t part, uint32_t own, uint32_t keep_lo, uint32_t src_lo, uint32_t keep_hi,
                                  uint32_t src_hi, uint32_t& out_lo, uint32_t& out_hi, bool) {
        uint32_t v;
        asm("{.reg .pred pl, ph; \n\t"
            ".reg .s16 rs0, rs1, rs2, rs3; \n\t"
            "max.s16x2 %0, %4, %3; \n\t"
            "mov.b32 {rs0, rs1}, %0; \n\t"
            "mov.b32 {rs2, rs3}, %4; \n\t"
            "setp.eq.s16 pl, rs0, rs2; \n\t"
            "setp.eq.s16 ph, rs1, rs3; \n\t"
            "selp.b32 %1, %5, %6, pl; \n\t"
            "selp.b32 %2, %7, %8, ph;} \n\t"
            : "=r"(v), "=r"(out_lo), "=r"(out_hi)
            : "r"(part), "r"(own), "r"(keep_lo), "r"(src_lo), "r"(keep_hi), "r"(src_hi));
        return v;
    }
    static VIT_D uint32_t acs_mov(uint32_t part, uint32_t own, uint32_t& io_lo, uint32_t src_lo, uint32_t& io_hi,
                                  uint32_t src_hi, uint32_t one, bool) {
        uint32_t v;
        asm("{.reg .pred pl, ph; \n\t"
            ".reg .s16 rs0, rs1, rs2, rs3; \n\t"
            "max.s16x2 %0, %4, %3; \n\t"
            "mov.b32 {rs0, rs1}, %0; \n\t"
            "mov.b32 {rs2, rs3}, %4; \n\t"
            "setp.eq.s16 pl, rs0, rs2; \n\t"
            "setp.eq.s16 ph, rs1, rs3; \n\t"
            "@!pl mad.lo.u32 %1, %5, %7, 0; \n\t"
            "@!ph mad.lo.u32 %2, %6, %7, 0;} \n\t"
            : "=r"(v), "+r"(io_lo), "+r"(io_hi)
            : "r"(part), "r"(own), "r"(src_lo), "r"(src_hi), "r"(one));
        return v;
    }
    static VIT_D uint32_t vmin(uint32_t a, uint32_t b) { return u32(__hmin2(h2(a), h2(b))); }
    static VIT_D uint32_t sub(uint32_t a, uint32_t m) { return u32(__hsub2(h2(a), h2(m))); }
    static VIT_D uint32_t enc(int v) { return (uint32_t)__half_as_ushort(__int2half_rn(v + f16_offset<IN>())); }
#else
    static uint32_t plus(uint32_t pm, uint32_t w, uint32_t) { EmuH2 a = emu_h2_unpack(pm), b = emu_h2_unpack(w); return emu_h2_pack(EmuH2{a.lo + b.lo, a.hi + b.hi}); }
    static uint32_t neg(uint32_t w, uint32_t) { EmuH2 b = emu_h2_unpack(w); return emu_h2_pack(EmuH2{-b.lo, -b.hi}); }
    static uint32_t emu_max(uint32_t part, uint32_t own, bool& p_lo, bool& p_hi) {
        EmuH2 a = emu_h2_unpack(part), b = emu_h2_unpack(own);
        p_lo = a.lo > b.lo; p_hi = a.hi > b.hi;
        return emu_h2_pack(EmuH2{p_lo ? a.lo : b.lo, p_hi ? a.hi : b.hi});
    }
    static uint32_t acs_sel(uint32_t part, uint32_t own, uint32_t keep_lo, uint32_t src_lo, uint32_t keep_hi,
                            uint32_t src_hi, uint32_t& out_lo, uint32_t& out_hi, bool) {
        bool pl, ph;
        uint32_t v = emu_max(part, own, pl, ph);
        out_lo = pl ? src_lo : keep_lo; out_hi = ph ? src_hi : keep_hi;
        return v;
    }
    static uint32_t acs_mov(uint32_t part, uint32_t own, uint32_t& io_lo, uint32_t src_lo, uint32_t& io_hi,
                            uint32_t src_hi, uint32_t, bool) {
        bool pl, ph;
        uint32_t v = emu_max(part, own, pl, ph);
        if (pl) io_lo = src_lo;
        if (ph) io_hi = src_hi;
        return v;
    }
    static uint32_t vmin(uint32_t a, uint32_t b) { EmuH2 x = emu_h2_unpack(a), y = emu_h2_unpack(b); return emu_h2_pack(EmuH2{x.lo < y.lo ? x.lo : y.lo, x.hi < y.hi ? x.hi : y.hi}); }
    static uint32_t sub(uint32_t a, uint32_t m) { return plus(a, neg(m, 1), 1); }
    static uint32_t enc(int v) { return (uint32_t)(uint16_t)(int16_t)(v + f16_offset<IN>()); }
#endif
    static VIT_HD uint32_t zero() { return 0; }
};

// ---- int32: one state per register --------------------------------------------------------------
template <int IN> struct Core<MET_B32, IN> {
    static constexpr bool PACKED = false;
    static VIT_HD uint32_t plus(uint32_t pm, uint32_t w, uint32_t one) { return pm * one + w; }      // IMAD
    // One state per register.  partner wins ties except where the reference's phase-0 rule makes the odd
    // predecessor win (viterbiACS.cuh:136-142: both selfPM compares are "odd-candidate >= even-candidate").
    // The candidate that loses ties is never materialised: max(loser_metric + loser_operand, winner_candidate)
    // is one fused VIADDMNMX, and the winner is chosen iff the result equals its candidate (ISETP.EQ) -- 4
    // instructions per state (IMAD, VIADDMNMX, ISETP, select) instead of 5 (2 IMAD, ISETP, 2 SEL).
    // acs1_sel writes the survivor to a new register (SEL, ALU pipe); acs1_mov moves it in place with a
    // predicated IMAD (FMA pipe).
    template <bool OWN_WINS>
    static VIT_HD uint32_t acs1_sel(uint32_t pm_part, uint32_t op_part, uint32_t pm_own, uint32_t op_own, uint32_t keep,
                                    uint32_t src, uint32_t& out, uint32_t one) {
        uint32_t v;
#if defined(__CUDA_ARCH__)
        if (OWN_WINS)
            asm("{.reg .pred p; .reg .s32 w, t; \n\t"
                "mad.lo.s32 w, %2, %6, %3; \n\t"
                "add.s32 t, %4, %5; \n\t"
                "max.s32 %0, t, w; \n\t"
                "setp.ne.s32 p, %0, w; \n\t"
                "selp.b32 %1, %8, %7, p;} \n\t"
                : "=r"(v), "=r"(out) : "r"(pm_own), "r"(op_own), "r"(pm_part), "r"(op_part), "r"(one), "r"(keep), "r"(src));
        else
            asm("{.reg .pred p; .reg .s32 w, t; \n\t"
                "mad.lo.s32 w, %4, %6, %5; \n\t"
                "add.s32 t, %2, %3; \n\t"
                "max.s32 %0, t, w; \n\t"
                "setp.eq.s32 p, %0, w; \n\t"
                "selp.b32 %1, %8, %7, p;} \n\t"
                : "=r"(v), "=r"(out) : "r"(pm_own), "r"(op_own), "r"(pm_part), "r"(op_part), "r"(one), "r"(keep), "r"(src));
#else
        const int a = (int)(pm_part + op_part), b = (int)(pm_own + op_own);
        const bool p = OWN_WINS ? (a > b) : (a >= b);
        out = p ? src : keep;
        v = (uint32_t)(p ? a : b);
        (void)one;
#endif
        return v;
    }
    template <bool OWN_WINS>
    static VIT_HD uint32_t acs1_mov(uint32_t pm_part, uint32_t op_part, uint32_t pm_own, uint32_t op_own, uint32_t& io,
                                    uint32_t src, uint32_t one) {
        uint32_t v;
#if defined(__CUDA_ARCH__)
        if (OWN_WINS)
            asm("{.reg .pred p; .reg .s32 w, t; \n\t"
                "mad.lo.s32 w, %2, %6, %3; \n\t"
                "add.s32 t, %4, %5; \n\t"
                "max.s32 %0, t, w; \n\t"
                "setp.ne.s32 p, %0, w; \n\t"
                "@p mad.lo.u32 %1, %7, %6, 0;} \n\t"
                : "=r"(v), "+r"(io) : "r"(pm_own), "r"(op_own), "r"(pm_part), "r"(op_part), "r"(one), "r"(src));
        else
            asm("{.reg .pred p; .reg .s32 w, t; \n\t"
                "mad.lo.s32 w, %4, %6, %5; \n\t"
                "add.s32 t, %2, %3; \n\t"
                "max.s32 %0, t, w; \n\t"
                "setp.eq.s32 p, %0, w; \n\t"
                "@p mad.lo.u32 %1, %7, %6, 0;} \n\t"
                : "=r"(v), "+r"(io) : "r"(pm_own), "r"(op_own), "r"(pm_part), "r"(op_part), "r"(one), "r"(src));
#else
        const int a = (int)(pm_part + op_part), b = (int)(pm_own + op_own);
        const bool p = OWN_WINS ? (a > b) : (a >= b);
        if (p) io = src;
        v = (uint32_t)(p ? a : b);
        (void)one;
#endif
        return v;
    }
    static VIT_HD uint32_t vmin(uint32_t a, uint32_t b) { return (uint32_t)((int)a < (int)b ? (int)a : (int)b); }
    static VIT_HD uint32_t sub(uint32_t a, uint32_t m) { return a - m; }
    static VIT_HD uint32_t enc(int v) { return (uint32_t)v; }
    static VIT_HD uint32_t zero() { return 0; }
};
template <int IN> struct Core<MET_B32D, IN> : Core<MET_B32, IN> {};

// candidate of `metric` along the branch whose operand type is `type` (see bm_type), given the
// table entry (w0 = X word, w1 = Y word); `opposite` = the other branch into the same state
struct Operands { uint32_t x, y, nx, ny, one; };   // +X, +Y, -X, -Y in the core's operand encoding


template <int TYPE, bool OPPOSITE>
VIT_HD uint32_t opnd(const Operands& o) {
    constexpr bool useY = (TYPE == 1 || TYPE == 2);
    constexpr bool neg = ((TYPE == 2 || TYPE == 3) != OPPOSITE);
    return neg ? (useY ? o.ny : o.nx) : (useY ? o.y : o.x);
}
template <class C, int TYPE, bool OPPOSITE>
VIT_HD uint32_t cand(uint32_t metric, const Operands& o) {
    constexpr bool useY = (TYPE == 1 || TYPE == 2);
    constexpr bool neg = ((TYPE == 2 || TYPE == 3) != OPPOSITE);
    return C::plus(metric, neg ? (useY ? o.ny : o.nx) : (useY ? o.y : o.x), o.one);
}

// pp | (lf ^ SF), or lf ^ SF when ASSIGN: one LOP3 with an immediate.  Written as PTX so that the loop-invariant
// (lf ^ SF) is not hoisted into a register per (register, batch) pair -- that costs ~90 registers.
template <bool ASSIGN, uint32_t SF>
VIT_HD uint32_t or_xor(uint32_t pp, uint32_t lf) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    if (ASSIGN) asm("lop3.b32 %0, %1, %2, %3, 0x66;" : "=r"(r) : "r"(pp), "r"(lf), "n"(SF));
    else asm("lop3.b32 %0, %1, %2, %3, 0xF6;" : "=r"(r) : "r"(pp), "r"(lf), "n"(SF));
    return r;
#else
    return ASSIGN ? (lf ^ SF) : (pp | (lf ^ SF));
#endif
}

// ------------------------------------------------------------------------------------------------
// branch-metric table build: one lane unpacks one stage and writes the four lane-class entries {W0,W1} for it
// (HV = how the high half of a packed operand differs from the low half at that stage's phase, see half_variant)
// ------------------------------------------------------------------------------------------------
// the raw channel word(s) of one stage; loaded separately from their decoding so that a builder can issue all its loads
// before its first table store (the compiler will not move a shared-memory load above a store it cannot disambiguate)
struct RawSym { uint32_t w; float f0, f1; };
template <int IN>
VIT_HD RawSym load_raw(const uint8_t* raw, int rel_stage) {
    // raw points at the byte holding the first stage of this superchunk (word aligned)
    RawSym r{0u, 0.f, 0.f};
    if constexpr (IN == IN_HARD) r.w = *reinterpret_cast<const uint32_t*>(raw + 4 * (rel_stage >> 4));
    else if constexpr (IN == IN_S4) r.w = *reinterpret_cast<const uint32_t*>(raw + 4 * (rel_stage >> 2));
    else if constexpr (IN == IN_S8) r.w = *reinterpret_cast<const uint32_t*>(raw + 4 * (rel_stage >> 1));
    else if constexpr (IN == IN_S16) r.w = *reinterpret_cast<const uint32_t*>(raw + 4 * rel_stage);
    else {
        const float* p = reinterpret_cast<const float*>(raw + 8 * rel_stage);
        r.f0 = p[0]; r.f1 = p[1];
    }
    return r;
}
template <int MET, int IN>
VIT_HD void decode_symbols(const RawSym& r, int rel_stage, int& d0, int& d1, float& f0, float& f1) {
    d0 = d1 = 0; f0 = f1 = 0.f;
    const uint32_t w = r.w;
    if constexpr (IN == IN_HARD) {
        // 32 symbols per int32, MSB first (reference viterbiBM.cuh:33-40)
        uint32_t rx = (w >> (30 - 2 * (rel_stage & 15))) & 3u;
        d0 = 2 * (int)(rx >> 1) - 1; d1 = 2 * (int)(rx & 1) - 1;   // +-1, halved after the sum
    } else if constexpr (IN == IN_S4) {
        int sh = 24 - 8 * (rel_stage & 3);                          // viterbiBM.cuh:64-75
        d0 = ((int)(w << (24 - sh))) >> 28;
        d1 = ((int)(w << (28 - sh))) >> 28;
    } else if constexpr (IN == IN_S8) {
        int sh = 16 - 16 * (rel_stage & 1);                         // viterbiBM.cuh:97-100
        d0 = (int)(int8_t)(w >> (sh + 8));
        d1 = (int)(int8_t)(w >> sh);
        if constexpr (MET == MET_F16) { d0 >>= 3; d1 >>= 3; }       // extension: keep half2 exact
    } else if constexpr (IN == IN_S16) {
        d0 = (int)(int16_t)(w >> 16);                               // viterbiBM.cuh:121-124
        d1 = (int)(int16_t)(w & 0xffff);
        if constexpr (MET == MET_F16) { d0 >>= 11; d1 >>= 11; }
    } else {
        f0 = fminf(fmaxf(r.f0, -8.0f), 7.0f);                       // viterbiBM.cuh:146-153
        f1 = fminf(fmaxf(r.f1, -8.0f), 7.0f);
    }
}

template <int MET, int IN, int HV>
VIT_HD void build_step(const RawSym& rs, int rel_stage, uint32_t* entry /* 8 words: class-major */) {
    using C = Core<MET, IN>;
    int d0, d1; float f0, f1;
    decode_symbols<MET, IN>(rs, rel_stage, d0, d1, f0, f1);
    int A, B;
    if constexpr (IN == IN_F32) {
        A = (int)(f0 + f1); B = (int)(f0 - f1);                     // truncation is odd-symmetric
    } else if constexpr (IN == IN_HARD) {
        A = (d0 + d1) >> 1; B = (d0 - d1) >> 1;
    } else {
        A = d0 + d1; B = d0 - d1;
    }
    // class (c0,c1): D0 = c0 ? d0 : -d0, D1 = c1 ? d1 : -d1;  X = D0+D1, Y = D0-D1
    //   cl 3: ( A,  B)   cl 2: ( B,  A)   cl 1: (-B, -A)   cl 0: (-A, -B)
    const int X[4] = {-A, -B, B, A};
    const int Y[4] = {-B, -A, A, B};
#pragma unroll
    for (int cl = 0; cl < 4; cl++) {
        uint32_t w0, w1;
        if constexpr (C::PACKED) {
            // high half: (0) same, (3) negated, (1) X<->Y, (2) X->-Y, Y->-X
            int xh = HV == 0 ? X[cl] : HV == 3 ? -X[cl] : HV == 1 ? Y[cl] : -Y[cl];
            int yh = HV == 0 ? Y[cl] : HV == 3 ? -Y[cl] : HV == 1 ? X[cl] : -X[cl];
            w0 = C::enc(X[cl]) | (C::enc(xh) << 16);
            w1 = C::enc(Y[cl]) | (C::enc(yh) << 16);
        } else {
            w0 = C::enc(X[cl]); w1 = C::enc(Y[cl]);
        }
        entry[2 * cl] = w0; entry[2 * cl + 1] = w1;
    }
}


// kp.gate_super[i] without dynamic indexing of the kernel parameters (that would copy them to local memory)
VIT_HD unsigned gate_super_at(const KParams& kp, unsigned i) {
    unsigned v = kp.gate_super[0];
#pragma unroll
    for (unsigned k = 1; k < 8; k++) v = (i == k) ? kp.gate_super[k] : v;
    return v;
}

// ------------------------------------------------------------------------------------------------
// The decoder proper (state map, shared-memory layout, stage code, warp body, kernel) is written once in
// vit_kernel_map.inc for a lane geometry given by VIT_NLB (lane bits per segment):
//   vitk::l8  8 lanes per segment, 4 segments per warp, 8 states per lane  -- the product kernels;
//   vitk::l4  4 lanes per segment, 8 segments per warp, 16 states per lane -- 3 half exchanges per 6 stages instead of 4,
//             10 % fewer instructions per decoded bit, but half as many warps;
//   vitk::l16 16 lanes per segment, 2 segments per warp, 4 states per lane -- twice as many warps, 5 half exchanges per 6
//             stages, 17 % more instructions per decoded bit.
// Both alternatives were measured slower than l8, so the library does not instantiate their kernels.  The
// host emulator runs all three (tests/test_emu_kernel.py): the map algebra is the same code.
// ------------------------------------------------------------------------------------------------
#define VIT_NLB 3
namespace l8 {
#include "vit_kernel_map.inc"
}  // namespace l8
#undef VIT_NLB
#define VIT_NLB 2
namespace l4 {
#include "vit_kernel_map.inc"
}  // namespace l4
#undef VIT_NLB
#define VIT_NLB 4
namespace l16 {
#include "vit_kernel_map.inc"
}  // namespace l16
#undef VIT_NLB
// vitk::l1  ONE lane per segment, 32 segments per warp, 64 states per lane: no exchanges and no operand table at all, ~30 %
//           fewer instructions per decoded bit, but only 200 warps per stream -- the geometry for launches of many streams
//           (vit_kernel_l1.inc; the map algebra with no lane bits is the same code).
#define VIT_NLB 0
namespace l1 {
#include "vit_kernel_map.inc"
#include "vit_kernel_l1.inc"
}  // namespace l1
#undef VIT_NLB

}  // namespace vitk

