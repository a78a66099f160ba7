// vit_kernel.cuh -- fused, persistent-per-segment Viterbi decode kernel for sm_100a (B200).
//
// Replaces the reference's viterbi_core kernel and its three device headers
// (reference src/viterbi/viterbi.cu:144-207, viterbiBM.cuh, viterbiACS.cuh, viterbiTB.cuh) with a
// different mapping of the same algorithm (K=7, 64 states, polys 0171/0133, 6400 stream segments,
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
//     are its state index, so they are merged in as a 6-bit field (one LOP3) every 6 stages.
//   * The one-pointer ring (3 x 64 words per segment) lives in shared memory, not global.
//   * Optional upload gates let one launch start before its input has arrived (vit_run's time-sliced copy-in).
//
// The same source compiles for the host (VIT_HOST_EMU) where 32 fibers run the warp in lockstep;
// tests/emu uses that to check the kernel logic against the oracle without a GPU.  The host build
// is test scaffolding only and is never linked into the product library.
#pragma once

#include <stddef.h>
#include <stdint.h>

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
enum { MET_B32 = 0, MET_B16 = 1, MET_F16 = 2 };

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
};

// ------------------------------------------------------------------------------------------------
// compile-time trellis geometry
// ------------------------------------------------------------------------------------------------
constexpr int mod6(int x) { return ((x % 6) + 6) % 6; }
constexpr int LANES_PER_SEG = 8;
constexpr int SEGS_PER_WARP = 4;
constexpr int SUPER = 96;       // unroll period

// ------------------------------------------------------------------------------------------------
// trellis state <-> (lane, register, half) map
//
// A position is a 6-bit vector v = [l0 l1 l2 | r0 r1 r2]: l = lane within the 8-lane group, r0 r1 = packed
// register index, r2 = half of the packed register (int32 core: third register-index bit).  At every
// stage the map position -> state is GF(2)-linear: state bit i = parity(v & f_i).  A trellis stage pairs
// the two positions that differ only in state bit 0 (the butterfly) and writes the two new states back
// in place, the new MSB u taking the value the position's old state bit 0 had, so the six functionals
// just rotate (f_i <- f_{i+1}, f_5 <- f_0).  The map is chosen so that every butterfly joins two
// registers (or the two halves of one register) of the SAME lane:
//
//   phase 0: butterfly along r1            (no data movement)
//   phase 1: butterfly along r2            (half swap for the packed cores)
//   phase 2..5: butterfly along r0, each preceded by a "half exchange": every lane sends its r0=1 registers
//               (2 metric + 4 survivor registers) to lane ^ m, m = 1, 2, 4, 7, and keeps its r0=0 registers.
//
// The half exchange is the transvection v -> v + r0(v)*m of the position space: it substitutes
// l -> l + r0*m in every functional, so that the functional about to become state bit 0 has a dual vector
// without lane component.  The masks sum to zero (1^2^4^7), so the map has period 6 and everything below
// is a compile-time function of the phase.  24 shuffles per 6 stages instead of the 36 of a map that
// exchanges all 12 registers on the three lane-bit stages.
// ------------------------------------------------------------------------------------------------
constexpr int PB_R0 = 8, PB_R1 = 16, PB_R2 = 32;
constexpr int par6(int v) { return ((v >> 0) ^ (v >> 1) ^ (v >> 2) ^ (v >> 3) ^ (v >> 4) ^ (v >> 5)) & 1; }
constexpr int xmask_of(int p) { return p == 2 ? 1 : p == 3 ? 2 : p == 4 ? 4 : p == 5 ? 7 : 0; }   // exchange before stage p
constexpr int cum_of(int p) { return p == 2 ? 1 : p == 3 ? 3 : p == 4 ? 7 : 0; }                   // sum of the masks so far
constexpr int lprime(int j, int c) { return (1 << j) | (((c >> j) & 1) ? PB_R0 : 0); }              // l_j + r0*c_j
// slot q = the functional whose butterfly is done at phase q, under accumulated exchange offset c
constexpr int func_of_slot(int q, int c) {
    return q == 0 ? PB_R1 : q == 1 ? PB_R2 : q == 2 ? (lprime(0, c) ^ lprime(1, c)) : q == 3 ? (lprime(1, c) ^ lprime(2, c))
         : q == 4 ? lprime(2, c) : (PB_R0 ^ lprime(0, c));
}
// ... and its dual position vector (flips that state bit only)
constexpr int dual_of_slot(int q, int c) {
    return q == 0 ? PB_R1 : q == 1 ? PB_R2 : q == 2 ? ((1 ^ c) | PB_R0) : q == 3 ? ((3 ^ c) | PB_R0) : q == 4 ? ((7 ^ c) | PB_R0) : (c | PB_R0);
}
constexpr int f_before(int p, int i) { return func_of_slot(mod6(p + i), cum_of(p)); }      // state bit i entering stage p
constexpr int f_after(int p, int i) { return func_of_slot(mod6(p + 1 + i), cum_of(p)); }   // state bit i leaving stage p

// Own-branch symbol of a position entering stage p.  The encoder register is (u<<6)|S with u = S bit 0,
// and both polynomials tap bits 0 and 6, so symbol bit 0 (poly 0171) = S3^S4^S5, symbol bit 1 (poly 0133) = S1^S3^S4.
constexpr int sym_func(int which, int p) {
    return which == 0 ? (f_before(p, 3) ^ f_before(p, 4) ^ f_before(p, 5)) : (f_before(p, 1) ^ f_before(p, 3) ^ f_before(p, 4));
}
constexpr int reg_mask(int which, int p) { return (sym_func(which, p) >> 3) & 7; }
constexpr int par3(int v) { return ((v >> 0) ^ (v >> 1) ^ (v >> 2)) & 1; }

// operand type of the own-branch metric for register index s8 at phase p:
//   0: +X   1: +Y   2: -Y   3: -X      (X = D0+D1, Y = D0-D1 of the lane-class adjusted symbols)
constexpr int bm_type(int s8, int p) {
    int st0 = par3(s8 & reg_mask(0, p)), st1 = par3(s8 & reg_mask(1, p));
    return st0 == 0 ? (st1 == 0 ? 0 : 1) : (st1 == 0 ? 2 : 3);
}
// how the high half of a packed operand differs from the low half (does r2 enter symbol bit 0 / 1)
constexpr int half_variant(int p) { return ((sym_func(0, p) & PB_R2) ? 2 : 0) | ((sym_func(1, p) & PB_R2) ? 1 : 0); }

// stage kinds
enum { KIND_HALF = 0, KIND_REG = 2 };
template <int MET>
constexpr int stage_kind(int p) { return (p == 1 && MET != MET_B32) ? KIND_HALF : KIND_REG; }
constexpr int stage_bit(int p) { return p == 0 ? 1 : p == 1 ? 2 : 0; }  // register-index bit of the butterfly

// 6-bit survivor field (oldest message bit = state bit 0 in the MSB) of register s8 leaving stage p; the
// lane's part (lane_field_of) is XOR-ed in at run time
constexpr int static_field(int s8, int p) {
    int f = 0;
    for (int i = 0; i < 6; i++) if (par3(s8 & (f_after(p, i) >> 3))) f |= 1 << (5 - i);
    return f;
}

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

// shared memory carve-up (per warp; one warp per block)
// Operand table: row (stage) = 4 segments x 4 lane classes x 8 B = 128 B; rows are grouped by
// j = stage/6 and each j-block is padded by 16 B so that the 8 lanes of a group, which build rows
// 6 apart, store to 8 different 16-byte bank columns (conflict-free STS.128).
constexpr int BM_ROW = SEGS_PER_WARP * 4 * 8;            // 128
constexpr int BM_JBLOCK = 6 * BM_ROW + 16;               // 784
// stage within the super-step -> table row.  TBL = stages the table holds: 96 (built once per super-step) or 32
// (rebuilt before every slide: 1/3 of the shared memory, twice the resident warps, for multi-stream launches)
template <int TBL>
constexpr int row_off(int s) { return ((s % TBL) / 6) * BM_JBLOCK + ((s % TBL) % 6) * BM_ROW; }
// Ring: per (slot, segment) 64 words + 16 B pad; lane l stores its 8 words at l*32 + (l>>2)*16 so that
// the two STS.128 of a flush are conflict-free as well.
constexpr int RING_SEG = 64 * 4 + 16;                    // 272
constexpr int ring_word_of(int l, int r) { return l * 8 + (l >> 2) * 4 + r; }

template <int IN, int TBL = 96> struct Smem {
    static constexpr int BM_BYTES = ((TBL + 5) / 6) * BM_JBLOCK;                   // 12544 | 4704
    static constexpr int RAW_SEG = raw_pieces<IN>() * 16;
    static constexpr int RAW_BYTES = SEGS_PER_WARP * RAW_SEG;      // single buffer: consumed whole by the table build
    static constexpr int RING_BYTES = 3 * SEGS_PER_WARP * RING_SEG;                // 3264
    static constexpr int LUT_BYTES = 3 * 64;
    static constexpr int OFF_BM = 0;
    static constexpr int OFF_RAW = OFF_BM + BM_BYTES;
    static constexpr int OFF_RING = OFF_RAW + RAW_BYTES;
    static constexpr int OFF_LUT = OFF_RING + RING_BYTES;
    static constexpr int TOTAL = OFF_LUT + LUT_BYTES;
};

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
// wait until *flag == epoch (written by the host's copy stream after the slice's bytes); gives up after ~2 s
VIT_D bool gate_spin(const unsigned* flag, unsigned epoch) {
    unsigned long long t0 = 0;
    for (unsigned it = 1;; it++) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == epoch) return true;
        __nanosleep(it < 8 ? 200 : 1000);
        if ((it & 255u) == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 2000000000ull) return false;
        }
    }
}
#else
uint32_t emu_shfl_xor(uint32_t v, int m);
uint32_t emu_shfl_idx(uint32_t v, int src);
void emu_syncwarp();
inline uint32_t shfl_xor(uint32_t v, int m) { return emu_shfl_xor(v, m); }
inline uint32_t shfl_idx(uint32_t v, int src) { return emu_shfl_idx(v, src); }
inline void syncwarp() { emu_syncwarp(); }
inline void cp_async16(void* dst, const void* src, unsigned n) {
    uint8_t* d = (uint8_t*)dst;
    const uint8_t* s = (const uint8_t*)src;
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
inline bool gate_spin(const unsigned* flag, unsigned epoch) { return *flag == epoch; }
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

// ------------------------------------------------------------------------------------------------
// per-lane decoder state
// ------------------------------------------------------------------------------------------------
template <int MET> struct LaneState {
    static constexpr int NPM = (MET == MET_B32) ? 8 : 4;
    uint32_t pm[NPM];
    uint32_t pp[8];
};

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

// one trellis stage at phase P (static).  lane bits are the low 3 bits of the lane id.
// Survivor selects alternate between the SEL form (ALU pipe) and the predicated-IMAD form (FMA pipe).
template <int MET, int IN, int P>
VIT_HD void acs_stage(LaneState<MET>& s, const Operands& ops) {
    using C = Core<MET, IN>;
    constexpr int KIND = stage_kind<MET>(P);
    constexpr int BIT = stage_bit(P);
    const uint32_t one = ops.one;
    if constexpr (KIND == KIND_HALF) {
        // packed cores, phase 0: the two predecessors are the two halves of the same register
#define VIT_HALF(r)                                                                               \
    {                                                                                             \
        uint32_t sw = prmt(s.pm[r], 0, 0x1032);                                                   \
        uint32_t oc = cand<C, bm_type(r, P), false>(s.pm[r], ops);                                \
        uint32_t pc = cand<C, bm_type(r, P), true>(sw, ops);                                      \
        const uint32_t a = s.pp[r], b = s.pp[r + 4];                                              \
        s.pm[r] = C::acs_sel(pc, oc, a, b, b, a, s.pp[r], s.pp[r + 4], false);                    \
    }
        VIT_HALF(0) VIT_HALF(1) VIT_HALF(2) VIT_HALF(3)
#undef VIT_HALF
    } else {
        // register-local butterfly on register-index bit BIT: the even register's survivors use the SEL
        // form (new registers), the odd register's the in-place predicated form
        if constexpr (C::PACKED) {
#define VIT_REG_PACKED(ra)                                                                        \
    if constexpr (((ra >> BIT) & 1) == 0) {                                                       \
        constexpr int rb = ra | (1 << BIT);                                                       \
        const uint32_t ea = s.pm[ra], eb = s.pm[rb];                                              \
        const uint32_t pa0 = s.pp[ra], pb0 = s.pp[rb], pa1 = s.pp[ra + 4], pb1 = s.pp[rb + 4];    \
        s.pm[ra] = C::acs_sel(cand<C, bm_type(ra, P), true>(eb, ops), cand<C, bm_type(ra, P), false>(ea, ops), \
                              pa0, pb0, pa1, pb1, s.pp[ra], s.pp[ra + 4], false);                 \
        s.pm[rb] = C::acs_mov(cand<C, bm_type(rb, P), true>(ea, ops), cand<C, bm_type(rb, P), false>(eb, ops), \
                              s.pp[rb], pa0, s.pp[rb + 4], pa1, one, false);                      \
    }
            VIT_REG_PACKED(0) VIT_REG_PACKED(1) VIT_REG_PACKED(2) VIT_REG_PACKED(3)
#undef VIT_REG_PACKED
        } else {
            // reference int32 core, phase 0: the odd predecessor wins ties for both new states
            // (viterbiACS.cuh:136-142); all other phases: partner wins ties (viterbiACS.cuh:238-245)
            constexpr bool ODD_WINS = (P == 0);
#define VIT_REG_B32(ra)                                                                           \
    if constexpr (((ra >> BIT) & 1) == 0) {                                                       \
        constexpr int rb = ra | (1 << BIT);                                                       \
        const uint32_t ea = s.pm[ra], eb = s.pm[rb];                                              \
        const uint32_t pa0 = s.pp[ra], pb0 = s.pp[rb];                                            \
        s.pm[ra] = C::template acs1_sel<false>(eb, opnd<bm_type(ra, P), true>(ops), ea, opnd<bm_type(ra, P), false>(ops), \
                                               pa0, pb0, s.pp[ra], one);                          \
        s.pm[rb] = C::template acs1_mov<ODD_WINS>(ea, opnd<bm_type(rb, P), true>(ops), eb, opnd<bm_type(rb, P), false>(ops), \
                                                  s.pp[rb], pa0, one);                            \
    }
            VIT_REG_B32(0) VIT_REG_B32(1) VIT_REG_B32(2) VIT_REG_B32(3)
            VIT_REG_B32(4) VIT_REG_B32(5) VIT_REG_B32(6) VIT_REG_B32(7)
#undef VIT_REG_B32
        }
    }
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

// OR the survivors' newest message bits (== state index bits) into the register-exchange words.
// NB bits (6 or 2) taken from the top of the 6-bit field, placed at bit SHIFT.
template <int MET, int P, int SHIFT, int NB, bool ASSIGN>
VIT_HD void insert_field(LaneState<MET>& s, uint32_t lane_field) {
    const uint32_t lf = (NB == 6 ? lane_field : (lane_field >> 4)) << SHIFT;
#define VIT_INS(k)                                                                                \
    {                                                                                             \
        constexpr uint32_t sf = (uint32_t)(NB == 6 ? static_field(k, P) : (static_field(k, P) >> 4)) << SHIFT; \
        s.pp[k] = or_xor<ASSIGN, sf>(s.pp[k], lf);                                                \
    }
    VIT_INS(0) VIT_INS(1) VIT_INS(2) VIT_INS(3) VIT_INS(4) VIT_INS(5) VIT_INS(6) VIT_INS(7)
#undef VIT_INS
}

// position index (lane*8 + reg) of the position that holds state `st` after the stage of phase p
VIT_HD int pos_index_of_state(int st, int p) {
    int v = 0;
    for (int i = 0; i < 6; i++)
        if ((st >> i) & 1) v ^= dual_of_slot(mod6(p + 1 + i), cum_of(p));
    return (v & 7) * 8 + (v >> 3);
}

// the lane's part of the survivor field after the stage of phase p (XOR-ed with static_field)
VIT_HD uint32_t lane_field_of(int l, int p) {
    uint32_t f = 0;
    for (int i = 0; i < 6; i++)
        if (par3(l & f_after(p, i) & 7)) f |= 1u << (5 - i);
    return f;
}
// the lane's part of the own-branch symbol entering stage p: operand class = 2*c0 + c1
VIT_HD int lane_class_of(int l, int p) {
    return par3(l & sym_func(0, p) & 7) * 2 + par3(l & sym_func(1, p) & 7);
}

// Half exchange before the stages of phases 2..5: the r0=1 registers (metrics and survivors) swap with lane ^ XM.
template <int MET, int XM>
VIT_HD void exchange_half(LaneState<MET>& s) {
#pragma unroll
    for (int r = 1; r < LaneState<MET>::NPM; r += 2) s.pm[r] = shfl_xor(s.pm[r], XM);
#pragma unroll
    for (int k = 1; k < 8; k += 2) s.pp[k] = shfl_xor(s.pp[k], XM);
}

// ------------------------------------------------------------------------------------------------
// branch-metric table build: lane j of the group unpacks stage base+P+6j (P static) and writes the
// four lane-class entries {W0,W1} for it.
// ------------------------------------------------------------------------------------------------
template <int MET, int IN>
VIT_HD void load_symbols(const uint8_t* raw, int rel_stage, int& d0, int& d1, float& f0, float& f1) {
    // raw points at the byte holding the first stage of this superchunk (word aligned)
    d0 = d1 = 0; f0 = f1 = 0.f;
    if constexpr (IN == IN_HARD) {
        // 32 symbols per int32, MSB first (reference viterbiBM.cuh:33-40)
        uint32_t w = *reinterpret_cast<const uint32_t*>(raw + 4 * (rel_stage >> 4));
        uint32_t rx = (w >> (30 - 2 * (rel_stage & 15))) & 3u;
        d0 = 2 * (int)(rx >> 1) - 1; d1 = 2 * (int)(rx & 1) - 1;   // +-1, halved after the sum
    } else if constexpr (IN == IN_S4) {
        uint32_t w = *reinterpret_cast<const uint32_t*>(raw + 4 * (rel_stage >> 2));   // viterbiBM.cuh:64-75
        int sh = 24 - 8 * (rel_stage & 3);
        d0 = ((int)(w << (24 - sh))) >> 28;
        d1 = ((int)(w << (28 - sh))) >> 28;
    } else if constexpr (IN == IN_S8) {
        uint32_t w = *reinterpret_cast<const uint32_t*>(raw + 4 * (rel_stage >> 1));   // viterbiBM.cuh:97-100
        int sh = 16 - 16 * (rel_stage & 1);
        d0 = (int)(int8_t)(w >> (sh + 8));
        d1 = (int)(int8_t)(w >> sh);
        if constexpr (MET == MET_F16) { d0 >>= 3; d1 >>= 3; }       // extension: keep half2 exact
    } else if constexpr (IN == IN_S16) {
        uint32_t w = *reinterpret_cast<const uint32_t*>(raw + 4 * rel_stage);          // viterbiBM.cuh:121-124
        d0 = (int)(int16_t)(w >> 16);
        d1 = (int)(int16_t)(w & 0xffff);
        if constexpr (MET == MET_F16) { d0 >>= 11; d1 >>= 11; }
    } else {
        const float* p = reinterpret_cast<const float*>(raw + 8 * rel_stage);          // viterbiBM.cuh:146-153
        f0 = fminf(fmaxf(p[0], -8.0f), 7.0f);
        f1 = fminf(fmaxf(p[1], -8.0f), 7.0f);
    }
}

template <int MET, int IN, int P>
VIT_HD void build_step(const uint8_t* raw, int rel_stage, uint32_t* entry /* 8 words: class-major */) {
    using C = Core<MET, IN>;
    int d0, d1; float f0, f1;
    load_symbols<MET, IN>(raw, rel_stage, d0, d1, f0, f1);
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
    constexpr int HV = half_variant(P);
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

// ------------------------------------------------------------------------------------------------
// the warp body
// ------------------------------------------------------------------------------------------------
template <int MET, int IN, int BPP, int TBL>
struct WarpCtx {
    LaneState<MET> st;
    uint8_t* smem;
    int lane, g, l;
    uint32_t one;
    uint32_t bm_off[6];        // byte offset of this lane's class entry within a stage row, per phase
    uint32_t bm_offn[6];       // ... and of the complementary class (3 - class): the same operands negated
    uint32_t lane_field[3];    // phases 1,3,5
    // segment geometry
    unsigned long long seg_byte0;   // byte offset of the segment's first stage in the stream
    const uint8_t* in;
    uint8_t* out;
    unsigned long long in_bytes;
    unsigned long long out_word0;   // first decoded pack of this segment
    unsigned seg_bits;              // L
    unsigned t0;                    // absolute stage index of the current superchunk's stage 0
};

template <int MET, int IN, int BPP, int TBL>
VIT_HD void issue_raw_copy(WarpCtx<MET, IN, BPP, TBL>& c, unsigned super_idx) {
    using S = Smem<IN, TBL>;
    constexpr int B96 = InTraits<IN>::B96;
    unsigned long long b0 = c.seg_byte0 + (unsigned long long)super_idx * B96;
    unsigned long long al = b0 & ~15ull;
    uint8_t* dst = c.smem + S::OFF_RAW + c.g * S::RAW_SEG;
#pragma unroll
    for (int i = 0; i < (raw_pieces<IN>() + 7) / 8; i++) {
        int piece = c.l + 8 * i;
        if (piece < raw_pieces<IN>()) {
            unsigned long long off = al + 16ull * piece;
            unsigned valid = off >= c.in_bytes ? 0u : (c.in_bytes - off >= 16 ? 16u : (unsigned)(c.in_bytes - off));
            const uint8_t* src = c.in + (valid ? off : 0);
            cp_async16(dst + 16 * piece, src, valid);
        }
    }
    cp_async_commit();
}

VIT_HD void store_row(uint32_t* row, const uint32_t (&e)[8]) {
#if defined(__CUDA_ARCH__)
    reinterpret_cast<uint4*>(row)[0] = make_uint4(e[0], e[1], e[2], e[3]);
    reinterpret_cast<uint4*>(row)[1] = make_uint4(e[4], e[5], e[6], e[7]);
#else
    for (int i = 0; i < 8; i++) row[i] = e[i];
#endif
}

// lane l of the group builds the rows of stages P + 6*(l + 8*round): j-block l + 8*round, row P
template <int MET, int IN, int BPP, int TBL, int P>
VIT_HD void build_phase(WarpCtx<MET, IN, BPP, TBL>& c, int round, unsigned skew) {
    using S = Smem<IN, TBL>;
    const uint8_t* raw = c.smem + S::OFF_RAW + c.g * S::RAW_SEG + skew;
    const int j = c.l + 8 * round;
    uint32_t e[8];
    build_step<MET, IN, P>(raw, P + 6 * j, e);
    uint32_t* row = reinterpret_cast<uint32_t*>(c.smem + S::OFF_BM + j * BM_JBLOCK + P * BM_ROW + c.g * 32);
    store_row(row, e);
}

template <int MET, int IN, int BPP, int TBL>
VIT_HD void build_table(WarpCtx<MET, IN, BPP, TBL>& c, unsigned skew) {
#pragma unroll 1
    for (int round = 0; round < 2; round++) {
        build_phase<MET, IN, BPP, TBL, 0>(c, round, skew);
        build_phase<MET, IN, BPP, TBL, 1>(c, round, skew);
        build_phase<MET, IN, BPP, TBL, 2>(c, round, skew);
        build_phase<MET, IN, BPP, TBL, 3>(c, round, skew);
        build_phase<MET, IN, BPP, TBL, 4>(c, round, skew);
        build_phase<MET, IN, BPP, TBL, 5>(c, round, skew);
    }
}

// TBL == 32: the rows of slide W (stages 32W .. 32W+31) are built just before it runs.  Lane l builds stage
// 32W + 6l + r in step r (phase (32W + r) % 6 is static per step), i.e. j-block l, row r; lanes 6,7 idle.
template <int MET, int IN, int BPP, int TBL, int W>
VIT_HD void build_slide(WarpCtx<MET, IN, BPP, TBL>& c, unsigned skew) {
    using S = Smem<IN, TBL>;
    const uint8_t* raw = c.smem + S::OFF_RAW + c.g * S::RAW_SEG + skew;
    int l = c.l;
#if defined(__CUDA_ARCH__)
    // opaque copy of the lane index: the per-row shift amounts and offsets derived from it are loop invariant, and
    // hoisting all 18 sets of them out of the super-step loop costs ~50 registers (occupancy is the point of TBL=32)
    asm volatile("" : "+r"(l));
#endif
    uint8_t* blk = c.smem + S::OFF_BM + l * BM_JBLOCK + c.g * 32;
#define VIT_BUILD_ROW(r)                                                                          \
    if (6 * l + r < 32) {                                                                         \
        uint32_t e[8];                                                                            \
        build_step<MET, IN, (32 * W + r) % 6>(raw, 32 * W + 6 * l + r, e);                        \
        uint32_t* row = reinterpret_cast<uint32_t*>(blk + r * BM_ROW);                            \
        store_row(row, e);                                                                        \
    }
    VIT_BUILD_ROW(0) VIT_BUILD_ROW(1) VIT_BUILD_ROW(2) VIT_BUILD_ROW(3) VIT_BUILD_ROW(4) VIT_BUILD_ROW(5)
#undef VIT_BUILD_ROW
    syncwarp();
}

// subtract the segment-wide minimum metric (a common offset never changes a decision; the
// reference does the same on a threshold, viterbiACS.cuh:307-378)
template <int MET, int IN>
VIT_HD void normalize(LaneState<MET>& s) {
    using C = Core<MET, IN>;
    uint32_t m = s.pm[0];
#pragma unroll
    for (int r = 1; r < LaneState<MET>::NPM; r++) m = C::vmin(m, s.pm[r]);
    if constexpr (C::PACKED) m = C::vmin(m, prmt(m, 0, 0x1032));
    m = C::vmin(m, shfl_xor(m, 1));
    m = C::vmin(m, shfl_xor(m, 2));
    m = C::vmin(m, shfl_xor(m, 4));
#pragma unroll
    for (int r = 0; r < LaneState<MET>::NPM; r++) s.pm[r] = C::sub(s.pm[r], m);
}

// stages per normalization: 96 (super-step start), 32 (slide end) or 16 (slide end and after its third loop iteration).
// int16x2: offsets grow a metric by <= 2*offset per stage and it must stay below 2^15; half2: below 2048 (exact integers):
// hard 2*96 + 12, s4/fp32 32*32 + 192, s8/s16 (5-bit symbols, offset 32) 64*18 + 384.
template <int MET, int IN> constexpr int norm_period() {
    return (MET == MET_B16) ? (IN == IN_S8 ? 32 : 96)
         : (MET == MET_F16) ? (IN == IN_HARD ? 96 : (IN == IN_S8 || IN == IN_S16) ? 16 : 32)
         : 96;
}

// End of a 32-stage slide (superchunk stage S = 32*SLOT + 31, phase P = S % 6): flush the
// register-exchange words to ring slot SLOT, trace back from state 0, emit 32 decoded bits, start
// the next survivor word.
template <int MET, int IN, int BPP, int TBL, int SLOT>
VIT_HD void slide_end(WarpCtx<MET, IN, BPP, TBL>& c) {
    using SM = Smem<IN, TBL>;
    constexpr int S = 32 * SLOT + 31;
    constexpr int P = S % 6;                      // 1, 3, 5 for slots 0, 1, 2
    uint8_t* ringb = c.smem + SM::OFF_RING;
    uint32_t* mine = reinterpret_cast<uint32_t*>(ringb + (SLOT * SEGS_PER_WARP + c.g) * RING_SEG) + ring_word_of(c.l, 0);
#if defined(__CUDA_ARCH__)
    reinterpret_cast<uint4*>(mine)[0] = make_uint4(c.st.pp[0], c.st.pp[1], c.st.pp[2], c.st.pp[3]);
    reinterpret_cast<uint4*>(mine)[1] = make_uint4(c.st.pp[4], c.st.pp[5], c.st.pp[6], c.st.pp[7]);
#else
    for (int k = 0; k < 8; k++) mine[k] = c.st.pp[k];
#endif
    // state 0 sits at position 0 in every phase: lane 0 of the group, register 0
    uint32_t w_e = shfl_idx(c.st.pp[0], c.g * 8);
    syncwarp();
    const unsigned e = c.t0 + S;
    if (e >= 95) {
        unsigned st1 = brev32(w_e) & 63u;                                   // reference viterbiTB.cuh:9-12
        const uint8_t* lut = c.smem + SM::OFF_LUT + SLOT * 64;             // word at e-32 has phase (P+4)%6
        unsigned idx = lut[st1];
        const uint32_t* prev = reinterpret_cast<const uint32_t*>(ringb + (((SLOT + 2) % 3) * SEGS_PER_WARP + c.g) * RING_SEG);
        uint32_t word = prev[idx];                                          // viterbiTB.cuh:14-19
        unsigned k = (e - 95) / 32;                                         // slide index
        if (c.l == 0) {
            if constexpr (BPP == 32) {
                if (k * 32 < c.seg_bits) reinterpret_cast<uint32_t*>(c.out)[c.out_word0 + k] = word;
            } else {
                uint16_t* o = reinterpret_cast<uint16_t*>(c.out) + c.out_word0 + 2ull * k;
                if (k * 32 < c.seg_bits) o[0] = (uint16_t)(word >> 16);
                if (k * 32 + 16 < c.seg_bits) o[1] = (uint16_t)(word & 0xffff);
            }
        }
    }
    // start the next word: message bits e-5..e are the state index
    insert_field<MET, P, 26, 6, true>(c.st, c.lane_field[P / 2]);
    if constexpr (norm_period<MET, IN>() <= 32) normalize<MET, IN>(c.st);
}

// i-th insertion batch of a survivor word (i = 0..4), all at phase P: 6 bits at 20,14,8,2, then 2 bits at 0
template <int MET, int P>
VIT_HD void insert_batch(LaneState<MET>& st, uint32_t lane_field, int i) {
    switch (i) {
        case 0: insert_field<MET, P, 20, 6, false>(st, lane_field); break;
        case 1: insert_field<MET, P, 14, 6, false>(st, lane_field); break;
        case 2: insert_field<MET, P, 8, 6, false>(st, lane_field); break;
        case 3: insert_field<MET, P, 2, 6, false>(st, lane_field); break;
        default: insert_field<MET, P, 0, 2, false>(st, lane_field); break;
    }
}

// stage S of the super-step; tbl = table base advanced by the loop iteration (i * BM_JBLOCK)
template <int MET, int IN, int BPP, int TBL, int S>
VIT_HD void one_stage(WarpCtx<MET, IN, BPP, TBL>& c, const uint8_t* tbl) {
    constexpr int P = S % 6;
    // class c holds (X, Y) in the core's operand encoding; class 3-c holds exactly (-X, -Y), so the
    // operands of the opposite branches cost a second LDS.64 instead of arithmetic
    const uint32_t* ent = reinterpret_cast<const uint32_t*>(tbl + row_off<TBL>(S) + c.bm_off[P]);
    const uint32_t* entn = reinterpret_cast<const uint32_t*>(tbl + row_off<TBL>(S) + c.bm_offn[P]);
    if constexpr (xmask_of(P) != 0) exchange_half<MET, xmask_of(P)>(c.st);
#if defined(__CUDA_ARCH__)
    const uint2 w = *reinterpret_cast<const uint2*>(ent);
    const uint2 n = *reinterpret_cast<const uint2*>(entn);
    acs_stage<MET, IN, P>(c.st, Operands{w.x, w.y, n.x, n.y, c.one});
#else
    acs_stage<MET, IN, P>(c.st, Operands{ent[0], ent[1], entn[0], entn[1], c.one});
#endif
}

// One 32-stage slide = survivor word W of the super-step (stages 32W .. 32W+31):
//   5 x { 6 stages ; insertion batch i }   (batches land on stages 32W+5, +11, +17, +23, +29)
//   2 stages ; flush + traceback + emit     (stage 32W+31)
// One loop branch and one batch dispatch per 6 stages; everything else is straight-line.
// Returns true when the segment group has emitted its last word.
template <int MET, int IN, int BPP, int TBL, int W>
VIT_HD bool slide(WarpCtx<MET, IN, BPP, TBL>& c, unsigned Tmax) {
    using SM = Smem<IN, TBL>;
    constexpr int S0 = 32 * W;
    constexpr int PB = (S0 + 5) % 6;              // phase of the batch stages: 5, 1, 3
    const uint8_t* tbl = c.smem + SM::OFF_BM;
#pragma unroll 1
    for (int i = 0; i < 5; i++, tbl += BM_JBLOCK) {
        one_stage<MET, IN, BPP, TBL, S0 + 0>(c, tbl);
        one_stage<MET, IN, BPP, TBL, S0 + 1>(c, tbl);
        one_stage<MET, IN, BPP, TBL, S0 + 2>(c, tbl);
        one_stage<MET, IN, BPP, TBL, S0 + 3>(c, tbl);
        one_stage<MET, IN, BPP, TBL, S0 + 4>(c, tbl);
        one_stage<MET, IN, BPP, TBL, S0 + 5>(c, tbl);
        insert_batch<MET, PB>(c.st, c.lane_field[PB / 2], i);
        if constexpr (norm_period<MET, IN>() == 16) {
            if (i == 2) normalize<MET, IN>(c.st);
        }
    }
    one_stage<MET, IN, BPP, TBL, S0 + 30>(c, c.smem + SM::OFF_BM);
    one_stage<MET, IN, BPP, TBL, S0 + 31>(c, c.smem + SM::OFF_BM);
    slide_end<MET, IN, BPP, TBL, W>(c);
    return c.t0 + 32 * (W + 1) >= Tmax;
}

// kp.gate_super[i] without dynamic indexing of the kernel parameters (that would copy them to local memory)
VIT_HD unsigned gate_super_at(const KParams& kp, unsigned i) {
    unsigned v = kp.gate_super[0];
#pragma unroll
    for (unsigned k = 1; k < 8; k++) v = (i == k) ? kp.gate_super[k] : v;
    return v;
}

// Decode the 4 segments owned by warp `warp_id` of stream `stream`.
template <int MET, int IN, int BPP, int TBL = 96>
VIT_HD void warp_body(const KParams& kp, unsigned warp_id, unsigned stream, int lane, uint8_t* smem) {
    using SM = Smem<IN, TBL>;
    WarpCtx<MET, IN, BPP, TBL> c;
    c.smem = smem; c.lane = lane; c.g = lane >> 3; c.l = lane & 7;
    c.one = kp.one;

    // segment partition, reference viterbi.cu:156-165
    const unsigned W = kp.segments;
    const unsigned long long q = kp.packs / W, rem = kp.packs % W;
    const unsigned long long w = (unsigned long long)kp.seg_first + (unsigned long long)warp_id * SEGS_PER_WARP + c.g;
    unsigned long long Lp = (w < kp.seg_limit) ? q + (w < rem ? 1 : 0) : 0;
    const unsigned long long start_pack = q * w + (w < rem ? w : rem);
    c.seg_bits = (unsigned)(Lp * BPP);
    c.out_word0 = start_pack;
    const unsigned long long s0 = start_pack * BPP;                       // first message index
    c.seg_byte0 = s0 * InTraits<IN>::B96 / 96;
    c.in = kp.in + (unsigned long long)stream * kp.in_stride;
    c.out = kp.out + (unsigned long long)stream * kp.out_stride;
    c.in_bytes = kp.in_bytes;

    // the first segment of the warp is never shorter than the others
    const unsigned long long w_first = (unsigned long long)kp.seg_first + (unsigned long long)warp_id * SEGS_PER_WARP;
    const unsigned long long Lp_first = (w_first < kp.seg_limit) ? q + (w_first < rem ? 1 : 0) : 0;
    if (Lp_first == 0) return;
    const unsigned Lmax = (unsigned)(Lp_first * BPP);
    const unsigned Tmax = 64 + 32 * ((Lmax + 31) / 32);                   // viterbi.cu:176-197
    const unsigned nsuper = (Tmax + SUPER - 1) / SUPER;

#pragma unroll
    for (int p = 0; p < 6; p++) {
        c.bm_off[p] = (uint32_t)(c.g * 32 + lane_class_of(c.l, p) * 8);
        c.bm_offn[p] = (uint32_t)(c.g * 32 + (3 - lane_class_of(c.l, p)) * 8);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) c.lane_field[i] = lane_field_of(c.l, 2 * i + 1);
#pragma unroll
    for (int r = 0; r < LaneState<MET>::NPM; r++) c.st.pm[r] = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) c.st.pp[k] = 0;

    // traceback lookup: variant v (ring slot of e) -> ring word of a state at the phase of stage e-32
    {
        uint8_t* lut = smem + SM::OFF_LUT;
        for (int i = lane; i < 3 * 64; i += 32) {
            int v = i / 64, st = i % 64;
            int pw = mod6((v * 32 + 31) + 4);                           // phase of stage e-32
            int pos = pos_index_of_state(st, pw);                       // lane*8 + reg
            lut[i] = (uint8_t)ring_word_of(pos >> 3, pos & 7);
        }
    }

    // upload gates passed so far (uniform across the warp)
    unsigned gate_next = 0;
#define VIT_GATE(x)                                                                               \
    while (gate_next < kp.gate_n && (x) >= gate_super_at(kp, gate_next)) {                        \
        if (!gate_spin(kp.gate + gate_next, kp.gate_epoch)) { *kp.gate_err = 1u; return; }        \
        gate_next++;                                                                              \
    }
    VIT_GATE(0u)
    issue_raw_copy(c, 0);
#pragma unroll 1
    for (unsigned sc = 0; sc < nsuper; sc++) {
        c.t0 = sc * SUPER;
        const unsigned skew = (unsigned)((c.seg_byte0 + (unsigned long long)sc * InTraits<IN>::B96) & 15ull);
        cp_async_wait<0>();
        syncwarp();
        // the raw staging buffer is single: the next super-step's words are requested as soon as the last table
        // build of this one has consumed it, and land while the remaining stages run
#define VIT_PREFETCH                                                                              \
    if (sc + 1 < nsuper) {                                                                        \
        VIT_GATE(sc + 1)                                                                          \
        issue_raw_copy(c, sc + 1);                                                                \
    }
        if constexpr (TBL == 96) {
            build_table(c, skew);
            syncwarp();
            VIT_PREFETCH
        }
        if constexpr (norm_period<MET, IN>() == 96) normalize<MET, IN>(c.st);
        if constexpr (TBL == 32) build_slide<MET, IN, BPP, TBL, 0>(c, skew);
        if (slide<MET, IN, BPP, TBL, 0>(c, Tmax)) break;
        if constexpr (TBL == 32) build_slide<MET, IN, BPP, TBL, 1>(c, skew);
        if (slide<MET, IN, BPP, TBL, 1>(c, Tmax)) break;
        if constexpr (TBL == 32) {
            build_slide<MET, IN, BPP, TBL, 2>(c, skew);
            VIT_PREFETCH
        }
        if (slide<MET, IN, BPP, TBL, 2>(c, Tmax)) break;
        syncwarp();
    }
#undef VIT_PREFETCH
#undef VIT_GATE
}

#if defined(__CUDACC__)
template <int MET, int IN, int BPP, int TBL>
__global__ void __launch_bounds__(32) vit_decode_kernel(const KParams kp) {
    extern __shared__ __align__(16) uint8_t vit_smem[];
    warp_body<MET, IN, BPP, TBL>(kp, blockIdx.x, blockIdx.y, (int)threadIdx.x, vit_smem);
}
#endif

}  // namespace vitk
