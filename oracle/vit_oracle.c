/*
 * vit_oracle.c -- scalar C golden model (TEST INFRASTRUCTURE ONLY, see vit_oracle.h).
 *
 * Restates, state-indexed and without any lane/warp structure, what the reference's
 * viterbi_core kernel computes.  Citations are to /root/reference/src/viterbi/.
 *
 * Trellis (viterbi.h:61-63, viterbiDF.h:48-52): K=7, 64 states, rate 1/2, generators 0171/0133 (vo_set_polynomials: test hook).
 * The encoder buffer is (u<<6)|S with S the 6-bit state (newest bit at bit 5); next state
 * S' = (u<<5)|(S>>1); symbol k = (parity(buf&0171)<<1)|parity(buf&0133).  The two branches into a
 * state carry complementary symbols, so BM(odd predecessor) = -BM(even predecessor).
 *
 * Path metrics are kept in int64 here.  The reference's cores (int16x2 / int32 / half2) periodically
 * subtract the minimum (viterbiACS.cuh:307-378); a common offset never changes a decision and the
 * reference's own strides keep every core inside its exact range, so the only
 * core-specific behaviour is the tie rule, modelled below.
 */
#include "vit_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { NST = 64, EXTRA_L = 26, EXTRA_R = 38, SLIDE = 32, FWD = 96 };  /* viterbi.h:70-76 */

static inline int in_type(int o) { return o & 0xf; }
static inline int metric_type(int o) { return o & 0xf0; }
static inline int out_type(int o) { return o & 0xf00; }
static inline int bits_per_pack(int o) { return out_type(o) == VO_O_B16 ? 16 : 32; }

int vo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* viterbi.h:22-36 */
int vo_options_valid_ref(int o) {
    int it = in_type(o), mt = metric_type(o), cm = o & 0xf000;
    if (cm > VO_DPX) return 0;
    if (it == VO_SOFT8 && mt == VO_M_FP16) return 0;
    if (it == VO_SOFT16 && mt == VO_M_FP16) return 0;
    if (it == VO_SOFT16 && mt == VO_M_B16) return 0;
    if (mt == VO_M_FP16 && cm == VO_DPX) return 0;
    return 1;
}

/* viterbi.cu:63-84 */
size_t vo_input_size(int o, size_t n) {
    switch (in_type(o)) {
        case VO_HARD: return (n + 7) / 8;
        case VO_SOFT4: return (n + 1) / 2;
        case VO_SOFT8: return n;
        case VO_SOFT16: return n * 2;
        case VO_FP32: return n * 4;
        default: return 0;
    }
}
/* viterbi.cu:86-88 */
size_t vo_message_len(int o, size_t n) {
    size_t bpp = (size_t)bits_per_pack(o);
    if (n / 2 < (size_t)(EXTRA_L + EXTRA_R)) return 0;
    return (n / 2 - (EXTRA_L + EXTRA_R)) / bpp * bpp;
}
/* viterbi.cu:90-92 */
size_t vo_output_size(int o, size_t n) { return vo_message_len(o, n) / 8; }

/* Generator polynomials (viterbi.h:62-63: polyn1 0171, polyn2 0133).  Test hook vo_set_polynomials: another K=7 code, for
 * checking builds of the product library compiled for other polynomials (csrc/vit_code.h).  The restatement below, like the
 * reference's cores, needs both polynomials to tap encoder bits 0 and 6 (complementary symbols on the two branches into a
 * state); other values are rejected. */
static unsigned g_poly1 = 0171, g_poly2 = 0133;
int vo_set_polynomials(unsigned p1, unsigned p2) {
    if (p1 == 0 && p2 == 0) { p1 = 0171; p2 = 0133; }
    if (p1 > 0x7f || p2 > 0x7f || (p1 & 0101) != 0101 || (p2 & 0101) != 0101) return -1;
    g_poly1 = p1; g_poly2 = p2;
    return 0;
}
void vo_get_polynomials(unsigned* p1, unsigned* p2) { *p1 = g_poly1; *p2 = g_poly2; }

static inline int parity7(unsigned v) { return __builtin_popcount(v & 0x7f) & 1; }
static inline int sym_of(unsigned buf) { return (parity7(buf & g_poly1) << 1) | parity7(buf & g_poly2); }

/* ------------------------------------------------------------------------------------------ */
/* Branch metrics of one trellis stage.  `p` points at a zero-padded private copy of the       */
/* segment's input whose byte 0 is the byte holding message index `base`; `rel` = index-base.  */
/* viterbiBM.cuh:33-40 (HARD), 64-75 (SOFT4), 97-100 (SOFT8), 121-124 (SOFT16), 146-153 (FP32) */

static inline uint32_t ld32(const uint8_t* p, size_t word) {
    uint32_t v;
    memcpy(&v, p + 4 * word, 4);
    return v;
}

static inline void soft_bm(int d0, int d1, int bm[4]) {
    bm[0] = -d0 - d1; bm[1] = -d0 + d1; bm[2] = d0 - d1; bm[3] = d0 + d1;
}

static inline void stage_bm(int it, int f16_prescale, const uint8_t* p, size_t i, int bm[4]) {
    switch (it) {
        case VO_HARD: {
            uint32_t w = ld32(p, i >> 4);
            unsigned rx = (w >> (30 - 2 * (i & 15))) & 3u;
            for (int k = 0; k < 4; k++) bm[k] = 1 - __builtin_popcount(rx ^ (unsigned)k);
            break;
        }
        case VO_SOFT4: {
            uint32_t w = ld32(p, i >> 2);
            unsigned j = (unsigned)(i & 3);
            int d0 = (int)((w >> (28 - 8 * j)) & 0xf), d1 = (int)((w >> (24 - 8 * j)) & 0xf);
            d0 = (d0 ^ 8) - 8; d1 = (d1 ^ 8) - 8;
            soft_bm(d0, d1, bm);
            break;
        }
        case VO_SOFT8: {
            uint32_t w = ld32(p, i >> 1);
            unsigned j = (unsigned)(i & 1);
            int d0 = (int8_t)(w >> (24 - 16 * j)), d1 = (int8_t)(w >> (16 - 16 * j));
            if (f16_prescale) { d0 >>= 3; d1 >>= 3; }
            soft_bm(d0, d1, bm);
            break;
        }
        case VO_SOFT16: {
            uint32_t w = ld32(p, i);
            int d0 = (int16_t)(w >> 16), d1 = (int16_t)(w & 0xffff);
            if (f16_prescale) { d0 >>= 11; d1 >>= 11; }
            soft_bm(d0, d1, bm);
            break;
        }
        default: { /* FP32: clamp to [-8,7] (FPprecision=4), truncate toward zero */
            float b0, b1;
            memcpy(&b0, p + 8 * i, 4);
            memcpy(&b1, p + 8 * i + 4, 4);
            b0 = fminf(fmaxf(b0, -8.0f), 7.0f);
            b1 = fminf(fmaxf(b1, -8.0f), 7.0f);
            for (int k = 0; k < 4; k++) {
                float c0 = (k & 2) ? 1.0f : -1.0f, c1 = (k & 1) ? 1.0f : -1.0f;
                volatile float s = (c0 * b0) + (c1 * b1);
                bm[k] = (int)s;
            }
            break;
        }
    }
}

/* bytes of input per trellis stage, as a fraction num/den */
static inline void bytes_per_stage(int it, size_t* num, size_t* den) {
    switch (it) {
        case VO_HARD: *num = 1; *den = 4; break;
        case VO_SOFT4: *num = 1; *den = 1; break;
        case VO_SOFT8: *num = 2; *den = 1; break;
        case VO_SOFT16: *num = 4; *den = 1; break;
        default: *num = 8; *den = 1; break;
    }
}

static inline unsigned brev_state(uint32_t pp, int bpp) {
    /* viterbiTB.cuh:11,18: __brev(pp << (32-bpp)) & 63 */
    unsigned r = 0;
    for (int i = 0; i < 6; i++) r |= ((pp >> (bpp - 1 - i)) & 1u) << i;
    return r;
}

typedef struct {
    int options, flags;
    const uint8_t* in;
    size_t in_bytes;
    uint8_t* out;
    size_t M, P;          /* decoded bits, decoded packs */
    uint32_t* overrun;    /* [W][2] values, REF_OVERRUN only */
    uint8_t* overrun_valid;
    /* window mode (vo_decode_window): `in` holds stream bytes [in_off, in_off + in_avail) only and `out`
     * points at decoded pack number out_off; whole-stream calls have in_off = out_off = 0, in_avail = in_bytes */
    size_t in_off, in_avail, out_off;
} dec_ctx;

static size_t g_segments = VO_SEGMENTS;   /* test hook: vo_set_segments */
void vo_set_segments(size_t w) { g_segments = w ? w : VO_SEGMENTS; }

static inline void seg_range(size_t P, int bpp, size_t w, size_t* s0, size_t* L) {
    /* viterbi.cu:156-162 */
    size_t q = P / g_segments, r = P % g_segments;
    *L = (q + (w < r ? 1 : 0)) * (size_t)bpp;
    *s0 = (q * w + (w < r ? w : r)) * (size_t)bpp;
}

static void store_word(const dec_ctx* c, size_t widx, uint32_t v) {
    widx -= c->out_off;
    if (bits_per_pack(c->options) == 16) ((uint16_t*)c->out)[widx] = (uint16_t)v;
    else ((uint32_t*)c->out)[widx] = v;
}

static void decode_segment(const dec_ctx* c, size_t w) {
    const int o = c->options;
    const int it = in_type(o), mt = metric_type(o), bpp = bits_per_pack(o);
    const int f16_prescale = (mt == VO_M_FP16) && (it == VO_SOFT8 || it == VO_SOFT16);
    size_t s0, L;
    seg_range(c->P, bpp, w, &s0, &L);
    if (L == 0) return;

    const int ref_overrun = (c->flags & VO_FLAG_REF_OVERRUN) && (L % 32 != 0);
    const size_t nslides = (L + SLIDE - 1) / SLIDE;                  /* viterbi.cu:186 */
    size_t T = EXTRA_L + EXTRA_R + nslides * SLIDE;                  /* stages of the main loops */
    const size_t T_tail = ref_overrun ? (L % 32) : 0;                /* viterbi.cu:199-206 */

    /* private zero-padded copy of this segment's input */
    size_t num, den;
    bytes_per_stage(it, &num, &den);
    size_t byte0 = s0 * num / den;                                   /* s0 is a multiple of 16 */
    size_t need = (T + T_tail + 16) * num / den + 16;
    uint8_t* buf = (uint8_t*)calloc(need, 1);
    if (byte0 < c->in_bytes && byte0 >= c->in_off && byte0 < c->in_off + c->in_avail) {
        size_t avail = c->in_bytes - byte0;                            /* bytes past the stream read as zeros */
        if (avail > c->in_off + c->in_avail - byte0) avail = c->in_off + c->in_avail - byte0;
        memcpy(buf, c->in + (byte0 - c->in_off), avail < need ? avail : need);
    }

    int sym0[32];
    for (int j = 0; j < 32; j++) sym0[j] = sym_of((unsigned)(2 * j));   /* u=0, even predecessor */

    int64_t pm[NST], npm[NST];
    uint32_t pp[NST], npp[NST];
    uint32_t ring[FWD / 16][NST];
    memset(pm, 0, sizeof pm);                                          /* viterbi.cu:168-169 */
    memset(pp, 0, sizeof pp);
    memset(ring, 0, sizeof ring);
    const uint32_t ppmask = bpp == 32 ? 0xffffffffu : 0xffffu;
    const size_t out_w0 = s0 / (size_t)bpp;
    const size_t own_words = L / (size_t)bpp;

    for (size_t t = 0; t < T + T_tail; t++) {
        int bm[4];
        stage_bm(it, f16_prescale, buf, t, bm);
        const int phase = (int)(t % 6);
        /* tie rules, viterbiACS.cuh:112-157,215-256 (REG variants) */
        int tie_u0, tie_u1;
        if (mt == VO_M_FP16) { tie_u0 = 0; tie_u1 = 1; }
        else if (mt == VO_M_B16) { tie_u0 = 1; tie_u1 = 0; }
        else if ((o & 0xf000) == VO_DPX_TIES) { tie_u0 = 1; tie_u1 = 0; }   /* viterbiACS.cuh:123-134,224-236 (DPX variants) */
        else { tie_u0 = 1; tie_u1 = (phase == 0) ? 1 : 0; }

        for (int j = 0; j < 32; j++) {
            const int64_t b = bm[sym0[j]];
            const int64_t e = pm[2 * j], od = pm[2 * j + 1];
            int64_t ce = e + b, co = od - b;                           /* new state j   (u=0) */
            int x = co > ce ? 1 : (co < ce ? 0 : tie_u0);
            npm[j] = x ? co : ce;
            npp[j] = (pp[2 * j + x] << 1) | (uint32_t)x;               /* viterbiACS.cuh:161-198 */
            ce = e - b; co = od + b;                                   /* new state j+32 (u=1) */
            x = co > ce ? 1 : (co < ce ? 0 : tie_u1);
            npm[j + 32] = x ? co : ce;
            npp[j + 32] = (pp[2 * j + x] << 1) | (uint32_t)x;
        }
        memcpy(pm, npm, sizeof pm);
        memcpy(pp, npp, sizeof pp);

        const int ppInd = (int)(t % FWD);
        if ((ppInd + 1) % bpp == 0) {                                  /* viterbiACS.cuh:391-414,517 */
            for (int s = 0; s < NST; s++) { ring[ppInd / bpp][s] = pp[s] & ppmask; pp[s] = 0; }
        }

        int do_tb = 0, tb_len = SLIDE;
        size_t data_end = 0;
        if (t < T) {
            if (t >= FWD - 1 && (t - (FWD - 1)) % SLIDE == 0) { do_tb = 1; data_end = t - (FWD - 1) + SLIDE - 1; }
        } else if (t == T + T_tail - 1) {                              /* remainder block */
            do_tb = 1; tb_len = (int)T_tail; data_end = nslides * SLIDE + T_tail - 1;
        }
        if (do_tb) {                                                   /* viterbiTB.cuh:4-21 */
            const size_t e = t;
            unsigned st = 0;
            for (int s = 0; s < EXTRA_R - bpp; s += bpp)
                st = brev_state(ring[((e - (size_t)s) % FWD) / (size_t)bpp][st], bpp) & 63u;
            for (int s = 0; s < tb_len; s += bpp) {
                size_t i = (e - EXTRA_R - (size_t)s) % FWD;
                uint32_t word = ring[i / (size_t)bpp][st];
                size_t lw = (data_end - (size_t)s) / (size_t)bpp;      /* local word index */
                if (lw < own_words) store_word(c, out_w0 + lw, word);
                else if (ref_overrun && lw - own_words < 2 && out_w0 + lw < c->P) {
                    c->overrun[2 * w + (lw - own_words)] = word;
                    c->overrun_valid[2 * w + (lw - own_words)] = 1;
                }
                st = brev_state(word, bpp) & 63u;
            }
        }
    }
    free(buf);
}

static int decode_range(int options, const void* in, size_t in_off, size_t in_avail, void* out, size_t out_off,
                        size_t inputNum, size_t seg_begin, size_t seg_end, int nthreads, int flags) {
    int it = in_type(options), mt = metric_type(options);
    if (it > VO_FP32) return -1;
    if (mt != VO_M_B32 && mt != VO_M_B16 && mt != VO_M_FP16) return -1;
    if (mt == VO_M_B16 && it == VO_SOFT16) return -1;                  /* viterbi.h:28-29 */
    if ((options & 0xf000) == VO_DPX_TIES && mt == VO_M_FP16) return -1; /* no half2 DPX code in the reference */
    if ((options & 0xf000) > VO_DPX_TIES) return -1;
    dec_ctx c;
    c.options = options; c.flags = flags;
    c.in = (const uint8_t*)in; c.in_bytes = vo_input_size(options, inputNum);
    c.out = (uint8_t*)out;
    c.M = vo_message_len(options, inputNum);
    c.P = c.M / (size_t)bits_per_pack(options);
    c.overrun = NULL; c.overrun_valid = NULL;
    c.in_off = in_off; c.in_avail = in_avail < c.in_bytes ? in_avail : c.in_bytes; c.out_off = out_off;
    if (seg_end > g_segments) seg_end = g_segments;
    if (flags & VO_FLAG_REF_OVERRUN) {
        c.overrun = (uint32_t*)calloc(2 * g_segments, sizeof(uint32_t));
        c.overrun_valid = (uint8_t*)calloc(2 * g_segments, 1);
    }
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 8) num_threads(nthreads)
#endif
    for (long w = (long)seg_begin; w < (long)seg_end; w++) decode_segment(&c, (size_t)w);

    if (flags & VO_FLAG_REF_OVERRUN) {
        /* the over-running warp stores last; lowest segment wins where several overlap */
        int bpp = bits_per_pack(options);
        for (long w = (long)seg_end - 1; w >= (long)seg_begin; w--) {
            size_t s0, L;
            seg_range(c.P, bpp, (size_t)w, &s0, &L);
            for (int k = 1; k >= 0; k--)
                if (c.overrun_valid[2 * w + k]) store_word(&c, (s0 + L) / (size_t)bpp + (size_t)k, c.overrun[2 * w + k]);
        }
        free(c.overrun); free(c.overrun_valid);
    }
    return 0;
}

int vo_decode_segments(int options, const void* in, void* out, size_t inputNum,
                       size_t seg_begin, size_t seg_end, int nthreads, int flags) {
    return decode_range(options, in, 0, (size_t)-1, out, 0, inputNum, seg_begin, seg_end, nthreads, flags);
}

/* the byte range of the input that segments [seg_begin, seg_end) read, and the range of decoded packs they own */
void vo_segment_window(int options, size_t inputNum, size_t seg_begin, size_t seg_end,
                       size_t* in_byte0, size_t* in_bytes, size_t* out_word0, size_t* out_words) {
    const int bpp = bits_per_pack(options);
    const size_t P = vo_message_len(options, inputNum) / (size_t)bpp, total = vo_input_size(options, inputNum);
    size_t num, den, s0a, La, s0b, Lb;
    bytes_per_stage(in_type(options), &num, &den);
    if (seg_end > g_segments) seg_end = g_segments;
    seg_range(P, bpp, seg_begin, &s0a, &La);
    seg_range(P, bpp, seg_end - 1, &s0b, &Lb);
    size_t b0 = s0a * num / den;
    size_t b1 = (s0b + EXTRA_L + EXTRA_R + (Lb + SLIDE - 1) / SLIDE * SLIDE + SLIDE + 16) * num / den + 16;
    if (b1 > total) b1 = total;
    if (b0 > b1) b0 = b1;
    *in_byte0 = b0; *in_bytes = b1 - b0;
    *out_word0 = s0a / (size_t)bpp; *out_words = (s0b + Lb - s0a) / (size_t)bpp;
}

/* Decode segments [seg_begin, seg_end) of a stream too long to hold on the host: in_window holds the stream's bytes
 * [in_byte0, in_byte0 + in_bytes) (as vo_segment_window reports them) and out_window receives the packs from out_word0 on. */
int vo_decode_window(int options, const void* in_window, size_t in_byte0, size_t in_bytes, void* out_window,
                     size_t out_word0, size_t inputNum, size_t seg_begin, size_t seg_end, int nthreads) {
    return decode_range(options, in_window, in_byte0, in_bytes, out_window, out_word0, inputNum, seg_begin, seg_end, nthreads, 0);
}

int vo_decode(int options, const void* in, void* out, size_t inputNum, int nthreads, int flags) {
    return vo_decode_segments(options, in, out, inputNum, 0, g_segments, nthreads, flags);
}

size_t vo_overrun_words(int options, size_t inputNum, uint64_t* idx, size_t cap) {
    int bpp = bits_per_pack(options);
    size_t P = vo_message_len(options, inputNum) / (size_t)bpp, cnt = 0;
    if (bpp != 16) return 0;
    for (size_t w = 0; w < g_segments; w++) {
        size_t s0, L;
        seg_range(P, bpp, w, &s0, &L);
        if (L == 0 || L % 32 == 0) continue;
        for (size_t k = 0; k < 2; k++) {
            size_t g = (s0 + L) / 16 + k;
            if (g < P) { if (cnt < cap && idx) idx[cnt] = g; cnt++; }
        }
    }
    return cnt;
}

/* ------------------------------------------------------------------------------------------ */
/* host pipeline twin, reference src/viterbiDF.h                                               */

void vo_encode(const uint8_t* bits, size_t n, uint8_t* coded) {        /* viterbiDF.h:36-63 */
    unsigned buffer = 0;
    for (size_t i = 0; i < n; i++) {
        buffer >>= 1;
        buffer |= (unsigned)(bits[i] & 1) << 6;
        coded[2 * i] = (uint8_t)parity7(buffer & g_poly1);
        coded[2 * i + 1] = (uint8_t)parity7(buffer & g_poly2);
    }
}

static inline int32_t quant(int it, float v) {                          /* viterbiDF.h:105-125 */
    long q;
    switch (it) {
        case VO_HARD: return v > 0.0f ? 1 : 0;
        case VO_SOFT4: q = lrintf(v); if (q < -8) q = -8; if (q > 7) q = 7; return (int32_t)(q & 0xF);
        case VO_SOFT8: q = lrintf(v); if (q < -128) q = -128; if (q > 127) q = 127; return (int32_t)(q & 0xFF);
        default: q = lrintf(v); if (q < -32768) q = -32768; if (q > 32767) q = 32767; return (int32_t)(q & 0xFFFF);
    }
}

void vo_pack(int it, const float* soft, size_t nsym, float scale, void* out) { /* viterbiDF.h:139-166 */
    if (it == VO_FP32) {
        float* o = (float*)out;
        for (size_t i = 0; i < nsym; i++) o[i] = soft[i] * scale;
        return;
    }
    int width = it == VO_HARD ? 1 : it == VO_SOFT4 ? 4 : it == VO_SOFT8 ? 8 : 16;
    size_t per = 32 / (size_t)width;
    uint32_t* o = (uint32_t*)out;
    for (size_t i = 0; i + per <= nsym; i += per) {
        uint32_t b = 0;
        for (size_t j = i; j < i + per; j++) {
            b = (width == 32) ? 0 : (b << width);
            b |= (uint32_t)quant(it, soft[j] * scale);
        }
        o[i / per] = b;
    }
}

void vo_prbs31(uint32_t seed, uint8_t* bits, size_t n) {
    uint32_t s = seed & 0x7fffffffu;
    if (!s) s = 0x7fffffffu;
    for (size_t i = 0; i < n; i++) {
        uint32_t nb = ((s >> 30) ^ (s >> 27)) & 1u;
        s = ((s << 1) | nb) & 0x7fffffffu;
        bits[i] = (uint8_t)nb;
    }
}

uint64_t vo_count_errors(int options, const void* out, size_t messageLen, const uint8_t* bits) {
    int bpp = bits_per_pack(options);                                   /* main.cpp:153-169 */
    uint64_t errs = 0;
    for (size_t i = 0; i < messageLen; i++) {
        uint32_t word = bpp == 16 ? ((const uint16_t*)out)[i / 16] : ((const uint32_t*)out)[i / 32];
        int d = (int)((word >> (bpp - 1 - (i % (size_t)bpp))) & 1u);
        errs += (uint64_t)(d != (bits[i + EXTRA_L] & 1));
    }
    return errs;
}
