/*
 * vit_oracle.h -- scalar C golden model of the reference's K=7 rate-1/2 Viterbi decode path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (csrc/, the C-ABI library, the C++ shim,
 * the harness) may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / CPU baseline.
 *
 * It is a CPU *restatement* of the algorithm in /root/reference/src/viterbi (the reference ships no
 * CPU decoder and no tests).  Parity pin: the restatement is checked
 * word-for-word against the reference's own CUDA decoder (oracle/_ref/libvitref.so, built from the
 * reference sources by oracle/Makefile) on a B200, and against tests/golden/ vectors that were
 * produced by that library (tests/golden/make_golden.py).
 *
 * Option bitfield values follow reference src/viterbi/viterbi.h:7-20.
 */
#ifndef VIT_ORACLE_H
#define VIT_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* reference src/viterbi/viterbi.h:17-20 */
enum {
    VO_HARD = 0x0, VO_SOFT4 = 0x1, VO_SOFT8 = 0x2, VO_SOFT16 = 0x3, VO_FP32 = 0x4,
    VO_M_B32 = 0x00, VO_M_B16 = 0x10, VO_M_FP16 = 0x20,
    VO_O_B32 = 0x000, VO_O_B16 = 0x100,
    VO_REG = 0x0000, VO_DPX = 0x1000,
    /* extension (not a reference option value): the tie rule the reference's DPX code paths define (viterbiACS.cuh:101-110,
     * 123-134,205-213,224-236) but which the reference never instantiates, because viterbi_core does not forward compMode to
     * forwardACS (viterbi.cu:181,192,204).  int16x2: same as REG.  int32: the partner/even predecessor wins ties for the
     * u=1 state at phase 0 too (REG: the odd one).  Pinned by oracle/_ref/libvitref_dpx.so, a build of the reference sources
     * with the flag forwarded (oracle/ref_dpx_shim.cu). */
    VO_DPX_TIES = 0x2000
};

/* flags for vo_decode */
enum {
    /* emulate the reference's O_B16 over-run stores (SURVEY.md 8a "O_B16 over-run quirk",
     * reference viterbi.cu:186,199-206): a segment whose length is 16 mod 32 also decodes the first
     * two 16-bit words of the following segment, and (running last) overwrites them. */
    VO_FLAG_REF_OVERRUN = 1
};

/* reference viterbi.h:22-36 (OptionsValid) */
int vo_options_valid_ref(int options);

/* reference viterbi.cu:63-92 */
size_t vo_input_size(int options, size_t inputNum);
size_t vo_message_len(int options, size_t inputNum);
size_t vo_output_size(int options, size_t inputNum);

/* number of stream segments, reference viterbi.cu:19 (blocksNum_total = 16*400) */
#define VO_SEGMENTS 6400
/* test hook: decode with a different segment count (0 restores 6400); not thread safe */
void vo_set_segments(size_t w);

/*
 * Decode `inputNum` coded symbols (packed as the reference's encPack_t stream for the option's input
 * type) into vo_output_size() bytes of packed bits (decPack_t words, MSB = earliest).
 * nthreads <= 0 -> all cores (OpenMP over the 6400 independent segments).
 * Returns 0, or -1 for an option combination this model does not define.
 */
int vo_decode(int options, const void* in, void* out, size_t inputNum, int nthreads, int flags);

/* decode only segments [seg_begin, seg_end) -- used by the bounded-sample CPU baseline */
int vo_decode_segments(int options, const void* in, void* out, size_t inputNum,
                       size_t seg_begin, size_t seg_end, int nthreads, int flags);

/* Windowed form for streams too long to hold on the host (BASELINE config 3, 4 Gbit of fp32 input = 32 GB):
 * vo_segment_window reports the input byte range segments [seg_begin, seg_end) read and the decoded packs they
 * own; vo_decode_window decodes them from a buffer holding just that byte range into a buffer holding just those
 * packs.  Same arithmetic as vo_decode_segments. */
void vo_segment_window(int options, size_t inputNum, size_t seg_begin, size_t seg_end,
                       size_t* in_byte0, size_t* in_bytes, size_t* out_word0, size_t* out_words);
int vo_decode_window(int options, const void* in_window, size_t in_byte0, size_t in_bytes, void* out_window,
                     size_t out_word0, size_t inputNum, size_t seg_begin, size_t seg_end, int nthreads);

/* word indices (in decPack_t units) whose value the reference leaves to a store race
 * (O_B16 over-run); writes up to cap indices, returns the count. */
size_t vo_overrun_words(int options, size_t inputNum, uint64_t* idx, size_t cap);

/* ---- host pipeline twin (reference src/viterbiDF.h) -- input generation for tests ---- */

/* Test hook: the generator polynomials the decoder model and vo_encode use (default 0171, 0133 = reference viterbi.h:62-63;
 * 0, 0 restores them).  Both must be 7-bit values tapping bits 0 and 6 (the reference's cores assume complementary branch
 * symbols); returns -1 otherwise.  Process-wide. */
int vo_set_polynomials(unsigned polyn1, unsigned polyn2);
void vo_get_polynomials(unsigned* polyn1, unsigned* polyn2);

/* K=7 encoder (polynomials as set above), reference viterbiDF.h:36-63.  bits[n] in {0,1} -> coded[2n] in {0,1} */
void vo_encode(const uint8_t* bits, size_t n, uint8_t* coded);

/* quantise+pack, reference viterbiDF.h:98-167: soft[nsym] floats (already noise-added, +-1 based),
 * multiplied by `scale`, quantised per input type and packed MSB-first into int32 words
 * (FP32: scaled floats).  nsym must be a multiple of the symbols-per-word. */
void vo_pack(int inputType, const float* soft, size_t nsym, float scale, void* out);

/* PRBS-31 (x^31 + x^28 + 1) message source used by the bench; state != 0 */
void vo_prbs31(uint32_t seed, uint8_t* bits, size_t n);

/* bit errors between decoded output and the message: out bit j <-> message bit j+26
 * (reference main.cpp:153-169) */
uint64_t vo_count_errors(int options, const void* out, size_t messageLen, const uint8_t* bits);

int vo_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
