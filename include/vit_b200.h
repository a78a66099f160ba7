/*
 * vit_b200.h -- C ABI of the B200-native Viterbi decoder (libvitb200.so).
 *
 * Drop-in boundary for the reference's host class `template<int options> class ViterbiCUDA`
 * (reference src/viterbi/viterbi.h:91-152, implemented in src/viterbi/viterbi.cu:10-139,210-238)
 * as consumed by ViterbiDecoder<options>::process (src/viterbiDF.h:186-196) and runPipeline
 * (src/main.cpp:119-138).  Plain pointers and sizes only; no C++ or torch types.
 *
 * `options` is the reference's bitfield (viterbi.h:7-20):
 *   bits 0-3  ChannelIn : HARD=0 SOFT4=1 SOFT8=2 SOFT16=3 FP32=4
 *   bits 4-7  Metric    : B32=0x00 B16=0x10 FP16=0x20
 *   bits 8-11 DecodeOut : O_B32=0x000 O_B16=0x100
 *   bits 12-15 CompMode : REG=0x0000 DPX=0x1000 (both select the same core, as in the reference
 *                         where the flag is never forwarded: viterbi.cu:181,192,204);
 *                         extension DPX_TIES=0x2000: the tie rule the reference's DPX code paths define
 *                         (viterbiACS.cuh:101-110,123-134,205-213,224-236) but never run -- identical to REG for
 *                         the int16x2 core, "the partner wins ties in every phase" for the int32 core, rejected
 *                         for the half2 core (the reference has no half2 DPX code)
 * Every size argument called `inputNum` is the number of CODED SYMBOLS (2 x message bits), exactly
 * as in the reference (viterbi.cu:63-92,210-215).
 *
 * All functions returning int return 0 on success and a non-zero code on failure;
 * vit_last_error() then describes the failure (thread local).  Nothing here prints or exits: the
 * C++ shim (host/viterbi.h) turns a failure into the reference's "print + exit(EXIT_FAILURE)"
 * convention (reference src/viterbi/gpuerrors.h:8-17).
 */
#ifndef VIT_B200_H
#define VIT_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vit_handle vit_handle;

enum {
    VIT_OK = 0,
    VIT_ERR_OPTIONS = 1,   /* option combination not supported */
    VIT_ERR_CUDA = 2,      /* a CUDA runtime call failed */
    VIT_ERR_ARG = 3,       /* bad pointer / alignment / size */
    VIT_ERR_NCCL = 4       /* an NCCL call failed, or libnccl.so.2 could not be loaded */
};

/* replaces ViterbiCUDA<options>::ViterbiCUDA() and ViterbiCUDA(size_t inputNum)
 * (viterbi.cu:23-36).  prealloc_inputNum > 0 sizes the device buffers up front; they grow on
 * demand otherwise.  `device` is the CUDA ordinal (the reference hard-codes 0, viterbi.cu:134). */
int vit_create(vit_handle** out, int options, int device, size_t prealloc_inputNum);

/* replaces ~ViterbiCUDA (viterbi.cu:38-42) */
void vit_destroy(vit_handle* h);

/* replaces ViterbiCUDA<options>::run(input_h, output_h, inputNum, kernelTime) (viterbi.cu:210-238):
 * host buffers in and out, synchronous.  in_h holds vit_input_size() bytes, out_h receives
 * vit_output_size() bytes.  kernel_ms (optional) = device time of the decode kernel in ms,
 * measured with CUDA events around the launch only, as the reference does (viterbi.cu:224-232). */
int vit_run(vit_handle* h, const void* in_h, void* out_h, size_t inputNum, float* kernel_ms);

/* How vit_run moves host buffers (default VIT_UPLOAD_AUTO; the environment variable VIT_RUN_MODE=<value> sets the
 * initial mode of every handle).  Results are identical in every mode.
 *   AUTO        time-sliced upload (one decode launch that waits, inside the kernel, for column blocks of its input)
 *               when copies can run beside kernels, otherwise as CHUNKED.  Pinned buffers are read/written in place;
 *               pageable buffers (the reference's calling convention, viterbiDF.h:188-193) are staged through pinned
 *               buffers by worker threads (VIT_STAGE_THREADS, default min(16, cores / local ranks); the caller is blocked meanwhile).  A run with kernel_ms != NULL
 *               always takes the SEQUENTIAL path, so that the reported time is the decode kernel alone.
 *   SEQUENTIAL  the reference's copy -> launch -> copy sequence (viterbi.cu:219-235).
 *   CHUNKED     segment-range pipeline: pinned buffers only, pageable buffers fall back to SEQUENTIAL.
 *   GATED       insist on the time-sliced upload (fails instead of falling back if an upload gate times out).
 * AUTO avoids the time-sliced upload up front under CUDA_DEVICE_MAX_CONNECTIONS=1, CUDA_LAUNCH_BLOCKING=1 and under
 * tools injected into the CUDA driver (ncu, nsys, compute-sanitizer), where a copy issued after a kernel cannot run
 * before that kernel ends; should a gate still time out (copy size / 1 GB/s + 0.2 s), the call decodes again through
 * the fallback and the handle stops using gates. */
enum { VIT_UPLOAD_AUTO = 0, VIT_UPLOAD_SEQUENTIAL = 1, VIT_UPLOAD_CHUNKED = 2, VIT_UPLOAD_GATED = 3 };
int vit_set_upload_mode(vit_handle* h, int mode);
/* the mode AUTO currently resolves to (GATED or CHUNKED), or the mode that was set */
int vit_upload_mode_in_effect(const vit_handle* h);

/* device-resident variant (no copies): in_d must be 16-byte aligned and hold vit_input_size()
 * bytes; out_d receives vit_output_size() bytes.  cuda_stream is a cudaStream_t (NULL = default
 * stream).  Asynchronous unless kernel_ms is non-NULL.  New relative to the reference: needed for
 * device-generated input and multi-GPU stream sharding. */
int vit_run_device(vit_handle* h, const void* in_d, void* out_d, size_t inputNum,
                   void* cuda_stream, float* kernel_ms);

/* nstreams independent codeword streams of inputNum symbols each, one launch.  Stream s lives at
 * in_d + s*in_stride (16-byte multiple) and decodes to out_d + s*out_stride. */
int vit_run_device_batch(vit_handle* h, const void* in_d, void* out_d, size_t inputNum,
                         size_t nstreams, size_t in_stride, size_t out_stride,
                         void* cuda_stream, float* kernel_ms);

/* Chunked decode of an endless stream (new; SURVEY.md 8f: the reference decodes one buffer per run() and restarts from
 * all-zero metrics every call, viterbi.cu:168-169, so consecutive calls lose 64 stages at every seam).
 * Every push appends `inputNum` coded symbols (whole 32-bit channel packs) to the stream.  The decoder forms the window
 * "symbols carried from the previous push ++ this chunk", decodes it exactly as vit_run would (same 6400-segment
 * partition OF THE WINDOW, same warm-up, same tie rules: bit-identical to the reference's run() on that window), emits
 * M = ((window stages - 64) / bitsPerPack) * bitsPerPack bits and carries the window's symbols from stage M on (64 to
 * 64 + bitsPerPack - 1 stages, kept on the device) into the next window.  Window k+1 therefore starts at stage M of
 * window k and its first decoded bit is the one after window k's last: the concatenated outputs are the contiguous
 * message bits 26, 27, ... of the endless stream, and equal a one-shot decode of the concatenated input in which the
 * segment partition is applied per window.  The last 38..69 stages pushed are always pending (the reference likewise
 * never decodes its last extraR bits).  out_cap: capacity of the output buffer in bytes; *out_bytes: bytes written. */
int vit_stream_reset(vit_handle* h);
int vit_stream_push(vit_handle* h, const void* in_h, size_t inputNum, void* out_h, size_t out_cap, size_t* out_bytes);
/* device-resident chunk and output (out_d aligned to its packs); asynchronous on cuda_stream */
int vit_stream_push_device(vit_handle* h, const void* in_d, size_t inputNum, void* out_d, size_t out_cap,
                           size_t* out_bytes, void* cuda_stream);
size_t vit_stream_pending(const vit_handle* h);            /* carried coded symbols */
unsigned long long vit_stream_bits(const vit_handle* h);   /* bits emitted since vit_stream_reset */

/* replace getInputSize / getMessageLen / getOutputSize (viterbi.cu:63-92); bytes, bits, bytes */
size_t vit_input_size(int options, size_t inputNum);
size_t vit_message_len(int options, size_t inputNum);
size_t vit_output_size(int options, size_t inputNum);

/* Code parameters this build of the library decodes (and its device source encodes): constraint length and the two
 * generator polynomials, octal-style bit masks over the encoder register with bit 6 = newest bit -- the reference's
 * ViterbiCUDA<>::constLen / polyn1 / polyn2 (viterbi.h:61-63: 7, 0171, 0133).  K = 7 is fixed; the polynomials are
 * compile-time parameters of the library (csrc/vit_code.h, `make EXTRA="-DVIT_POLY1=0117 -DVIT_POLY2=0155" ...`) and
 * must both tap bits 0 and 6, as the reference's own cores assume.  Any pointer may be NULL. */
void vit_code_parameters(int* constLen, int* polyn1, int* polyn2);

/* 1 if this library decodes the combination.  Superset of the reference's OptionsValid
 * (viterbi.h:22-36): FP16 metric with SOFT8/SOFT16 input is accepted (symbols are pre-scaled to
 * 5 bits); B16 metric with SOFT16 input is rejected as in the reference. */
int vit_options_valid(int options);
/* exactly the reference's OptionsValid<options>::value */
int vit_options_valid_ref(int options);

/* static kernel facts for reporting: registers per thread, dynamic shared memory per block,
 * threads per block, stream segments per block.  Returns non-zero if the kernel is missing. */
int vit_kernel_info(int options, int* regs, int* smem_bytes, int* block_threads, int* segs_per_block);

/* Lane geometry of the decode kernel for this handle's gate-free launches (vit_run_device, vit_run_device_batch, the stream
 * job; vit_run's time-sliced upload always uses L8).
 *   L8 (default)  8 lanes per segment, 4 segments per warp: the measured product kernel.
 *   L1            one lane per segment, 32 segments per warp (csrc/vit_kernel_l1.inc): no lane exchanges, no operand table.
 *                 EXPERIMENTAL: bit-exact against the golden model in the emulator and against L8 on a B200, measured once
 *                 (119.2 against 108.8 Gb/s on 16 streams per launch, profiles/r2_l1_bench.txt).  A stream is only 200 such
 *                 warps, so it pays from about a dozen streams per launch; packed cores (b16, f16) and local output buffers
 *                 only -- a launch that does not qualify uses L8.
 * The environment variable VIT_GEOMETRY=l1 sets the initial value of every handle.  Results are identical. */
enum { VIT_GEOMETRY_L8 = 0, VIT_GEOMETRY_L1 = 1 };
int vit_set_geometry(vit_handle* h, int geometry);
/* the geometry the last launch actually ran with */
int vit_last_launch_geometry(const vit_handle* h);

/* launches issued by this handle so far (one decode kernel per run) */
unsigned long long vit_launch_count(const vit_handle* h);

/* 1 if the last launch stored its packs 32 bytes per segment at a time (output buffer on another GPU: a mapping from
 * vit_comm_shared_alloc or any peer pointer), 0 if one pack per slide (local output) */
int vit_last_launch_staged_output(const vit_handle* h);

/* test hook: number of stream segments (reference: 6400, viterbi.cu:19); 0 restores 6400 */
int vit_set_segments(vit_handle* h, unsigned segments);

/* Device-side synthetic channel source (new): the GPU twin of the reference harness's host
 * chain RandBitGen | ConvolutionalEncoder | AddNoise | SoftDecisionPacker (src/viterbiDF.h:20-167), counter
 * based and integer only so a CPU twin reproduces it bit for bit.  Writes the packed received stream for
 * n_bits message bits to packed_d (whole 32-bit packs) and, if bits_d != NULL, one byte per message bit.
 * amp = symbol amplitude in quantiser units (<= 0: default per input type), sigma = noise sd / amp,
 * zero != 0: all-zero symbols (tie stress). */
int vit_synth_device(int input_type, size_t n_bits, unsigned seed, int amp, double sigma, int zero,
                     void* packed_d, void* bits_d, void* cuda_stream);

/* the same with a choice of message source: 0 = counter hash (vit_synth_device), 1 = PRBS-31 (x^31 + x^28 + 1 started
 * from state `seed`; the bench's source) */
int vit_synth_device_ex(int input_type, size_t n_bits, unsigned seed, int amp, double sigma, int zero, int source,
                        void* packed_d, void* bits_d, void* cuda_stream);

/* Bit errors of a decoded stream against the message bits of the synthetic source itself, regenerated on the fly (no
 * bits buffer: for multi-Gbit streams it would be 8x the decoded output).  Synchronous. */
int vit_count_errors_synth_device(int options, const void* out_d, size_t messageLen, unsigned seed, int source,
                                  unsigned long long* errors, void* cuda_stream);

/* Depuncturing pre-pass (new; the reference has no punctured modes): expands a punctured rate-k/n stream of soft symbols to
 * the rate-1/2 stream vit_run_device takes, writing zero-valued symbols (erasures: 0 in every branch metric) where the
 * transmitter dropped one.  keep0 / keep1: bit t set = the 0171 / 0133 symbol of stage t of the `period`-stage pattern is
 * transmitted (e.g. DVB-S rate 3/4: period 3, keep0 = 0b101, keep1 = 0b011; 2/3: period 2, 0b01, 0b11).  Symbols are packed
 * like the decoder's input type (SOFT4 / SOFT8 / SOFT16 / FP32; hard decisions have no erasure value).  Note that the
 * decoder keeps the reference's window constants (32-stage warm-up, 38-stage traceback merge), which are sized for the
 * rate-1/2 code: rates above 2/3 decode with a truncation error floor. */
int vit_depuncture_device(int input_type, const void* in_d, size_t n_in_syms, unsigned period, unsigned keep0, unsigned keep1,
                          void* out_d, size_t n_out_stages, void* cuda_stream);

/* Bit errors of a decoded stream counted on the device: #{ j < messageLen : out bit j != bits[j + 26] }
 * (the reference's BER loop, src/main.cpp:153-169); bits_d holds one byte per message bit.  Synchronous. */
int vit_count_errors_device(int options, const void* out_d, const void* bits_d, size_t messageLen,
                            unsigned long long* errors, void* cuda_stream);

/* device-memory helpers for host-only callers of the device-resident entry points (no CUDA headers needed) */
int vit_dev_alloc(void** ptr, size_t bytes);
void vit_dev_free(void* ptr);
int vit_dev_sync(void);
int vit_dev_set(int device);                 /* cudaSetDevice for the calling thread */
int vit_dev_copy_to_host(void* dst_h, const void* src_d, size_t bytes);      /* synchronous */
int vit_dev_copy_from_host(void* dst_d, const void* src_h, size_t bytes);
int vit_dev_count(void);
/* page-locked (pinned) host memory: vit_run reads and writes such buffers in place (no staging copy); pageable buffers are
 * staged through the handle's own pinned buffers by worker threads (see vit_set_upload_mode) */
int vit_host_alloc(void** ptr, size_t bytes);
void vit_host_free(void* ptr);

/* ------------------------------------------------------------------------------------------------------------------
 * Multi-GPU (new; the reference is single-GPU, cudaSetDevice(0) at src/viterbi/viterbi.cu:134).  Independent codeword
 * streams are sharded over the GPUs of one box -- one decoder per GPU, a stream is never split -- and only the packed
 * output bits cross GPUs: they are gathered into one stream-major buffer on a root GPU (BASELINE.json
 * configs[4]).  NCCL is loaded at run time (dlopen of libnccl.so.2), the library has no link dependency on it.
 * A communicator is either one of N created by ONE process for N GPUs (vit_comm_init_all; one host thread per GPU then
 * drives its own communicator) or this process's rank in a job of N processes (vit_comm_init_rank; the 128-byte id comes
 * from vit_comm_get_unique_id on one rank and is distributed by the caller, e.g. over torch.distributed or MPI).
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct vit_comm vit_comm;
#define VIT_COMM_ID_BYTES 128
int vit_comm_available(void);                 /* 1 if libnccl.so.2 could be loaded */
int vit_comm_nccl_version(void);              /* e.g. 22809, 0 if unavailable */
int vit_comm_get_unique_id(void* id128);
int vit_comm_init_rank(vit_comm** out, int nranks, int rank, const void* id128, int device);
int vit_comm_init_all(vit_comm** out /* [ndev] */, int ndev, const int* devices /* NULL: 0..ndev-1 */);
void vit_comm_destroy(vit_comm* c);
int vit_comm_rank(const vit_comm* c);
int vit_comm_size(const vit_comm* c);
/* host-side: wait for this rank's gathers, then meet the other ranks; afterwards the root's buffer is complete */
int vit_comm_barrier(vit_comm* c);
/* make a stream wait for the gathers this rank has issued so far */
int vit_comm_stream_wait(vit_comm* c, void* cuda_stream);
/* markers 0..7: remember how far the gather stream has got / make a stream wait for that point only (output-slot rings) */
int vit_comm_mark(vit_comm* c, int k);
int vit_comm_stream_wait_mark(vit_comm* c, int k, void* cuda_stream);
/* the stream the gathers run on (a cudaStream_t) */
void* vit_comm_stream(vit_comm* c);

/* contiguous block partition of nstreams over nranks: rank r owns [first, first + count); blocks differ by <= 1 */
void vit_shard_range(size_t nstreams, int nranks, int rank, size_t* first, size_t* count);
int vit_shard_owner(size_t nstreams, int nranks, size_t stream);

/* how packed output bits reach the root */
enum {
    VIT_GATHER_NONE = 0,    /* they stay on the decoding GPU */
    VIT_GATHER_NCCL = 1,    /* grouped ncclSend/ncclRecv to the root on a side stream, overlapping the next decode */
    VIT_GATHER_COPY = 2,    /* device-to-device copy into the root's buffer (copy engines over NVLink, no SMs) */
    VIT_GATHER_DIRECT = 3   /* the decode kernel stores straight into the root's buffer over NVLink (no gather step) */
};
/* collective: a device buffer on the root's GPU addressable by every rank (root: the allocation, others: a mapping
 * through CUDA IPC / peer access).  Owned by the communicator. */
int vit_comm_shared_alloc(vit_comm* c, void** ptr, size_t bytes, int root);
/* rank p's block (sizes[p] bytes at send_d on rank p) lands at recv_base + offsets[p] on the root; every rank passes the
 * same offsets/sizes (one entry per rank).  Ordered after the work queued on producer_stream; asynchronous. */
int vit_comm_gatherv(vit_comm* c, int mode, const void* send_d, void* recv_base, const size_t* offsets,
                     const size_t* sizes, int root, void* producer_stream);

/* The sharded stream job: `nstreams` independent streams of n_bits message bits each (stream s: synthetic source with
 * seed + s), dealt to the ranks in contiguous blocks; per round every rank generates `batch` of its streams on the device
 * (untimed), decodes them `wave` streams per launch and gathers each finished wave while the next one decodes. */
typedef struct {
    int options;            /* option bitfield */
    size_t n_bits;          /* message bits per stream */
    unsigned nstreams;      /* streams of the whole job, all ranks together */
    unsigned wave;          /* streams per decode launch */
    unsigned batch;         /* streams generated ahead of each timed decode phase (device memory: batch x (in + out)) */
    unsigned seed;
    int source;             /* 0: counter hash, 1: PRBS-31 */
    int amp;                /* symbol amplitude in quantiser units, <= 0: default per input type */
    double sigma;           /* noise sd relative to amp */
    int gather;             /* VIT_GATHER_* */
    int root;
} vit_job_config;
typedef struct {
    double decode_ms;       /* this rank: device time of the decode launches, summed over the rounds */
    double job_ms;          /* this rank: ... up to the end of the round's last gather (== decode_ms without gather) */
    double synth_ms;        /* generation, not part of the job time */
    unsigned long long decoded_bits, bit_errors, max_stream_errors, launches;
    unsigned streams;       /* streams this rank decoded */
} vit_job_result;
typedef struct vit_job vit_job;
int vit_job_create(vit_job** out, vit_comm* comm /* NULL: one GPU */, int device, const vit_job_config* cfg);
int vit_job_run(vit_job* job, vit_job_result* result);
/* root: the stream-major gathered outputs (stream s at + s * out_stride), a device pointer; NULL with VIT_GATHER_NONE */
const void* vit_job_gathered(const vit_job* job, size_t* out_stride);
int vit_job_stream_range(const vit_job* job, size_t* first, size_t* count);
/* bit errors of each of this rank's streams in the last run */
int vit_job_stream_errors(const vit_job* job, unsigned long long* errs, size_t cap);
void vit_job_destroy(vit_job* job);

const char* vit_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
