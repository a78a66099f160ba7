// viterbi.h -- header-only C++ shim with the reference's public interface
// (`template<int options> class ViterbiCUDA`, reference src/viterbi/viterbi.h:7-152) over the C ABI
// of libvitb200.so (include/vit_b200.h).  A caller written against the reference header -- e.g.
// ViterbiDecoder<options> (reference src/viterbiDF.h:170-209) or runPipeline (src/main.cpp:119-172)
// -- compiles against this one with a host compiler only (no nvcc, no CUDA headers).
//
// Error convention kept from the reference (src/viterbi/gpuerrors.h:8-17): a failing call prints
// "<message> in <file> at line <n>" to stderr and exits with EXIT_FAILURE.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "../../include/vit_b200.h"
#include "../csrc/vit_code.h"   // VIT_CONST_LEN, VIT_POLY1, VIT_POLY2: the code parameters (defaults = the reference's)

// ---- option bitfield: values are the contract (reference viterbi.h:7-20) ----------------------
constexpr int CHANNEL_SHIFT = 0, METRIC_SHIFT = 4, DECODE_SHIFT = 8, COMP_SHIFT = 12;
constexpr int CHANNEL_MASK = 0xf << CHANNEL_SHIFT;
constexpr int METRIC_MASK = 0xf << METRIC_SHIFT;
constexpr int DECODE_MASK = 0xf << DECODE_SHIFT;
constexpr int COMP_MASK = 0xf << COMP_SHIFT;

enum ChannelIn { HARD = 0 << CHANNEL_SHIFT, SOFT4 = 1 << CHANNEL_SHIFT, SOFT8 = 2 << CHANNEL_SHIFT,
                 SOFT16 = 3 << CHANNEL_SHIFT, FP32 = 4 << CHANNEL_SHIFT };
enum Metric { M_B32 = 0 << METRIC_SHIFT, M_B16 = 1 << METRIC_SHIFT, M_FP16 = 2 << METRIC_SHIFT };
enum DecodeOut { O_B32 = 0 << DECODE_SHIFT, O_B16 = 1 << DECODE_SHIFT };
enum CompMode { REG = 0 << COMP_SHIFT, DPX = 1 << COMP_SHIFT,
                // extension (not a reference value): the tie rule the reference's DPX code paths define but never run
                // (viterbiACS.cuh:123-134; viterbi.cu:181,192,204 do not forward compMode) -- int32 core only differs
                DPX_TIES = 2 << COMP_SHIFT };

// Which combinations exist.  The reference (viterbi.h:22-36) excludes f16 x {s8,s16}, b16 x s16 and
// f16 x dpx; this library adds f16 x {s8,s16} (symbols pre-scaled to 5 bits) and treats dpx as an
// alias of reg for every core, so only b16 x s16 stays invalid.
template <int options>
struct OptionsValid {
    static constexpr int in = options & CHANNEL_MASK, met = options & METRIC_MASK,
                         out = options & DECODE_MASK, cmp = options & COMP_MASK;
    static constexpr bool known = in <= FP32 && met <= M_FP16 && out <= O_B16 && cmp <= DPX_TIES && (options >> 16) == 0;
    static constexpr bool value = known && !(met == M_B16 && in == SOFT16) && !(met == M_FP16 && cmp == DPX_TIES);
    // the reference's own table, for callers that want to stay inside it
    static constexpr bool reference_value =
        known && cmp <= DPX && !((in == SOFT8 || in == SOFT16) && met == M_FP16) && !(in == SOFT16 && met == M_B16) &&
        !(met == M_FP16 && cmp == DPX);
};

namespace vit_detail {
[[noreturn]] inline void die(const char* what, const char* file, int line) {
    std::fprintf(stderr, "%s in %s at line %d\n", what, file, line);
    std::exit(EXIT_FAILURE);
}
inline void check(int rc, const char* file, int line) {
    if (rc != VIT_OK) die(vit_last_error(), file, line);
}
}  // namespace vit_detail
#define VIT_HANDLE_ERROR(rc) (::vit_detail::check((rc), __FILE__, __LINE__))

template <int options = 0, bool enable = OptionsValid<options>::value>
class ViterbiCUDA;

// constants and types: available for every option value, as in the reference (viterbi.h:46-89)
template <int options>
struct ViterbiCUDA<options, false> {
    static constexpr ChannelIn inputType = static_cast<ChannelIn>(options & CHANNEL_MASK);
    static constexpr Metric metricType = static_cast<Metric>(options & METRIC_MASK);
    static constexpr DecodeOut outputType = static_cast<DecodeOut>(options & DECODE_MASK);
    static constexpr CompMode compMode = static_cast<CompMode>(options & COMP_MASK);

    // metric_t is informational on the host (the reference exposes __half here; a host-only build
    // has no CUDA headers, so the half2 core's storage type is shown as uint16_t)
    using metric_t = std::conditional_t<metricType == M_B16, int16_t,
                     std::conditional_t<metricType == M_B32, int32_t, uint16_t>>;
    using decPack_t = std::conditional_t<outputType == O_B16, uint16_t, uint32_t>;
    using encPack_t = std::conditional_t<inputType == FP32, float, int32_t>;

    // reference viterbi.h:61-63: 7, 0171, 0133.  A caller compiled with -DVIT_POLY1= / -DVIT_POLY2= must be linked with a
    // library built with the same values (checked when a decoder is constructed).
    static constexpr int constLen = VIT_CONST_LEN;
    static constexpr int polyn1 = VIT_POLY1;
    static constexpr int polyn2 = VIT_POLY2;
    static constexpr int roundup(int a, int b) { return a <= 0 ? 0 : (a + b - 1) / b * b; }
    static constexpr size_t roundup(size_t a, size_t b) { return a == 0 ? 0 : (a + b - 1) / b * b; }
    static constexpr int bitsPerMetric = metricType == M_B16 ? 16 : metricType == M_B32 ? 32 : 11;
    static constexpr int bitsPerPack = outputType == O_B16 ? 16 : 32;
    static constexpr int extraL_raw = 32, extraR_raw = 32, slideSize_raw = 32;
    static constexpr int extraL = roundup(extraL_raw, bitsPerPack) - (constLen - 1);   // 26
    static constexpr int extraR = roundup(extraR_raw, bitsPerPack) + (constLen - 1);   // 38
    static constexpr int slideSize = roundup(slideSize_raw, bitsPerPack);              // 32
    static constexpr int forwardLen = extraL + slideSize + extraR;                     // 96
    static constexpr int bmMemWidth = 32;
    static constexpr int blockDimY = 2;
    static constexpr int FPprecision = 4;
    static constexpr int encDataPerPack = inputType == HARD ? 32 : inputType == SOFT4 ? 8 : inputType == SOFT8 ? 4
                                        : inputType == SOFT16 ? 2 : 1;
    static constexpr int encDataWidth = inputType == HARD ? 1 : inputType == SOFT4 ? 4 : inputType == SOFT8 ? 8
                                      : inputType == SOFT16 ? 16 : FPprecision;
};

template <int options>
class ViterbiCUDA<options, true> : public ViterbiCUDA<options, false> {
    using Base = ViterbiCUDA<options, false>;

public:
    using typename Base::decPack_t;
    using typename Base::encPack_t;
    using typename Base::metric_t;

    // reference viterbi.cu:23-36.  `device`: the reference always uses 0 (viterbi.cu:134).
    ViterbiCUDA() { checkCode(); VIT_HANDLE_ERROR(vit_create(&h_, options, 0, 0)); }
    explicit ViterbiCUDA(size_t inputNum, int device = 0) { checkCode(); VIT_HANDLE_ERROR(vit_create(&h_, options, device, inputNum)); }
    ~ViterbiCUDA() { vit_destroy(h_); }
    ViterbiCUDA(const ViterbiCUDA&) = delete;
    ViterbiCUDA& operator=(const ViterbiCUDA&) = delete;

    // reference viterbi.cu:210-238.  The third argument is the number of CODED SYMBOLS (the
    // reference header calls it messageLen, viterbi.h:130, but passes symbols, viterbiDF.h:190-193).
    void run(encPack_t* input_h, decPack_t* output_h, size_t inputNum, float* kernelTime = nullptr) {
        VIT_HANDLE_ERROR(vit_run(h_, input_h, output_h, inputNum, kernelTime));
    }
    // device-resident variants (new)
    void runDevice(const void* in_d, void* out_d, size_t inputNum, void* stream = nullptr, float* kernelTime = nullptr) {
        VIT_HANDLE_ERROR(vit_run_device(h_, in_d, out_d, inputNum, stream, kernelTime));
    }
    void runDeviceBatch(const void* in_d, void* out_d, size_t inputNum, size_t nstreams, size_t inStride,
                        size_t outStride, void* stream = nullptr, float* kernelTime = nullptr) {
        VIT_HANDLE_ERROR(vit_run_device_batch(h_, in_d, out_d, inputNum, nstreams, inStride, outStride, stream, kernelTime));
    }

    // chunked decode of an endless stream (new: vit_stream_* in include/vit_b200.h); returns the bytes written to output_h
    void streamReset() { VIT_HANDLE_ERROR(vit_stream_reset(h_)); }
    size_t streamPush(const encPack_t* input_h, size_t inputNum, decPack_t* output_h, size_t outputCapacityBytes) {
        size_t written = 0;
        VIT_HANDLE_ERROR(vit_stream_push(h_, input_h, inputNum, output_h, outputCapacityBytes, &written));
        return written;
    }
    size_t streamPending() const { return vit_stream_pending(h_); }

    size_t getInputSize(size_t inputNum) { return vit_input_size(options, inputNum); }      // viterbi.cu:63-84
    size_t getMessageLen(size_t inputNum) { return vit_message_len(options, inputNum); }    // viterbi.cu:86-88
    size_t getOutputSize(size_t inputNum) { return vit_output_size(options, inputNum); }    // viterbi.cu:90-92

private:
    // the library decodes the code it was compiled for: it must be the one this header's constants describe
    static void checkCode() {
        int k = 0, p1 = 0, p2 = 0;
        vit_code_parameters(&k, &p1, &p2);
        if (k != Base::constLen || p1 != Base::polyn1 || p2 != Base::polyn2)
            vit_detail::die("libvitb200 was built for other code parameters (constLen / polyn1 / polyn2) than this header", __FILE__, __LINE__);
    }
    vit_handle* h_ = nullptr;
};
