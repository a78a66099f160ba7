/*
 * ref_shim.cu -- C entry points around the UNMODIFIED reference decoder (TEST INFRASTRUCTURE ONLY).
 *
 * Compiled by oracle/Makefile together with /root/reference/src/viterbi/viterbi.cu (in place, never
 * copied) into oracle/_ref/libvitref.so.  Used to (1) pin the C golden model against the reference's
 * own CUDA decoder on a B200, (2) generate tests/golden/ vectors, (3) time the reference in
 * `bench.py --impl reference`.  The reference class is instantiated for every option combination by
 * viterbi.cu:240-262; we only dispatch a runtime `options` value to ViterbiCUDA<options>::run
 * (viterbi.cu:210-238).
 */
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>
#include "viterbi.h"

namespace {

template <int O>
int run_one(const void* in, void* out, size_t inputNum, float* ms) {
    if constexpr (OptionsValid<O>::value) {
        ViterbiCUDA<O> dec;
        dec.run((typename ViterbiCUDA<O>::encPack_t*)in, (typename ViterbiCUDA<O>::decPack_t*)out, inputNum, ms);
        return 0;
    } else {
        return -1;
    }
}

template <int O>
int sizes_one(size_t inputNum, size_t* v) {
    if constexpr (OptionsValid<O>::value) {
        ViterbiCUDA<O> dec;
        v[0] = dec.getInputSize(inputNum);
        v[1] = dec.getMessageLen(inputNum);
        v[2] = dec.getOutputSize(inputNum);
        return 0;
    } else {
        return -1;
    }
}

#define REF_CASE(o) case (o): return FN<(o)>(ARGS);
#define REF_COMP(o) REF_CASE((o) | CompMode::REG) REF_CASE((o) | CompMode::DPX)
#define REF_OUT(o) REF_COMP((o) | DecodeOut::O_B32) REF_COMP((o) | DecodeOut::O_B16)
#define REF_MET(o) REF_OUT((o) | Metric::M_B32) REF_OUT((o) | Metric::M_B16) REF_OUT((o) | Metric::M_FP16)
#define REF_ALL REF_MET(ChannelIn::HARD) REF_MET(ChannelIn::SOFT4) REF_MET(ChannelIn::SOFT8) \
                REF_MET(ChannelIn::SOFT16) REF_MET(ChannelIn::FP32)

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int ref_run(int options, const void* in_h, void* out_h, size_t inputNum, float* kernel_ms) {
#define FN run_one
#define ARGS in_h, out_h, inputNum, kernel_ms
    switch (options) { REF_ALL default: return -1; }
#undef FN
#undef ARGS
}

int ref_sizes(int options, size_t inputNum, size_t* in_msg_out) {
#define FN sizes_one
#define ARGS inputNum, in_msg_out
    switch (options) { REF_ALL default: return -1; }
#undef FN
#undef ARGS
}

int ref_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

#pragma GCC visibility pop
}  // extern "C"
