/*
 * ref_dpx_shim.cu -- the reference decoder with its DPX code paths actually instantiated (TEST INFRASTRUCTURE ONLY).
 *
 * The reference's kernel never forwards its compMode template argument to forwardACS (viterbi.cu:181,192,204), so
 * `-c dpx` runs the REG code and the __viaddmax-based selfPM/pairPM variants (viterbiACS.cuh:101-110,123-134,205-213,
 * 224-236) are dead.  This translation unit compiles the UNMODIFIED reference sources (included in place, never copied)
 * with one macro that makes those three calls pass the kernel's own compMode on, so that the tie rule of the DPX
 * variants can be observed on a GPU: it pins the oracle's VO_DPX_TIES table and this repo's CompMode value 2.
 * Built by oracle/Makefile into oracle/_ref/libvitref_dpx.so (hidden visibility: it defines the same class templates as
 * libvitref.so and must not share them).
 */
#include <cstddef>
#include <cstdint>
#include <stdio.h>
#include <iostream>
#include <cuda_runtime.h>
#include "viterbi.h"
#include "viterbiConsts.h"
#include "viterbiBM.cuh"
#include "viterbiACS.cuh"

template <CompMode comp>
struct forwardACS_with {
    template <Metric metric, DecodeOut out, typename... A>
    __device__ static void call(A&&... a) { forwardACS<metric, out, comp>(static_cast<A&&>(a)...); }
};
/* inside viterbi_core `compMode` names the kernel's template argument */
#define forwardACS forwardACS_with<compMode>::template call
#include "viterbi.cu"
#undef forwardACS

#define ref_run ref_run_dpx
#define ref_sizes ref_sizes_dpx
#define ref_device_count ref_device_count_dpx
#include "ref_shim.cu"
