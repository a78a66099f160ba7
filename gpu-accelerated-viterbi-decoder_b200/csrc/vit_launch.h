// vit_launch.h -- internal: one launcher per (metric core, input type, output pack) kernel.
#pragma once
#include <cuda_runtime.h>
#include "vit_kernel.cuh"

namespace vitk {
struct KernelEntry {
    // decodes segments [kp.seg_first, kp.seg_limit) of kp.nstreams streams; picks the kernel build and the grid itself
    cudaError_t (*launch)(const KParams& kp, cudaStream_t stream);
    const void* func;
    int smem_bytes;
    // the same launch with the one-lane-per-segment geometry (vit_kernel_l1.inc), or nullptr where it is not built; ignores
    // upload gates and staged output stores (the caller only uses it where neither is needed)
    cudaError_t (*launch_l1)(const KParams& kp, cudaStream_t stream);
};
// met: MET_*, in: IN_*, bpp16: 0/1.  Returns nullptr for combinations that are not built.
const KernelEntry* kernel_entry(int met, int in, int bpp16);

// defined by the instantiation units
const KernelEntry* kernel_entry_b32(int in, int bpp16);
const KernelEntry* kernel_entry_b32d(int in, int bpp16);
const KernelEntry* kernel_entry_b16(int in, int bpp16);
const KernelEntry* kernel_entry_f16(int in, int bpp16);
}  // namespace vitk
