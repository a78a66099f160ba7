// l1_probe.cu -- cross-compile check and SASS census input for the one-lane-per-segment geometry (csrc/vit_kernel_l1.inc),
// which the library does not instantiate yet (it has not been measured on a GPU):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Xptxas -v \
//        -I gpu-accelerated-viterbi-decoder_b200/csrc -c scripts/l1_probe.cu -o /tmp/l1_probe.o
//   python scripts/sass_loop_census.py /tmp/l1_probe.o vit_decode_kernel_l1ILi1ELi2ELi32EE
#include "vit_kernel.cuh"
namespace vitk {
template __global__ void l1::vit_decode_kernel_l1<MET_B16, IN_S8, 32>(const KParams);     // BASELINE configs[4] shape
template __global__ void l1::vit_decode_kernel_l1<MET_B16, IN_S8, 16>(const KParams);
template __global__ void l1::vit_decode_kernel_l1<MET_B16, IN_S4, 32>(const KParams);
template __global__ void l1::vit_decode_kernel_l1<MET_B16, IN_HARD, 32>(const KParams);
template __global__ void l1::vit_decode_kernel_l1<MET_F16, IN_S8, 16>(const KParams);
template __global__ void l1::vit_decode_kernel_l1<MET_B32, IN_F32, 32>(const KParams);
}  // namespace vitk
