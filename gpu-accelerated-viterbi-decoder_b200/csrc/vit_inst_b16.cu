// int16x2 core; SOFT16 input is rejected as in the reference (viterbi.h:28-29)
#define VIT_INST_MET MET_B16
#define VIT_INST_FN kernel_entry_b16
#define VIT_INST_HAS_S16 0
#define VIT_INST_HAS_L1 1
#include "vit_inst.inc"
