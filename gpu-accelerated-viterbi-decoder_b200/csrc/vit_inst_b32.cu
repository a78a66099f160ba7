#define VIT_INST_MET MET_B32
#define VIT_INST_FN kernel_entry_b32
#define VIT_INST_HAS_S16 1
#define VIT_INST_HAS_L1 0
#include "vit_inst.inc"
