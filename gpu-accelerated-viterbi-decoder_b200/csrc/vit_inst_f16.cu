#define VIT_INST_MET MET_F16
#define VIT_INST_FN kernel_entry_f16
#define VIT_INST_HAS_S16 1
#define VIT_INST_HAS_L1 1
#include "vit_inst.inc"
