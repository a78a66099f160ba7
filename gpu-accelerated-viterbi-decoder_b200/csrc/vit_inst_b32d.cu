// int32 core with the reference's DPX tie rule (CompMode value 2, an extension: see MET_B32D in vit_kernel.cuh)
#define VIT_INST_MET MET_B32D
#define VIT_INST_FN kernel_entry_b32d
#define VIT_INST_HAS_S16 1
#define VIT_INST_HAS_L1 0
#include "vit_inst.inc"
