// Statically scheduled fp32 decoder kernels, BG1, CTAs of at most 8 warps, three per SM (decode_inst.cuh).
#define NR_INST_NAME nr_launch_static_bg1_w8
#define NR_INST_BG 1
#define NR_INST_ES 0
#define NR_INST_W8 1
#include "decode_inst.cuh"
