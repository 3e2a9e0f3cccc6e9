// Statically scheduled fp32 decoder kernels, BG1, without the early-termination code (decode_inst.cuh).
#define NR_INST_NAME nr_launch_static_bg1
#define NR_INST_BG 1
#define NR_INST_ES 0
#include "decode_inst.cuh"
