// Statically scheduled fp32 decoder kernels, BG1, with the early-termination code (decode_inst.cuh).
#define NR_INST_NAME nr_launch_static_bg1_es
#define NR_INST_BG 1
#define NR_INST_ES 1
#include "decode_inst.cuh"
