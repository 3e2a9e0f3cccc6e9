// Statically scheduled fp32 decoder kernels, BG2, with the early-termination code (decode_inst.cuh).
#define NR_INST_NAME nr_launch_static_bg2_es
#define NR_INST_BG 2
#define NR_INST_ES 1
#include "decode_inst.cuh"
