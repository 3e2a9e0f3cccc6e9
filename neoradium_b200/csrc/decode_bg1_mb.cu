// Statically scheduled fp32 decoder kernels, BG1, several code blocks per CTA (decode_inst.cuh).
#define NR_INST_NAME nr_launch_static_bg1_mb
#define NR_INST_BG 1
#define NR_INST_ES 0
#define NR_INST_MB 1
#include "decode_inst.cuh"
