// Launchers of the statically scheduled fp32 decoder kernels.  The instantiations live in decode_bg1.cu, decode_bg1_es.cu,
// decode_bg2.cu and decode_bg2_es.cu (decode_inst.cuh) so that ptxas compiles the ~90 KB row schedules in parallel;
// decode.cu holds the launch policy.  NrDecGraph / DecArgs are file-local types, hence the untyped pointers.
#pragma once
#include <cuda_runtime.h>

// allt: 0 tiered state (TMEM, shared planes, L2 scratch), 1 all rows in Tensor Memory, 2 split (21 rows TMEM + planes)
// esm: early-termination code compiled in;  zs: compile-time lifting size (0 = run-time table)
// cudaErrorNotSupported: the combination is not instantiated
cudaError_t nr_launch_static_bg1(int allt, int esm, int zs, const void* decGraph, const void* decArgs, unsigned grid, int nT,
                                 size_t smem, cudaStream_t s);
cudaError_t nr_launch_static_bg1_es(int allt, int esm, int zs, const void* decGraph, const void* decArgs, unsigned grid, int nT,
                                    size_t smem, cudaStream_t s);
cudaError_t nr_launch_static_bg2(int allt, int esm, int zs, const void* decGraph, const void* decArgs, unsigned grid, int nT,
                                 size_t smem, cudaStream_t s);
cudaError_t nr_launch_static_bg2_es(int allt, int esm, int zs, const void* decGraph, const void* decArgs, unsigned grid, int nT,
                                    size_t smem, cudaStream_t s);
// several code blocks per CTA (lifting sizes below 224 and those that are no multiple of 32): allt 0 | 1 | 2, esm 0, zs 0 only
cudaError_t nr_launch_static_bg1_mb(int allt, int esm, int zs, const void* decGraph, const void* decArgs, unsigned grid, int nT,
                                    size_t smem, cudaStream_t s);
cudaError_t nr_launch_static_bg2_mb(int allt, int esm, int zs, const void* decGraph, const void* decArgs, unsigned grid, int nT,
                                    size_t smem, cudaStream_t s);
// CTAs of at most 8 warps (Zc <= 256), three per SM, 128 Tensor-Memory columns each: allt 1 (<= 16 rows) | 2 (16 rows in TMEM, the
// rest in shared planes); esm 0, zs 0; `oneCb` selects the one-block-per-CTA or the multi-block instantiation
cudaError_t nr_launch_static_bg1_w8(int allt, int oneCb, const void* decGraph, const void* decArgs, unsigned grid, int nT,
                                    size_t smem, cudaStream_t s);
cudaError_t nr_launch_static_bg2_w8(int allt, int oneCb, const void* decGraph, const void* decArgs, unsigned grid, int nT,
                                    size_t smem, cudaStream_t s);
