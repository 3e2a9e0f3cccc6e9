// One group of statically scheduled decoder kernels (see decode_launch.cuh).  The including .cu defines
//   NR_INST_NAME  launcher name      NR_INST_BG  1 | 2      NR_INST_ES  0 | 1      NR_INST_MB  (defined: several blocks per CTA)
#include "decode_kernel.cuh"
#include "decode_launch.cuh"

namespace {
template <typename K>
cudaError_t launch_one(K kern, const void* dg, const void* da, unsigned grid, int nT, size_t smem, cudaStream_t s)
{
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, nT, smem, s>>>(*reinterpret_cast<const NrDecGraph*>(dg), *reinterpret_cast<const DecArgs*>(da));
    return cudaGetLastError();
}
}   // namespace

#if defined(NR_INST_W8)
cudaError_t NR_INST_NAME(int allt, int oneCb, const void* dg, const void* da, unsigned grid, int nT, size_t smem, cudaStream_t s)
{
    constexpr int BG = NR_INST_BG;
    if (oneCb && allt == 1) return launch_one(nr_decode_kernel<float, true, BG, 1, 0, 0, 2>, dg, da, grid, nT, smem, s);
    if (oneCb && allt == 2) return launch_one(nr_decode_kernel<float, true, BG, 2, 0, 0, 2>, dg, da, grid, nT, smem, s);
    if (!oneCb && allt == 1) return launch_one(nr_decode_kernel<float, false, BG, 1, 0, 0, 2>, dg, da, grid, nT, smem, s);
    if (!oneCb && allt == 2) return launch_one(nr_decode_kernel<float, false, BG, 2, 0, 0, 2>, dg, da, grid, nT, smem, s);
    if (oneCb && allt == 0) return launch_one(nr_decode_kernel<float, true, BG, 0, 0, 0, 2>, dg, da, grid, nT, smem, s);    // tiered state
    if (!oneCb && allt == 0) return launch_one(nr_decode_kernel<float, false, BG, 0, 0, 0, 2>, dg, da, grid, nT, smem, s);
    return cudaErrorNotSupported;
}
#else
cudaError_t NR_INST_NAME(int allt, int esm, int zs, const void* dg, const void* da, unsigned grid, int nT, size_t smem,
                         cudaStream_t s)
{
    constexpr int BG = NR_INST_BG;
#if defined(NR_INST_MB)
    if (esm != 0 || zs != 0) return cudaErrorNotSupported;
    if (allt == 1) return launch_one(nr_decode_kernel<float, false, BG, 1, 0, 0>, dg, da, grid, nT, smem, s);
    if (allt == 2) return launch_one(nr_decode_kernel<float, false, BG, 2, 0, 0>, dg, da, grid, nT, smem, s);
    if (allt == 0) return launch_one(nr_decode_kernel<float, false, BG, 0, 0, 0>, dg, da, grid, nT, smem, s);   // tiered state (low rates)
#elif NR_INST_ES
    if (esm != 1) return cudaErrorNotSupported;
    if (allt == 1 && zs == 384) return launch_one(nr_decode_kernel<float, true, BG, 1, 1, 384>, dg, da, grid, nT, smem, s);
    if (zs != 0) return cudaErrorNotSupported;
    if (allt == 2) return launch_one(nr_decode_kernel<float, true, BG, 2, 1, 0>, dg, da, grid, nT, smem, s);
    if (allt == 1) return launch_one(nr_decode_kernel<float, true, BG, 1, 1, 0>, dg, da, grid, nT, smem, s);
    if (allt == 0) return launch_one(nr_decode_kernel<float, true, BG, 0, 1, 0>, dg, da, grid, nT, smem, s);
#else
    if (esm != 0) return cudaErrorNotSupported;
    if (allt == 1 && zs == 384) return launch_one(nr_decode_kernel<float, true, BG, 1, 0, 384>, dg, da, grid, nT, smem, s);
    if (allt == 2 && zs == 384) return launch_one(nr_decode_kernel<float, true, BG, 2, 0, 384>, dg, da, grid, nT, smem, s);
    if (allt == 2 && zs == 0) return launch_one(nr_decode_kernel<float, true, BG, 2, 0, 0>, dg, da, grid, nT, smem, s);
    if (allt == 1 && zs == 0) return launch_one(nr_decode_kernel<float, true, BG, 1, 0, 0>, dg, da, grid, nT, smem, s);
    if (allt == 0 && zs == 0) return launch_one(nr_decode_kernel<float, true, BG, 0, 0, 0>, dg, da, grid, nT, smem, s);   // tiered state (low rates)
#endif
    return cudaErrorNotSupported;
}
#endif
