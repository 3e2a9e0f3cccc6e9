// K3b (standalone form): rate recovery with HARQ soft combining.  Replaces LdpcDecoder.recoverRate
// (neoradium/ldpc.py:1365-1418).  HBM-bound gather: one thread per circular-buffer position sums its wraps in ascending
// stream order (the reference's chunked `+=`, ldpc.py:1407-1410), so no atomics and a deterministic result.
// The decoder (decode.cu) contains the same gather fused into its load phase; this kernel serves the drop-in
// recoverRate() call that has to hand a [C, N] array back to the caller.
#include "nrldpc_internal.cuh"

namespace {

template <typename T>
struct Add;
template <>
struct Add<float> {
    static __device__ __forceinline__ float f(float a, float b) { return __fadd_rn(a, b); }
};
template <>
struct Add<double> {
    static __device__ __forceinline__ double f(double a, double b) { return __dadd_rn(a, b); }
};

template <typename T>
__global__ void __launch_bounds__(256)
    nr_rate_recover_kernel(const T* llr, long long numCb, long long llrLen, long long llrStride, int C, int K, int F,
                           int Z, int ncb, int k0, int qm, int E0, int nShort, int fStep, T* softBuf, T* out)
{
    const int L = ncb - F;
    const int sysLen = K - 2 * Z - F;
    for (long long cb = blockIdx.x; cb < numCb; cb += gridDim.x) {
        const long long tb = cb / C;
        const int r = (int)(cb - tb * C);
        const int E = E0 + (r >= nShort ? fStep : 0);
        const long long off = (long long)r * E0 + (long long)(r > nShort ? r - nShort : 0) * fStep;
        const int Eq = E / qm;
        const T* x = llr + tb * llrStride + off;
        const long long avail = llrLen - off;   // missing tail LLRs count as zeros (ldpc.py:1402-1403)
        T* sb = softBuf ? softBuf + cb * (long long)L : nullptr;
        T* o = out + cb * (long long)ncb;
        for (int n = threadIdx.x; n < ncb; n += blockDim.x) {
            if (n >= sysLen && n < sysLen + F) {
                o[n] = (T)1e20;   // LARGE_LLR, chancodebase.py:52
                continue;
            }
            const int q = (n < sysLen) ? n : n - F;
            T acc = sb ? sb[q] : (T)0;
            int i = q - k0;
            if (i < 0) i += L;
            for (; i < E; i += L) {
                const int s = i % Eq, b = i / Eq;   // de-interleave (ldpc.py:1405): stream index s*qm + b
                const long long xi = (long long)s * qm + b;
                acc = Add<T>::f(acc, (xi < avail) ? x[xi] : (T)0);
            }
            if (sb) sb[q] = acc;
            o[n] = acc;
        }
    }
}

}   // namespace

extern "C" int nrldpc_rate_recover(nrldpc_handle* h, const nrldpc_tb_config* cfg, int dtype, const void* llr,
                                   int64_t num_tb, int64_t llr_len, int64_t llr_stride, void* soft_buffer, void* out,
                                   nrldpc_stream stream)
{
    if (!h || !cfg) { nr_set_error("rate_recover: null argument"); return NRLDPC_ERR_ARG; }
    NrGraph g;
    if (nr_build_graph(cfg->bg, cfg->zc, &g)) return NRLDPC_ERR_ARG;
    int rc = nr_check_tb_config(cfg, g, "rate_recover");
    if (rc) return rc;
    if (num_tb <= 0 || llr_len < 0 || llr_stride < llr_len) { nr_set_error("rate_recover: bad shape"); return NRLDPC_ERR_ARG; }
    const int N = (g.ncols - 2) * cfg->zc;
    int E0, nShort, fStep, k0;
    nr_tb_split(cfg, N, &E0, &nShort, &fStep, &k0);
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const long long numCb = num_tb * cfg->C;
    const int grid = (int)min(numCb, (long long)h->numSMs * 8);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == NRLDPC_F32)
        nr_rate_recover_kernel<float><<<grid, 256, 0, s>>>((const float*)llr, numCb, llr_len, llr_stride, cfg->C, cfg->K,
                                                          cfg->F, cfg->zc, cfg->ncb, k0, cfg->qm, E0, nShort, fStep,
                                                          (float*)soft_buffer, (float*)out);
    else if (dtype == NRLDPC_F64)
        nr_rate_recover_kernel<double><<<grid, 256, 0, s>>>((const double*)llr, numCb, llr_len, llr_stride, cfg->C, cfg->K,
                                                           cfg->F, cfg->zc, cfg->ncb, k0, cfg->qm, E0, nShort, fStep,
                                                           (double*)soft_buffer, (double*)out);
    else { nr_set_error("rate_recover: bad dtype"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}
