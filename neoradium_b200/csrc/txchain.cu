// K2 + K3a: systematic QC-LDPC encoder, rate matching, and the full parity check.
// Replaces LdpcEncoder.encode (neoradium/ldpc.py:1057-1090), LdpcEncoder.rateMatch (ldpc.py:1128-1159) and a correct
// LdpcBase.isValidCodedBlock (ldpc.py:825-843).  All three are HBM-bound byte kernels: coalesced int8 loads/stores
// along the lifted index, the code block staged in shared memory, circulant shifts as index arithmetic.
#include "nrldpc_internal.cuh"

namespace {

constexpr int TX_THREADS = 384;

__device__ __forceinline__ int wrapZ(int p, int Z) { return (p >= Z) ? p - Z : p; }

// XOR over the edges [e0, e1) of bits s[col*Z + (m + shift) % Z]
__device__ __forceinline__ uint32_t row_xor(const NrGraph& g, int e0, int e1, const unsigned char* s, int m, int Z, int maxCol)
{
    uint32_t acc = 0;
    for (int e = e0; e < e1; e++) {
        const uint32_t ew = g.edge[e];
        const int col = (int)(ew >> 16);
        if (col >= maxCol) break;   // columns ascend inside a row
        acc ^= s[col * Z + wrapZ(m + (int)(ew & 0xffffu), Z)];
    }
    return acc;
}

__device__ __forceinline__ int edge_shift(const NrGraph& g, int row, int col)
{
    for (int e = g.rowEdge0[row]; e < g.rowEdge0[row + 1]; e++)
        if ((int)(g.edge[e] >> 16) == col) return (int)(g.edge[e] & 0xffffu);
    return -1;
}

// one thread per lifted position; a CTA hosts floor(384/Z) code blocks
__global__ void __launch_bounds__(TX_THREADS)
    nr_encode_kernel(const __grid_constant__ NrGraph g, const signed char* in, long long numCb, int cbPerCta,
                     signed char* out, int puncture)
{
    extern __shared__ unsigned char sm[];
    const int Z = g.Z, k = g.ksys, ncore = g.ncore;
    const int tid = threadIdx.x;
    const int cbl = tid / Z, m = tid - cbl * Z;
    unsigned char* s = sm + (size_t)cbl * (ncore + 1) * Z;   // core columns + one scratch column
    const int firstCol = puncture ? 2 : 0;
    const long long outLen = (long long)(g.ncols - firstCol) * Z;
    const long long numGroups = (numCb + cbPerCta - 1) / cbPerCta;
    for (long long grp = blockIdx.x; grp < numGroups; grp += gridDim.x) {
        const long long cb = grp * cbPerCta + cbl;
        const bool active = cbl < cbPerCta && cb < numCb;
        if (active)
            for (int col = 0; col < k; col++) s[col * Z + m] = (unsigned char)(in[cb * (long long)k * Z + col * Z + m] & 1);
        __syncthreads();
        uint32_t lam[4] = {0, 0, 0, 0};
        if (active) {
            for (int i = 0; i < 4; i++) lam[i] = row_xor(g, g.rowEdge0[i], g.rowEdge0[i + 1], s, m, Z, k);
            s[ncore * Z + m] = (unsigned char)(lam[0] ^ lam[1] ^ lam[2] ^ lam[3]);
        }
        __syncthreads();
        if (active) {
            // p0 = rot(sum, Z - b), b = shift of column k in row 1, or in row 2 when row 1 has none (ldpc.py:1068)
            int b = edge_shift(g, 1, k);
            if (b < 0) b = edge_shift(g, 2, k);
            s[k * Z + m] = s[ncore * Z + wrapZ(m + Z - b, Z)];
        }
        __syncthreads();
        for (int i = 0; i < 3; i++) {   // p1..p3 through the double diagonal (ldpc.py:1077-1080)
            if (active) {
                uint32_t acc = lam[i];
                for (int e = g.rowEdge0[i]; e < g.rowEdge0[i + 1]; e++) {
                    const uint32_t ew = g.edge[e];
                    const int col = (int)(ew >> 16);
                    if (col >= k && col <= k + i) acc ^= s[col * Z + wrapZ(m + (int)(ew & 0xffffu), Z)];
                }
                s[(k + i + 1) * Z + m] = (unsigned char)acc;
            }
            __syncthreads();
        }
        if (active) {
            signed char* o = out + cb * outLen;
            for (int col = firstCol; col < ncore; col++) o[(long long)(col - firstCol) * Z + m] = (signed char)s[col * Z + m];
            for (int r = 4; r < g.P; r++) {   // extension rows (ldpc.py:1083-1084): XOR over the core columns only
                const uint32_t p = row_xor(g, g.rowEdge0[r], g.rowEdge0[r + 1], s, m, Z, ncore);
                o[(long long)(k + r - firstCol) * Z + m] = (signed char)p;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(TX_THREADS)
    nr_parity_kernel(const __grid_constant__ NrGraph g, const signed char* coded, long long numCb, int cbPerCta,
                     unsigned char* ok)
{
    extern __shared__ unsigned char sm[];
    __shared__ int bad[TX_THREADS];
    const int Z = g.Z;
    const int tid = threadIdx.x;
    const int cbl = tid / Z, m = tid - cbl * Z;
    unsigned char* s = sm + (size_t)cbl * g.ncols * Z;
    const long long numGroups = (numCb + cbPerCta - 1) / cbPerCta;
    for (long long grp = blockIdx.x; grp < numGroups; grp += gridDim.x) {
        const long long cb = grp * cbPerCta + cbl;
        const bool active = cbl < cbPerCta && cb < numCb;
        if (tid < cbPerCta) bad[tid] = 0;
        if (active)
            for (int col = 0; col < g.ncols; col++) s[col * Z + m] = (unsigned char)(coded[cb * (long long)g.ncols * Z + col * Z + m] & 1);
        __syncthreads();
        if (active) {
            uint32_t any = 0;
            for (int r = 0; r < g.P; r++) any |= row_xor(g, g.rowEdge0[r], g.rowEdge0[r + 1], s, m, Z, g.ncols);
            if (any) bad[cbl] = 1;
        }
        __syncthreads();
        if (active && m == 0) ok[cb] = bad[cbl] ? 0 : 1;
        __syncthreads();
    }
}

// rateMatch: one CTA per code block; thread per OUTPUT bit (coalesced stores), gather from the coded block
__global__ void __launch_bounds__(256)
    nr_rate_match_kernel(const signed char* coded, long long numCb, int C, int N, int K, int F, int Z, int ncb, int k0,
                         int qm, int E0, int nShort, int fStep, signed char* out, long long outStride)
{
    const int L = ncb - F;
    const int sysLen = K - 2 * Z - F;
    for (long long cb = blockIdx.x; cb < numCb; cb += gridDim.x) {
        const long long tb = cb / C;
        const int r = (int)(cb - tb * C);
        const int E = E0 + (r >= nShort ? fStep : 0);
        const long long off = (long long)r * E0 + (long long)(r > nShort ? r - nShort : 0) * fStep;
        const int Eq = E / qm;
        const signed char* src = coded + cb * (long long)N;
        signed char* dst = out + tb * outStride + off;
        for (int gI = threadIdx.x; gI < E; gI += blockDim.x) {
            const int sIdx = gI / qm, b = gI - sIdx * qm;   // interleaver: out[s*qm + b] = e[b*Eq + s]  (ldpc.py:1155)
            const int i = b * Eq + sIdx;
            const int q = (k0 + i) % L;                      // circular buffer WITHOUT fillers (ldpc.py:1139-1142)
            const int n = (q < sysLen) ? q : q + F;
            dst[gI] = src[n];
        }
    }
}

void tb_split(const nrldpc_tb_config* c, int N, int* E0, int* nShort, int* fStep, int* k0)
{
    const long long f = (long long)c->nl * c->qm;
    const long long gBase = (c->G + f - 1) / f;
    *fStep = (int)f;
    *E0 = (int)((gBase / c->C) * f);
    *nShort = (int)(c->C - gBase % c->C);
    static const int k0n1[4] = {0, 17, 33, 56}, k0n2[4] = {0, 13, 25, 43};
    const int num = (c->bg == 1 ? k0n1 : k0n2)[c->rv];
    *k0 = (int)(((long long)num * c->ncb / N) * c->zc);
}

}   // namespace

int nr_tb_split(const nrldpc_tb_config* c, int N, int* E0, int* nShort, int* fStep, int* k0)
{
    tb_split(c, N, E0, nShort, fStep, k0);
    return 0;
}

int nr_check_tb_config(const nrldpc_tb_config* cfg, const NrGraph& g, const char* who)
{
    const int Z = cfg->zc, N = (g.ncols - 2) * Z;
    if (cfg->rv < 0 || cfg->rv > 3) { nr_set_error("Invalid 'rv' value! It must be one of 0, 1, 2, or 3."); return NRLDPC_ERR_ARG; }
    if (cfg->C < 1 || cfg->K != g.ksys * Z || cfg->F < 0 || cfg->F >= cfg->K - 2 * Z || cfg->ncb > N ||
        cfg->ncb <= cfg->K - 2 * Z || cfg->qm < 1 || cfg->nl < 1 || cfg->G <= 0) {
        nr_set_error("%s: bad transport-block configuration", who);
        return NRLDPC_ERR_ARG;
    }
    return NRLDPC_OK;
}

extern "C" int nrldpc_encode(nrldpc_handle* h, int bg, int zc, const int8_t* code_blocks, int64_t num_cb,
                             int8_t* coded, int puncture, nrldpc_stream stream)
{
    if (!h) { nr_set_error("encode: null handle"); return NRLDPC_ERR_ARG; }
    NrGraph g;
    if (nr_build_graph(bg, zc, &g)) return NRLDPC_ERR_ARG;
    if (num_cb <= 0) { nr_set_error("encode: bad shape"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    int cbPerCta = max(1, TX_THREADS / zc);
    if (cbPerCta > num_cb) cbPerCta = (int)num_cb;
    const int nT = (cbPerCta * zc + 31) & ~31;
    const size_t smem = (size_t)cbPerCta * (g.ncore + 1) * zc;
    const long long groups = (num_cb + cbPerCta - 1) / cbPerCta;
    const int grid = (int)min(groups, (long long)h->numSMs * 4);
    nr_encode_kernel<<<grid, nT, smem, (cudaStream_t)stream>>>(g, (const signed char*)code_blocks, num_cb, cbPerCta,
                                                               (signed char*)coded, puncture);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_parity_check(nrldpc_handle* h, int bg, int zc, const int8_t* coded_full, int64_t num_cb,
                                   uint8_t* ok, nrldpc_stream stream)
{
    if (!h) { nr_set_error("parity_check: null handle"); return NRLDPC_ERR_ARG; }
    NrGraph g;
    if (nr_build_graph(bg, zc, &g)) return NRLDPC_ERR_ARG;
    if (num_cb <= 0) { nr_set_error("parity_check: bad shape"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    int cbPerCta = max(1, TX_THREADS / zc);
    if (cbPerCta > num_cb) cbPerCta = (int)num_cb;
    const int nT = (cbPerCta * zc + 31) & ~31;
    const size_t smem = (size_t)cbPerCta * g.ncols * zc;
    const long long groups = (num_cb + cbPerCta - 1) / cbPerCta;
    const int grid = (int)min(groups, (long long)h->numSMs * 4);
    nr_parity_kernel<<<grid, nT, smem, (cudaStream_t)stream>>>(g, (const signed char*)coded_full, num_cb, cbPerCta, ok);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_rate_match(nrldpc_handle* h, const nrldpc_tb_config* cfg, const int8_t* coded, int64_t num_tb,
                                 int8_t* out, int64_t out_stride, nrldpc_stream stream)
{
    if (!h || !cfg) { nr_set_error("rate_match: null argument"); return NRLDPC_ERR_ARG; }
    NrGraph g;
    if (nr_build_graph(cfg->bg, cfg->zc, &g)) return NRLDPC_ERR_ARG;
    int rc = nr_check_tb_config(cfg, g, "rate_match");
    if (rc) return rc;
    if (num_tb <= 0) { nr_set_error("rate_match: bad shape"); return NRLDPC_ERR_ARG; }
    const int N = (g.ncols - 2) * cfg->zc;
    int E0, nShort, fStep, k0;
    tb_split(cfg, N, &E0, &nShort, &fStep, &k0);
    if (E0 % cfg->qm) { nr_set_error("rate_match: E not a multiple of qm"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const long long numCb = num_tb * cfg->C;
    const int grid = (int)min(numCb, (long long)h->numSMs * 8);
    nr_rate_match_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const signed char*)coded, numCb, cfg->C, N, cfg->K,
                                                                 cfg->F, cfg->zc, cfg->ncb, k0, cfg->qm, E0, nShort, fStep,
                                                                 (signed char*)out, out_stride);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}
