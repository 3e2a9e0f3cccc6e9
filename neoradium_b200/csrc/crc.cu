// K4 (standalone form): CRC attach / check for the six TS 38.212 polynomials, code-block segmentation and
// CRC-check + merge.  Replaces ChanCodeBase.getCrc/checkCrc/appendCrc (neoradium/chancodebase.py:83-189),
// LdpcEncoder.doSegmentation (neoradium/ldpc.py:1011-1030) and LdpcDecoder.checkCrcAndMerge (ldpc.py:1610-1619).
// HBM-bound byte work: one CTA per bit stream, coalesced int8 traffic, per-thread chunk division + log-depth merge.
#include "crc_device.cuh"
#include "nrldpc_internal.cuh"

namespace {

constexpr int CRC_THREADS = 256;

struct GlobalBits {
    const signed char* p;
    __device__ __forceinline__ uint32_t operator()(long long i) const { return (uint32_t)(p[i] & 1); }
};

struct PaddedBits {   // transport block zero-extended past its end
    const signed char* p;
    long long base, B;
    __device__ __forceinline__ uint32_t operator()(long long i) const { return (base + i < B) ? (uint32_t)(p[base + i] & 1) : 0u; }
};

// mode 0: crc bits / remainder; mode 1: attach (copy + crc); mode 2: check
__global__ void __launch_bounds__(CRC_THREADS)
    nr_crc_kernel(const signed char* bits, long long numStreams, long long len, long long stride, uint32_t poly, int c,
                  int mode, signed char* crcBits, uint32_t* rem, signed char* attachOut, unsigned char* ok)
{
    __shared__ uint32_t tree[CRC_THREADS];
    for (long long sidx = blockIdx.x; sidx < numStreams; sidx += gridDim.x) {
        const signed char* src = bits + sidx * stride;
        const uint32_t r = nr_group_crc(GlobalBits{src}, len, CRC_THREADS, threadIdx.x, tree, poly, c);
        if (mode == 1) {
            signed char* dst = attachOut + sidx * (len + c);
            for (long long i = threadIdx.x; i < len; i += CRC_THREADS) dst[i] = src[i];
            if ((int)threadIdx.x < c) dst[len + threadIdx.x] = (signed char)((r >> (c - 1 - threadIdx.x)) & 1u);
        } else if (mode == 2) {
            if (threadIdx.x == 0) ok[sidx] = (r == 0);
        } else {
            if (crcBits && (int)threadIdx.x < c) crcBits[sidx * c + threadIdx.x] = (signed char)((r >> (c - 1 - threadIdx.x)) & 1u);
            if (rem && threadIdx.x == 0) rem[sidx] = r;
        }
    }
}

// doSegmentation: one CTA per code block
__global__ void __launch_bounds__(CRC_THREADS)
    nr_segment_kernel(const signed char* tb, long long numTb, long long B, long long tbStride, int C, int K, int per,
                      signed char* out)
{
    __shared__ uint32_t tree[CRC_THREADS];
    const NrCrcPoly pb = nr_crc_poly(NRLDPC_CRC24B);
    const long long numCb = numTb * C;
    for (long long cb = blockIdx.x; cb < numCb; cb += gridDim.x) {
        const long long t = cb / C;
        const int r = (int)(cb - t * C);
        const signed char* src = tb + t * tbStride;
        const long long base = (long long)r * per;
        signed char* dst = out + cb * (long long)K;
        // payload, zero pad at the END of the transport block (ldpc.py:1014-1016)
        for (int i = threadIdx.x; i < per; i += CRC_THREADS) dst[i] = (base + i < B) ? (signed char)(src[base + i] & 1) : 0;
        int filled = per;
        if (C > 1) {
            const uint32_t rem = nr_group_crc(PaddedBits{src, base, B}, per, CRC_THREADS, threadIdx.x, tree, pb.poly, pb.len);
            if (threadIdx.x < 24) dst[per + threadIdx.x] = (signed char)((rem >> (23 - threadIdx.x)) & 1u);
            filled += 24;
        }
        for (int i = filled + threadIdx.x; i < K; i += CRC_THREADS) dst[i] = 0;   // filler bits are zeros (ldpc.py:1026)
    }
}

// checkCrcAndMerge: one CTA per code block
__global__ void __launch_bounds__(CRC_THREADS)
    nr_merge_kernel(const signed char* decoded, long long numTb, int C, int K, int F, signed char* tbBits,
                    long long tbStride, unsigned char* cbOk)
{
    __shared__ uint32_t tree[CRC_THREADS];
    const int Lk = K - F;
    const int per = (C > 1) ? Lk - 24 : Lk;
    const NrCrcPoly p = nr_crc_poly(C > 1 ? NRLDPC_CRC24B : NRLDPC_CRC24A);
    const long long numCb = numTb * C;
    for (long long cb = blockIdx.x; cb < numCb; cb += gridDim.x) {
        const signed char* src = decoded + cb * (long long)K;
        const uint32_t rem = nr_group_crc(GlobalBits{src}, Lk, CRC_THREADS, threadIdx.x, tree, p.poly, p.len);
        if (cbOk && threadIdx.x == 0) cbOk[cb] = (rem == 0);
        if (tbBits) {
            const long long t = cb / C;
            const int r = (int)(cb - t * C);
            signed char* dst = tbBits + t * tbStride + (long long)r * per;
            for (int i = threadIdx.x; i < per; i += CRC_THREADS) dst[i] = src[i];
        }
    }
}

__global__ void nr_counters_kernel(long long numTb, int C, const unsigned char* cbOk, const unsigned char* tbOk,
                                   const int* iters, const signed char* tbBits, const signed char* refBits,
                                   long long bitsPerTb, long long bitsStride, unsigned long long* counters)
{
    unsigned long long cbFail = 0, tbFail = 0, bitErr = 0, itSum = 0;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gsz = (long long)gridDim.x * blockDim.x;
    const long long numCb = numTb * C;
    for (long long i = gtid; i < numCb; i += gsz) {
        if (cbOk) cbFail += cbOk[i] ? 0 : 1;
        if (iters) itSum += (unsigned long long)iters[i];
    }
    if (tbOk)
        for (long long i = gtid; i < numTb; i += gsz) tbFail += tbOk[i] ? 0 : 1;
    if (tbBits && refBits) {
        const long long total = numTb * bitsPerTb;
        for (long long i = gtid; i < total; i += gsz) {
            const long long t = i / bitsPerTb, j = i - t * bitsPerTb;
            bitErr += ((tbBits[t * bitsStride + j] ^ refBits[t * bitsStride + j]) & 1) ? 1 : 0;
        }
    }
    // warp reduce then one atomic per warp
    for (int o = 16; o; o >>= 1) {
        cbFail += __shfl_xor_sync(0xffffffffu, cbFail, o);
        tbFail += __shfl_xor_sync(0xffffffffu, tbFail, o);
        bitErr += __shfl_xor_sync(0xffffffffu, bitErr, o);
        itSum += __shfl_xor_sync(0xffffffffu, itSum, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (cbFail) atomicAdd(&counters[1], cbFail);
        if (tbFail) atomicAdd(&counters[3], tbFail);
        if (bitErr) atomicAdd(&counters[4], bitErr);
        if (itSum) atomicAdd(&counters[5], itSum);
    }
    if (gtid == 0) {
        atomicAdd(&counters[0], (unsigned long long)numCb);
        atomicAdd(&counters[2], (unsigned long long)numTb);
    }
}

int crc_common(nrldpc_handle* h, const int8_t* bits, int64_t n, int64_t len, int64_t stride, int poly, int mode,
               int8_t* crcBits, uint32_t* rem, int8_t* attachOut, uint8_t* ok, nrldpc_stream stream)
{
    if (!h) { nr_set_error("crc: null handle"); return NRLDPC_ERR_ARG; }
    if (poly < 0 || poly > 5) { nr_set_error("crc: unknown polynomial id %d", poly); return NRLDPC_ERR_ARG; }
    if (n <= 0 || len < 0 || stride < len) { nr_set_error("crc: bad shape"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const NrCrcPoly p = nr_crc_poly(poly);
    const int grid = (int)min((long long)n, (long long)h->numSMs * 8);
    nr_crc_kernel<<<grid, CRC_THREADS, 0, (cudaStream_t)stream>>>((const signed char*)bits, n, len, stride, p.poly, p.len,
                                                                  mode, (signed char*)crcBits, rem,
                                                                  (signed char*)attachOut, ok);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

}   // namespace

extern "C" int nrldpc_crc(nrldpc_handle* h, const int8_t* bits, int64_t num_streams, int64_t len, int64_t stride,
                          int poly, int8_t* crc_bits, uint32_t* rem, nrldpc_stream stream)
{
    return crc_common(h, bits, num_streams, len, stride, poly, 0, crc_bits, rem, nullptr, nullptr, stream);
}

extern "C" int nrldpc_crc_attach(nrldpc_handle* h, const int8_t* bits, int64_t num_streams, int64_t len,
                                 int64_t stride, int poly, int8_t* out, nrldpc_stream stream)
{
    return crc_common(h, bits, num_streams, len, stride, poly, 1, nullptr, nullptr, out, nullptr, stream);
}

extern "C" int nrldpc_crc_check(nrldpc_handle* h, const int8_t* bits, int64_t num_streams, int64_t len,
                                int64_t stride, int poly, uint8_t* ok, nrldpc_stream stream)
{
    return crc_common(h, bits, num_streams, len, stride, poly, 2, nullptr, nullptr, nullptr, ok, stream);
}

extern "C" int nrldpc_segment(nrldpc_handle* h, const nrldpc_tb_config* cfg, const int8_t* tb, int64_t num_tb,
                              int64_t B, int64_t tb_stride, int8_t* code_blocks, nrldpc_stream stream)
{
    if (!h || !cfg) { nr_set_error("segment: null argument"); return NRLDPC_ERR_ARG; }
    const int C = cfg->C, K = cfg->K;
    const int64_t per = (B + C - 1) / C;
    if (C < 1 || num_tb <= 0 || B <= 0 || tb_stride < B || per + (C > 1 ? 24 : 0) + cfg->F != K) {
        nr_set_error("segment: inconsistent configuration (B=%lld C=%d K=%d F=%d)", (long long)B, C, K, cfg->F);
        return NRLDPC_ERR_ARG;
    }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const int grid = (int)min((long long)num_tb * C, (long long)h->numSMs * 8);
    nr_segment_kernel<<<grid, CRC_THREADS, 0, (cudaStream_t)stream>>>((const signed char*)tb, num_tb, B, tb_stride, C, K,
                                                                      (int)per, (signed char*)code_blocks);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_check_crc_and_merge(nrldpc_handle* h, const nrldpc_tb_config* cfg, const int8_t* decoded,
                                          int64_t num_tb, int8_t* tb_bits, int64_t tb_bits_stride, uint8_t* cb_crc_ok,
                                          nrldpc_stream stream)
{
    if (!h || !cfg) { nr_set_error("merge: null argument"); return NRLDPC_ERR_ARG; }
    if (cfg->C < 1 || num_tb <= 0 || cfg->K - cfg->F <= (cfg->C > 1 ? 24 : 0)) {
        nr_set_error("merge: bad configuration");
        return NRLDPC_ERR_ARG;
    }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const int grid = (int)min((long long)num_tb * cfg->C, (long long)h->numSMs * 8);
    nr_merge_kernel<<<grid, CRC_THREADS, 0, (cudaStream_t)stream>>>((const signed char*)decoded, num_tb, cfg->C, cfg->K,
                                                                    cfg->F, (signed char*)tb_bits, tb_bits_stride,
                                                                    cb_crc_ok);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_accumulate_counters(nrldpc_handle* h, int64_t num_tb, int C, const uint8_t* cb_crc_ok,
                                          const uint8_t* tb_crc_ok, const int32_t* iters, const int8_t* tb_bits,
                                          const int8_t* ref_bits, int64_t bits_per_tb, int64_t bits_stride,
                                          int64_t* counters, nrldpc_stream stream)
{
    if (!h || !counters || num_tb <= 0 || C < 1) { nr_set_error("counters: bad argument"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const long long work = max((long long)num_tb * C, (tb_bits && ref_bits) ? num_tb * bits_per_tb : 0LL);
    const int grid = (int)max(1LL, min((work + 255) / 256, (long long)h->numSMs * 8));
    nr_counters_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(num_tb, C, cb_crc_ok, tb_crc_ok, iters,
                                                               (const signed char*)tb_bits, (const signed char*)ref_bits,
                                                               bits_per_tb, bits_stride, (unsigned long long*)counters);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}
