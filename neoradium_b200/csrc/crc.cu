// K4 (standalone form): CRC attach / check for the six TS 38.212 polynomials, code-block segmentation and
// CRC-check + merge.  Replaces ChanCodeBase.getCrc/checkCrc/appendCrc (neoradium/chancodebase.py:83-189),
// LdpcEncoder.doSegmentation (neoradium/ldpc.py:1011-1030) and LdpcDecoder.checkCrcAndMerge (ldpc.py:1610-1619).
// HBM-bound byte work: a warp per stream segment, 8-byte coalesced traffic, bit-packed table-driven CRC (see below).
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "crc_device.cuh"
#include "nrldpc_internal.cuh"

namespace {

constexpr int CRC_THREADS = 256;
constexpr int CRC_WARPS = CRC_THREADS / 32;
constexpr int CRC_MAX_SEG = 16384;            // bytes (= bits) of a stream handled by one warp at a time
constexpr int CRC_MAX_SEGS = 128;             // segments per stream (factor table in shared memory)

// ---------------------------------------------------------------------------------------------------------------
// One kernel for every CRC-carrying byte stream of the chain (getCrc / appendCrc / checkCrc, doSegmentation,
// checkCrcAndMerge).  A "stream" is `len` values of one bit each (int8, bit = value & 1), optionally zero beyond
// `avail`.  HBM-bound design:
//   * a WARP owns one segment of one stream: 8-byte coalesced loads (256 B per request), each lane folds its 8 bytes
//     into 8 bits with one integer multiply, the same registers are stored to the copy destination (attach /
//     segmentation / merge move the stream anyway), the packed bits go to a per-warp shared-memory buffer;
//   * the CRC runs on the PACKED bits: the segment is cut into 32 right-aligned lane chunks, a lane walks its chunk
//     one packed byte per step through a 256-entry table (rem = (rem << 8) ^ T[top ^ byte]), multiplies its remainder
//     by x^(bits behind its chunk) (per-lane factor, computed once per kernel) and the lanes XOR-reduce with
//     redux.sync -- CRC is GF(2)-linear (chancodebase.py:120-128 is the bit-serial form of the same remainder);
//   * long streams are cut into several segments (work for every SM even with few transport blocks): partial
//     remainders, already multiplied by x^(bits behind the segment), meet in a global XOR accumulator and the warp
//     that arrives last writes the result.
// ---------------------------------------------------------------------------------------------------------------
enum { BS_CRC = 0, BS_ATTACH = 1, BS_CHECK = 2, BS_SEGMENT = 3, BS_MERGE = 4 };

struct BsArgs {
    int mode;
    long long numStreams;
    long long len;         // bits per stream entering the CRC
    int segBytes, numSegs; // segmentation of a stream over warps (segBytes multiple of 256)
    uint32_t poly;
    int c;
    int writeCrc;          // BS_SEGMENT: C > 1 (CRC24B appended), else no CRC at all
    // sources
    const signed char* src;
    long long srcStride;   // BS_CRC/ATTACH/CHECK/MERGE: stream s at src + s * srcStride
    // BS_SEGMENT: stream cb = (t, r) at src + t * srcStride + r * per, valid up to B
    int C, per, K;
    long long B;
    // destinations
    signed char* dst;      // ATTACH: [s][len + c]; SEGMENT: [cb][K]; MERGE: tbBits
    long long dstStride;   // MERGE: transport-block pitch
    signed char* crcBits;  // BS_CRC
    uint32_t* rem;         // BS_CRC
    unsigned char* ok;     // BS_CHECK / BS_MERGE
    // multi-segment combine
    unsigned int* acc;     // [numStreams] XOR accumulators, zeroed by the host
    unsigned int* cnt;     // [numStreams] arrival counters, zeroed by the host
    // GF(2) constants of this call, computed on the host (a few hundred integer operations)
    uint32_t T[256];                  // T[i] = i(x) * x^c mod g
    uint32_t laneFac[32];             // x^(8 q (31 - lane)) mod g, q = segBytes / 256 packed bytes per lane chunk
    uint32_t segFac[CRC_MAX_SEGS];    // x^(bits of the stream behind segment g) mod g
};

uint32_t host_gf_mulmod(uint32_t a, uint32_t b, uint32_t poly, int c)
{
    const uint32_t mask = (1u << c) - 1u;
    uint32_t r = 0;
    for (int i = c - 1; i >= 0; i--) {
        const uint32_t top = (r >> (c - 1)) & 1u;
        r = (r << 1) & mask;
        if (top) r ^= poly;
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}
uint32_t host_gf_xpow(long long e, uint32_t poly, int c)
{
    uint32_t f = 1, base = host_gf_mulmod(1u, 2u, poly, c);   // x
    while (e > 0) {
        if (e & 1) f = host_gf_mulmod(f, base, poly, c);
        base = host_gf_mulmod(base, base, poly, c);
        e >>= 1;
    }
    return f;
}

// 8 bytes (one bit each, first byte = first bit of the stream) -> 8 bits, first bit in bit 7
__device__ __forceinline__ uint32_t pack8(uint2 w)
{
    // (w & 0x01010101) * 0x08040201 puts byte0..byte3 into bits 27..24 (no carries: the partial products hit distinct bits)
    const uint32_t lo = ((w.x & 0x01010101u) * 0x08040201u) >> 24;
    const uint32_t hi = ((w.y & 0x01010101u) * 0x08040201u) >> 24;
    return ((lo & 0xFu) << 4) | (hi & 0xFu);
}

// 16 bytes -> 16 bits (first byte in bit 7 of the low byte, ninth byte in bit 7 of the high byte): two words share one
// multiply -- ((x & m) * 16 + (y & m)) * 0x08040201 has the bits of x in 31..28 and those of y in 27..24, no carries
__device__ __forceinline__ uint32_t pack16(uint4 w)
{
    const uint32_t m = 0x01010101u;
    const uint32_t a = (((w.x & m) * 16u + (w.y & m)) * 0x08040201u) >> 24;
    const uint32_t b = (((w.z & m) * 16u + (w.w & m)) * 0x08040201u) >> 16;
    return a | (b & 0xFF00u);
}

// `steps` whole warp steps of 512 values (16 per lane): load, copy out (COPY; values masked to their bit with MASK), pack
template <bool COPY, bool MASK>
__device__ __forceinline__ void move_pack_steps(const uint4* __restrict__ sp, uint4* __restrict__ dp, unsigned short* pp, int steps,
                                                bool wantCrc)
{
    constexpr uint32_t cm = MASK ? 0x01010101u : 0xFFFFFFFFu;
    int k = 0;
    for (; k + 4 <= steps; k += 4) {
        uint4 w[4];
#pragma unroll
        for (int u = 0; u < 4; u++) w[u] = __ldcs(sp + (k + u) * 32);
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (COPY) __stcs(dp + (k + u) * 32, make_uint4(w[u].x & cm, w[u].y & cm, w[u].z & cm, w[u].w & cm));
            if (wantCrc) pp[(k + u) * 32] = (unsigned short)pack16(w[u]);
        }
    }
    for (; k < steps; k++) {
        const uint4 w = __ldcs(sp + k * 32);
        if (COPY) __stcs(dp + k * 32, make_uint4(w.x & cm, w.y & cm, w.z & cm, w.w & cm));
        if (wantCrc) pp[k * 32] = (unsigned short)pack16(w);
    }
}

__device__ __forceinline__ uint32_t crc_byte_step(uint32_t rem, uint32_t byte, const uint32_t* T, int c, uint32_t mask)
{
    if (c >= 8) return ((rem << 8) & mask) ^ T[((rem >> (c - 8)) ^ byte) & 0xFFu];
    return T[((rem << (8 - c)) ^ byte) & 0xFFu];   // c < 8: rem * x^8 + byte * x^c = x^c * (rem * x^(8-c) + byte)
}

template <int U>
__global__ void __launch_bounds__(CRC_THREADS)
    nr_bitstream_kernel(const __grid_constant__ BsArgs a)
{
    __shared__ uint32_t T[256];
    __shared__ __align__(16) unsigned char pk[CRC_WARPS][CRC_MAX_SEG / 8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = a.c;
    const uint32_t poly = a.poly, mask = (c >= 32) ? 0xFFFFFFFFu : ((1u << c) - 1u);
    const bool wantCrc = (a.mode != BS_SEGMENT) || a.writeCrc;
    T[tid] = a.T[tid];   // lane-divergent lookups later: shared memory, not the constant bank
    const int q = a.segBytes >> 8;                       // packed bytes per lane chunk of a full segment
    const uint32_t laneFac = a.laneFac[lane];
    __syncthreads();

    const long long numJobs = a.numStreams * a.numSegs;
    for (long long job = (long long)blockIdx.x * CRC_WARPS + warp; job < numJobs; job += (long long)gridDim.x * CRC_WARPS) {
        const long long s = job / a.numSegs;
        const int g = (int)(job - s * a.numSegs);
        // geometry of the stream
        const signed char* src;
        long long avail = a.len;       // values beyond `avail` are zeros (segmentation pads the transport block)
        signed char* dst = nullptr;
        long long copyLen = 0;
        if (a.mode == BS_SEGMENT) {
            const long long t = s / a.C;
            const int r = (int)(s - t * a.C);
            src = a.src + t * a.srcStride + (long long)r * a.per;
            avail = max(0LL, min((long long)a.per, a.B - (long long)r * a.per));
            dst = a.dst + s * (long long)a.K;
            copyLen = a.per;
        } else {
            src = a.src + s * a.srcStride;
            if (a.mode == BS_ATTACH) { dst = a.dst + s * (a.len + c); copyLen = a.len; }
            if (a.mode == BS_MERGE && a.dst) {
                const long long t = s / a.C;
                const int r = (int)(s - t * a.C);
                dst = a.dst + t * a.dstStride + (long long)r * a.per;
                copyLen = a.per;
            }
        }
        const long long o0 = (long long)g * a.segBytes;
        const long long o1 = min(a.len, o0 + a.segBytes);
        const int nbits = (int)(o1 - o0);
        const bool srcAligned = ((reinterpret_cast<uintptr_t>(src) & 7) == 0);
        const bool dstAligned = dst && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
        const bool maskCopy = (a.mode == BS_SEGMENT);   // doSegmentation writes bit values (ldpc.py:1011-1030)
        // ---- phase 1, 16-byte form (source and destination 16-byte aligned at the segment start: every block-structured
        //      caller): a lane moves 16 values per step with one 128-bit load / store and packs them into two bytes -- half
        //      the steps of the 8-byte form below (the kernel is bound by instruction issue, not by HBM) ----
        const bool wide = ((reinterpret_cast<uintptr_t>(src + o0) & 15) == 0) && (!dst || ((reinterpret_cast<uintptr_t>(dst + o0) & 15) == 0));
        int base = 0;   // values of the segment already moved by the check-free loop
        if (wide) {
            // whole 512-value warp steps that lie inside the stream, the valid source and the copy: no bounds in the loop,
            // four 128-bit loads in flight per lane, ~12 instructions per 16 values
            long long lim = min((long long)nbits, avail - o0);
            if (dst) lim = min(lim, copyLen - o0);
            const int nFast = lim > 0 ? (int)(lim >> 9) : 0;
            const uint4* sp = reinterpret_cast<const uint4*>(src + o0) + lane;
            uint4* dp = dst ? reinterpret_cast<uint4*>(dst + o0) + lane : nullptr;
            unsigned short* pp = reinterpret_cast<unsigned short*>(&pk[warp][0]) + lane;
            if (!dp) move_pack_steps<false, false>(sp, dp, pp, nFast, wantCrc);
            else if (maskCopy) move_pack_steps<true, true>(sp, dp, pp, nFast, wantCrc);
            else move_pack_steps<true, false>(sp, dp, pp, nFast, wantCrc);
            base = nFast << 9;
        }
        if (wide) {
            constexpr int U2 = (U > 1) ? U / 2 : 1;
            for (int i0 = base + lane * 16; i0 < nbits; i0 += 512 * U2) {
                uint4 w[U2];
#pragma unroll
                for (int u = 0; u < U2; u++) {
                    const int i = i0 + u * 512;
                    const long long o = o0 + i;
                    w[u] = make_uint4(0u, 0u, 0u, 0u);
                    if (i < nbits) {
                        if (o + 16 <= avail) {
                            w[u] = *reinterpret_cast<const uint4*>(src + o);
                        } else {
                            unsigned char* wb = reinterpret_cast<unsigned char*>(&w[u]);
                            for (int k = 0; k < 16; k++)
                                if (o + k < avail) wb[k] = (unsigned char)src[o + k];
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U2; u++) {
                    const int i = i0 + u * 512;
                    const long long o = o0 + i;
                    if (i >= nbits) break;
                    if (dst && o < copyLen) {
                        uint4 wo = w[u];
                        if (maskCopy) { wo.x &= 0x01010101u; wo.y &= 0x01010101u; wo.z &= 0x01010101u; wo.w &= 0x01010101u; }
                        if (o + 16 <= copyLen) {
                            *reinterpret_cast<uint4*>(dst + o) = wo;
                        } else {
                            const unsigned char* wb = reinterpret_cast<const unsigned char*>(&wo);
                            for (int k = 0; k < 16 && o + k < copyLen; k++) dst[o + k] = (signed char)wb[k];
                        }
                    }
                    if (wantCrc) {
                        uint32_t pa = pack8(make_uint2(w[u].x, w[u].y)), pb = pack8(make_uint2(w[u].z, w[u].w));
                        if (i + 8 > nbits) pa &= 0xFF00u >> (nbits - i);           // bits beyond the stream end never enter the CRC
                        if (i + 8 >= nbits) pb = 0u;
                        else if (i + 16 > nbits) pb &= 0xFF00u >> (nbits - i - 8);
                        *reinterpret_cast<unsigned short*>(&pk[warp][i >> 3]) = (unsigned short)((pa & 0xFFu) | ((pb & 0xFFu) << 8));
                    }
                }
            }
        } else
        // ---- phase 1: load 8 values per lane, copy out, pack; four loads in flight per lane (a warp works through its
        //      segment alone, so the memory-level parallelism has to come from here) ----
        for (int i0 = lane * 8; i0 < nbits; i0 += 256 * U) {
            uint2 w[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * 256;
                const long long o = o0 + i;
                w[u] = make_uint2(0u, 0u);
                if (i < nbits) {
                    if (srcAligned && o + 8 <= avail) {
                        w[u] = *reinterpret_cast<const uint2*>(src + o);
                    } else {
                        unsigned char* wb = reinterpret_cast<unsigned char*>(&w[u]);
                        for (int k = 0; k < 8; k++)
                            if (o + k < avail) wb[k] = (unsigned char)src[o + k];
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int i = i0 + u * 256;
                const long long o = o0 + i;
                if (i >= nbits) break;
                if (dst && o < copyLen) {
                    uint2 wo = w[u];
                    if (maskCopy) { wo.x &= 0x01010101u; wo.y &= 0x01010101u; }
                    if (dstAligned && o + 8 <= copyLen) {
                        *reinterpret_cast<uint2*>(dst + o) = wo;
                    } else {
                        const unsigned char* wb = reinterpret_cast<const unsigned char*>(&wo);
                        for (int k = 0; k < 8 && o + k < copyLen; k++) dst[o + k] = (signed char)wb[k];
                    }
                }
                if (wantCrc) {
                    uint32_t p8 = pack8(w[u]);
                    if (i + 8 > nbits) p8 &= 0xFF00u >> (nbits - i);   // bits beyond the stream end never enter the CRC
                    pk[warp][i >> 3] = (unsigned char)p8;
                }
            }
        }
        if (a.mode == BS_SEGMENT && g == a.numSegs - 1) {
            // filler bits are zeros (ldpc.py:1026); the CRC bytes (if any) are written by the finishing warp
            for (int i = a.per + (a.writeCrc ? 24 : 0) + lane; i < a.K; i += 32) dst[i] = 0;
        }
        if (!wantCrc) continue;
        __syncwarp();
        // ---- phase 2: table-driven CRC over the packed bits, 32 right-aligned lane chunks ----
        const int nFull = nbits >> 3, tail = nbits & 7;
        int b0 = nFull - (32 - lane) * q, b1 = b0 + q;
        b0 = max(b0, 0);
        uint32_t rem = 0;
        for (int j = b0; j < b1; j++) rem = crc_byte_step(rem, pk[warp][j], T, c, mask);
        uint32_t part = (b1 > b0) ? nr_gf_mulmod(rem, laneFac, poly, c) : 0u;
        part = __reduce_xor_sync(0xffffffffu, part);
        if (tail) {   // the last (< 8) bits of the stream, bit-serially
            const uint32_t tb = pk[warp][nFull];
            for (int k = 0; k < tail; k++) {
                const uint32_t fb = ((part >> (c - 1)) & 1u) ^ ((tb >> (7 - k)) & 1u);
                part = (part << 1) & mask;
                if (fb) part ^= poly;
            }
        }
        __syncwarp();   // pk[warp] is rewritten by the next job
        // ---- phase 3: combine segments, write the result ----
        uint32_t R = part;
        bool finisher = true;
        if (a.numSegs > 1) {
            R = nr_gf_mulmod(part, a.segFac[g], poly, c);
            unsigned int ticket = 0;
            if (lane == 0) {
                atomicXor(&a.acc[s], R);
                __threadfence();
                ticket = atomicAdd(&a.cnt[s], 1u);
            }
            ticket = __shfl_sync(0xffffffffu, ticket, 0);
            finisher = (ticket == (unsigned int)(a.numSegs - 1));
            if (finisher) {
                __threadfence();
                if (lane == 0) R = atomicXor(&a.acc[s], 0u);
                R = __shfl_sync(0xffffffffu, R, 0);
            }
        }
        if (!finisher) continue;
        const signed char bit = (lane < c) ? (signed char)((R >> (c - 1 - lane)) & 1u) : 0;
        switch (a.mode) {
            case BS_CRC:
                if (a.crcBits && lane < c) a.crcBits[s * c + lane] = bit;
                if (a.rem && lane == 0) a.rem[s] = R;
                break;
            case BS_ATTACH:
                if (lane < c) dst[a.len + lane] = bit;
                break;
            case BS_SEGMENT:
                if (lane < c) dst[a.per + lane] = bit;
                break;
            default:   // BS_CHECK, BS_MERGE
                if (a.ok && lane == 0) a.ok[s] = (R == 0);
                break;
        }
    }
}

// GF(2) constants of a call: T (per polynomial), the lane factors (per segment length) and the segment factors (per stream
// length).  Computing them costs 20-200 us of host time (32 + numSegs square-and-multiply chains of bit-serial products) --
// more than the kernel itself on a few thousand code blocks -- so the last few combinations are kept.
struct BsTabEntry {
    uint32_t poly;
    int c, numSegs;
    long long len, seg;
    unsigned long long stamp;
    uint32_t T[256], laneFac[32], segFac[CRC_MAX_SEGS];
};
void bs_fill_tables(BsArgs& a, long long seg)
{
    static BsTabEntry cache[8];
    static unsigned long long clock = 0;
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    BsTabEntry* e = nullptr;
    BsTabEntry* victim = &cache[0];
    for (BsTabEntry& x : cache) {
        if (x.stamp && x.poly == a.poly && x.c == a.c && x.len == a.len && x.seg == seg && x.numSegs == a.numSegs) { e = &x; break; }
        if (x.stamp < victim->stamp) victim = &x;
    }
    if (!e) {
        e = victim;
        e->poly = a.poly; e->c = a.c; e->numSegs = a.numSegs; e->len = a.len; e->seg = seg;
        const uint32_t mask = (1u << a.c) - 1u;
        for (int i = 0; i < 256; i++) {   // i(x) * x^c mod g: the 8 bits of i through the bit-serial divider
            uint32_t rem = 0;
            for (int b = 7; b >= 0; b--) {
                const uint32_t fb = ((rem >> (a.c - 1)) & 1u) ^ (((uint32_t)i >> b) & 1u);
                rem = (rem << 1) & mask;
                if (fb) rem ^= a.poly;
            }
            e->T[i] = rem;
        }
        const long long q = seg >> 8;
        for (int l = 0; l < 32; l++) e->laneFac[l] = host_gf_xpow(8LL * q * (31 - l), a.poly, a.c);
        for (int g = 0; g < a.numSegs; g++) {
            const long long end = (a.len < (long long)(g + 1) * seg) ? a.len : (long long)(g + 1) * seg;
            e->segFac[g] = host_gf_xpow(a.len - end, a.poly, a.c);
        }
    }
    e->stamp = ++clock;
    memcpy(a.T, e->T, sizeof(a.T));
    memcpy(a.laneFac, e->laneFac, sizeof(a.laneFac));
    memcpy(a.segFac, e->segFac, sizeof(uint32_t) * (size_t)a.numSegs);
}

// segmentation of the streams over warps and launch
int launch_bitstream(nrldpc_handle* h, BsArgs& a, cudaStream_t st)
{
    // enough warps for every SM (>= 16 per SM) even with few long streams; a segment is at most CRC_MAX_SEG bytes
    long long want = ((long long)h->numSMs * 16 + a.numStreams - 1) / a.numStreams;
    long long segs = max(1LL, min(want, (a.len + 4095) / 4096));
    segs = max(segs, (a.len + CRC_MAX_SEG - 1) / CRC_MAX_SEG);
    if (segs > CRC_MAX_SEGS) { nr_set_error("crc: stream of %lld bits is too long", a.len); return NRLDPC_ERR_ARG; }
    long long seg = (a.len + segs - 1) / segs;
    seg = max(256LL, (seg + 255) & ~255LL);
    if (seg > CRC_MAX_SEG) { nr_set_error("crc: stream of %lld bits is too long", a.len); return NRLDPC_ERR_ARG; }
    a.segBytes = (int)seg;
    a.numSegs = (int)max(1LL, (a.len + seg - 1) / seg);
    a.acc = nullptr;
    a.cnt = nullptr;
    if (a.numSegs > 1) {
        void* p = nullptr;
        const size_t bytes = (size_t)a.numStreams * 2 * sizeof(unsigned int);
        const int rc = nr_reserve_tmp2(h, bytes, &p);
        if (rc) return rc;
        NR_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st));
        a.acc = (unsigned int*)p;
        a.cnt = a.acc + a.numStreams;
    }
    bs_fill_tables(a, seg);
    const long long jobs = a.numStreams * a.numSegs;
    static const int perSM = nr_ctas_per_sm(nr_bitstream_kernel<4>, CRC_THREADS, 0);
    const int grid = (int)max(1LL, min((jobs + CRC_WARPS - 1) / CRC_WARPS, (long long)h->numSMs * perSM));
    // four loads in flight per lane: measured on B200 at 16384 code blocks, U = 1 / 2 / 4 -> 2.9 / 3.1 / 3.8 TB/s (merge)
    nr_bitstream_kernel<4><<<grid, CRC_THREADS, 0, st>>>(a);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

__global__ void nr_counters_kernel(long long numTb, int C, const unsigned char* cbOk, const unsigned char* tbOk,
                                   const int* iters, const signed char* tbBits, const signed char* refBits,
                                   long long bitsPerTb, long long bitsStride, long long refStride,
                                   unsigned long long* counters)
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
    if (tbBits && refBits) {   // four bits (bytes) per step where both rows allow an aligned word, single bytes elsewhere
        const long long W = (bitsPerTb + 3) >> 2, total = numTb * W;
        for (long long i = gtid; i < total; i += gsz) {
            const long long t = i / W, j = (i - t * W) << 2;
            const signed char* pa = tbBits + t * bitsStride + j;
            const signed char* pb = refBits + t * refStride + j;
            if (j + 4 <= bitsPerTb && ((reinterpret_cast<uintptr_t>(pa) | reinterpret_cast<uintptr_t>(pb)) & 3) == 0) {
                bitErr += __popc((*reinterpret_cast<const unsigned int*>(pa) ^ *reinterpret_cast<const unsigned int*>(pb)) & 0x01010101u);
            } else {
                for (long long k = j; k < bitsPerTb && k < j + 4; k++) bitErr += ((pa[k - j] ^ pb[k - j]) & 1) ? 1 : 0;
            }
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
    BsArgs a{};
    a.mode = mode;
    a.numStreams = n;
    a.len = len;
    a.poly = p.poly;
    a.c = p.len;
    a.src = (const signed char*)bits;
    a.srcStride = stride;
    a.dst = (signed char*)attachOut;
    a.crcBits = (signed char*)crcBits;
    a.rem = rem;
    a.ok = ok;
    return launch_bitstream(h, a, (cudaStream_t)stream);
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
    const NrCrcPoly pb = nr_crc_poly(NRLDPC_CRC24B);
    BsArgs a{};
    a.mode = BS_SEGMENT;
    a.numStreams = num_tb * C;
    a.len = per;
    a.poly = pb.poly;
    a.c = pb.len;
    a.writeCrc = (C > 1);
    a.src = (const signed char*)tb;
    a.srcStride = tb_stride;
    a.C = C;
    a.per = (int)per;
    a.K = K;
    a.B = B;
    a.dst = (signed char*)code_blocks;
    return launch_bitstream(h, a, (cudaStream_t)stream);
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
    const int Lk = cfg->K - cfg->F;
    const NrCrcPoly p = nr_crc_poly(cfg->C > 1 ? NRLDPC_CRC24B : NRLDPC_CRC24A);
    BsArgs a{};
    a.mode = BS_MERGE;
    a.numStreams = num_tb * cfg->C;
    a.len = Lk;
    a.poly = p.poly;
    a.c = p.len;
    a.src = (const signed char*)decoded;
    a.srcStride = cfg->K;
    a.C = cfg->C;
    a.per = (cfg->C > 1) ? Lk - 24 : Lk;
    a.K = cfg->K;
    a.dst = (signed char*)tb_bits;
    a.dstStride = tb_bits_stride;
    a.ok = cb_crc_ok;
    return launch_bitstream(h, a, (cudaStream_t)stream);
}

extern "C" int nrldpc_accumulate_counters(nrldpc_handle* h, int64_t num_tb, int C, const uint8_t* cb_crc_ok,
                                          const uint8_t* tb_crc_ok, const int32_t* iters, const int8_t* tb_bits,
                                          const int8_t* ref_bits, int64_t bits_per_tb, int64_t bits_stride,
                                          int64_t* counters, nrldpc_stream stream)
{
    return nrldpc_accumulate_counters_ref(h, num_tb, C, cb_crc_ok, tb_crc_ok, iters, tb_bits, bits_stride, ref_bits, bits_stride,
                                          bits_per_tb, counters, stream);
}

extern "C" int nrldpc_accumulate_counters_ref(nrldpc_handle* h, int64_t num_tb, int C, const uint8_t* cb_crc_ok,
                                              const uint8_t* tb_crc_ok, const int32_t* iters, const int8_t* tb_bits,
                                              int64_t tb_stride, const int8_t* ref_bits, int64_t ref_stride,
                                              int64_t bits_per_tb, int64_t* counters, nrldpc_stream stream)
{
    if (!h || !counters || num_tb <= 0 || C < 1) { nr_set_error("counters: bad argument"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const long long work = max((long long)num_tb * C, (tb_bits && ref_bits) ? num_tb * ((bits_per_tb + 3) / 4) : 0LL);
    const int grid = (int)max(1LL, min((work + 255) / 256, (long long)h->numSMs * 8));
    nr_counters_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(num_tb, C, cb_crc_ok, tb_crc_ok, iters,
                                                               (const signed char*)tb_bits, (const signed char*)ref_bits,
                                                               bits_per_tb, tb_stride, ref_stride, (unsigned long long*)counters);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}
