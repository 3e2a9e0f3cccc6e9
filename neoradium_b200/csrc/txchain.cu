// K2 + K3a: systematic QC-LDPC encoder, rate matching, and the full parity check.
// Replaces LdpcEncoder.encode (neoradium/ldpc.py:1057-1090), LdpcEncoder.rateMatch (ldpc.py:1128-1159) and a correct
// LdpcBase.isValidCodedBlock (ldpc.py:825-843).  All three are HBM-bound byte kernels: coalesced int8 loads/stores
// along the lifted index, the code block staged in shared memory, circulant shifts as index arithmetic.
#include <stdlib.h>

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

// ---------------------------------------------------------------------------------------------------------------
// Bit-packed encoder (Z a multiple of 16: every lifting size >= 128 and 16/32/48/64/80/96/112).
//
// One code block per 128-thread CTA (5-6 KB of shared memory, so an SM keeps 16 of them in flight).  The K input
// bytes are read as 16-byte vectors and squeezed to 16 bits each with an integer multiply; every column of degree > 1
// is kept DOUBLED in shared memory (bits 0..Z-1 followed by the same Z bits), so the circulant rotation rot(x, s) is
// "read Z bits at bit offset s" -- two consecutive words and one funnel shift per 32 result bits, no wrap-around
// logic, for every Z.  A row is the XOR of such words; the result bits are spread back to one byte per bit with one
// multiply per 4 bytes and leave as 16-byte vector stores.
// ---------------------------------------------------------------------------------------------------------------
constexpr int ENC_THREADS = 128;

__device__ __forceinline__ uint32_t pack4(uint32_t w)   // 4 bytes (0/1) -> 4 bits, byte 0 at bit 0
{
    return ((w & 0x01010101u) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t pack16(uint4 v)
{
    return pack4(v.x) | (pack4(v.y) << 4) | (pack4(v.z) << 8) | (pack4(v.w) << 12);
}
__device__ __forceinline__ uint32_t spread4(uint32_t nib)   // 4 bits -> 4 bytes (0/1)
{
    return ((nib & 0xFu) * 0x00204081u) & 0x01010101u;
}
// 32 bits of rot(x, s) starting at result bit 32*i, from the doubled column d
__device__ __forceinline__ uint32_t rot_word(const uint32_t* d, int s, int i)
{
    const int o = s + 32 * i;
    return __funnelshift_r(d[o >> 5], d[(o >> 5) + 1], o & 31);
}
// write result word `v` (bits 32*i .. of a Z-bit column) into the doubled column d (16-bit granularity: Z % 16 == 0)
__device__ __forceinline__ void store_doubled(uint32_t* d, int i, uint32_t v, int Zh)
{
    unsigned short* dh = reinterpret_cast<unsigned short*>(d);
    const int h0 = 2 * i;
    if (h0 < Zh) { dh[h0] = (unsigned short)v; dh[h0 + Zh] = (unsigned short)v; }
    if (h0 + 1 < Zh) { dh[h0 + 1] = (unsigned short)(v >> 16); dh[h0 + 1 + Zh] = (unsigned short)(v >> 16); }
}

__global__ void __launch_bounds__(ENC_THREADS)
    nr_encode_packed_kernel(const __grid_constant__ NrGraph g, const signed char* __restrict__ in, long long numCb,
                            signed char* __restrict__ out, int puncture)
{
    extern __shared__ uint32_t esm[];
    const int Z = g.Z, k = g.ksys, ncore = g.ncore, P = g.P;
    const int Zh = Z >> 4;                  // 16-bit units per column
    const int W = (Z + 31) >> 5;            // 32-bit words per column
    const int DW = ((2 * Z + 31) >> 5) + 2; // words per doubled column (padded)
    uint32_t* dbl = esm;                            // [ncore][DW]
    uint32_t* lam = dbl + ncore * DW;               // [4][W]   row sums over the systematic part
    uint32_t* sum = lam + 4 * W;                    // [DW]     doubled lam0^lam1^lam2^lam3
    uint32_t* ext = sum + DW;                       // [P-4][W] extension parity
    const int tid = threadIdx.x;
    const int firstCol = puncture ? 2 : 0;
    const int outCols = g.ncols - firstCol;
    const uint32_t rcpZh = (65536u + Zh - 1) / Zh;  // exact floor(n / Zh) for n < 2730 (n <= 68 * 24)
    const uint32_t rcpW = (65536u + W - 1) / W;
    // shifts of the core-parity double diagonal (ldpc.py:1068-1080)
    int b = edge_shift(g, 1, k);
    if (b < 0) b = edge_shift(g, 2, k);

    for (long long cb = blockIdx.x; cb < numCb; cb += gridDim.x) {
        // ---- load + pack the systematic columns (doubled) ----------------------------------------------------
        const uint4* src = reinterpret_cast<const uint4*>(in + cb * (long long)k * Z);
        for (int i = tid; i < ncore * DW + 4 * W + DW; i += ENC_THREADS) esm[i] = 0;
        __syncthreads();
        for (int gi = tid; gi < k * Zh; gi += ENC_THREADS) {
            const uint32_t h = pack16(__ldg(src + gi));
            const int col = (int)(((uint32_t)gi * rcpZh) >> 16), i = gi - col * Zh;
            unsigned short* dh = reinterpret_cast<unsigned short*>(dbl + col * DW);
            dh[i] = (unsigned short)h;
            dh[i + Zh] = (unsigned short)h;
        }
        __syncthreads();
        // ---- rows 0..3 over the systematic columns ------------------------------------------------------------
        for (int t = tid; t < 4 * W; t += ENC_THREADS) {
            const int r = (int)(((uint32_t)t * rcpW) >> 16), i = t - r * W;
            uint32_t acc = 0;
            for (int e = g.rowEdge0[r]; e < g.rowEdge0[r + 1]; e++) {
                const uint32_t ew = g.edge[e];
                const int col = (int)(ew >> 16);
                if (col >= k) break;   // columns ascend inside a row
                acc ^= rot_word(dbl + col * DW, (int)(ew & 0xffffu), i);
            }
            lam[r * W + i] = acc;
        }
        __syncthreads();
        if (tid < W) store_doubled(sum, tid, lam[tid] ^ lam[W + tid] ^ lam[2 * W + tid] ^ lam[3 * W + tid], Zh);
        __syncthreads();
        // p0 = rot(sum, Z - b)
        if (tid < W) store_doubled(dbl + k * DW, tid, rot_word(sum, (b == 0) ? 0 : Z - b, tid), Zh);
        __syncthreads();
        for (int r = 0; r < 3; r++) {   // p1..p3 through the double diagonal
            if (tid < W) {
                uint32_t acc = lam[r * W + tid];
                for (int e = g.rowEdge0[r]; e < g.rowEdge0[r + 1]; e++) {
                    const uint32_t ew = g.edge[e];
                    const int col = (int)(ew >> 16);
                    if (col >= k && col <= k + r) acc ^= rot_word(dbl + col * DW, (int)(ew & 0xffffu), tid);
                }
                store_doubled(dbl + (k + r + 1) * DW, tid, acc, Zh);
            }
            __syncthreads();
        }
        // ---- extension rows: XOR over the core columns only (ldpc.py:1083-1084) ------------------------------
        for (int t = tid; t < (P - 4) * W; t += ENC_THREADS) {
            const int r4 = (int)(((uint32_t)t * rcpW) >> 16), i = t - r4 * W;
            uint32_t acc = 0;
            for (int e = g.rowEdge0[r4 + 4]; e < g.rowEdge0[r4 + 5]; e++) {
                const uint32_t ew = g.edge[e];
                const int col = (int)(ew >> 16);
                if (col >= ncore) break;
                acc ^= rot_word(dbl + col * DW, (int)(ew & 0xffffu), i);
            }
            ext[t] = acc;
        }
        __syncthreads();
        // ---- spread to one byte per bit, 16 bytes per store ---------------------------------------------------
        uint4* dst = reinterpret_cast<uint4*>(out + cb * (long long)outCols * Z);
        for (int gi = tid; gi < outCols * Zh; gi += ENC_THREADS) {
            const int oc = (int)(((uint32_t)gi * rcpZh) >> 16), i = gi - oc * Zh;
            const int col = oc + firstCol;
            const unsigned short* hp = (col < ncore) ? reinterpret_cast<const unsigned short*>(dbl + col * DW)
                                                     : reinterpret_cast<const unsigned short*>(ext + (col - ncore) * W);
            const uint32_t h = hp[i];
            uint4 v;
            v.x = spread4(h);
            v.y = spread4(h >> 4);
            v.z = spread4(h >> 8);
            v.w = spread4(h >> 12);
            dst[gi] = v;
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

// Parity check on bit-packed columns (Z % 16 == 0), same machinery as the packed encoder: the core columns are kept
// doubled so that a circulant rotation is a funnel shift, the degree-1 extension columns (identity circulant) are read
// in place.  Syndrome word (row r, word i) = XOR over the row's edges; the block is a code word iff every word is 0.
__global__ void __launch_bounds__(ENC_THREADS)
    nr_parity_packed_kernel(const __grid_constant__ NrGraph g, const signed char* __restrict__ coded, long long numCb,
                            unsigned char* ok)
{
    extern __shared__ uint32_t esm[];
    __shared__ int badFlag;
    const int Z = g.Z, ncore = g.ncore, P = g.P, ncols = g.ncols;
    const int Zh = Z >> 4;
    const int W = (Z + 31) >> 5;
    const int DW = ((2 * Z + 31) >> 5) + 2;
    uint32_t* dbl = esm;                  // [ncore][DW] doubled core columns
    uint32_t* ext = dbl + ncore * DW;     // [P-4][W]    extension columns
    const int tid = threadIdx.x;
    const uint32_t rcpZh = (65536u + Zh - 1) / Zh;
    const uint32_t rcpW = (65536u + W - 1) / W;
    for (long long cb = blockIdx.x; cb < numCb; cb += gridDim.x) {
        const uint4* src = reinterpret_cast<const uint4*>(coded + cb * (long long)ncols * Z);
        for (int i = tid; i < ncore * DW + (P - 4) * W; i += ENC_THREADS) esm[i] = 0;
        if (tid == 0) badFlag = 0;
        __syncthreads();
        for (int gi = tid; gi < ncols * Zh; gi += ENC_THREADS) {
            const uint32_t h = pack16(__ldg(src + gi));
            const int col = (int)(((uint32_t)gi * rcpZh) >> 16), i = gi - col * Zh;
            if (col < ncore) {
                unsigned short* dh = reinterpret_cast<unsigned short*>(dbl + col * DW);
                dh[i] = (unsigned short)h;
                dh[i + Zh] = (unsigned short)h;
            } else {
                reinterpret_cast<unsigned short*>(ext + (col - ncore) * W)[i] = (unsigned short)h;
            }
        }
        __syncthreads();
        uint32_t any = 0;
        for (int t = tid; t < P * W; t += ENC_THREADS) {
            const int r = (int)(((uint32_t)t * rcpW) >> 16), i = t - r * W;
            uint32_t acc = (r >= 4) ? ext[(r - 4) * W + i] : 0u;
            for (int e = g.rowEdge0[r]; e < g.rowEdge0[r + 1]; e++) {
                const uint32_t ew = g.edge[e];
                const int col = (int)(ew >> 16);
                if (col >= ncore) break;   // columns ascend inside a row; the extension edge is the identity
                acc ^= rot_word(dbl + col * DW, (int)(ew & 0xffffu), i);
            }
            if (32 * i + 32 > Z) acc &= (1u << (Z - 32 * i)) - 1u;   // bits beyond Z in the last word
            any |= acc;
        }
        if (any) badFlag = 1;
        __syncthreads();
        if (tid == 0) ok[cb] = badFlag ? 0 : 1;
        __syncthreads();
    }
}

// rateMatch (ldpc.py:1128-1159): one CTA per code block.  The first Ncb values of the coded block are staged in shared
// memory with 16-byte coalesced loads; every thread then produces EIGHT consecutive values of the rate-matched stream
// (one 8-byte store, groups aligned to the destination address).  The interleaver / circular-buffer index is advanced
// incrementally inside a group -- one division and two modulos per 8 outputs instead of two of each per output.
__global__ void __launch_bounds__(256)
    nr_rate_match_kernel(const signed char* coded, long long numCb, int C, int N, int K, int F, int Z, int ncb, int k0,
                         int qm, int E0, int nShort, int fStep, signed char* out, long long outStride)
{
    extern __shared__ __align__(16) signed char cbS[];
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
        __syncthreads();   // the previous block's gathers are done
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const int n16 = ncb >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(cbS);
            for (int i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = s4[i];
            for (int i = (n16 << 4) + threadIdx.x; i < ncb; i += blockDim.x) cbS[i] = src[i];
        } else {
            for (int i = threadIdx.x; i < ncb; i += blockDim.x) cbS[i] = src[i];
        }
        __syncthreads();
        const int mis = (int)(reinterpret_cast<uintptr_t>(dst) & 7);
        signed char* dstA = dst - mis;                    // 8-byte aligned
        const int nGroups = (mis + E + 7) >> 3;
        for (int v = threadIdx.x; v < nGroups; v += blockDim.x) {
            const int g0 = 8 * v - mis;
            const int gFirst = max(g0, 0);
            // interleaver: out[s*qm + b] = e[b*Eq + s] (ldpc.py:1155); e[i] = circ[(k0 + i) mod L], the circular
            // buffer WITHOUT fillers (ldpc.py:1139-1142)
            int sIdx = gFirst / qm, b = gFirst - sIdx * qm;
            int qb0 = (k0 + sIdx) % L;                    // position of (sIdx, b = 0)
            int q = (int)(((long long)qb0 + (long long)b * Eq) % L);
            unsigned long long w = 0;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int g = g0 + k;
                if (g >= 0 && g < E) {
                    const int n = (q < sysLen) ? q : q + F;
                    w |= (unsigned long long)(unsigned char)cbS[n] << (8 * k);
                    if (++b == qm) {
                        b = 0;
                        qb0 = (qb0 + 1 == L) ? 0 : qb0 + 1;
                        q = qb0;
                    } else {
                        q += Eq;
                        while (q >= L) q -= L;
                    }
                }
            }
            if (g0 >= 0 && g0 + 8 <= E) {
                *reinterpret_cast<unsigned long long*>(dstA + 8 * v) = w;
            } else {
                for (int k = 0; k < 8; k++) {
                    const int g = g0 + k;
                    if (g >= 0 && g < E) dst[g] = (signed char)(w >> (8 * k));
                }
            }
        }
    }
}

// Staged scatter form of rateMatch (the one that normally runs).  The coded block is staged in shared memory as above;
// the circular buffer is then walked in up to three segments (cut at k0 and at the filler gap) inside which the stream
// index i = q + c and the source index n = q + d are affine in the buffer position q.  A warp takes 32*U consecutive
// positions at a time; when they stay inside one row of the interleaver (same b = i / Eq) the rate-matched index
// s*qm + b is affine too, so a byte costs one LDS.U8 + one STS.U8 (both conflict-free) into a shared-memory image of the
// output slice, which is finally written with aligned 16-byte stores.  Chunks that straddle a row, the end of the data
// or a repeated buffer (E > L) take the generic per-byte path.
template <int U>
__global__ void __launch_bounds__(256)
    nr_rate_match_staged_kernel(const signed char* __restrict__ coded, long long numCb, int C, int N, int K, int F, int Z,
                                int ncb, int k0, int qm, int E0, int nShort, int fStep, signed char* __restrict__ out,
                                long long outStride, int ncbPad)
{
    extern __shared__ __align__(16) signed char rmSmem[];
    signed char* cbS = rmSmem;
    signed char* outS = rmSmem + ncbPad;
    const int L = ncb - F;
    const int sysLen = K - 2 * Z - F;
    const int k0m = k0 % L;
    const int nT = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nWarps = nT >> 5;
    constexpr int CH = 32 * U;
    for (long long cb = blockIdx.x; cb < numCb; cb += gridDim.x) {
        const long long tb = cb / C;
        const int r = (int)(cb - tb * C);
        const int E = E0 + (r >= nShort ? fStep : 0);
        const long long off = (long long)r * E0 + (long long)(r > nShort ? r - nShort : 0) * fStep;
        const int Eq = E / qm;
        const signed char* __restrict__ src = coded + cb * (long long)N;
        signed char* __restrict__ dst = out + tb * outStride + off;
        __syncthreads();   // the previous block's copy-out is done
        if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
            const int n16 = ncb >> 4;
            const uint4* __restrict__ s4 = reinterpret_cast<const uint4*>(src);
            uint4* d4 = reinterpret_cast<uint4*>(cbS);
#pragma unroll 2
            for (int i = tid; i < n16; i += nT) d4[i] = __ldg(s4 + i);
            for (int i = (n16 << 4) + tid; i < ncb; i += nT) cbS[i] = src[i];
        } else {
            for (int i = tid; i < ncb; i += nT) cbS[i] = src[i];
        }
        const int mis = (int)(reinterpret_cast<uintptr_t>(dst) & 15);
        signed char* outB = outS + mis;   // outB[g] <-> dst[g]: 16-byte units of outS map to aligned units of the destination
        __syncthreads();
        // e[i] = circ[(k0 + i) mod L] (ldpc.py:1145-1151), out[s*qm + b] = e[b*Eq + s] (ldpc.py:1155)
        auto emit = [&](int q, int c, int d) {
            const signed char v = cbS[q + d];
            for (int i = q + c; i < E; i += L) {
                const int b = i / Eq, s2 = i - b * Eq;
                outB[s2 * qm + b] = v;
            }
        };
        auto seg = [&](int qlo, int qhi) {
            if (qlo >= qhi) return;
            const int c = (qlo < k0m) ? L - k0m : -k0m;
            const int d = (qlo < sysLen) ? 0 : F;
            const int qData = min(qhi, E - c);   // positions that are transmitted at least once
            int qw = qlo + warp * CH;
            if (qw >= qData) return;
            int bw = (qw + c) / Eq, sw = (qw + c) - bw * Eq;   // interleaver row / column of the chunk's first position
            const int step = nWarps * CH, stepB = step / Eq, stepS = step - stepB * Eq;
            const bool noRep = E <= L;
            for (; qw < qData; qw += step) {
                if (noRep && qw + CH <= qData && sw + CH <= Eq) {
                    const signed char* __restrict__ p = cbS + qw + d + lane;
                    signed char* __restrict__ o = outB + (sw + lane) * qm + bw;
                    signed char v[U];
#pragma unroll
                    for (int k = 0; k < U; k++) v[k] = p[32 * k];
#pragma unroll
                    for (int k = 0; k < U; k++) o[32 * k * qm] = v[k];
                } else {
#pragma unroll
                    for (int k = 0; k < U; k++) {
                        const int q = qw + lane + 32 * k;
                        if (q < qData) emit(q, c, d);
                    }
                }
                sw += stepS;
                bw += stepB;
                if (sw >= Eq) {
                    sw -= Eq;
                    bw++;
                }
            }
        };
        const int cutA = min(k0m, sysLen), cutB = max(k0m, sysLen);
        seg(0, cutA);
        seg(cutA, cutB);
        seg(cutB, L);
        __syncthreads();
        {   // copy-out: 16-byte units of outS; the first / last unit may be partial
            const int nUnits = (mis + E + 15) >> 4;
            const uint4* s4 = reinterpret_cast<const uint4*>(outS);
            signed char* dstA = dst - mis;   // 16-byte aligned
            for (int u = tid; u < nUnits; u += nT) {
                const int g0 = 16 * u - mis;
                if (g0 >= 0 && g0 + 16 <= E) {
                    *reinterpret_cast<uint4*>(dstA + 16 * u) = s4[u];
                } else {
                    for (int k = 0; k < 16; k++) {
                        const int g = g0 + k;
                        if (g >= 0 && g < E) dst[g] = outB[g];
                    }
                }
            }
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
    // start of the reads in the filler-less circular buffer, reduced once here: (arange + start) % cirBufSize of
    // ldpc.py:1148/1407 lets start exceed the buffer (LBRM + small BG2 block + rv 3: k0 >= Ncb - F)
    const int L = c->ncb - c->F;
    *k0 = (int)(((long long)num * c->ncb / N) * c->zc);
    if (L > 0) *k0 %= L;
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
    if (zc % 16 == 0 && ((uintptr_t)code_blocks & 15) == 0 && ((uintptr_t)coded & 15) == 0 && !getenv("NRLDPC_ENC_BYTEWISE")) {
        // bit-packed path: vector loads/stores need 16-byte granularity (column length Z is a multiple of 16)
        const int W = (zc + 31) / 32, DW = (2 * zc + 31) / 32 + 2;
        const size_t smem = (size_t)(g.ncore * DW + 4 * W + DW + (g.P - 4) * W) * sizeof(uint32_t);
        const int grid = (int)min((long long)num_cb, (long long)h->numSMs * nr_ctas_per_sm(nr_encode_packed_kernel, ENC_THREADS, smem));
        nr_encode_packed_kernel<<<grid, ENC_THREADS, smem, (cudaStream_t)stream>>>(g, (const signed char*)code_blocks, num_cb,
                                                                                  (signed char*)coded, puncture);
        NR_CUDA_CHECK(cudaGetLastError());
        return NRLDPC_OK;
    }
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
    return nrldpc_parity_check_rows(h, bg, zc, coded_full, num_cb, 0, ok, stream);
}

extern "C" int nrldpc_parity_check_rows(nrldpc_handle* h, int bg, int zc, const int8_t* coded_full, int64_t num_cb,
                                        int rows, uint8_t* ok, nrldpc_stream stream)
{
    if (!h) { nr_set_error("parity_check: null handle"); return NRLDPC_ERR_ARG; }
    NrGraph g;
    if (nr_build_graph(bg, zc, &g)) return NRLDPC_ERR_ARG;
    if (num_cb <= 0 || rows < 0) { nr_set_error("parity_check: bad shape"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const bool allRows = rows == 0 || rows >= g.P;
    if (!allRows) g.P = rows;   // the first `rows` base-graph rows only (rows = 1: the reference's isValidCodedBlock, ldpc.py:841-843)
    if (allRows && zc % 16 == 0 && ((uintptr_t)coded_full & 15) == 0 && !getenv("NRLDPC_ENC_BYTEWISE")) {
        const int W = (zc + 31) / 32, DW = (2 * zc + 31) / 32 + 2;
        const size_t smemP = (size_t)(g.ncore * DW + (g.P - 4) * W) * sizeof(uint32_t);
        const int gridP = (int)min((long long)num_cb, (long long)h->numSMs * nr_ctas_per_sm(nr_parity_packed_kernel, ENC_THREADS, smemP));
        nr_parity_packed_kernel<<<gridP, ENC_THREADS, smemP, (cudaStream_t)stream>>>(g, (const signed char*)coded_full, num_cb, ok);
        NR_CUDA_CHECK(cudaGetLastError());
        return NRLDPC_OK;
    }
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
    {   // staged scatter kernel: input block + output slice in shared memory
        const int ncbPad = (cfg->ncb + 15) & ~15;
        const size_t smemS = (size_t)ncbPad + (size_t)((E0 + fStep + 15) & ~15) + 32;
        if (smemS <= (size_t)h->maxSmemOptin && E0 >= cfg->qm && !getenv("NRLDPC_RM_GENERIC")) {
            const int nThr = 256;   // 512-thread CTAs measured slower (2571 vs 3440 GB/s at 16k blocks)
            int perSM = (int)((size_t)h->smemPerSM / (smemS + 1024));
            perSM = perSM < 1 ? 1 : (perSM > 2048 / nThr ? 2048 / nThr : perSM);
            const int gridS = (int)min(numCb, (long long)h->numSMs * perSM);
            const char* ue = getenv("NRLDPC_RM_U");   // A/B measurements: positions per lane per chunk
            const int U = ue ? atoi(ue) : 4;
#define NR_RM_LAUNCH(UU)                                                                                                \
    do {                                                                                                                \
        if (smemS > 48 * 1024)                                                                                          \
            NR_CUDA_CHECK(cudaFuncSetAttribute(nr_rate_match_staged_kernel<UU>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                               (int)smemS));                                                            \
        nr_rate_match_staged_kernel<UU><<<gridS, nThr, smemS, (cudaStream_t)stream>>>(                                   \
            (const signed char*)coded, numCb, cfg->C, N, cfg->K, cfg->F, cfg->zc, cfg->ncb, k0, cfg->qm, E0, nShort, fStep, \
            (signed char*)out, out_stride, ncbPad);                                                                     \
    } while (0)
            if (U == 8) NR_RM_LAUNCH(8);
            else if (U == 2) NR_RM_LAUNCH(2);
            else NR_RM_LAUNCH(4);
#undef NR_RM_LAUNCH
            NR_CUDA_CHECK(cudaGetLastError());
            return NRLDPC_OK;
        }
    }
    const int grid = (int)min(numCb, (long long)h->numSMs * 8);
    const size_t smem = (size_t)((cfg->ncb + 15) & ~15);
    if (smem > 48 * 1024) NR_CUDA_CHECK(cudaFuncSetAttribute(nr_rate_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nr_rate_match_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const signed char*)coded, numCb, cfg->C, N, cfg->K,
                                                                 cfg->F, cfg->zc, cfg->ncb, k0, cfg->qm, E0, nShort, fStep,
                                                                 (signed char*)out, out_stride);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}
