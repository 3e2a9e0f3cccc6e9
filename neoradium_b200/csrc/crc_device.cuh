// Cooperative CRC over a bit stream for one group of threads (device side, shared by crc.cu / txchain.cu / rxchain.cu).
//
// Algorithm of ChanCodeBase.getCrc (neoradium/chancodebase.py:120-128): remainder of bits(x) * x^c modulo g(x), MSB
// first, zero initial state.  The reference walks the stream one bit at a time; CRC is GF(2)-linear, so here the
// stream is right-aligned into `nThr` equal chunks (leading zeros do not change a zero-initialised CRC), every thread
// divides its own chunk, and the partial remainders are merged pairwise as rem = left * x^(chunk span) + right.
#pragma once
#include <stdint.h>

__host__ __device__ __forceinline__ uint32_t nr_gf_shift1(uint32_t r, uint32_t poly, int c)
{
    const uint32_t top = (r >> (c - 1)) & 1u;
    r = (r << 1) & ((1u << c) - 1u);
    return top ? (r ^ poly) : r;
}

__host__ __device__ __forceinline__ uint32_t nr_gf_mulmod(uint32_t a, uint32_t b, uint32_t poly, int c)
{
    uint32_t r = 0;
    for (int i = c - 1; i >= 0; i--) {
        r = nr_gf_shift1(r, poly, c);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}

// x^e mod g
__host__ __device__ __forceinline__ uint32_t nr_gf_xpow(long long e, uint32_t poly, int c)
{
    uint32_t f = 1, base = nr_gf_shift1(1u, poly, c);   // x (also right when c == 1.. never: c >= 6)
    while (e > 0) {
        if (e & 1) f = nr_gf_mulmod(f, base, poly, c);
        base = nr_gf_mulmod(base, base, poly, c);
        e >>= 1;
    }
    return f;
}

// All `nThr` threads of the group (nThr a power of two, tid in [0, nThr)) must call this; `tree` is shared memory of
// nThr words private to the group; barriers are CTA-wide (__syncthreads), so every thread of the CTA must take part.
// fetch(i) returns bit i (0/1) for 0 <= i < len.
template <typename Fetch>
__device__ uint32_t nr_group_crc(Fetch fetch, long long len, int nThr, int tid, uint32_t* tree, uint32_t poly, int c)
{
    const long long B = (len + nThr - 1) / nThr;
    const long long lead = B * nThr - len;
    uint32_t rem = 0;
    const long long i0 = (long long)tid * B - lead;
    for (long long b = 0; b < B; b++) {
        const long long i = i0 + b;
        const uint32_t bit = (i >= 0) ? fetch(i) : 0u;
        const uint32_t fb = ((rem >> (c - 1)) & 1u) ^ bit;
        rem = (rem << 1) & ((1u << c) - 1u);
        if (fb) rem ^= poly;
    }
    tree[tid] = rem;
    uint32_t f = nr_gf_xpow(B, poly, c);
    __syncthreads();
    for (int span = 1; span < nThr; span <<= 1) {
        const int right = (tid + 1) * 2 * span - 1;
        if (right < nThr) tree[right] = nr_gf_mulmod(tree[right - span], f, poly, c) ^ tree[right];
        f = nr_gf_mulmod(f, f, poly, c);
        __syncthreads();
    }
    const uint32_t out = tree[nThr - 1];
    __syncthreads();   // tree may be reused immediately by the caller
    return out;
}
