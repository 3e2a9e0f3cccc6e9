// K1: batched layered normalised min-sum decoder for the lifted QC parity-check matrices of TS 38.212 (BG1/BG2, every
// Zc <= 384), with rate recovery fused into its load phase (K3b) and the CRC fused into its epilogue (K4).
//
// Replaces LdpcDecoder.decode (neoradium/ldpc.py:1535-1581) and, in fused mode, the chain
// recoverRate -> decode -> checkCrcAndMerge (ldpc.py:1365-1418, 1610-1619; harq.py:165-173).
//
// Mapping (B200: 148 SMs, 227 KB shared memory / CTA, no tensor cores -- the work is not a contraction):
//   * one THREAD per lifted check: thread (cb, m) owns check m of EVERY layer of code block cb.  A CTA hosts
//     floor(384 / Zc) code blocks (1 at Zc >= 193), persistent over code-block groups.
//   * posteriors of the `ncore` = k+4 columns of degree > 1 live in shared memory ([cb][col][Zc], conflict-free:
//     consecutive m hit consecutive words (m + s) mod Zc).  Circulant shifts are index arithmetic only.
//   * everything else is THREAD-PRIVATE and never needs a barrier: the posterior of the degree-1 extension-parity column
//     of row i (its circulant is the identity, so lifted position m belongs to thread m) and the compressed
//     check-to-variable messages (alpha*min1, alpha*min2, sign bits + argmin).  They sit in per-row state planes,
//     in shared memory for as many rows as fit and in an L2-resident scratch for the rest.
//   * per layer: gather t_j = r - old message, two-min/sign/argmin over the <= 19 edges in registers (row bodies are
//     unrolled per degree, the (column, shift) table comes from the constant bank), scatter r = t + new, ONE barrier.
//   * extension rows whose parity LLRs are all zero can never change any other column (their min1 is 0), so the
//     schedule stops at the last row with a non-zero extension LLR; the beliefs of the skipped columns are produced
//     in closed form in the epilogue.  This is exact, not an approximation (tests/test_decode_gpu.py).
//
// Bit-exactness discipline ("a10 formula"): every add/sub/mul is an explicit round-to-nearest intrinsic
// (no FMA contraction), operation order t = r - old; new = (mag*sign)*0.75; r = t + new, sign(+-0) = +, first-index
// argmin, the "+100000" second-minimum quirk, clip to +-1e10.  -0.0 inputs are canonicalised to +0.0 at load, which
// makes the raw sign bit equal to (t < 0) for every t the recursion can produce.
#include <string.h>

#include "nr_bg_tables.h"
#include "nrldpc_internal.cuh"

#ifndef NR_DEC_MIN_CTAS
#define NR_DEC_MIN_CTAS 2   // fp32: cap registers at 80 so that two 384-thread CTAs share an SM
#endif

namespace {

// decoder view of the lifted graph: byte offsets instead of (column, shift), see process_row
struct __align__(16) NrDecGraph {
    int P, ncols, ksys, ncore, Z;
    uint32_t S;                // ceil(2^32 / Z): lifted positions are tracked as 32-bit fixed-point fractions of Z
    uint32_t one;              // 1, opaque to the compiler: keeps the column-base add an IMAD (FMA pipe) instead of an ALU add
    int pad[1];
    uint16_t rowEdge0[NR_MAX_ROWS + 2];
    uint2 tab[NR_MAX_EDGES];   // x = (shift * S) mod 2^32, y = col*Z*sizeof(T)
};

// per-thread "argmin so far" record of a row pass: written with a predicated 64-bit (128-bit for fp64) shared-memory
// store whenever a new strict minimum appears, read back once after the pass (LSU work instead of two ALU selects per edge)
template <typename T>
struct MinSlot;
template <>
struct __align__(8) MinSlot<float> {
    float t;
    uint32_t off;
};
template <>
struct __align__(16) MinSlot<double> {
    double t;
    uint32_t off, pad;
};
// The record is written and read with inline PTX only, so that the compiler neither forwards it through registers
// (which would bring the two selects per edge back as predicated moves) nor orders it against the posterior traffic.
__device__ __forceinline__ void slot_init(uint32_t sa, float t, uint32_t off)
{
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(sa), "r"(__float_as_uint(t)), "r"(off));
}
__device__ __forceinline__ void slot_update(uint32_t sa, float a, float min1, float t, uint32_t off)
{
    asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; @p st.shared.v2.b32 [%2], {%3, %4};}" ::"f"(a), "f"(min1), "r"(sa),
                 "r"(__float_as_uint(t)), "r"(off));
}
__device__ __forceinline__ MinSlot<float> slot_read(uint32_t sa, float)
{
    MinSlot<float> r;
    uint32_t tb;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(tb), "=r"(r.off) : "r"(sa));
    r.t = __uint_as_float(tb);
    return r;
}
__device__ __forceinline__ void slot_init(uint32_t sa, double t, uint32_t off)
{
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(sa), "l"(__double_as_longlong(t)), "l"((long long)off));
}
__device__ __forceinline__ void slot_update(uint32_t sa, double a, double min1, double t, uint32_t off)
{
    asm volatile("{.reg .pred p; setp.lt.f64 p, %0, %1; @p st.shared.v2.b64 [%2], {%3, %4};}" ::"d"(a), "d"(min1), "r"(sa),
                 "l"(__double_as_longlong(t)), "l"((long long)off));
}
__device__ __forceinline__ MinSlot<double> slot_read(uint32_t sa, double)
{
    MinSlot<double> r;
    long long tb, ob;
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(tb), "=l"(ob) : "r"(sa));
    r.t = __longlong_as_double(tb);
    r.off = (uint32_t)ob;
    r.pad = 0;
    return r;
}

// ---------------------------------------------------------------------------------------------------------------
// exact arithmetic helpers
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
struct FP;
template <>
struct FP<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
    static __device__ __forceinline__ float from_u32(uint32_t v) { return __uint_as_float(v); }
    static __device__ __forceinline__ uint32_t to_u32(float v) { return __float_as_uint(v); }
    static __device__ __forceinline__ uint32_t sign(float a) { return __float_as_uint(a) >> 31; }
    static __device__ __forceinline__ float flip(float mag, uint32_t bit)
    {
        return __uint_as_float(__float_as_uint(mag) ^ (bit << 31));
    }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ uint32_t hibits(float a) { return __float_as_uint(a); }
    static __device__ __forceinline__ void opaque(float& a) { asm volatile("" : "+f"(a)); }
    // mag with its sign flipped when bit 31 of `w` is set (the other bits of w are ignored)
    static __device__ __forceinline__ float flipbits(float mag, uint32_t w)
    {
        return __uint_as_float(__float_as_uint(mag) ^ (w & 0x80000000u));
    }
};
template <>
struct FP<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ double mn(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double mx(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ double from_u32(uint32_t v) { return __hiloint2double(0, (int)v); }
    static __device__ __forceinline__ uint32_t to_u32(double v) { return (uint32_t)__double2loint(v); }
    static __device__ __forceinline__ uint32_t sign(double a) { return ((uint32_t)__double2hiint(a)) >> 31; }
    static __device__ __forceinline__ double flip(double mag, uint32_t bit)
    {
        return __hiloint2double(__double2hiint(mag) ^ (int)(bit << 31), __double2loint(mag));
    }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ uint32_t hibits(double a) { return (uint32_t)__double2hiint(a); }
    static __device__ __forceinline__ void opaque(double& a) { asm volatile("" : "+d"(a)); }
    static __device__ __forceinline__ double flipbits(double mag, uint32_t w)
    {
        return __hiloint2double(__double2hiint(mag) ^ (int)(w & 0x80000000u), __double2loint(mag));
    }
};

// ---------------------------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------------------------
struct DecArgs {
    // batch
    long long numCb;
    int cbPerCta;       // code blocks hosted by one CTA
    int numIter;
    int flags;
    int numRows;        // rows scheduled (>= 4); rows >= numRows have all-zero extension LLRs
    int tmemRows;       // rows [0, tmemRows): state in Tensor Memory (ONE_CB kernels only)
    int tmemCols;       // TMEM columns to allocate (power of two >= 32), 0 = none
    int smemRows;       // next smemRows rows: state planes in shared memory; the rest go to `scratch`
    int outCols;        // columns written to bits / beliefs
    // mode A: rate-recovered input
    const void* llr;
    long long llrStride;
    int inCols;
    int inF64;          // element type of `llr` (compute type T is the kernel's template parameter)
    // mode B: fused rate recovery (rm != 0)
    int rm;
    int K, F, C, qm, ncb, k0, E0, nShort, fStep;   // per-TB split: first nShort blocks have E0, the rest E0+fStep
    long long llrLen;   // valid LLRs per TB
    void* softBuf;      // NULL or [numCb, ncb-F]
    // outputs
    signed char* bits;
    long long bitsStride;
    void* beliefs;
    int* iters;
    // fused CRC / merge (mode B)
    signed char* tbBits;
    long long tbBitsStride;
    unsigned char* cbCrcOk;
    unsigned int* cbRemA;    // per-CB CRC24A remainder of its payload (combined per TB by a second kernel)
    // overflow state
    void* scratch;
    unsigned int* workCounter;
};

// per-row thread-private state planes (SoA: plane p of row slot s = base + (4 s + p) * nThreads elements of T)
enum { PL_M1 = 0, PL_M2 = 1, PL_SW = 2, PL_REXT = 3, NPLANES = 4 };

template <typename T>
struct RowState {   // register copy of one check's state
    T m1s, m2s, rext;
    uint32_t sw;
};

// plane access through a pointer whose address space (shared / global) is known at the call site
template <typename T>
__device__ __forceinline__ void load_state(RowState<T>& st, const T* base, int nT)
{
    st.m1s = base[(size_t)PL_M1 * nT];
    st.m2s = base[(size_t)PL_M2 * nT];
    st.sw = *reinterpret_cast<const uint32_t*>(base + (size_t)PL_SW * nT);
    st.rext = base[(size_t)PL_REXT * nT];
}
template <typename T>
__device__ __forceinline__ void store_state(const RowState<T>& st, T* base, int nT)
{
    base[(size_t)PL_M1 * nT] = st.m1s;
    base[(size_t)PL_M2 * nT] = st.m2s;
    *reinterpret_cast<uint32_t*>(base + (size_t)PL_SW * nT) = st.sw;
    base[(size_t)PL_REXT * nT] = st.rext;
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor Memory as thread-private state storage (B200: 256 KB / SM next to the 227 KB of shared memory).
// The per-check state is touched by exactly one thread, once per iteration, and never needs a barrier -- it only
// needs CAPACITY.  TMEM is addressed as 128 lanes x 512 columns of 32 bits; with the 32x32b access shape a warp
// reads/writes, for each of its 32 threads, consecutive columns of the lane (warp % 4) * 32 + laneid.  Row slot s of
// warp w therefore lives in columns base + (s * warpsPerQuad + w / 4) * RW .. + RW-1 of the warp's lane quadrant,
// RW = 4 words (fp32 state) or 8 (fp64).  This frees ~100 KB of shared memory per code block, which is what lets
// two BG1/Zc=384 code blocks share one SM.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld(RowState<float>& st, uint32_t taddr)
{
    uint32_t a, b, c, d;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    st.m1s = __uint_as_float(a); st.m2s = __uint_as_float(b); st.sw = c; st.rext = __uint_as_float(d);
}
__device__ __forceinline__ void tmem_st(const RowState<float>& st, uint32_t taddr)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(st.m1s)),
                 "r"(__float_as_uint(st.m2s)), "r"(st.sw), "r"(__float_as_uint(st.rext)) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld(RowState<double>& st, uint32_t taddr)
{
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    st.m1s = __hiloint2double((int)r[1], (int)r[0]);
    st.m2s = __hiloint2double((int)r[3], (int)r[2]);
    st.rext = __hiloint2double((int)r[5], (int)r[4]);
    st.sw = r[6];
}
__device__ __forceinline__ void tmem_st(const RowState<double>& st, uint32_t taddr)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"((uint32_t)__double2loint(st.m1s)), "r"((uint32_t)__double2hiint(st.m1s)),
                 "r"((uint32_t)__double2loint(st.m2s)), "r"((uint32_t)__double2hiint(st.m2s)),
                 "r"((uint32_t)__double2loint(st.rext)), "r"((uint32_t)__double2hiint(st.rext)), "r"(st.sw), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------------------------
// one layer for one lifted check.  D = row degree, EXT = last edge is the thread-private extension column.
//
// State of a check between iterations: alpha*min1, alpha*min2, the sign bits of its D messages (bit D-1-j = edge j)
// and the shared-memory byte offset of the edge that received min2 (the argmin).  EXT rows keep the offset in bits
// 12.. of the sign word, core rows (D = 19, no private column) in the otherwise unused `rext` word.
//
// Pipe budget per edge (measured on B200, scripts/pipe_ubench.cu: the ALU pipe issues LOP3/SHF/SEL/ISETP/FSETP every
// 2nd clock per SM sub-partition, 2-input FMNMX every clock; FADD/FMUL run every clock and IMAD every 2nd on the FMA
// pipe; the kernel was ALU-pipe bound, so work is moved off that pipe wherever arithmetic allows):
//   address  : w = m*S + shift*S (IMAD) is the lifted position (m + shift) mod Z as a 32-bit fixed-point fraction --
//              the wrap-around is the integer overflow; byte offset = hi32(w * Z*sizeof(T)) + column base (IMAD.HI).
//              No compare/select, nothing on the ALU pipe.
//   gather   : LDS
//   old msg  : (offset == old argmin offset ? m2 : m1) ^ (sign bit moved to bit 31), FADD
//   signs    : one funnel shift collects the sign bit of t
//   two-min  : min1/min2 VALUES by three FMNMX; the argmin (signed t and offset) is not tracked in registers: a
//              predicated STS.64 drops it into the thread's MinSlot whenever |t| < min1 (strict: first minimum)
//   new msg  : every edge gets m1' ^ sign(t) (LOP3, FADD, STS); afterwards the argmin edge alone is re-written with
//              m2' from the MinSlot record -- no per-edge index compare/select.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lifted_offset(uint32_t m, uint32_t S, uint32_t ZB, uint32_t one, uint2 tb)
{
    // (a multiply-high WITH addend needs a zeroed even/odd register pair in SASS: two extra moves per edge)
    uint32_t w, p, off;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(w) : "r"(m), "r"(S), "r"(tb.x));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(p) : "r"(w), "r"(ZB));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(off) : "r"(p), "r"(one), "r"(tb.y));
    return off;
}

template <typename T, int D, bool EXT>
__device__ __forceinline__ void process_row(const NrDecGraph& g, int e0, char* __restrict__ rb, uint32_t m,
                                            uint32_t ZB, RowState<T>& st, uint32_t slot, uint32_t dummyOff)
{
    constexpr int OFF_SHIFT = 12;   // EXT rows: D <= 10 sign bits, then the argmin offset
    T t[D];
    uint32_t off[D];
    T m1s = st.m1s, m2s = st.m2s;
    const uint32_t sw = st.sw;
    const uint32_t oldOff = EXT ? (sw >> OFF_SHIFT) : FP<T>::to_u32(st.rext);
    const uint32_t S = g.S, one = g.one;
    T min1 = (T)0, min2 = FP<T>::inf();
    uint32_t nsw = 0;
#pragma unroll
    for (int j = 0; j < D; j++) {
        T rv;
        if (EXT && j == D - 1) {
            rv = st.rext;
            off[j] = dummyOff;
        } else {
            off[j] = lifted_offset(m, S, ZB, one, g.tab[e0 + j]);
            rv = *reinterpret_cast<const T*>(rb + off[j]);
        }
        {   // in the first iteration the state is all zero: r - (+0) == r exactly
            const T mag = (off[j] == oldOff) ? m2s : m1s;
            t[j] = FP<T>::sub(rv, FP<T>::flipbits(mag, sw << (31 - (D - 1 - j))));
        }
        const T a = FP<T>::abs(t[j]);
        nsw = __funnelshift_l(FP<T>::hibits(t[j]), nsw, 1);   // (nsw << 1) | sign(t_j)
        if (j == 0) {
            min1 = a;
            slot_init(slot, t[j], off[j]);
        } else {
            slot_update(slot, a, min1, t[j], off[j]);   // strict a < min1: keeps the FIRST minimum (np.argmin)
            min2 = FP<T>::mn(min2, FP<T>::mx(min1, a));
            min1 = FP<T>::mn(min1, a);
        }
    }
    const MinSlot<T> best = slot_read(slot, (T)0);
    // the reference bumps the signed minimum by 1e5 and takes |.| before searching the second minimum (ldpc.py:1563)
    min2 = FP<T>::mn(min2, FP<T>::abs(FP<T>::add(best.t, (T)100000)));
    const uint32_t par = __popc(nsw) & 1u;
    const uint32_t msw = par ? (~nsw & ((1u << D) - 1u)) : nsw;   // sign of new message j = sign_j * parity
    m1s = FP<T>::mul(min1, (T)0.75);
    m2s = FP<T>::mul(min2, (T)0.75);
    // parity folded into the two candidates by an exact multiplication with +-1 (an XOR here would be re-associated
    // by ptxas into one extra LOP3 per edge)
    const T psign = FP<T>::flip((T)1, par);
    const T m1p = FP<T>::mul(m1s, psign), m2p = FP<T>::mul(m2s, psign);
    T rext = (T)0;
#pragma unroll
    for (int j = 0; j < D; j++) {
        const T nv = FP<T>::add(t[j], FP<T>::flipbits(m1p, FP<T>::hibits(t[j])));
        if (EXT && j == D - 1)
            rext = nv;
        else
            *reinterpret_cast<T*>(rb + off[j]) = nv;
    }
    {   // the argmin edge takes the second minimum (program order after the generic store to the same word)
        const T nv = FP<T>::add(best.t, FP<T>::flipbits(m2p, FP<T>::hibits(best.t)));
        *reinterpret_cast<T*>(rb + best.off) = nv;   // lands in the thread's dummy word when the argmin is private
        if (EXT) rext = (best.off == dummyOff) ? nv : rext;
    }
    st.m1s = m1s;
    st.m2s = m2s;
    if (EXT) {
        st.sw = msw | (best.off << OFF_SHIFT);
        st.rext = rext;
    } else {
        st.sw = msw;
        st.rext = FP<T>::from_u32(best.off);
    }
}

template <typename T>
__device__ __forceinline__ void dispatch_row(const NrDecGraph& g, int row, char* rb, uint32_t m, uint32_t ZB,
                                             RowState<T>& st, uint32_t slot, uint32_t dummyOff)
{
    const int e0 = g.rowEdge0[row];
    const int deg = g.rowEdge0[row + 1] - e0;
    if (row >= 4) {
        switch (deg) {
            case 3: process_row<T, 3, true>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
            case 4: process_row<T, 4, true>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
            case 5: process_row<T, 5, true>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
            case 6: process_row<T, 6, true>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
            case 7: process_row<T, 7, true>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
            case 8: process_row<T, 8, true>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
            case 9: process_row<T, 9, true>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
            default: process_row<T, 10, true>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
        }
    } else {
        switch (deg) {
            case 8: process_row<T, 8, false>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
            case 10: process_row<T, 10, false>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
            default: process_row<T, 19, false>(g, e0, rb, m, ZB, st, slot, dummyOff); break;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Static schedule (fp32, one block per CTA): the rows of the base graph are unrolled at compile time, so the edge
// table entries are constant-bank operands of the address IMADs (no LDC, no degree dispatch) and the Tensor-Memory
// address of a row's state is an immediate.  The code of one iteration is ~90 KB for BG1; all warps of an SM walk it
// in step (one barrier per row), so it streams through the instruction cache once per iteration.
// ---------------------------------------------------------------------------------------------------------------
template <int BG>
struct BgRows {
    static constexpr int P = BG == 1 ? NR_BG1_ROWS : NR_BG2_ROWS;
    static __host__ __device__ constexpr int deg(int r) { return BG == 1 ? NR_BG1_ROW_DEG[r] : NR_BG2_ROW_DEG[r]; }
    static __host__ __device__ constexpr int e0(int r)
    {
        int e = 0;
        for (int i = 0; i < r; i++) e += deg(i);
        return e;
    }
};

template <typename T, int BG, int ROW, typename Store>
__device__ __forceinline__ void run_rows_static(const NrDecGraph& g, int numRows, char* rb, uint32_t m, uint32_t ZB,
                                                const Store& store, uint32_t slot, uint32_t dummyOff)
{
    if constexpr (ROW < BgRows<BG>::P) {
        if (ROW >= 4 && ROW >= numRows) return;   // numRows >= 4 always
        constexpr int D = BgRows<BG>::deg(ROW);
        constexpr int E0 = BgRows<BG>::e0(ROW);
        RowState<T> st;
        store.load(ROW, st);
        process_row<T, D, (ROW >= 4)>(g, E0, rb, m, ZB, st, slot, dummyOff);
        store.store(ROW, st);
        __syncthreads();
        run_rows_static<T, BG, ROW + 1>(g, numRows, rb, m, ZB, store, slot, dummyOff);
    }
}

// posterior addressed by edge `e` for lifted index m
template <typename T>
__device__ __forceinline__ T edge_posterior(const NrDecGraph& g, int e, const char* rb, uint32_t m, uint32_t ZB)
{
    return *reinterpret_cast<const T*>(rb + lifted_offset(m, g.S, ZB, g.one, g.tab[e]));
}

// ---------------------------------------------------------------------------------------------------------------
// GF(2) helpers for the fused CRC
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t gf_shift1(uint32_t r, uint32_t poly, int c)
{
    const uint32_t top = (r >> (c - 1)) & 1u;
    r = (r << 1) & ((1u << c) - 1u);
    return top ? (r ^ poly) : r;
}
__device__ __forceinline__ uint32_t gf_mulmod(uint32_t a, uint32_t b, uint32_t poly, int c)
{
    uint32_t r = 0;
    for (int i = c - 1; i >= 0; i--) {
        r = gf_shift1(r, poly, c);
        if ((b >> i) & 1u) r ^= a;
    }
    return r;
}

// CRC remainder of `len` hard-decision bits of one code block, cooperatively by its Z threads.
// Bit i is the sign of posterior i of the block (core columns are contiguous in shared memory).  The message is
// right-aligned in Z chunks of B bits (leading zeros do not change a zero-initialised CRC); per-thread remainders are
// merged pairwise, rem = left * x^(B*span) + right, with the factors x^(B*2^l) mod g precomputed in fac[].
// `tree` is per-block scratch of P2 = nextPow2(Z) words.  Every thread of the CTA must call this (barriers inside).
template <typename T>
__device__ uint32_t cb_crc(const T* rcb, int len, int Z, int P2, int m, bool active, uint32_t* tree,
                           const uint32_t* fac, uint32_t poly, int c)
{
    const int B = (len + Z - 1) / Z;
    const int lead = B * Z - len;
    if (active) {
        uint32_t rem = 0;
        const int i0 = m * B - lead;
        for (int b = 0; b < B; b++) {
            const int i = i0 + b;
            const uint32_t bit = (i >= 0) ? FP<T>::sign(rcb[i]) : 0u;
            const uint32_t fb = ((rem >> (c - 1)) & 1u) ^ bit;
            rem = (rem << 1) & ((1u << c) - 1u);
            if (fb) rem ^= poly;
        }
        tree[(P2 - Z) + m] = rem;
        if (m < P2 - Z) tree[m] = 0;   // virtual leading chunks
    }
    __syncthreads();
    int lvl = 0;
    for (int span = 1; span < P2; span <<= 1, lvl++) {
        const int right = (m + 1) * 2 * span - 1;
        if (active && right < P2) tree[right] = gf_mulmod(tree[right - span], fac[lvl], poly, c) ^ tree[right];
        __syncthreads();
    }
    return active ? tree[P2 - 1] : 0u;
}

// fac[l] = x^(B * 2^l) mod g for l = 0..nl-1, written by the first nl threads
__device__ __forceinline__ void crc_factors(uint32_t* fac, int len, int Z, int P2, uint32_t poly, int c, int tid)
{
    const int B = (len + Z - 1) / Z;
    int nl = 0;
    for (int span = 1; span < P2; span <<= 1) nl++;
    if (tid < nl) {
        uint32_t f = 1;
        for (int b = 0; b < B; b++) f = gf_shift1(f, poly, c);
        for (int l = 0; l < tid; l++) f = gf_mulmod(f, f, poly, c);
        fac[tid] = f;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// three-tier state storage: rows [0, tmemRows) in Tensor Memory (ONE_CB kernels), the next smemRows rows in shared
// memory planes, the rest in the per-CTA global scratch (stays in L2).  All branches are on the (uniform) row index.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, bool ONE_CB>
struct StateStore {
    uint32_t tbase;     // this thread's TMEM address of row slot 0 (lane quadrant and warp column offset folded in)
    uint32_t tstride;   // TMEM columns per row slot
    T* sS;              // shared planes, already offset by tid
    T* sG;              // global planes, already offset by tid
    int tmemRows, smemRows, nT;
    __device__ __forceinline__ void load(int row, RowState<T>& st) const
    {
        if (ONE_CB && row < tmemRows)
            tmem_ld(st, tbase + (uint32_t)row * tstride);
        else if (row < tmemRows + smemRows)
            load_state(st, sS + (size_t)(row - tmemRows) * NPLANES * nT, nT);
        else
            load_state(st, sG + (size_t)(row - tmemRows - smemRows) * NPLANES * nT, nT);
    }
    __device__ __forceinline__ void store(int row, const RowState<T>& st) const
    {
        if (ONE_CB && row < tmemRows)
            tmem_st(st, tbase + (uint32_t)row * tstride);
        else if (row < tmemRows + smemRows)
            store_state(st, sS + (size_t)(row - tmemRows) * NPLANES * nT, nT);
        else
            store_state(st, sG + (size_t)(row - tmemRows - smemRows) * NPLANES * nT, nT);
    }
};

// one input LLR, widened / narrowed to the compute type
template <typename T>
__device__ __forceinline__ T load_llr(const void* p, long long i, int f64)
{
    return f64 ? (T) reinterpret_cast<const double*>(p)[i] : (T) reinterpret_cast<const float*>(p)[i];
}

// ---------------------------------------------------------------------------------------------------------------
// the kernel.  ONE_CB: exactly one code block per CTA and blockDim.x == Z (Z a multiple of 32): no thread is ever
// idle, so the row bodies run in convergent code and the (column, shift) table is read through the uniform datapath.
// ---------------------------------------------------------------------------------------------------------------
template <typename T, bool ONE_CB, int SBG>
__global__ void __launch_bounds__(384, (sizeof(T) == 4 ? NR_DEC_MIN_CTAS : 1))
    nr_decode_kernel(const __grid_constant__ NrDecGraph g, const __grid_constant__ DecArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int Z = g.Z;
    const int ncore = g.ncore;
    const int nT = blockDim.x;
    const int tid = threadIdx.x;
    const int cbl = ONE_CB ? 0 : tid / Z;            // local code block
    const int m = ONE_CB ? tid : tid - cbl * Z;      // lifted check / position
    const bool lane_ok = ONE_CB ? true : (cbl < a.cbPerCta);
    const int cbPerCta = ONE_CB ? 1 : a.cbPerCta;

    T* rs = reinterpret_cast<T*>(smemRaw);                                   // [cbPerCta][ncore][Z]
    T* stateS = rs + (size_t)cbPerCta * ncore * Z;                           // [smemRows][NPLANES][nT]
    uint32_t* misc = reinterpret_cast<uint32_t*>(stateS + (size_t)a.smemRows * NPLANES * nT);
    // misc: [0, flagsLen) per-block flags | 32 words CRC factors (2 x 16) | per-block CRC trees
    const int flagsLen = (cbPerCta + 31) & ~31;
    uint32_t* fac = misc + flagsLen;
    int P2 = 1;
    while (P2 < Z) P2 <<= 1;
    uint32_t* tree = misc + flagsLen + 32 + (size_t)cbl * P2;
    // per-thread argmin record + dummy word (16-byte aligned region after the CRC trees)
    const size_t slotOfs = ((size_t)(reinterpret_cast<unsigned char*>(misc + flagsLen + 32 + (size_t)cbPerCta * P2) - smemRaw) + 15) & ~(size_t)15;
    MinSlot<T>* slotP = reinterpret_cast<MinSlot<T>*>(smemRaw + slotOfs) + tid;
    const uint32_t slot = (uint32_t)__cvta_generic_to_shared(slotP);
    T* dummyW = reinterpret_cast<T*>(slotP - tid + nT) + tid;
    const int globRows = a.numRows - a.tmemRows - a.smemRows;
    T* stateG = reinterpret_cast<T*>(a.scratch) + (size_t)blockIdx.x * (size_t)globRows * NPLANES * nT;
    T* rcb = rs + (size_t)cbl * ncore * Z;
    // Tensor Memory for the thread-private row state (see tmem_ld above)
    __shared__ uint32_t tmemBaseSh;
    const bool useTmem = ONE_CB && a.tmemCols > 0;
    if (useTmem) {
        if (tid < 32) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&tmemBaseSh);
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"((uint32_t)a.tmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    StateStore<T, ONE_CB> store;
    {
        const int warp = tid >> 5;
        const uint32_t RW = sizeof(T) == 4 ? 4u : 8u;
        const uint32_t wpq = (uint32_t)((nT >> 5) + 3) >> 2;              // warps per lane quadrant
        store.tstride = wpq * RW;
        store.tbase = useTmem ? (tmemBaseSh + ((uint32_t)(warp & 3) << 21) + (uint32_t)(warp >> 2) * RW) : 0u;   // lane (warp%4)*32 in bits 31..16
        store.sS = stateS + tid;
        store.sG = stateG + tid;
        store.tmemRows = ONE_CB ? a.tmemRows : 0;
        store.smemRows = a.smemRows;
        store.nT = nT;
    }
    char* rb = reinterpret_cast<char*>(rcb);
    const uint32_t mU = (uint32_t)m;
    const uint32_t ZB = (uint32_t)Z * (uint32_t)sizeof(T);
    const uint32_t dummyOff = (uint32_t)(reinterpret_cast<char*>(dummyW) - rb);
    const int ksys = g.ksys;

    const bool wantCrc = a.rm && (a.tbBits || a.cbCrcOk || a.cbRemA);
    const int Lk = a.K - a.F;                       // code block without fillers
    const int per = (a.C > 1) ? Lk - 24 : Lk;       // payload copied into the merged transport block
    const NrCrcPoly polyCb = nr_crc_poly(a.C > 1 ? NRLDPC_CRC24B : NRLDPC_CRC24A);
    const NrCrcPoly polyA = nr_crc_poly(NRLDPC_CRC24A);
    if (wantCrc) {
        crc_factors(fac, Lk, Z, P2, polyCb.poly, polyCb.len, tid);
        if (a.C > 1) crc_factors(fac + 16, per, Z, P2, polyA.poly, polyA.len, tid);
    }

    const long long numGroups = (a.numCb + cbPerCta - 1) / cbPerCta;
    for (long long grp = blockIdx.x; grp < numGroups; grp += gridDim.x) {
        const long long cb = grp * cbPerCta + cbl;
        const bool active = ONE_CB ? true : (lane_ok && cb < a.numCb);

        // -------------------------------------------------------------------------------------------------------
        // load phase: column block `col` (un-punctured index), position m.  Punctured columns 0,1 start at 0.
        // -------------------------------------------------------------------------------------------------------
        if (active) {
            rcb[m] = (T)0;
            rcb[Z + m] = (T)0;
            const int lastCol = ksys + a.numRows;   // exclusive; numRows >= 4
            int E = 0, L = 0, sysLen = 0, Eq = 1;
            long long xBase = 0, xAvail = 0;
            T* sb = nullptr;
            if (a.rm) {
                const long long tb = cb / a.C;
                const int r = (int)(cb - tb * a.C);
                E = a.E0 + (r >= a.nShort ? a.fStep : 0);
                const long long off = (long long)r * a.E0 + (long long)(r > a.nShort ? (r - a.nShort) : 0) * a.fStep;
                xBase = tb * a.llrStride + off;
                xAvail = a.llrLen - off;   // LLRs actually present for this block (rest are zeros, ldpc.py:1402)
                L = a.ncb - a.F;
                sysLen = a.K - a.F - 2 * Z;
                Eq = E / a.qm;
                if (a.softBuf) sb = reinterpret_cast<T*>(a.softBuf) + cb * (long long)L;
            }
            const int colEnd = (a.rm && sb) ? g.ncols : lastCol;   // a soft buffer is combined over its whole length
            // de-interleaver division i / Eq: float reciprocal + one correction step (exact for i < 2^24)
            const bool smallE = E < (1 << 24);
            const float rcpEq = 1.0f / (float)Eq;
            const int xAvailI = (int)(xAvail < 0 ? 0 : (xAvail > (long long)E ? (long long)E : xAvail));
            auto load_cols = [&](auto tin) {
                using TIn = decltype(tin);
                const TIn* __restrict__ x = reinterpret_cast<const TIn*>(a.llr) + (a.rm ? xBase : cb * a.llrStride);
                int n = m;   // index in the punctured coded block
                for (int col = 2; col < colEnd; col++, n += Z) {
                    T v = (T)0;
                    if (!a.rm) {
                        if (col - 2 < a.inCols) v = (T)x[n];
                    } else if (n < a.ncb) {
                        if (n >= sysLen && n < sysLen + a.F) {
                            v = (T)1e20;   // filler: LARGE_LLR (chancodebase.py:52), clipped below like any input
                        } else {
                            const int q = (n < sysLen) ? n : n - a.F;   // index in the filler-less circular buffer
                            T acc = sb ? sb[q] : (T)0;
                            int i = q - a.k0;
                            if (i < 0) i += L;
                            for (; i < E; i += L) {       // one term per wrap, ascending => the reference's += order
                                int b;                    // de-interleave: stream index s*qm + b, i = b*Eq + s
                                if (smallE) {
                                    b = (int)((float)i * rcpEq);
                                    const int r = i - b * Eq;
                                    b += (r >= Eq) ? 1 : 0;
                                    b -= (r < 0) ? 1 : 0;
                                } else {
                                    b = i / Eq;
                                }
                                const int xi = (i - b * Eq) * a.qm + b;
                                const T xv = (xi < xAvailI) ? (T)x[xi] : (T)0;
                                acc = FP<T>::add(acc, xv);
                            }
                            if (sb) sb[q] = acc;
                            v = acc;
                        }
                    }
                    if (col >= lastCol) continue;           // beyond the scheduled rows: only the soft buffer is updated
                    v = (v > (T)1e10) ? (T)1e10 : v;        // np.clip(., -1e10, 1e10), ldpc.py:1536
                    v = (v < (T)-1e10) ? (T)-1e10 : v;
                    v = FP<T>::add(v, (T)0);                 // -0.0 -> +0.0 (see header)
                    if (col < ncore) {
                        rcb[col * Z + m] = v;
                    } else {
                        RowState<T> st0;   // messages start at +0 (ldpc.py:1543), posterior of the extension column = its LLR
                        st0.m1s = (T)0; st0.m2s = (T)0; st0.sw = 0; st0.rext = v;
                        store.store(col - ksys, st0);
                    }
                }
            };
            if (a.rm && !sb && !a.inF64 && smallE) {
                // common case (fp32 stream, no HARQ history): same arithmetic, none of the generic bookkeeping
                const float* __restrict__ x = reinterpret_cast<const float*>(a.llr) + xBase;
                const int ncb = a.ncb, F = a.F, k0 = a.k0, qm = a.qm;
                // stream index of circular-buffer position i (de-interleaver), i < E
                auto stream_index = [&](int i) {
                    int b = (int)((float)i * rcpEq);
                    const int r = i - b * Eq;
                    b += (r >= Eq) ? 1 : 0;
                    b -= (r < 0) ? 1 : 0;
                    return (i - b * Eq) * qm + b;
                };
                constexpr int CH = 8;   // columns in flight: the HBM latency of the gather is paid once per chunk
                for (int col0 = 2; col0 < lastCol; col0 += CH) {
                    T v[CH];
                    int inext[CH];
#pragma unroll
                    for (int c = 0; c < CH; c++) {
                        const int n = (col0 + c - 2) * Z + m;
                        v[c] = (T)0;
                        inext[c] = E;   // nothing more to add
                        if (col0 + c < lastCol && n < ncb) {
                            const int nf = n - sysLen;   // >= 0: at or behind the filler gap
                            if ((unsigned)nf < (unsigned)F) {
                                v[c] = (T)1e10;          // LARGE_LLR after the clip
                            } else {
                                int i = n - (nf >= 0 ? F : 0) - k0;
                                i += (i < 0) ? L : 0;
                                if (i < E) {
                                    const int xi = stream_index(i);
                                    v[c] = (xi < xAvailI) ? (T)x[xi] : (T)0;   // 0 + x == x exactly
                                    inext[c] = i + L;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int c = 0; c < CH; c++) {
                        if (col0 + c < lastCol) {
                            T acc = v[c];
                            for (int i = inext[c]; i < E; i += L) {   // further wraps (E > Ncb - F), ascending order
                                const int xi = stream_index(i);
                                acc = FP<T>::add(acc, (xi < xAvailI) ? (T)x[xi] : (T)0);
                            }
                            acc = FP<T>::mn(acc, (T)1e10);
                            acc = FP<T>::mx(acc, (T)-1e10);
                            acc = FP<T>::add(acc, (T)0);   // -0.0 -> +0.0 (see header)
                            const int col = col0 + c;
                            if (col < ncore) {
                                rcb[col * Z + m] = acc;
                            } else {
                                RowState<T> st0;
                                st0.m1s = (T)0; st0.m2s = (T)0; st0.sw = 0; st0.rext = acc;
                                store.store(col - ksys, st0);
                            }
                        }
                    }
                }
            } else if (a.inF64) {
                load_cols(double());
            } else {
                load_cols(float());
            }
            for (int row = 0; row < 4; row++) {
                RowState<T> st0;
                st0.m1s = (T)0; st0.m2s = (T)0; st0.sw = 0; st0.rext = (T)0;
                store.store(row, st0);
            }
        }
        __syncthreads();

        // -------------------------------------------------------------------------------------------------------
        // iterations
        // -------------------------------------------------------------------------------------------------------
        int itersDone = 0;
        bool cbDone = false;
        for (int it = 0; it < a.numIter; it++) {
            if constexpr (SBG != 0) {
                run_rows_static<T, SBG, 0>(g, a.numRows, rb, mU, ZB, store, slot, dummyOff);
            } else {
                for (int row = 0; row < a.numRows; row++) {
                    if (ONE_CB || (active && !cbDone)) {
                        RowState<T> st;
                        store.load(row, st);
                        dispatch_row<T>(g, row, rb, mU, ZB, st, slot, dummyOff);
                        store.store(row, st);
                    }
                    __syncthreads();
                }
            }
            if (!cbDone) itersDone = it + 1;
            if (a.flags & NRLDPC_DEC_EARLY_STOP) {
                // syndrome of the hard decisions after a COMPLETE iteration over the scheduled rows (skipped rows are
                // satisfied by construction: their parity bit is the parity of the rest)
                uint32_t bad = 0;
                if (active && !cbDone) {
                    for (int row = 0; row < a.numRows; row++) {
                        const int e0 = g.rowEdge0[row];
                        const int e1 = g.rowEdge0[row + 1] - (row >= 4 ? 1 : 0);
                        uint32_t par = 0;
                        for (int e = e0; e < e1; e++) par ^= FP<T>::sign(edge_posterior<T>(g, e, rb, mU, ZB));
                        if (row >= 4) {
                            RowState<T> st;
                            store.load(row, st);
                            par ^= FP<T>::sign(st.rext);
                        }
                        bad |= par;
                    }
                }
                if (tid < cbPerCta) misc[tid] = 0;
                __syncthreads();
                if (bad) misc[cbl] = 1;
                __syncthreads();
                if (lane_ok && misc[cbl] == 0) cbDone = true;
                const int anyLeft = __syncthreads_or((active && !cbDone) ? 1 : 0);
                if (!anyLeft) break;
            }
        }

        // -------------------------------------------------------------------------------------------------------
        // epilogue: hard decisions / beliefs, closed form for skipped extension columns, fused CRC + merge
        // -------------------------------------------------------------------------------------------------------
        if (active) {
            if (a.iters && m == 0) a.iters[cb] = itersDone;
            const int outCore = min(a.outCols, ncore);
            if (a.bits) {
                signed char* o = a.bits + cb * a.bitsStride;
                for (int col = 0; col < outCore; col++) o[col * Z + m] = (signed char)FP<T>::sign(rcb[col * Z + m]);
            }
            if (a.beliefs) {
                T* o = reinterpret_cast<T*>(a.beliefs) + cb * (long long)a.outCols * Z;
                for (int col = 0; col < outCore; col++) o[col * Z + m] = rcb[col * Z + m];
            }
            for (int col = ncore; col < a.outCols; col++) {
                const int row = col - ksys;
                T v;
                if (row < a.numRows) {
                    RowState<T> st;
                    store.load(row, st);
                    v = st.rext;
                } else {
                    // skipped row: t_ext == 0 in every iteration, so its belief after the last iteration is
                    // 0.75 * parity * min(min_j |r_j|, 1e5) over the row's core edges evaluated on the final posteriors
                    const int e0 = g.rowEdge0[row];
                    const int e1 = g.rowEdge0[row + 1] - 1;
                    T mn = (T)100000;
                    uint32_t par = 0;
                    for (int e = e0; e < e1; e++) {
                        const T rv = edge_posterior<T>(g, e, rb, mU, ZB);
                        mn = FP<T>::mn(mn, FP<T>::abs(rv));
                        par ^= FP<T>::sign(rv);
                    }
                    v = (a.numIter > 0) ? FP<T>::flip(FP<T>::mul(mn, (T)0.75), par) : (T)0;
                    v = FP<T>::add(v, (T)0);
                }
                if (a.bits) a.bits[cb * a.bitsStride + col * Z + m] = (signed char)(v < (T)0);
                if (a.beliefs) reinterpret_cast<T*>(a.beliefs)[cb * (long long)a.outCols * Z + col * Z + m] = v;
            }
        }
        if (wantCrc) {
            // checkCrcAndMerge (ldpc.py:1610-1619) on the hard decisions still in shared memory
            const uint32_t remCb = cb_crc<T>(rcb, Lk, Z, P2, m, active, tree, fac, polyCb.poly, polyCb.len);
            __syncthreads();
            uint32_t remA = remCb;
            if (a.C > 1) remA = cb_crc<T>(rcb, per, Z, P2, m, active, tree, fac + 16, polyA.poly, polyA.len);
            if (active) {
                if (m == 0) {
                    if (a.cbCrcOk) a.cbCrcOk[cb] = (remCb == 0);
                    if (a.cbRemA) a.cbRemA[cb] = remA;
                }
                if (a.tbBits) {
                    const long long tb = cb / a.C;
                    const int r = (int)(cb - tb * a.C);
                    signed char* o = a.tbBits + tb * a.tbBitsStride + (long long)r * per;
                    for (int i = m; i < per; i += Z) o[i] = (signed char)FP<T>::sign(rcb[i]);
                }
            }
        }
        __syncthreads();   // shared memory is reused by the next group
    }
    if (useTmem) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBaseSh), "r"((uint32_t)a.tmemCols) : "memory");
    }
}

// combine per-code-block CRC24A remainders into the transport-block check: rem = sum_r rem_r * x^(per*(C-1-r))
__global__ void nr_tb_crc_kernel(const unsigned int* cbRemA, long long numTb, int C, int per, unsigned char* tbOk)
{
    const long long tb = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tb >= numTb) return;
    const NrCrcPoly pa = nr_crc_poly(NRLDPC_CRC24A);
    uint32_t f = 1;   // x^per mod g by square and multiply
    {
        uint32_t base = 2;   // x
        int e = per;
        while (e) {
            if (e & 1) f = gf_mulmod(f, base, pa.poly, pa.len);
            base = gf_mulmod(base, base, pa.poly, pa.len);
            e >>= 1;
        }
    }
    uint32_t rem = 0;
    for (int r = 0; r < C; r++) rem = gf_mulmod(rem, f, pa.poly, pa.len) ^ cbRemA[tb * C + r];
    tbOk[tb] = (rem == 0);
}

// last column (punctured frame) holding a non-zero LLR, max over the batch -> numRows for mode A
template <typename TIn>
__global__ void nr_last_nonzero_kernel(const TIn* llr, long long numCb, long long stride, int len, int Z, int* lastCol)
{
    int best = -1;
    const long long total = numCb * (long long)len;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long cb = i / len;
        const int n = (int)(i - cb * len);
        if (llr[cb * stride + n] != (TIn)0) best = max(best, n / Z);
    }
    for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0 && best >= 0) atomicMax(lastCol, best);
}

// host: byte-offset edge table for compute type T
template <typename T>
void build_dec_graph(const NrGraph& g, NrDecGraph* d)
{
    memset(d, 0, sizeof(*d));
    d->P = g.P; d->ncols = g.ncols; d->ksys = g.ksys; d->ncore = g.ncore; d->Z = g.Z;
    for (int i = 0; i < NR_MAX_ROWS + 2; i++) d->rowEdge0[i] = g.rowEdge0[i];
    d->one = 1;
    d->S = (uint32_t)((0x100000000ULL + (uint64_t)g.Z - 1) / (uint64_t)g.Z);   // ceil(2^32 / Z); Z >= 2
    for (int e = 0; e < g.rowEdge0[g.P]; e++) {
        const uint32_t col = g.edge[e] >> 16, sh = g.edge[e] & 0xffffu;
        d->tab[e].x = (uint32_t)((uint64_t)sh * d->S);   // mod 2^32
        d->tab[e].y = col * g.Z * (uint32_t)sizeof(T);
    }
}

template <typename T>
int launch_decode(nrldpc_handle* h, const NrGraph& g, DecArgs& a, cudaStream_t s)
{
    const int Z = g.Z;
    a.cbPerCta = max(1, 384 / Z);
    if ((long long)a.cbPerCta > a.numCb) a.cbPerCta = (int)a.numCb;
    int nT = a.cbPerCta * Z;
    nT = (nT + 31) & ~31;
    const bool oneCb = (a.cbPerCta == 1 && nT == Z);
    int P2 = 1;
    while (P2 < Z) P2 <<= 1;
    const size_t rBytes = (size_t)a.cbPerCta * g.ncore * Z * sizeof(T);
    const size_t rowBytes = (size_t)NPLANES * nT * sizeof(T);
    const size_t miscBytes = ((size_t)((a.cbPerCta + 31) & ~31) + 32 + (size_t)a.cbPerCta * P2) * sizeof(uint32_t) + 16 +
                             (size_t)nT * (sizeof(MinSlot<T>) + sizeof(T));
    // target resident CTAs per SM (env NRLDPC_DEC_OCC overrides): two for the fp32 one-block-per-CTA kernel, whose
    // registers are capped at 80 and whose row state lives in Tensor Memory; one otherwise
    int occ = h->decOcc > 0 ? h->decOcc : ((oneCb && sizeof(T) == 4) ? 2 : 1);
    occ = max(1, min(occ, 2048 / nT));
    if (sizeof(T) == 8) occ = 1;
    // Tensor Memory rows (ONE_CB kernels): 512 columns per SM shared by the resident CTAs
    a.tmemRows = 0;
    a.tmemCols = 0;
    if (oneCb && !h->noTmem) {
        int cols = 32;
        while (cols * 2 <= 512 / occ) cols *= 2;
        const int RW = sizeof(T) == 4 ? 4 : 8;
        const int wpq = ((nT >> 5) + 3) >> 2;
        int rowsFit = cols / (wpq * RW);
        if (rowsFit > a.numRows) rowsFit = a.numRows;
        if (rowsFit > 0) {
            int need = 32;
            while (need < rowsFit * wpq * RW) need *= 2;
            a.tmemRows = rowsFit;
            a.tmemCols = need;
        }
    }
    const int restRows = a.numRows - a.tmemRows;
    size_t budget = (size_t)h->smemPerSM / occ - 1024;
    budget = min(budget, (size_t)h->maxSmemOptin);
    if (rBytes + miscBytes > budget) {
        occ = 1;
        budget = (size_t)h->maxSmemOptin;
        if (rBytes + miscBytes > budget) {
            nr_set_error("decode: posteriors do not fit shared memory");
            return NRLDPC_ERR_ARG;
        }
    }
    int smemRows = (int)((budget - rBytes - miscBytes) / rowBytes);
    if (smemRows > restRows) smemRows = restRows;
    a.smemRows = smemRows;
    const size_t smem = rBytes + (size_t)smemRows * rowBytes + miscBytes;
    const long long numGroups = (a.numCb + a.cbPerCta - 1) / a.cbPerCta;
    int perSM = (int)((size_t)h->smemPerSM / (smem + 1024));
    perSM = max(1, min(min(perSM, 2048 / nT), occ));
    long long grid = min(numGroups, (long long)h->numSMs * perSM);
    const size_t needScratch = (size_t)grid * (size_t)(restRows - smemRows) * rowBytes;
    if (needScratch > h->scratchBytes) {
        if (h->scratch) NR_CUDA_CHECK(cudaFree(h->scratch));
        h->scratch = nullptr;
        h->scratchBytes = 0;
        NR_CUDA_CHECK(cudaMalloc(&h->scratch, needScratch));
        h->scratchBytes = needScratch;
    }
    a.scratch = h->scratch;
    NrDecGraph dg;
    build_dec_graph<T>(g, &dg);
    const bool staticRows = oneCb && sizeof(T) == 4 && !h->noStaticRows;
    if (staticRows && g.P == NR_BG1_ROWS) {
        auto kern = nr_decode_kernel<T, true, (sizeof(T) == 4 ? 1 : 0)>;
        NR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)grid, nT, smem, s>>>(dg, a);
    } else if (staticRows) {
        auto kern = nr_decode_kernel<T, true, (sizeof(T) == 4 ? 2 : 0)>;
        NR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)grid, nT, smem, s>>>(dg, a);
    } else if (oneCb) {
        auto kern = nr_decode_kernel<T, true, 0>;
        NR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)grid, nT, smem, s>>>(dg, a);
    } else {
        auto kern = nr_decode_kernel<T, false, 0>;
        NR_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)grid, nT, smem, s>>>(dg, a);
    }
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

int dispatch_decode(nrldpc_handle* h, const NrGraph& g, DecArgs& a, int inDtype, int computeDtype, cudaStream_t s)
{
    if (inDtype != NRLDPC_F32 && inDtype != NRLDPC_F64) { nr_set_error("decode: bad input dtype"); return NRLDPC_ERR_ARG; }
    a.inF64 = (inDtype == NRLDPC_F64);
    if (computeDtype == NRLDPC_F32) return launch_decode<float>(h, g, a, s);
    if (computeDtype == NRLDPC_F64) return launch_decode<double>(h, g, a, s);
    nr_set_error("decode: bad compute dtype");
    return NRLDPC_ERR_ARG;
}

}   // namespace

// =================================================================================================================
// C-ABI
// =================================================================================================================
extern "C" int nrldpc_decode(nrldpc_handle* h, int bg, int zc, int in_dtype, int compute_dtype, const void* llr,
                             int64_t num_cb, int64_t llr_stride, int in_cols, int num_iter, int flags, int out_cols,
                             int8_t* bits, void* beliefs, int32_t* iters, nrldpc_stream stream)
{
    if (!h) { nr_set_error("decode: null handle"); return NRLDPC_ERR_ARG; }
    NrGraph g;
    if (nr_build_graph(bg, zc, &g)) return NRLDPC_ERR_ARG;
    if (num_cb <= 0 || in_cols < 0 || in_cols > g.ncols - 2 || out_cols < 1 || out_cols > g.ncols || num_iter < 0 ||
        llr_stride < (int64_t)in_cols * zc) {
        nr_set_error("decode: bad shape (num_cb=%lld in_cols=%d out_cols=%d stride=%lld)", (long long)num_cb, in_cols,
                     out_cols, (long long)llr_stride);
        return NRLDPC_ERR_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    DecArgs a{};
    a.numCb = num_cb;
    a.numIter = num_iter;
    a.flags = flags;
    a.llr = llr;
    a.llrStride = llr_stride;
    a.inCols = in_cols;
    a.outCols = out_cols;
    a.bits = (signed char*)bits;
    a.bitsStride = (long long)out_cols * zc;
    a.beliefs = beliefs;
    a.iters = iters;
    a.numRows = g.P;
    if (!(flags & NRLDPC_DEC_ALL_ROWS)) {
        // exact row skipping: find the last column with any non-zero LLR (one tiny pass over the input, 4 B D2H)
        int* d = reinterpret_cast<int*>(h->workCounter) + 1;
        int init = -1;
        NR_CUDA_CHECK(cudaMemcpyAsync(d, &init, sizeof(int), cudaMemcpyHostToDevice, s));
        const int len = in_cols * zc;
        if (len > 0) {
            const long long total = num_cb * (long long)len;
            const int blocks = (int)min((long long)h->numSMs * 8, (total + 255) / 256);
            if (in_dtype == NRLDPC_F32)
                nr_last_nonzero_kernel<float><<<blocks, 256, 0, s>>>((const float*)llr, num_cb, llr_stride, len, zc, d);
            else
                nr_last_nonzero_kernel<double><<<blocks, 256, 0, s>>>((const double*)llr, num_cb, llr_stride, len, zc, d);
            NR_CUDA_CHECK(cudaGetLastError());
        }
        int last = -1;
        NR_CUDA_CHECK(cudaMemcpyAsync(&last, d, sizeof(int), cudaMemcpyDeviceToHost, s));
        NR_CUDA_CHECK(cudaStreamSynchronize(s));
        const int lastFull = last + 2;                     // un-punctured column index
        int rows = lastFull - g.ksys + 1;                  // row owning that extension column
        a.numRows = max(4, min(g.P, rows));
    }
    return dispatch_decode(h, g, a, in_dtype, compute_dtype, s);
}

extern "C" int nrldpc_decode_tb(nrldpc_handle* h, const nrldpc_tb_config* cfg, int in_dtype, int compute_dtype,
                                const void* llr, int64_t num_tb, int64_t llr_len, int64_t llr_stride,
                                void* soft_buffer, int num_iter, int flags, int8_t* tb_bits, int64_t tb_bits_stride,
                                uint8_t* cb_crc_ok, uint8_t* tb_crc_ok, int32_t* iters, nrldpc_stream stream)
{
    if (!h || !cfg) { nr_set_error("decode_tb: null argument"); return NRLDPC_ERR_ARG; }
    NrGraph g;
    if (nr_build_graph(cfg->bg, cfg->zc, &g)) return NRLDPC_ERR_ARG;
    const int Z = cfg->zc, N = (g.ncols - 2) * Z;
    {
        const int rc = nr_check_tb_config(cfg, g, "decode_tb");
        if (rc) return rc;
    }
    if (num_tb <= 0 || llr_len < 0 || llr_stride < llr_len) { nr_set_error("decode_tb: bad shape"); return NRLDPC_ERR_ARG; }
    cudaStream_t s = (cudaStream_t)stream;
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    DecArgs a{};
    a.numCb = num_tb * cfg->C;
    a.numIter = num_iter;
    a.flags = flags;
    a.llr = llr;
    a.llrStride = llr_stride;
    a.llrLen = llr_len;
    a.rm = 1;
    a.K = cfg->K; a.F = cfg->F; a.C = cfg->C; a.qm = cfg->qm; a.ncb = cfg->ncb;
    nr_tb_split(cfg, N, &a.E0, &a.nShort, &a.fStep, &a.k0);
    a.softBuf = soft_buffer;
    a.outCols = g.ksys;
    a.iters = iters;
    a.tbBits = (signed char*)tb_bits;
    a.tbBitsStride = tb_bits_stride;
    a.cbCrcOk = cb_crc_ok;
    const int Lk = cfg->K - cfg->F, per = cfg->C > 1 ? Lk - 24 : Lk;
    // per-CB CRC24A partials for the transport-block check
    unsigned int* remA = nullptr;
    if (tb_crc_ok) {
        void* p = nullptr;
        int rc0 = nr_reserve_tmp(h, (size_t)a.numCb * sizeof(unsigned int), &p);
        if (rc0) return rc0;
        remA = (unsigned int*)p;
    }
    a.cbRemA = remA;
    // rows to schedule: with no soft buffer the LLR support is known in closed form; with a soft buffer (HARQ
    // history unknown to the host) every row is scheduled.
    a.numRows = g.P;
    if (!soft_buffer && !(flags & NRLDPC_DEC_ALL_ROWS)) {
        const int L = cfg->ncb - cfg->F;
        const int Emax = a.E0 + ((a.nShort < cfg->C) ? a.fStep : 0);
        int lastQ;   // last circular-buffer index written
        if (a.k0 + Emax >= L) lastQ = L - 1; else lastQ = a.k0 + Emax - 1;
        const int sysLen = cfg->K - cfg->F - 2 * Z;
        const int lastN = (lastQ < sysLen) ? lastQ : lastQ + cfg->F;
        const int lastFull = lastN / Z + 2;
        a.numRows = max(4, min(g.P, lastFull - g.ksys + 1));
    }
    int rc = dispatch_decode(h, g, a, in_dtype, compute_dtype, s);
    if (rc) return rc;
    if (tb_crc_ok) {
        const int blocks = (int)((num_tb + 127) / 128);
        nr_tb_crc_kernel<<<blocks, 128, 0, s>>>(remA, num_tb, cfg->C, per, tb_crc_ok);
        NR_CUDA_CHECK(cudaGetLastError());
    }
    return NRLDPC_OK;
}
