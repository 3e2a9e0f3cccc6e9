// Device/host building blocks of the decoder kernels of decode.cu (arithmetic helpers, row body, split barrier,
// CRC and TMA staging helpers).
#pragma once
#include <cuda_fp16.h>
#include <string.h>

#include "nr_bg_tables.h"
#include "nrldpc_internal.cuh"

#ifndef NR_DEC_PRED_PATH
#define NR_DEC_PRED_PATH 1   // fp32: predicated FMA-pipe selects in the row body (see sub_sel / twomin_update)
#endif
#ifndef NR_DEC_PREGATHER
#define NR_DEC_PREGATHER 1   // gather next-row posteriors of columns the current row does not write ahead of the barrier
#endif
#ifndef NR_DEC_LOADS_FIRST
#define NR_DEC_LOADS_FIRST 1
#endif
#ifndef NR_DEC_LIFT_REGS
#define NR_DEC_LIFT_REGS 1
#endif
#ifndef NR_DEC_MIN_CTAS
#define NR_DEC_MIN_CTAS 2   // fp32: cap registers at 80 so that two 384-thread CTAs share an SM
#endif

// ---------------------------------------------------------------------------------------------------------------
// kernel arguments
// ---------------------------------------------------------------------------------------------------------------
struct DecArgs {
    // batch
    long long numCb;
    int cbPerCta;       // code blocks hosted by one CTA
    int nAreas;         // multi-block static kernels: posterior areas in shared memory = ceil(threads / Zc) (phantom areas for padding threads)
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
    int inF16;          // llr holds IEEE half values (fused mode only); widened exactly to T on load
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
    // transport-block CRC24A, combined inside the kernel: every block XORs  remA_r * x^(per (C-1-r)) mod g  into tbAcc[2 tb]
    // and counts itself in tbAcc[2 tb + 1]; the block that completes the count writes tbOk[tb] and clears both words
    unsigned int* tbAcc;          // [numTb][2], all zero between launches
    const unsigned int* tbFac;    // [C]: x^(per (C-1-r)) mod g24A
    unsigned char* tbOk;
    // overflow state
    void* scratch;
    unsigned int* workCounter;
    unsigned int* esAuto;   // NRLDPC_DEC_ES_AUTO: [0] hint for this launch (first tested iteration), [1] running minimum of the iteration counts
    int inSym;          // fused chain: `llr` holds complex64 equalised SYMBOLS (re, im floats); llrLen / llrStride stay in LLR units
    double invN0;       // 1 / noise variance of the max-log demapper (inSym)
    // static fp32 kernels: the TMA staging buffer of the fused load phase
    int stageFloats;
    // fused CRC of the static kernels: per-thread factors x^(B (Z-1-m)) mod g for the code-block CRC [0, Z) and the
    // CRC24A partial [Z, 2Z); computed on the host once per configuration and cached in the handle
    const unsigned int* crcFacDev;    // capacity in floats (multiple of 4); 0 = gather straight from global memory
    // decode2 (ldpc.py:1421-1492): true second minimum and a caller-chosen alpha; generic kernels only
    int trueMin2;
    double alpha;
    double beta;        // offset min-sum (extension): |message| = max(alpha * min - beta, 0); 0 = normalised min-sum (the reference)
    int synRows;        // generic kernels: rows of the early-termination syndrome (0 = all scheduled rows; 1 = decode2 compatibility)
    // static kernels with NRLDPC_DEC_EARLY_STOP: words of the bit-packed hard decisions (multiple of 4), see the kernel
    int packWords;
};

namespace {

// decoder view of the lifted graph: byte offsets instead of (column, shift), see process_row
struct __align__(16) NrDecGraph {
    int P, ncols, ksys, ncore, Z;
    uint32_t S;                // ceil(2^32 / Z): lifted positions are tracked as 32-bit fixed-point fractions of Z
    uint32_t one;              // 1, opaque to the compiler: keeps the column-base add an IMAD (FMA pipe) instead of an ALU add
    float onef;                // 1.0f, opaque to the compiler: an exact register move issued as FMUL on the FMA pipe
    uint16_t rowEdge0[NR_MAX_ROWS + 2];
    uint2 tab[NR_MAX_EDGES];   // x = (shift * S) mod 2^32, y = col*Z*sizeof(T)
    uint16_t raw[NR_MAX_EDGES];   // (col << 9) | shift: the bit-packed syndrome of the early-termination test
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

// fp32 fast path of process_row_at: the two selects of an edge are predicated FMA-pipe instructions instead of ALU-pipe
// SEL/FMNMX (the kernel is bound by the half-rate ALU pipe, profiles/README.md).
//   t = rv - x1, or rv - x2 on the edge that received the second minimum in the previous iteration
__device__ __forceinline__ float sub_sel(float rv, float x1, float x2, uint32_t off, uint32_t oldOff)
{
    float t;
    asm("{.reg .pred p; setp.eq.u32 p, %4, %5; sub.rn.f32 %0, %1, %2; @p sub.rn.f32 %0, %1, %3;}"
        : "=&f"(t) : "f"(rv), "f"(x1), "f"(x2), "r"(off), "r"(oldOff));
    return t;
}
//   two-min update with the strict first-minimum record: p = |t| < min1;  p: min2 = min1 (exact FMUL by an opaque 1.0f),
//   record (t, off);  !p: min2 = min(min2, |t|);  min1 = min(min1, |t|).  Equal to min2 = min(min2, max(min1, |t|)) because
//   min1 <= min2 always.
#ifndef NR_DEC_PRED_MIN1
#define NR_DEC_PRED_MIN1 0   // 1: the running minimum too is a predicated FMUL (FMA pipe) instead of an FMNMX (ALU pipe)
#endif
__device__ __forceinline__ void twomin_update(uint32_t sa, float t, uint32_t off, float& min1, float& min2, float onef)
{
#if NR_DEC_PRED_MIN1
    asm volatile(
        "{.reg .pred p; .reg .f32 a;\n"
        "abs.f32 a, %2;\n"
        "setp.lt.f32 p, a, %0;\n"
        "@p st.shared.v2.b32 [%3], {%4, %5};\n"
        "@p mul.rn.f32 %1, %0, %6;\n"
        "@!p min.f32 %1, %1, a;\n"
        "@p mul.rn.f32 %0, a, %6;}"
        : "+f"(min1), "+f"(min2) : "f"(t), "r"(sa), "r"(__float_as_uint(t)), "r"(off), "f"(onef));
    return;
#endif
    asm volatile(
        "{.reg .pred p; .reg .f32 a;\n"
        "abs.f32 a, %2;\n"
        "setp.lt.f32 p, a, %0;\n"
        "@p st.shared.v2.b32 [%3], {%4, %5};\n"
        "@p mul.rn.f32 %1, %0, %6;\n"
        "@!p min.f32 %1, %1, a;\n"
        "min.f32 %0, %0, a;}"
        : "+f"(min1), "+f"(min2) : "f"(t), "r"(sa), "r"(__float_as_uint(t)), "r"(off), "f"(onef));
}

// per-thread copies of the three multipliers of lifted_offset.  The static kernels read them back from shared memory
// once: a value loaded per thread sits in an ordinary register, which lets the edge-table entry be the constant-bank
// operand of the IMAD (a uniform-register multiplier would force one LDC per edge to fetch the table entry).
struct Lift {
    uint32_t S, ZB, one;
};
__device__ __forceinline__ uint32_t lifted_offset(uint32_t m, Lift L, uint2 tb)
{
    const uint32_t S = L.S, ZB = L.ZB, one = L.one;
    // (a multiply-high WITH addend needs a zeroed even/odd register pair in SASS: two extra moves per edge)
    uint32_t w, p, off;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(w) : "r"(m), "r"(S), "r"(tb.x));
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(p) : "r"(w), "r"(ZB));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(off) : "r"(p), "r"(one), "r"(tb.y));
    return off;
}

// shared-memory byte offsets of the D edges of a row for lifted check m (the private extension edge of an EXT row
// points at the thread's dummy word).  Depends on (row, m) only -- the static schedule computes it for the NEXT row
// between the arrive and the wait of the split layer barrier.
template <int D, bool EXT>
__device__ __forceinline__ void row_offsets(const NrDecGraph& g, int e0, uint32_t m, Lift ZB, uint32_t dummyOff,
                                            uint32_t (&off)[D])
{
#pragma unroll
    for (int j = 0; j < D; j++) off[j] = (EXT && j == D - 1) ? dummyOff : lifted_offset(m, ZB, g.tab[e0 + j]);
}

template <typename T, int D, bool EXT, uint32_t PRE = 0>
__device__ __forceinline__ void process_row_at(const uint32_t (&off)[D], char* __restrict__ rb, RowState<T>& st,
                                               uint32_t slot, uint32_t dummyOff, float onef, bool stdRule = true,
                                               T alpha = (T)0.75, const float* pre = nullptr, T beta = (T)0)
{
    // PRE (static fp32 schedule): bit j set = the posterior of edge j was gathered ahead of the layer barrier into pre[j]
    // (its column is not touched by the previous row, see run_rows_static)
    // stdRule / alpha: LdpcDecoder.decode (alpha = 0.75 and the "+100000" second-minimum quirk, ldpc.py:1563).  The
    // decode2 variant (ldpc.py:1421-1492) passes its own alpha and the true second minimum; only the generic kernels do.
    constexpr int OFF_SHIFT = 12;   // EXT rows: D <= 10 sign bits, then the argmin offset
    if constexpr (sizeof(T) == 4 && NR_DEC_PRED_PATH) if (stdRule) {
        // fp32 state: m1s = alpha*min1 (unsigned), m2s = the SIGNED message of the argmin edge (alpha*min2 * sign),
        // sw / rext as below.  Same arithmetic as the generic body, value for value.
        float t[D];
        const uint32_t rbS = (uint32_t)__cvta_generic_to_shared(rb);
        const float m1o = st.m1s, x2 = st.m2s;
        const uint32_t sw = st.sw;
        const uint32_t oldOff = EXT ? (sw >> OFF_SHIFT) : __float_as_uint(st.rext);
        float min1 = 0.f, min2 = __int_as_float(0x7f800000);
        uint32_t nsw = 0;
#if NR_DEC_LOADS_FIRST
        // all gathers of the row first: t[] is live through both phases anyway, so the loads cost no extra registers and
        // their shared-memory latency overlaps instead of sitting in front of every subtraction
#pragma unroll
        for (int j = 0; j < D; j++) {
            if (EXT && j == D - 1)
                t[j] = st.rext;
            else if ((PRE >> j) & 1u)
                t[j] = pre[j];
            else
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t[j]) : "r"(rbS + off[j]));
        }
#endif
#pragma unroll
        for (int j = 0; j < D; j++) {
            float rv;
#if NR_DEC_LOADS_FIRST
            rv = t[j];
#else
            if (EXT && j == D - 1)
                rv = st.rext;
            else
                rv = *reinterpret_cast<const float*>(rb + off[j]);
#endif
            const float x1 = FP<float>::flipbits(m1o, sw << (31 - (D - 1 - j)));
            t[j] = sub_sel(rv, x1, x2, off[j], oldOff);
            nsw = __funnelshift_l(__float_as_uint(t[j]), nsw, 1);
            if (j == 0) {
                min1 = fabsf(t[j]);
                slot_init(slot, t[j], off[j]);
            } else {
                twomin_update(slot, t[j], off[j], min1, min2, onef);
            }
        }
        const MinSlot<float> best = slot_read(slot, 0.f);
        min2 = fminf(min2, fabsf(__fadd_rn(best.t, 100000.f)));   // ldpc.py:1563
        const uint32_t par = __popc(nsw) & 1u;
        const uint32_t msw = par ? (~nsw & ((1u << D) - 1u)) : nsw;
        const float m1s = __fmul_rn(min1, 0.75f);
        const float m2s = __fmul_rn(min2, 0.75f);
        const float psign = FP<float>::flip(1.f, par);
        const float m1p = __fmul_rn(m1s, psign), m2p = __fmul_rn(m2s, psign);
        float rext = 0.f;
#pragma unroll
        for (int j = 0; j < D; j++) {
            const float nv = __fadd_rn(t[j], FP<float>::flipbits(m1p, __float_as_uint(t[j])));
            if (EXT && j == D - 1)
                rext = nv;
            else
                *reinterpret_cast<float*>(rb + off[j]) = nv;
        }
        const float nm2 = FP<float>::flipbits(m2p, __float_as_uint(best.t));   // new message of the argmin edge
        {
            const float nv = __fadd_rn(best.t, nm2);
            *reinterpret_cast<float*>(rb + best.off) = nv;
            if (EXT) rext = (best.off == dummyOff) ? nv : rext;
        }
        st.m1s = m1s;
        st.m2s = nm2;
        if (EXT) {
            st.sw = msw | (best.off << OFF_SHIFT);
            st.rext = rext;
        } else {
            st.sw = msw;
            st.rext = __uint_as_float(best.off);
        }
        return;
    }
    T t[D];
    T m1s = st.m1s, m2s = st.m2s;
    const uint32_t sw = st.sw;
    const uint32_t oldOff = EXT ? (sw >> OFF_SHIFT) : FP<T>::to_u32(st.rext);
    T min1 = (T)0, min2 = FP<T>::inf();
    uint32_t nsw = 0;
#pragma unroll
    for (int j = 0; j < D; j++) {
        T rv;
        if (EXT && j == D - 1)
            rv = st.rext;
        else
            rv = *reinterpret_cast<const T*>(rb + off[j]);
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
    if (stdRule) min2 = FP<T>::mn(min2, FP<T>::abs(FP<T>::add(best.t, (T)100000)));
    const uint32_t par = __popc(nsw) & 1u;
    const uint32_t msw = par ? (~nsw & ((1u << D) - 1u)) : nsw;   // sign of new message j = sign_j * parity
    m1s = FP<T>::mul(min1, alpha);
    m2s = FP<T>::mul(min2, alpha);
    if (beta != (T)0) {   // offset min-sum (extension of decode2; the reference has only the normalised form)
        m1s = FP<T>::mx(FP<T>::sub(m1s, beta), (T)0);
        m2s = FP<T>::mx(FP<T>::sub(m2s, beta), (T)0);
    }
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

template <typename T, int D, bool EXT>
__device__ __forceinline__ void process_row(const NrDecGraph& g, int e0, char* __restrict__ rb, uint32_t m,
                                            Lift ZB, RowState<T>& st, uint32_t slot, uint32_t dummyOff, bool stdRule, T alpha, T beta = (T)0)
{
    uint32_t off[D];
    row_offsets<D, EXT>(g, e0, m, ZB, dummyOff, off);
    process_row_at<T, D, EXT>(off, rb, st, slot, dummyOff, g.onef, stdRule, alpha, nullptr, beta);
}

template <typename T>
__device__ __forceinline__ void dispatch_row(const NrDecGraph& g, int row, char* rb, uint32_t m, Lift ZB,
                                             RowState<T>& st, uint32_t slot, uint32_t dummyOff, bool stdRule, T alpha, T beta = (T)0)
{
    const int e0 = g.rowEdge0[row];
    const int deg = g.rowEdge0[row + 1] - e0;
    if (row >= 4) {
        switch (deg) {
            case 3: process_row<T, 3, true>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
            case 4: process_row<T, 4, true>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
            case 5: process_row<T, 5, true>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
            case 6: process_row<T, 6, true>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
            case 7: process_row<T, 7, true>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
            case 8: process_row<T, 8, true>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
            case 9: process_row<T, 9, true>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
            default: process_row<T, 10, true>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
        }
    } else {
        switch (deg) {
            case 8: process_row<T, 8, false>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
            case 10: process_row<T, 10, false>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
            default: process_row<T, 19, false>(g, e0, rb, m, ZB, st, slot, dummyOff, stdRule, alpha, beta); break;
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

// Compile-time copy of the decoder's edge table (build_dec_graph<float>) for ONE lifting size ZS: with the static schedule
// the two table words of an edge become immediates of the address IMADs -- no uniform constant loads in front of a row.
// Instantiated for ZS = 384 (the largest lifting size: the throughput case); ZS = 0 means "run-time table".
template <int BG, int ZS>
struct SpecTab {
    static __host__ __device__ constexpr int ils()
    {
        int a = ZS;
        while (a % 2 == 0 && a > 2) a /= 2;   // ZS = a * 2^j, a in {2, 3, 5, 7, 9, 11, 13, 15}
        return a == 2 ? 0 : a == 3 ? 1 : a == 5 ? 2 : a == 7 ? 3 : a == 9 ? 4 : a == 11 ? 5 : a == 13 ? 6 : 7;
    }
    static constexpr uint32_t S = (uint32_t)((0x100000000ULL + (unsigned long long)(ZS > 0 ? ZS : 1) - 1) / (unsigned long long)(ZS > 0 ? ZS : 1));
    static __host__ __device__ constexpr uint32_t col(int e) { return BG == 1 ? NR_BG1_COL[e] : NR_BG2_COL[e]; }
    static __host__ __device__ constexpr uint32_t shift(int e)
    {
        return (uint32_t)((BG == 1 ? NR_BG1_SHIFT[ils()][e] : NR_BG2_SHIFT[ils()][e]) % (ZS > 0 ? ZS : 1));
    }
    static __host__ __device__ constexpr uint32_t x(int e) { return (uint32_t)((unsigned long long)shift(e) * S); }   // mod 2^32
    static __host__ __device__ constexpr uint32_t y(int e) { return col(e) * (uint32_t)ZS * 4u; }
};
// Split layer barrier.  The posteriors written by layer i are read by other threads in layer i+1, so the layers of a
// code block are separated by a CTA-wide barrier -- but everything a thread does between its last posterior store of
// layer i and its first gather of layer i+1 is private (row state to Tensor Memory, next row's state back, the lifted
// addresses of the next row).  arrive() is placed after the last store, wait() before the first gather, so that work
// overlaps the barrier latency instead of following it.
//   mode 0: plain bar.sync at the wait point   mode 1: mbarrier, one arrival per warp   mode 2: hardware cluster
//   barrier of the (implicit 1-CTA) cluster, which is split-phase by construction
// Measured on B200 (BG1 Zc=384): for the all-Tensor-Memory kernels (every scheduled row's state in TMEM, the R >= ~0.5 case
// of the benchmark) mode 1 is +0.7 % over mode 0 (19.76 vs 19.63 Gbit/s), for the kernels whose state spills to shared-memory
// planes / the L2 scratch (low rates, all 46 rows) it is 5-8 % SLOWER (731 vs 775 G edge-updates/s): the mode is a template
// parameter chosen per kernel.  Mode 2 is slower everywhere.
#ifndef NR_DEC_BAR_MODE
#define NR_DEC_BAR_MODE 1   // mode of the all-TMEM kernels; the others use mode 0
#endif
#ifndef NR_DEC_BAR_MODE_SPLIT
#define NR_DEC_BAR_MODE_SPLIT 0   // mode of the split (TMEM + shared planes) kernels
#endif
template <int MODE>
struct LayerBarT {
    uint32_t bar;     // shared-memory address of the mbarrier (mode 1)
    uint32_t phase;
    static constexpr int mode = MODE;
    __device__ __forceinline__ void arrive() const
    {
        if (mode == 1) {
            __syncwarp();
            if ((threadIdx.x & 31) == 0)
                asm volatile("{.reg .b64 st; mbarrier.arrive.release.cta.shared::cta.b64 st, [%0];}" ::"r"(bar) : "memory");
        } else if (mode == 2) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        }
    }
    __device__ __forceinline__ void wait()
    {
        if (mode == 1) {
            asm volatile(
                "{.reg .pred p;\n"
                "LB_WAIT_%=:\n"
                "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n"
                "@!p bra LB_WAIT_%=;}" ::"r"(bar), "r"(phase) : "memory");
            phase ^= 1u;
        } else if (mode == 2) {
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        } else {
            __syncthreads();
        }
    }
};

// posterior addressed by edge `e` for lifted index m
template <typename T>
__device__ __forceinline__ T edge_posterior(const NrDecGraph& g, int e, const char* rb, uint32_t m, Lift ZB)
{
    return *reinterpret_cast<const T*>(rb + lifted_offset(m, ZB, g.tab[e]));
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

// One-block-per-CTA form of the fused CRC: no merge tree.  Thread m multiplies the remainder of its chunk by
// x^(B*(Z-1-m)) mod g (its entry of a per-CTA factor table, built once per kernel from fac[]), and the Z products are
// XOR-reduced with redux.sync + one shared-memory exchange.  Both CRCs of a code block share that exchange.
__device__ __forceinline__ uint32_t crc_thread_factor(const uint32_t* fac, int Z, int m, uint32_t poly, int c)
{
    uint32_t f = 1;
    int k = Z - 1 - m;
    for (int l = 0; k; l++, k >>= 1)
        if (k & 1) f = gf_mulmod(f, fac[l], poly, c);
    return f;
}
template <typename T>
__device__ __forceinline__ uint32_t crc_chunk_product(const T* rcb, int len, int Z, int m, uint32_t f, uint32_t poly, int c)
{
    const int B = (len + Z - 1) / Z;
    const int i0 = m * B - (B * Z - len);
    uint32_t rem = 0;
    for (int b = 0; b < B; b++) {
        const int i = i0 + b;
        const uint32_t bit = (i >= 0) ? FP<T>::sign(rcb[i]) : 0u;
        const uint32_t fb = ((rem >> (c - 1)) & 1u) ^ bit;
        rem = (rem << 1) & ((1u << c) - 1u);
        if (fb) rem ^= poly;
    }
    return gf_mulmod(rem, f, poly, c);
}

// ---------------------------------------------------------------------------------------------------------------
// TMA staging of the fused load phase (static fp32 kernels): the rate-matched LLR stream of the NEXT code block of
// this CTA is copied global -> shared by cp.async.bulk while the current block iterates, so the de-interleaving
// gather of the load phase reads shared memory instead of waiting on HBM.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_issue(uint32_t bar, uint32_t dst, const void* src, uint32_t bytes)
{
    if (bytes == 0) {
        asm volatile("{.reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0];}" ::"r"(bar) : "memory");
        return;
    }
    asm volatile("{.reg .b64 st; mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;}" ::"r"(bar), "r"(bytes) : "memory");
    const char* sp = reinterpret_cast<const char*>(src);
    for (uint32_t o = 0; o < bytes; o += 32768u) {
        const uint32_t n = min(32768u, bytes - o);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + o),
                     "l"(sp + o), "r"(n), "r"(bar) : "memory");
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase)
{
    asm volatile(
        "{.reg .pred p;\n"
        "MB_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra MB_WAIT_%=;}" ::"r"(bar), "r"(phase) : "memory");
}

// host: byte-offset edge table for compute type T
// taggedTab: the table of the static kernels (decode_static.cuh): x = (column << 16) | shift * sizeof(T), y = column * (Z * sizeof(T) - 65536)
template <typename T>
void build_dec_graph(const NrGraph& g, NrDecGraph* d, bool taggedTab = false)
{
    memset(d, 0, sizeof(*d));
    d->P = g.P; d->ncols = g.ncols; d->ksys = g.ksys; d->ncore = g.ncore; d->Z = g.Z;
    for (int i = 0; i < NR_MAX_ROWS + 2; i++) d->rowEdge0[i] = g.rowEdge0[i];
    d->one = 1;
    d->onef = 1.0f;
    d->S = (uint32_t)((0x100000000ULL + (uint64_t)g.Z - 1) / (uint64_t)g.Z);   // ceil(2^32 / Z); Z >= 2
    for (int e = 0; e < g.rowEdge0[g.P]; e++) {
        const uint32_t col = g.edge[e] >> 16, sh = g.edge[e] & 0xffffu;
        if (taggedTab) {
            d->tab[e].x = (col << 16) | (sh * (uint32_t)sizeof(T));
            d->tab[e].y = col * ((uint32_t)g.Z * (uint32_t)sizeof(T) - 65536u);   // mod 2^32
        } else {
            d->tab[e].x = (uint32_t)((uint64_t)sh * d->S);   // mod 2^32
            d->tab[e].y = col * g.Z * (uint32_t)sizeof(T);
        }
        d->raw[e] = (uint16_t)((col << 9) | sh);
    }
}


}   // namespace
