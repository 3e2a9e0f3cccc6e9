// Static row schedule of the fp32 decoder kernels (decode.cu): the rows of the base graph are unrolled at compile time,
// the Tensor-Memory address of a row's state is an immediate, and the lifted address of an edge costs TWO instructions.
//
// Tagged offsets.  Thread m keeps mB = m * 4 (byte offset of position m inside a column).  For an edge (column c,
// shift s) the word  w = mB + ((c << 16) | 4 s)  holds the un-wrapped byte position in its LOW half (< 2 * 4 Z, so no
// carry) and the column in its HIGH half; one packed VIADDMNMX.U16x2,  off = min.u16x2(w + (65536 - 4 Z), w),  wraps the
// low half modulo 4 Z (the subtraction underflows the halfword exactly when no wrap is due) and leaves the column
// untouched.  The byte address of the posterior is  rb + off + c * (4 Z - 65536); with a compile-time lifting size
// that last term is the immediate of the LDS / STS, otherwise one add with a constant-bank operand.  `off` is unique
// among the edges of a row (the column is in it), so it doubles as the identity of the argmin edge in the row state.
// (Round 1 used three IMADs per edge, one of them a quarter-rate IMAD.HI: profiles/r01_pipe_ubench.txt.)
//
// Row state between iterations (Tensor Memory / shared planes, 4 words per check):
//   m1s  = 0.75 * min1 * parity          (the message of every edge but the argmin one, up to the sign of t_j)
//   m2s  = signed message of the argmin edge
//   sw   = sign bits of t_j (bit D-1-j = edge j) [| argmin offset << 10 for extension rows]
//   rext = posterior of the private extension column (rows >= 4) | argmin offset (core rows)
// The first iteration is a separate instantiation (FIRST): all messages are +0 there, so t_j = r_j exactly and neither
// the state of the core rows nor the old messages are touched.
#pragma once
#include <type_traits>

#include "decode_common.cuh"

namespace {

template <int I, int N, typename F>
__device__ __forceinline__ void static_for_impl(F& f)
{
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for_impl<I + 1, N>(f);
    }
}
template <int N, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
    static_for_impl<0, N>(f);
}

struct Lift2 {
    uint32_t one;     // 1, opaque to the compiler: the position add stays an IMAD (FMA pipe), see NrDecGraph::one
    uint32_t negZB;   // 65536 - 4 Z  (low halfword; the high halfword is 0)
    uint32_t kz;      // 4 Z - 65536 (mod 2^32): byte address = rb + off + column * kz
};

// per-edge constants: compile-time lifting size (immediates) or the kernel's constant-bank table
template <int BG, int ZS>
struct EdgeTab {
    static constexpr bool kStatic = true;
    static constexpr uint32_t ZB = (uint32_t)ZS * 4u;
    template <int E>
    static __device__ __forceinline__ uint32_t x(const NrDecGraph&)
    {
        constexpr uint32_t v = (SpecTab<BG, ZS>::shift(E) * 4u) | (SpecTab<BG, ZS>::col(E) << 16);
        return v;
    }
    template <int E>
    static __device__ __forceinline__ uint32_t k(const NrDecGraph&)
    {
        constexpr uint32_t v = SpecTab<BG, ZS>::col(E) * (ZB - 65536u);
        return v;
    }
    static __device__ __forceinline__ uint32_t negZB(const Lift2&) { return 65536u - ZB; }
    static __device__ __forceinline__ uint32_t kz(const Lift2&) { return ZB - 65536u; }
};
template <int BG>
struct EdgeTab<BG, 0> {
    static constexpr bool kStatic = false;
    template <int E>
    static __device__ __forceinline__ uint32_t x(const NrDecGraph& g) { return g.tab[E].x; }
    template <int E>
    static __device__ __forceinline__ uint32_t k(const NrDecGraph& g) { return g.tab[E].y; }
    static __device__ __forceinline__ uint32_t negZB(const Lift2& L) { return L.negZB; }
    static __device__ __forceinline__ uint32_t kz(const Lift2& L) { return L.kz; }
};

#ifndef NR_DEC_FIRST_SPECIAL
#define NR_DEC_FIRST_SPECIAL 1   // separate instantiation of the first iteration (no old messages); 0 = one code path
#endif
__device__ __forceinline__ uint32_t tagged_offset(uint32_t mB, uint32_t one, uint32_t x, uint32_t negZB)
{
    uint32_t w;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(w) : "r"(mB), "r"(one), "r"(x));
    return __viaddmin_u16x2(w, negZB, w);
}

// shared-memory address of an edge's posterior: rb + off + column base.  Compile-time table: the base is the immediate of
// the LDS / STS.  Run-time table: one add with a constant-bank operand, issued as IMAD (x * 1 + c) to stay off the ALU pipe.
template <typename ET, int E>
__device__ __forceinline__ uint32_t edge_addr(const NrDecGraph& g, uint32_t rbS, uint32_t off, uint32_t one)
{
    if constexpr (ET::kStatic) {
        return rbS + off + ET::template k<E>(g);
    } else {
        uint32_t a;
        asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(off), "r"(one), "r"(ET::template k<E>(g)));
        return rbS + a;
    }
}

template <int BG, int ROW>
struct RowCtx2 {   // what a thread prepares for a row before it may touch the posteriors
    static constexpr int D = BgRows<BG>::deg(ROW);
    uint32_t off[D];
    RowState<float> st;
    float pre[D];   // posteriors gathered ahead of the barrier (edges of pregather_mask)
};

// Edges of row ROW whose column is NOT an edge of row ROW - 1: the previous layer does not write them, and every older
// write is already ordered by an earlier barrier, so they may be gathered BEFORE the barrier that ends row ROW - 1.
// Row 0 follows the last scheduled row of the previous iteration (a run-time quantity): nothing is pre-gathered there.
template <int BG, int ROW>
__host__ __device__ constexpr uint32_t pregather_mask()
{
#if NR_DEC_PREGATHER
    if (ROW == 0) return 0u;
    const int e0 = BgRows<BG>::e0(ROW), d = BgRows<BG>::deg(ROW) - (ROW >= 4 ? 1 : 0);
    const int p0 = BgRows<BG>::e0(ROW > 0 ? ROW - 1 : 0), pd = BgRows<BG>::deg(ROW > 0 ? ROW - 1 : 0);
    uint32_t mask = 0;
    for (int j = 0; j < d; j++) {
        const int col = BG == 1 ? NR_BG1_COL[e0 + j] : NR_BG2_COL[e0 + j];
        bool hit = false;
        for (int k = 0; k < pd; k++) hit = hit || ((BG == 1 ? NR_BG1_COL[p0 + k] : NR_BG2_COL[p0 + k]) == col);
        if (!hit) mask |= 1u << j;
    }
    return mask;
#else
    return 0u;
#endif
}

// state + lifted offsets of a row; runs between the arrive and the wait of the layer barrier of the previous row
template <int BG, int ROW, int ZS, bool FIRST, typename Store>
__device__ __forceinline__ void prep_row2(const NrDecGraph& g, uint32_t mB, const Lift2& L, const Store& store,
                                          uint32_t dummyOff, RowCtx2<BG, ROW>& c)
{
    using ET = EdgeTab<BG, ZS>;
    constexpr int D = BgRows<BG>::deg(ROW), e0 = BgRows<BG>::e0(ROW);
    if (!FIRST || ROW >= 4) store.load(ROW, c.st);   // first iteration: only the extension posterior (rext) is needed
    static_for<D>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (ROW >= 4 && j == D - 1)
            c.off[j] = dummyOff;
        else
            c.off[j] = tagged_offset(mB, L.one, ET::template x<e0 + j>(g), ET::negZB(L));
    });
}

template <int BG, int ROW, int ZS>
__device__ __forceinline__ void pregather_row2(const NrDecGraph& g, uint32_t rbS, const Lift2& L, RowCtx2<BG, ROW>& c)
{
    using ET = EdgeTab<BG, ZS>;
    constexpr uint32_t PRE = pregather_mask<BG, ROW>();
    constexpr int D = BgRows<BG>::deg(ROW), e0 = BgRows<BG>::e0(ROW);
    static_for<D>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr ((PRE >> j) & 1u)
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(c.pre[j]) : "r"(edge_addr<ET, e0 + j>(g, rbS, c.off[j], L.one)));
    });
}

// one layer for one lifted check (fp32, LdpcDecoder.decode rule: alpha = 0.75 and the "+100000" second minimum,
// ldpc.py:1545-1576); same arithmetic, value for value, as process_row_at in decode_common.cuh
template <int BG, int ROW, int ZS, bool FIRST>
__device__ __forceinline__ void process_row2(const NrDecGraph& g, RowCtx2<BG, ROW>& c, uint32_t rbS, uint32_t slot,
                                             uint32_t dummyOff, const Lift2& L)
{
    using ET = EdgeTab<BG, ZS>;
    constexpr int D = BgRows<BG>::deg(ROW), e0 = BgRows<BG>::e0(ROW);
    constexpr bool EXT = ROW >= 4;
    constexpr uint32_t PRE = pregather_mask<BG, ROW>();
    constexpr int OFF_SHIFT = 10;   // EXT rows: D <= 10 sign bits, then the 21-bit tagged argmin offset
    float t[D];
    const float m1o = FIRST ? 0.f : c.st.m1s, x2 = FIRST ? 0.f : c.st.m2s;
    const uint32_t sw = FIRST ? 0u : c.st.sw;
    const uint32_t oldOff = FIRST ? 0u : (EXT ? (sw >> OFF_SHIFT) : __float_as_uint(c.st.rext));
    // all gathers of the row first: their shared-memory latency overlaps instead of sitting in front of every subtraction
    static_for<D>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        if constexpr (EXT && j == D - 1)
            t[j] = c.st.rext;
        else if constexpr ((PRE >> j) & 1u)
            t[j] = c.pre[j];
        else
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t[j]) : "r"(edge_addr<ET, e0 + j>(g, rbS, c.off[j], L.one)));
    });
    float min1 = 0.f, min2 = __int_as_float(0x7f800000);
    uint32_t nsw = 0;
#pragma unroll
    for (int j = 0; j < D; j++) {
        if (!FIRST) {
            const float x1 = FP<float>::flipbits(m1o, sw << (31 - (D - 1 - j)));
            t[j] = sub_sel(t[j], x1, x2, c.off[j], oldOff);
        }
        nsw = __funnelshift_l(__float_as_uint(t[j]), nsw, 1);
        if (j == 0) {
            min1 = fabsf(t[j]);
            slot_init(slot, t[j], c.off[j]);
        } else {
            twomin_update(slot, t[j], c.off[j], min1, min2, g.onef);
        }
    }
    const MinSlot<float> best = slot_read(slot, 0.f);
    min2 = fminf(min2, fabsf(__fadd_rn(best.t, 100000.f)));   // ldpc.py:1563
    // (min * sign_j * parity) * 0.75: the parity goes into the constant, +-0.75 (one rounding either way)
    const float c75 = __uint_as_float(0x3f400000u + ((uint32_t)__popc(nsw) << 31));
    const float m1p = __fmul_rn(min1, c75), m2p = __fmul_rn(min2, c75);
    float rext = 0.f;
    static_for<D>([&](auto jc) {
        constexpr int j = decltype(jc)::value;
        const float nv = __fadd_rn(t[j], FP<float>::flipbits(m1p, __float_as_uint(t[j])));
        if constexpr (EXT && j == D - 1)
            rext = nv;
        else
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(edge_addr<ET, e0 + j>(g, rbS, c.off[j], L.one)), "f"(nv));
    });
    const float nm2 = FP<float>::flipbits(m2p, __float_as_uint(best.t));   // new message of the argmin edge
    {
        const float nv = __fadd_rn(best.t, nm2);
        // program order after the store of m1' to the same word; lands in the dummy row when the argmin is private
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(rbS + best.off + (best.off >> 16) * ET::kz(L)), "f"(nv));
        if (EXT) rext = (best.off == dummyOff) ? nv : rext;
    }
    c.st.m1s = m1p;
    c.st.m2s = nm2;
    if (EXT) {
        c.st.sw = nsw | (best.off << OFF_SHIFT);
        c.st.rext = rext;
    } else {
        c.st.sw = nsw;
        c.st.rext = __uint_as_float(best.off);
    }
}

template <int BG, int ROW, int ZS, bool FIRST, typename Store, typename LayerBar>
__device__ __forceinline__ void run_rows_static2(const NrDecGraph& g, int numRows, uint32_t rbS, uint32_t mB, const Lift2& L,
                                                 const Store& store, uint32_t slot, uint32_t dummyOff, LayerBar& lb,
                                                 RowCtx2<BG, ROW>& cur)
{
    process_row2<BG, ROW, ZS, FIRST>(g, cur, rbS, slot, dummyOff, L);
    lb.arrive();
    store.store(ROW, cur.st);
    if constexpr (ROW + 1 < BgRows<BG>::P) {
        if (ROW + 1 >= 4 && ROW + 1 >= numRows) {   // numRows >= 4 always
            lb.wait();
            return;
        }
        RowCtx2<BG, ROW + 1> nxt;
        prep_row2<BG, ROW + 1, ZS, FIRST>(g, mB, L, store, dummyOff, nxt);
        pregather_row2<BG, ROW + 1, ZS>(g, rbS, L, nxt);
        lb.wait();
        run_rows_static2<BG, ROW + 1, ZS, FIRST>(g, numRows, rbS, mB, L, store, slot, dummyOff, lb, nxt);
    } else {
        lb.wait();
    }
}

// posterior addressed by edge `e` for lifted index m (tagged-offset table of the static kernels)
__device__ __forceinline__ float edge_posterior2(const NrDecGraph& g, int e, uint32_t rbS, uint32_t mB, const Lift2& L)
{
    const uint32_t off = tagged_offset(mB, L.one, g.tab[e].x, L.negZB);
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(rbS + off + g.tab[e].y));
    return v;
}

}   // namespace
