// The decoder kernel template (see decode.cu for the mapping and the bit-exactness discipline).  Included by decode.cu
// (generic kernels, launch policy, C-ABI) and by decode_bg*.cu (the statically scheduled instantiations, one translation
// unit per group so that ptxas works on them in parallel).
#pragma once
#include "crc_device.cuh"
#include "decode_common.cuh"
#include "decode_static.cuh"
#include "demap_device.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// three-tier state storage: rows [0, tmemRows) in Tensor Memory (ONE_CB kernels), the next smemRows rows in shared
// memory planes, the rest in the per-CTA global scratch (stays in L2).  All branches are on the (uniform) row index.
// ---------------------------------------------------------------------------------------------------------------
//   ALLT = 1: every scheduled row in Tensor Memory (no tier branches);  ALLT = 2 ("split"): rows [0, kSplitRows) in Tensor
//   Memory at the same fixed stride, every further scheduled row in the shared-memory planes -- the tier of a row is then
//   known at compile time in the static schedule (22-33 scheduled rows with two resident CTAs, i.e. code rates ~0.4-0.54)
//   WSYNC (multi-block static kernels): a __syncwarp() in front of every Tensor-Memory access.  tcgen05.ld / st are
//   warp-collective (.sync.aligned) but reach the compiler as opaque inline asm: around thread-varying conditions (the padding
//   threads of a multi-block CTA differ from their warp mates in `active` and in the soft-buffer pointer) it is free to unswitch
//   or duplicate the code that holds them, which leaves part of a warp at one copy and the rest at another -- a hang.  The
//   convergent intrinsic pins the code (no unswitching across it) and re-converges the warp at run time.
//   WPQ: warps per Tensor-Memory lane quadrant of the ALLT / split layouts: 3 (CTAs of up to 12 warps, two per SM, 256 columns
//   each: 21 rows) or 2 (CTAs of up to 8 warps, THREE per SM, 128 columns each: 16 rows)
template <typename T, bool ONE_CB, int ALLT, bool WSYNC = false, int WPQ = 3>
struct StateStore {
    uint32_t tbase;     // this thread's TMEM address of row slot 0 (lane quadrant and warp column offset folded in)
    uint32_t tstride;   // TMEM columns per row slot
    T* sS;              // shared planes, already offset by tid
    T* sG;              // global planes, already offset by tid
    int tmemRows, smemRows, nT;
    // ALLT: every scheduled row lives in Tensor Memory at a compile-time stride (3 warps per lane quadrant): no tier
    // branches, and with a static row index the TMEM address is base + immediate
    static constexpr uint32_t kAllTStride = (uint32_t)WPQ * (sizeof(T) == 4 ? 4u : 8u);   // referenced by the ALLT instantiations only
    static constexpr int kSplitRows = WPQ == 3 ? 21 : 16;   // 256 / 128 TMEM columns / kAllTStride (fp32)
    __device__ __forceinline__ void load(int row, RowState<T>& st) const
    {
        if constexpr (WSYNC) __syncwarp();
        if constexpr (ALLT == 1) {
            tmem_ld(st, tbase + (uint32_t)row * kAllTStride);
            return;
        }
        if constexpr (ALLT == 2) {
            if (row < kSplitRows) tmem_ld(st, tbase + (uint32_t)row * kAllTStride);
            else load_state(st, sS + (size_t)(row - kSplitRows) * NPLANES * nT, nT);
            return;
        }
        if (ONE_CB && row < tmemRows)
            tmem_ld(st, tbase + (uint32_t)row * tstride);
        else if (row < tmemRows + smemRows)
            load_state(st, sS + (size_t)(row - tmemRows) * NPLANES * nT, nT);
        else
            load_state(st, sG + (size_t)(row - tmemRows - smemRows) * NPLANES * nT, nT);
    }
    __device__ __forceinline__ void store(int row, const RowState<T>& st) const
    {
        if constexpr (WSYNC) __syncwarp();
        if constexpr (ALLT == 1) {
            tmem_st(st, tbase + (uint32_t)row * kAllTStride);
            return;
        }
        if constexpr (ALLT == 2) {
            if (row < kSplitRows) tmem_st(st, tbase + (uint32_t)row * kAllTStride);
            else store_state(st, sS + (size_t)(row - kSplitRows) * NPLANES * nT, nT);
            return;
        }
        if (ONE_CB && row < tmemRows)
            tmem_st(st, tbase + (uint32_t)row * tstride);
        else if (row < tmemRows + smemRows)
            store_state(st, sS + (size_t)(row - tmemRows) * NPLANES * nT, nT);
        else
            store_state(st, sG + (size_t)(row - tmemRows - smemRows) * NPLANES * nT, nT);
    }
};

// input element -> compute type (half and float widen exactly)
template <typename T, typename TIn>
__device__ __forceinline__ T llr_cvt(TIn v)
{
    return (T)v;
}
template <>
__device__ __forceinline__ float llr_cvt<float, __half>(__half v)
{
    return __half2float(v);
}
template <>
__device__ __forceinline__ double llr_cvt<double, __half>(__half v)
{
    return (double)__half2float(v);
}

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
// ESM (static kernels): 1 = the early-termination code is compiled in (run-time flag), 0 = left out altogether
//      ZS (static kernels without the early-termination code): lifting size known at compile time (SpecTab), 0 = run time
template <typename T, bool ONE_CB, int SBG, int ALLT, int ESM = 1, int ZS = 0, int WPQ = 3>
__global__ void __launch_bounds__((WPQ == 2 ? 256 : 384), (sizeof(T) == 4 ? (WPQ == 2 ? 3 : NR_DEC_MIN_CTAS) : 1))
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

    // static kernels: one extra row per block, the per-thread dummy words of the private extension edge (decode_static.cuh)
    constexpr int XROW = (SBG != 0) ? 1 : 0;
    // MB: statically scheduled kernel with SEVERAL code blocks per CTA (Zc <= 192, and the lifting sizes that are no multiple
    // of 32).  The row code holds warp-collective instructions (Tensor-Memory loads / stores), so EVERY thread runs it: the
    // threads that pad the CTA to whole warps, and the blocks of a partly filled last group, work on real shared memory (the
    // padding threads on a phantom block area of their own) and simply never load or store anything global.
    constexpr bool MB = (SBG != 0) && !ONE_CB;
    const int nArea = MB ? a.nAreas : cbPerCta;
    T* rs = reinterpret_cast<T*>(smemRaw);                                   // [nArea][ncore + XROW][Z]
    T* stateS = rs + (size_t)nArea * (ncore + XROW) * Z;                     // [smemRows][NPLANES][nT]
    uint32_t* misc = reinterpret_cast<uint32_t*>(stateS + (size_t)a.smemRows * NPLANES * nT);
    // misc: [0, flagsLen) per-block flags | 32 words CRC factors (2 x 16) | per-block CRC trees
    const int flagsLen = (cbPerCta + 31) & ~31;
    uint32_t* fac = misc + flagsLen;
    int P2 = 1;
    while (P2 < Z) P2 <<= 1;
    uint32_t* tree = misc + flagsLen + 32 + (size_t)(cbl < cbPerCta ? cbl : cbPerCta - 1) * P2;   // (padding threads never write it)
    // per-thread argmin record + dummy word (16-byte aligned region after the CRC trees)
    const size_t slotOfs = ((size_t)(reinterpret_cast<unsigned char*>(misc + flagsLen + 32 + (size_t)cbPerCta * P2) - smemRaw) + 15) & ~(size_t)15;
    MinSlot<T>* slotP = reinterpret_cast<MinSlot<T>*>(smemRaw + slotOfs) + tid;
    const uint32_t slot = (uint32_t)__cvta_generic_to_shared(slotP);
    T* dummyW = reinterpret_cast<T*>(slotP - tid + nT) + tid;
    const int globRows = a.numRows - a.tmemRows - a.smemRows;
    T* stateG = reinterpret_cast<T*>(a.scratch) + (size_t)blockIdx.x * (size_t)globRows * NPLANES * nT;
    T* rcb = rs + (size_t)cbl * (ncore + XROW) * Z;
    // Tensor Memory for the thread-private row state (see tmem_ld above)
    __shared__ uint32_t tmemBaseSh;
    const bool useTmem = (ONE_CB || MB) && a.tmemCols > 0;
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
    StateStore<T, (ONE_CB || MB), ALLT, MB, WPQ> store;
    {
        const int warp = tid >> 5;
        const uint32_t RW = sizeof(T) == 4 ? 4u : 8u;
        const uint32_t wpq = ALLT != 0 ? (uint32_t)WPQ : ((uint32_t)((nT >> 5) + 3) >> 2);   // warps per lane quadrant
        store.tstride = wpq * RW;
        store.tbase = useTmem ? (tmemBaseSh + ((uint32_t)(warp & 3) << 21) + (uint32_t)(warp >> 2) * RW) : 0u;   // lane (warp%4)*32 in bits 31..16
        store.sS = stateS + tid;
        store.sG = stateG + tid;
        store.tmemRows = (ONE_CB || MB) ? a.tmemRows : 0;
        store.smemRows = a.smemRows;
        store.nT = nT;
    }
    char* rb = reinterpret_cast<char*>(rcb);
    const uint32_t mU = (uint32_t)m;
    Lift ZB;
    ZB.S = g.S;
    ZB.ZB = (uint32_t)Z * (uint32_t)sizeof(T);
    ZB.one = g.one;
    const uint32_t dummyOff = (uint32_t)(reinterpret_cast<char*>(dummyW) - rb);
    const int ksys = g.ksys;

    // static kernels: [2 mbarriers | XOR exchange 2 x 32 | packed hard decisions (early stop) | LLR staging buffer]
    unsigned char* extra = smemRaw + ((slotOfs + (size_t)nT * (sizeof(MinSlot<T>) + sizeof(T)) + 15) & ~(size_t)15);
    const uint32_t barLayer = (uint32_t)__cvta_generic_to_shared(extra);
    const uint32_t barStage = barLayer + 8;
    uint32_t* crcRed = reinterpret_cast<uint32_t*>(extra + 16);
    // early-termination test of the static kernels (a.packWords words in all, 0 when the test is off):
    //   pk [ (2 ncore + numRows - 4 + 1) W ] packed hard decisions | esSyn [numRows W] syndrome words | esTask [160] | esRaw [320 x u16]
    uint32_t* pk = crcRed + 64;
    const int esPkWords = (((2 * ncore + (a.numRows - 4) + 1) * (nT >> 5)) + 3) & ~3;
    const int esSynWords = ((a.numRows * (nT >> 5)) + 3) & ~3;
    uint32_t* esSyn = pk + esPkWords;
    uint32_t* esTask = esSyn + esSynWords;
    uint16_t* esRaw = reinterpret_cast<uint16_t*>(esTask + 160);
    int esNumTasks = 0;
    float* stage = reinterpret_cast<float*>(pk + ((SBG != 0) ? a.packWords : 0));
    const bool useStage = (SBG != 0) && a.stageFloats > 0;
    LayerBarT<(ALLT == 1 ? NR_DEC_BAR_MODE : (ALLT == 2 ? NR_DEC_BAR_MODE_SPLIT : 0))> lb;
    lb.bar = barLayer;
    lb.phase = 0;
    uint32_t stagePhase = 0;
    __shared__ uint32_t liftSh[4];
    Lift2 L2;
    L2.one = g.one;
    L2.negZB = 65536u - (uint32_t)Z * 4u;
    L2.kz = (uint32_t)Z * 4u - 65536u;
    const uint32_t mB = (uint32_t)m * 4u;                            // tagged offsets (decode_static.cuh)
    const uint32_t dummyOff2 = mB | ((uint32_t)ncore << 16);         // the thread's word of the dummy row
    const uint32_t rbS = (uint32_t)__cvta_generic_to_shared(rb);
    if (SBG != 0 && ESM != 0 && a.packWords > 0) {   // task list of the early-termination test: <= 8 core edges of one row each
        for (int row = 0; row < a.numRows; row++) {
            const int e0 = g.rowEdge0[row], e1 = g.rowEdge0[row + 1] - (row >= 4 ? 1 : 0);
            for (int e = e0; e < e1; e += 8) {
                if (tid == 0) esTask[esNumTasks] = (uint32_t)row | ((uint32_t)e << 8) | ((uint32_t)min(8, e1 - e) << 20) | ((row >= 4 && e == e0) ? (1u << 24) : 0u);
                esNumTasks++;
            }
        }
        for (int e = tid; e < g.rowEdge0[a.numRows]; e += nT) esRaw[e] = g.raw[e];
        for (int i = tid; i < a.numRows * (nT >> 5); i += nT) esSyn[i] = 0u;
    }
    if (SBG != 0) {
        if (tid == 0) {
            liftSh[0] = L2.negZB;
            liftSh[1] = L2.kz;
            liftSh[2] = L2.one;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barLayer), "r"((uint32_t)(nT >> 5)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barStage), "r"(1u) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (NR_DEC_LIFT_REGS) {   // per-thread register copies (see struct Lift)
            const volatile uint32_t* lv = liftSh;
            L2.negZB = lv[0];
            L2.kz = lv[1];
            L2.one = lv[2];
        }
    }

    // symbol input (nrldpc_decode_tb_symbols; one-block static fp32 kernels with a staged stream only): the per-axis amplitude
    // levels of the demapper, the same expression as nr_demap_kernel's
    constexpr bool SYM = (SBG != 0) && ONE_CB && sizeof(T) == 4;
    __shared__ double demapLevels[SYM ? 32 : 1];
    if constexpr (SYM) {
        if (a.inSym) {
            const int half = a.qm >> 1;
            if (a.qm > 1 && tid < (1 << half)) demapLevels[tid] = qam_scale(a.qm) * (double)pam_level((uint32_t)tid, half);
            __syncthreads();
        }
    }
    const bool wantCrc = a.rm && (a.tbBits || a.cbCrcOk || a.tbOk);
    const int Lk = a.K - a.F;                       // code block without fillers
    const int per = (a.C > 1) ? Lk - 24 : Lk;       // payload copied into the merged transport block
    const NrCrcPoly polyCb = nr_crc_poly(a.C > 1 ? NRLDPC_CRC24B : NRLDPC_CRC24A);
    const NrCrcPoly polyA = nr_crc_poly(NRLDPC_CRC24A);
    if (wantCrc) {
        if (SBG == 0 || MB) {
            crc_factors(fac, Lk, Z, P2, polyCb.poly, polyCb.len, tid);
            if (a.C > 1) crc_factors(fac + 16, per, Z, P2, polyA.poly, polyA.len, tid);
        }
    }

    // geometry of a code block's slice of the rate-matched stream (getRateMatchedCbLens, ldpc.py:846-856)
    auto stream_geom = [&](long long cbi, int& E, long long& xBase, long long& xAvail) {
        const long long tb = cbi / a.C;
        const int r = (int)(cbi - tb * a.C);
        E = a.E0 + (r >= a.nShort ? a.fStep : 0);
        const long long off = (long long)r * a.E0 + (long long)(r > a.nShort ? (r - a.nShort) : 0) * a.fStep;
        xBase = tb * a.llrStride + off;
        xAvail = a.llrLen - off;   // LLRs actually present for this block (rest are zeros, ldpc.py:1402)
    };
    // staging: copy the 16-byte aligned window [xBase - head, xBase - head + nCopy) of the stream; the (< 4) LLRs
    // behind the last whole 16 bytes are read from global memory by the gather
    auto stage_block = [&](long long cbi) {
        int E;
        long long xBase, xAvail;
        stream_geom(cbi, E, xBase, xAvail);
        const int n = (int)(xAvail < 0 ? 0 : (xAvail > (long long)E ? (long long)E : xAvail));
        if (a.inSym) {   // n / qm symbols of 8 bytes from symbol xBase / qm on (E, the offsets and the pitch are multiples of qm)
            const long long xs = xBase / a.qm;
            const int headS = (int)(xs & 1), nCopyS = (headS + n / a.qm) & ~1;
            stage_issue(barStage, (uint32_t)__cvta_generic_to_shared(stage),
                        reinterpret_cast<const char*>(a.llr) + (xs - headS) * 8, (uint32_t)(nCopyS * 8));
            return;
        }
        const int es = a.inF16 ? 2 : 4, epv = 16 / es;   // element size, elements per 16 bytes
        const int head = (int)(xBase & (epv - 1));
        const int nCopy = (head + n) & ~(epv - 1);
        stage_issue(barStage, (uint32_t)__cvta_generic_to_shared(stage),
                    reinterpret_cast<const char*>(a.llr) + (xBase - head) * es, (uint32_t)(nCopy * es));
    };
    if (useStage && tid == 0 && (long long)blockIdx.x < (a.numCb + cbPerCta - 1) / cbPerCta) stage_block((long long)blockIdx.x);

    // Work queue.  A CTA starts with groups blockIdx.x and blockIdx.x + gridDim.x; every further group index comes from an
    // atomic counter in the handle (workCounter[0], reset by the last CTA to leave), fetched ONE GROUP AHEAD by thread 0 so
    // that its latency hides behind a whole code block.  With early termination blocks differ in cost and the queue removes
    // the imbalance of a static stride; with a fixed iteration count it degenerates to the same assignment.
    const long long numGroups = (a.numCb + cbPerCta - 1) / cbPerCta;
    // first iteration (1-based) after which the syndrome is tested: the flags' value, raised by the hint of the previous launch
    int esFrom = (a.flags >> 8) & 0xff;
    if (a.esAuto) esFrom = max(esFrom, (int)*reinterpret_cast<volatile unsigned int*>(a.esAuto));
    __shared__ long long nextGrpSh;
    const bool dynQ = a.workCounter != nullptr;
    bool firstGroup = true;
    unsigned int grpPend = 0;   // thread 0: counter value fetched while the previous group was loaded
    for (long long grp = blockIdx.x; grp < numGroups;) {
        const long long cb = grp * cbPerCta + cbl;
        const bool active = ONE_CB ? true : (lane_ok && cb < a.numCb);

        // -------------------------------------------------------------------------------------------------------
        // load phase: column block `col` (un-punctured index), position m.  Punctured columns 0,1 start at 0.
        // -------------------------------------------------------------------------------------------------------
        if (active || MB) {
            rcb[m] = (T)0;
            rcb[Z + m] = (T)0;
            const int lastCol = ksys + a.numRows;   // exclusive; numRows >= 4
            int E = 0, L = 0, sysLen = 0, Eq = 1;
            long long xBase = 0, xAvail = 0;
            T* sb = nullptr;
            if (a.rm) {
                stream_geom(cb, E, xBase, xAvail);
                L = a.ncb - a.F;
                sysLen = a.K - a.F - 2 * Z;
                Eq = E / a.qm;
                if (a.softBuf && active) sb = reinterpret_cast<T*>(a.softBuf) + cb * (long long)L;
            }
            // (uniform conditions only: `sb` is null for the padding threads of a multi-block CTA, and the paths below hold
            // warp-collective Tensor-Memory stores)
            const int colEnd = (a.rm && a.softBuf) ? g.ncols : lastCol;   // a soft buffer is combined over its whole length
            // de-interleaver division i / Eq: float reciprocal + one correction step (exact for i < 2^24)
            const bool smallE = E < (1 << 24);
            const float rcpEq = 1.0f / (float)Eq;
            const int xAvailI = !active ? 0 : (int)(xAvail < 0 ? 0 : (xAvail > (long long)E ? (long long)E : xAvail));   // inactive: nothing is read
            auto load_cols = [&](auto tin) {
                using TIn = decltype(tin);
                const TIn* __restrict__ x = reinterpret_cast<const TIn*>(a.llr) + (a.rm ? xBase : cb * a.llrStride);
                int n = m;   // index in the punctured coded block
                for (int col = 2; col < colEnd; col++, n += Z) {
                    T v = (T)0;
                    if (!a.rm) {
                        if (active && col - 2 < a.inCols) v = llr_cvt<T, TIn>(x[n]);
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
                                const T xv = (xi < xAvailI) ? llr_cvt<T, TIn>(x[xi]) : (T)0;
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
            auto staged_load = [&](auto tin) {
                using TIn = decltype(tin);
                constexpr int EPV = 16 / (int)sizeof(TIn);   // elements per 16 bytes
                // staged stream, no repetition (E <= Ncb - F: a buffer position receives at most one LLR): one term per
                // position, read from shared memory; element xi of the stream sits at stage[head + xi] for head + xi <
                // nCopy, the (< EPV) LLRs behind the last whole 16 bytes come from global memory
                const TIn* __restrict__ x = reinterpret_cast<const TIn*>(a.llr) + xBase;
                const int ncb = a.ncb, F = a.F, k0 = a.k0, qm = a.qm;
                const int head = (int)(xBase & (EPV - 1));
                const int nCopy = (head + xAvailI) & ~(EPV - 1);
                mbar_wait(barStage, stagePhase);
                stagePhase ^= 1u;
                const TIn* __restrict__ sp = reinterpret_cast<const TIn*>(stage) + head;
                const int nStaged = nCopy - head;
                int n = m;
                for (int col = 2; col < lastCol; col++, n += Z) {
                    // branch-free: every thread computes an index, invalid ones read element 0 and drop it
                    const int nf = n - sysLen;                       // >= 0: at or behind the filler gap
                    const bool isFill = (unsigned)nf < (unsigned)F;  // LARGE_LLR (chancodebase.py:52) after the clip
                    int i = n - (nf >= 0 ? F : 0) - k0;
                    i += (i < 0) ? L : 0;
                    int b = (int)((float)i * rcpEq);                 // de-interleaver: stream index (i mod Eq) * qm + i / Eq
                    int r = i - b * Eq;
                    b += (r >= Eq) ? 1 : 0;
                    r -= (r >= Eq) ? Eq : 0;
                    b -= (r < 0) ? 1 : 0;
                    r += (r < 0) ? Eq : 0;
                    const int xi = r * qm + b;
                    const bool valid = (n < ncb) && !isFill && (i < E) && (xi < xAvailI);
                    T v = llr_cvt<T, TIn>(sp[(valid && xi < nStaged) ? xi : 0]);
                    if (valid && xi >= nStaged) v = llr_cvt<T, TIn>(x[xi]);   // behind the last whole 16 bytes
                    v = FP<T>::mn(v, (T)1e10);                        // np.clip(., -1e10, 1e10), ldpc.py:1536
                    v = FP<T>::mx(v, (T)-1e10);
                    v = FP<T>::add(v, (T)0);                          // -0.0 -> +0.0 (see header)
                    v = valid ? v : ((isFill && n < ncb) ? (T)1e10 : (T)0);
                    if (col < ncore) {
                        rcb[col * Z + m] = v;
                    } else {
                        RowState<T> st0;
                        st0.m1s = (T)0; st0.m2s = (T)0; st0.sw = 0; st0.rext = v;
                        store.store(col - ksys, st0);
                    }
                }
            };
            // staged SYMBOLS: stream element xi = r * qm + b is bit b of symbol r of the block; its LLR is computed here,
            // value for value as nr_demap_kernel<float, float> would have written it (no LLR buffer, no demapper launch)
            auto staged_load_sym = [&]() {
                if constexpr (SYM) {
                    const int ncb = a.ncb, F = a.F, k0 = a.k0, qm = a.qm;
                    const long long xs = xBase / qm;
                    const float2* __restrict__ x = reinterpret_cast<const float2*>(a.llr) + xs;
                    const int headS = (int)(xs & 1);
                    const int nCopyS = (headS + xAvailI / qm) & ~1;
                    mbar_wait(barStage, stagePhase);
                    stagePhase ^= 1u;
                    const float2* __restrict__ sp = reinterpret_cast<const float2*>(stage) + headS;
                    const int nStagedS = nCopyS - headS;
                    const double invN0 = a.invN0;
                    int n = m;
                    for (int col = 2; col < lastCol; col++, n += Z) {
                        const int nf = n - sysLen;
                        const bool isFill = (unsigned)nf < (unsigned)F;
                        int i = n - (nf >= 0 ? F : 0) - k0;
                        i += (i < 0) ? L : 0;
                        int b = (int)((float)i * rcpEq);
                        int r = i - b * Eq;
                        b += (r >= Eq) ? 1 : 0;
                        r -= (r >= Eq) ? Eq : 0;
                        b -= (r < 0) ? 1 : 0;
                        r += (r < 0) ? Eq : 0;
                        const bool valid = (n < ncb) && !isFill && (i < E) && (r * qm + b < xAvailI);
                        float2 y = sp[(valid && r < nStagedS) ? r : 0];
                        if (valid && r >= nStagedS) y = x[r];   // behind the last whole 16 bytes
                        T v = (T)demap_bit_f32(y.x, y.y, valid ? b : 0, qm, demapLevels, invN0);
                        v = FP<T>::mn(v, (T)1e10);
                        v = FP<T>::mx(v, (T)-1e10);
                        v = FP<T>::add(v, (T)0);
                        v = valid ? v : ((isFill && n < ncb) ? (T)1e10 : (T)0);
                        if (col < ncore) {
                            rcb[col * Z + m] = v;
                        } else {
                            RowState<T> st0;
                            st0.m1s = (T)0; st0.m2s = (T)0; st0.sw = 0; st0.rext = v;
                            store.store(col - ksys, st0);
                        }
                    }
                }
            };
            if (useStage && E <= L) {
                if (SYM && a.inSym) staged_load_sym();
                else if (a.inF16) staged_load(__half()); else staged_load(float());
            } else if (a.rm && !a.softBuf && !a.inF64 && !a.inF16 && smallE) {
                // common case (fp32 stream, no HARQ history): same arithmetic, none of the generic bookkeeping
                const float* __restrict__ x = reinterpret_cast<const float*>(a.llr) + xBase;
                const int ncb = a.ncb, F = a.F, k0 = a.k0, qm = a.qm;
                // staged stream (static kernels): element xi sits at stage[head + xi] for head + xi < nCopy
                const int head = (int)(xBase & 3);
                const int nCopy = useStage ? ((head + xAvailI) & ~3) : 0;
                if (useStage) {
                    mbar_wait(barStage, stagePhase);
                    stagePhase ^= 1u;
                }
                auto fetch = [&](int xi) -> T {
                    if (xi >= xAvailI) return (T)0;
                    return (head + xi < nCopy) ? (T)stage[head + xi] : (T)x[xi];
                };
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
                                    v[c] = fetch(stream_index(i));   // 0 + x == x exactly
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
                                acc = FP<T>::add(acc, fetch(stream_index(i)));
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
            } else if (a.inF16) {
                load_cols(__half());
            } else {
                load_cols(float());
            }
            for (int row = 0; row < ((SBG != 0 && NR_DEC_FIRST_SPECIAL) ? 0 : 4); row++) {   // static kernels: the first iteration never reads it
                RowState<T> st0;
                st0.m1s = (T)0; st0.m2s = (T)0; st0.sw = 0; st0.rext = (T)0;
                store.store(row, st0);
            }
        }
        __syncthreads();
        // the staging buffer is free again: fetch the stream of this CTA's next code block while this one iterates
        if (tid == 0) {
            const long long next = (!dynQ || firstGroup) ? grp + gridDim.x : 2ll * gridDim.x + grpPend;
            if (useStage && next < numGroups) stage_block(next);
            nextGrpSh = next;   // read by everybody after the barrier that ends this group
            if (dynQ && next < numGroups) grpPend = atomicAdd(a.workCounter, 1u);
        }
        firstGroup = false;

        // -------------------------------------------------------------------------------------------------------
        // iterations
        // -------------------------------------------------------------------------------------------------------
        int itersDone = 0;
        bool cbDone = false;
        for (int it = 0; it < a.numIter; it++) {
            if constexpr (SBG != 0) {
                RowCtx2<SBG, 0> c0;
                if (NR_DEC_FIRST_SPECIAL && it == 0) {   // all messages are +0: t = r, no state to read (decode_static.cuh)
                    prep_row2<SBG, 0, ZS, true>(g, mB, L2, store, dummyOff2, c0);
                    run_rows_static2<SBG, 0, ZS, true>(g, a.numRows, rbS, mB, L2, store, slot, dummyOff2, lb, c0);
                } else {
                    prep_row2<SBG, 0, ZS, false>(g, mB, L2, store, dummyOff2, c0);
                    run_rows_static2<SBG, 0, ZS, false>(g, a.numRows, rbS, mB, L2, store, slot, dummyOff2, lb, c0);
                }
            } else {
                for (int row = 0; row < a.numRows; row++) {
                    if (ONE_CB || (active && !cbDone)) {
                        RowState<T> st;
                        store.load(row, st);
                        dispatch_row<T>(g, row, rb, mU, ZB, st, slot, dummyOff, !a.trueMin2, a.trueMin2 ? (T)a.alpha : (T)0.75, a.trueMin2 ? (T)a.beta : (T)0);
                        store.store(row, st);
                    }
                    __syncthreads();
                }
            }
            if (!cbDone) itersDone = it + 1;
            if constexpr (SBG != 0) {
                if (ESM != 0 && (a.flags & NRLDPC_DEC_EARLY_STOP) && it + 1 >= esFrom) {
                    // Syndrome of the hard decisions after a COMPLETE iteration, bit-packed.  (1) Every warp ballots the sign of
                    // its 32 positions of each core column (stored twice, so a circulant shift is one funnel shift of two
                    // neighbouring words) and of each scheduled extension column (read back from the row state: the row loop
                    // carries no early-termination code at all).  (2) The XOR of a row's shifted words is cut into tasks of at
                    // most 8 edges x one 32-check word, one task per thread, merged with shared-memory atomics.  (3) All words 0?
                    const int W = nT >> 5, warp = tid >> 5, lane = tid & 31;
                    uint32_t* pe = pk + (size_t)ncore * 2 * W;   // extension columns, not doubled
                    constexpr int NC = (SBG == 1) ? 26 : 14;   // == ncore (k + 4 columns of degree > 1)
                    {   // all loads first, then ballot + one predicated store per column (lanes 0 and 1 write the two copies)
                        uint32_t hv[NC];
                        const uint32_t rAddr = (uint32_t)__cvta_generic_to_shared(rcb + m);
                        const uint32_t cStride = (uint32_t)Z * (uint32_t)sizeof(T);
#pragma unroll
                        for (int col = 0; col < NC; col++)
                            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hv[col]) : "r"(rAddr + (uint32_t)col * cStride + (sizeof(T) == 8 ? 4u : 0u)));
                        uint32_t pAddr = (uint32_t)__cvta_generic_to_shared(pk + (lane & 1) * W + warp);
                        const uint32_t pStride = 2u * (uint32_t)W * 4u;
#pragma unroll
                        for (int col = 0; col < NC; col++) {
                            const uint32_t w = __ballot_sync(0xffffffffu, (int)hv[col] < 0);
                            asm volatile("{.reg .pred p; setp.lt.u32 p, %2, 2; @p st.shared.b32 [%0], %1;}" ::"r"(pAddr), "r"(w), "r"((uint32_t)lane) : "memory");
                            pAddr += pStride;
                        }
                        uint32_t eAddr = (uint32_t)__cvta_generic_to_shared(pe + warp);
                        for (int row = 4; row < a.numRows; row++) {
                            RowState<T> st;
                            store.load(row, st);
                            const uint32_t w = __ballot_sync(0xffffffffu, FP<T>::sign(st.rext) != 0);
                            asm volatile("{.reg .pred p; setp.eq.u32 p, %2, 0; @p st.shared.b32 [%0], %1;}" ::"r"(eAddr), "r"(w), "r"((uint32_t)lane) : "memory");
                            eAddr += (uint32_t)W * 4u;
                        }
                    }
                    __syncthreads();
                    {
                        const int w = tid % W, k0 = tid / W, kStep = nT / W;
                        for (int k = k0; k < esNumTasks; k += kStep) {
                            const uint32_t tk = esTask[k];            // row | first edge << 8 | edges << 20 | first chunk << 24
                            const int row = tk & 0xff, e0 = (tk >> 8) & 0xfff, cnt = (tk >> 20) & 0xf;
                            uint32_t acc = ((tk >> 24) & 1u) ? pe[(row - 4) * W + w] : 0u;
                            for (int e = e0; e < e0 + cnt; e++) {
                                const uint32_t raw = esRaw[e];
                                const uint32_t bpos = 32u * (uint32_t)w + (raw & 511u);   // first position read by these 32 checks
                                const uint32_t* pc = pk + (raw >> 9) * 2 * W + (bpos >> 5);
                                acc ^= __funnelshift_r(pc[0], pc[1], bpos & 31u);
                            }
                            if (acc) atomicXor(&esSyn[row * W + w], acc);
                        }
                    }
                    __syncthreads();
                    uint32_t bad = 0;
                    for (int i = tid; i < a.numRows * W; i += nT) {
                        bad |= esSyn[i];
                        esSyn[i] = 0u;   // ready for the next test (ordered by the barriers around the accumulation)
                    }
                    const int anyBad = __syncthreads_or(bad != 0);
                    if (!anyBad) break;
                }
            } else
            if ((a.flags & NRLDPC_DEC_EARLY_STOP) && it + 1 >= esFrom) {
                // syndrome of the hard decisions after a COMPLETE iteration over the scheduled rows (skipped rows are
                // satisfied by construction: their parity bit is the parity of the rest)
                uint32_t bad = 0;
                if (active && !cbDone) {
                    const int synRows = (a.synRows > 0 && a.synRows < a.numRows) ? a.synRows : a.numRows;
                    for (int row = 0; row < synRows; row++) {
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
            if (a.esAuto && m == 0) atomicMin(a.esAuto + 1, (unsigned int)itersDone);
            const int outCore = min(a.outCols, ncore);
            if (a.bits) {
                signed char* o = a.bits + cb * a.bitsStride;
                for (int col = 0; col < outCore; col++) o[col * Z + m] = (signed char)FP<T>::sign(rcb[col * Z + m]);
            }
            if (a.beliefs) {
                T* o = reinterpret_cast<T*>(a.beliefs) + cb * (long long)a.outCols * Z;
                for (int col = 0; col < outCore; col++) o[col * Z + m] = rcb[col * Z + m];
            }
        }
        // extension columns: the state reads are warp-collective in the Tensor-Memory kernels, so every thread takes them
        for (int col = ncore; col < a.outCols; col++) {
            const int row = col - ksys;
            T v = (T)0;
            if (row < a.numRows) {
                RowState<T> st;
                store.load(row, st);
                v = st.rext;
            } else if (active) {
                // skipped row: t_ext == 0 in every iteration, so its belief after the last iteration is
                // 0.75 * parity * min(min_j |r_j|, 1e5) over the row's core edges evaluated on the final posteriors
                const int e0 = g.rowEdge0[row];
                const int e1 = g.rowEdge0[row + 1] - 1;
                T mn = (T)100000;
                uint32_t par = 0;
                for (int e = e0; e < e1; e++) {
                    T rv;
                    if constexpr (SBG != 0) rv = (T)edge_posterior2(g, e, rbS, mB, L2);
                    else rv = edge_posterior<T>(g, e, rb, mU, ZB);
                    mn = FP<T>::mn(mn, FP<T>::abs(rv));
                    par ^= FP<T>::sign(rv);
                }
                v = (a.numIter > 0) ? FP<T>::flip(FP<T>::mul(mn, (T)0.75), par) : (T)0;
                v = FP<T>::add(v, (T)0);
            }
            if (active) {
                if (a.bits) a.bits[cb * a.bitsStride + col * Z + m] = (signed char)(v < (T)0);
                if (a.beliefs) reinterpret_cast<T*>(a.beliefs)[cb * (long long)a.outCols * Z + col * Z + m] = v;
            }
        }
        if (wantCrc) {
            // checkCrcAndMerge (ldpc.py:1610-1619) on the hard decisions still in shared memory
            uint32_t remCb, remA;
            if constexpr (SBG != 0 && ONE_CB) {
                // CRC by linearity: remainder = XOR over the set bits i of x^(len-1-i) mod g.  Thread m owns bit col*Z + m of
                // every systematic column; the per-bit constants come from a per-configuration table in global memory
                // ([2][ksys][Z] words, L2-resident, coalesced; 0 beyond the message, so fillers and the CRC24A/B length
                // difference need no branches).  ~6 instructions per bit for both CRCs instead of a bit-serial division.
                constexpr int KS = (SBG == 1) ? 22 : 10;
                uint32_t pc = 0, pa = 0;
                {
                    const unsigned int* __restrict__ tc = a.crcFacDev + m;
                    const bool two = a.C > 1;
                    uint32_t cc[KS], ca[KS];
#pragma unroll
                    for (int col = 0; col < KS; col++) {
                        cc[col] = tc[col * Z];
                        ca[col] = two ? tc[(KS + col) * Z] : 0u;
                    }
#pragma unroll
                    for (int col = 0; col < KS; col++) {
                        const uint32_t sm = (uint32_t)((int)FP<T>::hibits(rcb[col * Z + m]) >> 31);   // all ones when the bit is 1
                        pc ^= sm & cc[col];
                        pa ^= sm & ca[col];
                    }
                }
                pc = __reduce_xor_sync(0xffffffffu, pc);
                pa = __reduce_xor_sync(0xffffffffu, pa);
                if ((tid & 31) == 0) {
                    crcRed[tid >> 5] = pc;
                    crcRed[32 + (tid >> 5)] = pa;
                }
                __syncthreads();
                remCb = 0;
                remA = 0;
                for (int w = 0; w < (nT >> 5); w++) {
                    remCb ^= crcRed[w];
                    remA ^= crcRed[32 + w];
                }
                if (a.C <= 1) remA = remCb;
            } else {
                remCb = cb_crc<T>(rcb, Lk, Z, P2, m, active, tree, fac, polyCb.poly, polyCb.len);
                __syncthreads();
                remA = remCb;
                if (a.C > 1) remA = cb_crc<T>(rcb, per, Z, P2, m, active, tree, fac + 16, polyA.poly, polyA.len);
            }
            if (active) {
                if (m == 0) {
                    if (a.cbCrcOk) a.cbCrcOk[cb] = (remCb == 0);
                    if (a.tbOk) {
                        if (a.C <= 1) {
                            a.tbOk[cb] = (remA == 0);
                        } else {   // checkCrc(txBlock, '24A') of harq.py:173 / ldpc.py:1245 by linearity over the C blocks
                            const long long tb = cb / a.C;
                            const int r = (int)(cb - tb * a.C);
                            unsigned int* acc = a.tbAcc + 2 * tb;
                            atomicXor(acc, gf_mulmod(remA, a.tbFac[r], polyA.poly, polyA.len));
                            __threadfence();
                            if (atomicAdd(acc + 1, 1u) == (unsigned)(a.C - 1)) {   // last block of this transport block
                                __threadfence();
                                a.tbOk[tb] = (atomicExch(acc, 0u) == 0u);
                                acc[1] = 0u;
                            }
                        }
                    }
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
        grp = nextGrpSh;
    }
    if (dynQ && tid == 0) {   // the last CTA to leave re-arms the queue for the next launch on this handle
        __threadfence();
        if (atomicAdd(a.workCounter + 2, 1u) == gridDim.x - 1) {
            a.workCounter[0] = 0u;
            a.workCounter[2] = 0u;
            if (a.esAuto) {   // every block of this launch has reported: next launch tests from (smallest count - 1)
                const unsigned int mn = atomicExch(a.esAuto + 1, 0x7fffffffu);
                a.esAuto[0] = (mn == 0x7fffffffu || mn < 2u) ? 0u : mn - 1u;
            }
            __threadfence();
        }
    }
    if (useTmem) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmemBaseSh), "r"((uint32_t)a.tmemCols) : "memory");
    }
}

}   // namespace
