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
#include "crc_device.cuh"
#include "decode_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------------------
// three-tier state storage: rows [0, tmemRows) in Tensor Memory (ONE_CB kernels), the next smemRows rows in shared
// memory planes, the rest in the per-CTA global scratch (stays in L2).  All branches are on the (uniform) row index.
// ---------------------------------------------------------------------------------------------------------------
//   ALLT = 1: every scheduled row in Tensor Memory (no tier branches);  ALLT = 2 ("split"): rows [0, kSplitRows) in Tensor
//   Memory at the same fixed stride, every further scheduled row in the shared-memory planes -- the tier of a row is then
//   known at compile time in the static schedule (22-33 scheduled rows with two resident CTAs, i.e. code rates ~0.4-0.54)
template <typename T, bool ONE_CB, int ALLT>
struct StateStore {
    uint32_t tbase;     // this thread's TMEM address of row slot 0 (lane quadrant and warp column offset folded in)
    uint32_t tstride;   // TMEM columns per row slot
    T* sS;              // shared planes, already offset by tid
    T* sG;              // global planes, already offset by tid
    int tmemRows, smemRows, nT;
    // ALLT: every scheduled row lives in Tensor Memory at a compile-time stride (3 warps per lane quadrant): no tier
    // branches, and with a static row index the TMEM address is base + immediate
    static constexpr uint32_t kAllTStride = 3u * (sizeof(T) == 4 ? 4u : 8u);   // referenced by the ALLT instantiations only
    static constexpr int kSplitRows = 21;   // 256 TMEM columns / kAllTStride (fp32)
    __device__ __forceinline__ void load(int row, RowState<T>& st) const
    {
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
template <typename T, bool ONE_CB, int SBG, int ALLT, int ESM = 1, int ZS = 0>
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
    StateStore<T, ONE_CB, ALLT> store;
    {
        const int warp = tid >> 5;
        const uint32_t RW = sizeof(T) == 4 ? 4u : 8u;
        const uint32_t wpq = ALLT != 0 ? 3u : ((uint32_t)((nT >> 5) + 3) >> 2);   // warps per lane quadrant
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
    uint32_t* pk = crcRed + 64;   // bit-packed hard decisions of the early-termination test (a.packWords words)
    float* stage = reinterpret_cast<float*>(pk + ((SBG != 0) ? a.packWords : 0));
    const bool useStage = (SBG != 0) && a.stageFloats > 0;
    LayerBarT<(ALLT == 1 ? NR_DEC_BAR_MODE : (ALLT == 2 ? NR_DEC_BAR_MODE_SPLIT : 0))> lb;
    lb.bar = barLayer;
    lb.phase = 0;
    uint32_t stagePhase = 0;
    __shared__ uint32_t liftSh[4];
    if (SBG != 0) {
        if (tid == 0) {
            liftSh[0] = ZB.S;
            liftSh[1] = ZB.ZB;
            liftSh[2] = ZB.one;
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barLayer), "r"((uint32_t)(nT >> 5)) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(barStage), "r"(1u) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();
        if (NR_DEC_LIFT_REGS) {   // per-thread register copies (see struct Lift)
            const volatile uint32_t* lv = liftSh;
            ZB.S = lv[0];
            ZB.ZB = lv[1];
            ZB.one = lv[2];
        }
    }

    const bool wantCrc = a.rm && (a.tbBits || a.cbCrcOk || a.cbRemA);
    const int Lk = a.K - a.F;                       // code block without fillers
    const int per = (a.C > 1) ? Lk - 24 : Lk;       // payload copied into the merged transport block
    const NrCrcPoly polyCb = nr_crc_poly(a.C > 1 ? NRLDPC_CRC24B : NRLDPC_CRC24A);
    const NrCrcPoly polyA = nr_crc_poly(NRLDPC_CRC24A);
    if (wantCrc) {
        if (SBG == 0) {
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
        const int es = a.inF16 ? 2 : 4, epv = 16 / es;   // element size, elements per 16 bytes
        const int head = (int)(xBase & (epv - 1));
        const int nCopy = (head + n) & ~(epv - 1);
        stage_issue(barStage, (uint32_t)__cvta_generic_to_shared(stage),
                    reinterpret_cast<const char*>(a.llr) + (xBase - head) * es, (uint32_t)(nCopy * es));
    };
    if (useStage && tid == 0 && (long long)blockIdx.x < (a.numCb + cbPerCta - 1) / cbPerCta) stage_block((long long)blockIdx.x);

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
                stream_geom(cb, E, xBase, xAvail);
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
                        if (col - 2 < a.inCols) v = llr_cvt<T, TIn>(x[n]);
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
            if (useStage && E <= L) {
                if (a.inF16) staged_load(__half()); else staged_load(float());
            } else if (a.rm && !sb && !a.inF64 && !a.inF16 && smallE) {
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
            for (int row = 0; row < 4; row++) {
                RowState<T> st0;
                st0.m1s = (T)0; st0.m2s = (T)0; st0.sw = 0; st0.rext = (T)0;
                store.store(row, st0);
            }
        }
        __syncthreads();
        // the staging buffer is free again: fetch the stream of this CTA's next code block while this one iterates
        if (useStage && tid == 0 && grp + gridDim.x < numGroups) stage_block(grp + gridDim.x);

        // -------------------------------------------------------------------------------------------------------
        // iterations
        // -------------------------------------------------------------------------------------------------------
        int itersDone = 0;
        bool cbDone = false;
        for (int it = 0; it < a.numIter; it++) {
            if constexpr (SBG != 0) {
                RowCtx<T, SBG, 0> c0;
                prep_row<T, SBG, 0, ZS>(g, mU, ZB, store, dummyOff, c0);
                run_rows_static<T, SBG, 0, (ESM != 0), ZS>(g, a.numRows, rb, mU, ZB, store, slot, dummyOff, lb, c0,
                                                       (ESM != 0 && (a.flags & NRLDPC_DEC_EARLY_STOP)) ? pk + (size_t)ncore * 2 * (nT >> 5) : nullptr);
            } else {
                for (int row = 0; row < a.numRows; row++) {
                    if (ONE_CB || (active && !cbDone)) {
                        RowState<T> st;
                        store.load(row, st);
                        dispatch_row<T>(g, row, rb, mU, ZB, st, slot, dummyOff, !a.trueMin2, a.trueMin2 ? (T)a.alpha : (T)0.75);
                        store.store(row, st);
                    }
                    __syncthreads();
                }
            }
            if (!cbDone) itersDone = it + 1;
            if constexpr (SBG != 0) {
                if (ESM != 0 && (a.flags & NRLDPC_DEC_EARLY_STOP) && it + 1 >= ((a.flags >> 8) & 0xff)) {
                    // Syndrome of the hard decisions after a COMPLETE iteration, bit-packed: every warp ballots the sign
                    // of its 32 positions of each core column (stored twice, so a circulant shift is one funnel shift of two
                    // neighbouring words); the scheduled extension columns were packed by their rows (run_rows_static); then one thread
                    // per (row, 32 checks) XORs the shifted words of the row's edges.  ~6 % of an iteration.
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
                    }
                    __syncthreads();
                    uint32_t bad = 0;
                    for (int task = tid; task < a.numRows * W; task += nT) {
                        const int row = task / W, w = task - row * W;
                        const int e0 = g.rowEdge0[row];
                        const int e1 = g.rowEdge0[row + 1] - (row >= 4 ? 1 : 0);
                        uint32_t acc = (row >= 4) ? pe[(row - 4) * W + w] : 0u;
#pragma unroll 4
                        for (int e = e0; e < e1; e++) {
                            const uint32_t raw = g.raw[e];
                            const uint32_t b = 32u * (uint32_t)w + (raw & 511u);   // first position read by these 32 checks
                            const uint32_t* pc = pk + (raw >> 9) * 2 * W + (b >> 5);
                            acc ^= __funnelshift_r(pc[0], pc[1], b & 31u);
                        }
                        bad |= acc;
                    }
                    const int anyBad = __syncthreads_or(bad != 0);
                    if (!anyBad) break;
                }
            } else
            if ((a.flags & NRLDPC_DEC_EARLY_STOP) && it + 1 >= ((a.flags >> 8) & 0xff)) {
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
            uint32_t remCb, remA;
            if constexpr (SBG != 0) {
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
    size_t miscBytes = ((size_t)((a.cbPerCta + 31) & ~31) + 32 + (size_t)a.cbPerCta * P2) * sizeof(uint32_t) + 16 +
                       (size_t)nT * (sizeof(MinSlot<T>) + sizeof(T));
    // static kernels: mbarriers, CRC factor table, XOR exchange (see the kernel's `extra` region)
    const bool staticRows = oneCb && sizeof(T) == 4 && !h->noStaticRows && !a.trueMin2;
    a.packWords = 0;
    if (staticRows && (a.flags & NRLDPC_DEC_EARLY_STOP))
        a.packWords = ((g.ncore * 2 + (a.numRows - 4) + 1) * (nT >> 5) + 3) & ~3;   // +1 row: the funnel shift reads one word past the end
    if (staticRows) miscBytes += 16 + 16 + 64 * sizeof(uint32_t) + (size_t)a.packWords * sizeof(uint32_t);
    // target resident CTAs per SM (env NRLDPC_DEC_OCC overrides): two for the fp32 one-block-per-CTA kernel, whose
    // registers are capped at 80 and whose row state lives in Tensor Memory; one otherwise
    // (measured at 17 scheduled rows, profiles/r01_decode_per_lifting_size.json: two resident CTAs lift the generic fp32
    // kernels by 11-45 %, a third one helps only the 7-8 warp CTAs of Zc = 208 / 240)
    int occ = h->decOcc > 0 ? h->decOcc : (sizeof(T) == 4 ? ((!oneCb && a.cbPerCta == 1 && nT <= 256) ? 3 : 2) : 1);
    occ = max(1, min(occ, 2048 / nT));
    if (sizeof(T) == 8) occ = 1;
    // static fp32 kernels: when the scheduled rows fit Tensor Memory only with ONE resident CTA (22-42 rows: low code rates,
    // every BG2 row), one all-TMEM CTA per SM beats two CTAs whose state spills to shared-memory planes / the L2 scratch
    // (measured: BG2 all rows 635 -> 773, 30 rows BG1 705 -> 775 G edge-updates/s); beyond 42 rows two spilling CTAs win
    // ... unless the rows beyond the 21 that fit 256 TMEM columns fit the shared-memory planes of two resident CTAs: the
    // "split" kernels keep both CTAs and know every row's tier at compile time
    bool split = false;
    if (staticRows && h->decOcc <= 0 && occ == 2 && !h->noTmem && a.numRows > 21 && !getenv("NRLDPC_NO_SPLIT")) {
        const size_t budget2 = min((size_t)h->smemPerSM / 2 - 1024, (size_t)h->maxSmemOptin);
        split = rBytes + miscBytes + (size_t)(a.numRows - 21) * rowBytes <= budget2;
    }
    if (!split && staticRows && h->decOcc <= 0 && occ == 2 && a.numRows * 3 * 4 > 256 && a.numRows * 3 * 4 <= 512) occ = 1;
    // Tensor Memory rows (ONE_CB kernels): 512 columns per SM shared by the resident CTAs
    a.tmemRows = 0;
    a.tmemCols = 0;
    bool allT = false;
    if (oneCb && !h->noTmem) {
        int cols = 32;
        while (cols * 2 <= 512 / occ) cols *= 2;
        const int RW = sizeof(T) == 4 ? 4 : 8;
        int wpq = ((nT >> 5) + 3) >> 2;
        // every scheduled row fits Tensor Memory at the fixed stride of the ALLT kernels (3 warps per lane quadrant)
        if (staticRows && cols / (3 * RW) >= a.numRows) {
            allT = true;
            wpq = 3;
        }
        if (split) wpq = 3;   // 21 rows at the fixed stride
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
    // TMA staging of the rate-matched stream (fused mode, fp32 stream, no HARQ history): one code block's E LLRs
    a.stageFloats = 0;
    a.crcFacDev = nullptr;
    if (staticRows && a.rm && (a.tbBits || a.cbCrcOk || a.cbRemA)) {
        // per-bit CRC constants x^(len-1-i) mod g, i = col*Z + m, laid out [which][col][Z]; which = 0: the code-block CRC over
        // the K-F bits, 1: the CRC24A partial over the payload part (C > 1).  0 beyond the message.
        const int Lk = a.K - a.F, per = (a.C > 1) ? Lk - 24 : Lk;
        const unsigned long long key = ((unsigned long long)(unsigned)Lk << 32) | ((unsigned)Z << 12) | ((unsigned)g.ksys << 1) | (unsigned)(a.C > 1);
        const size_t words = (size_t)2 * g.ksys * Z;
        if (!h->crcFacDev || h->crcFacKey != key) {
            if (!h->crcFacDev) NR_CUDA_CHECK(cudaMalloc(&h->crcFacDev, (size_t)2 * 22 * NR_MAX_Z * sizeof(unsigned int)));
            unsigned int* host = (unsigned int*)calloc(words, sizeof(unsigned int));
            if (!host) { nr_set_error("decode: out of host memory"); return NRLDPC_ERR_NOMEM; }
            const NrCrcPoly pc = nr_crc_poly(a.C > 1 ? NRLDPC_CRC24B : NRLDPC_CRC24A), pa = nr_crc_poly(NRLDPC_CRC24A);
            for (int which = 0; which < 2; which++) {
                const NrCrcPoly pp = which ? pa : pc;
                const int len = which ? per : Lk;
                uint32_t f = 1;   // x^0 for the last bit
                for (int i = len - 1; i >= 0; i--) {
                    host[(size_t)which * g.ksys * Z + i] = f;   // i = col*Z + m is exactly the [col][Z] layout
                    f = nr_gf_mulmod(f, 2u, pp.poly, pp.len);
                }
            }
            cudaError_t ce = cudaMemcpyAsync(h->crcFacDev, host, words * sizeof(unsigned int), cudaMemcpyHostToDevice, s);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);   // `host` is freed below
            free(host);
            NR_CUDA_CHECK(ce);
            h->crcFacKey = key;
        }
        a.crcFacDev = (const unsigned int*)h->crcFacDev;
    }
    if (staticRows && a.rm && !a.softBuf && !a.inF64 && !h->noStage && (reinterpret_cast<uintptr_t>(a.llr) & 15) == 0) {
        const int Emax = a.E0 + ((a.nShort < a.C) ? a.fStep : 0);
        const int es = a.inF16 ? 2 : 4, epv = 16 / es;
        const size_t need = ((size_t)((Emax + 2 * (epv - 1)) & ~(epv - 1)) * es + 15) & ~(size_t)15;
        const size_t used = rBytes + (size_t)smemRows * rowBytes + miscBytes;
        if (used + need <= budget && need <= (size_t)(1u << 19)) a.stageFloats = (int)(need / sizeof(float));
    }
    const size_t smem = rBytes + (size_t)smemRows * rowBytes + miscBytes + (size_t)a.stageFloats * sizeof(float);
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
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)grid, nT, smem, s>>>(dg, a);
        return cudaSuccess;
    };
    constexpr int S1 = sizeof(T) == 4 ? 1 : 0, S2 = sizeof(T) == 4 ? 2 : 0;   // static schedules exist in fp32 only
    constexpr int AT = sizeof(T) == 4 ? 1 : 0, SP = sizeof(T) == 4 ? 2 : 0;
    // the all-TMEM / split kernels exist with and without the early-termination code (its mere presence costs the row loop a few %)
    constexpr int Z384 = sizeof(T) == 4 ? 384 : 0;
    // compile-time edge table (SpecTab) for the largest lifting size: no uniform table loads in front of a row, +2-3 %
    // (bench 20.07 -> 20.67 Gbit/s)
    const bool z384 = sizeof(T) == 4 && Z == 384 && !getenv("NRLDPC_NO_SPECZ");
    const bool noEs = sizeof(T) == 4 && !(a.flags & NRLDPC_DEC_EARLY_STOP) && !getenv("NRLDPC_ES_CODE");
    if (split && (a.tmemRows != 21 || a.smemRows != a.numRows - 21)) { nr_set_error("decode: internal error (split state layout)"); return NRLDPC_ERR_ARG; }
    if (staticRows && g.P == NR_BG1_ROWS) {
        // (the BG1 split kernel spills with the compile-time table -- 861 vs 914 G edge-updates/s -- and keeps the run-time one)
        if (allT && noEs && z384) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S1, AT, 0, Z384>));
        else if (split && noEs) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S1, SP, 0>));
        else if (split) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S1, SP>));
        else if (allT && noEs) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S1, AT, 0>));
        else if (allT) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S1, AT>));
        else NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S1, 0>));
    } else if (staticRows) {
        if (split && noEs && z384) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S2, SP, 0, Z384>));
        else if (allT && noEs && z384) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S2, AT, 0, Z384>));
        else if (split && noEs) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S2, SP, 0>));
        else if (split) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S2, SP>));
        else if (allT && noEs) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S2, AT, 0>));
        else if (allT) NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S2, AT>));
        else NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, S2, 0>));
    } else if (oneCb) {
        NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, 0, 0>));
    } else {
        NR_CUDA_CHECK(launch(nr_decode_kernel<T, false, 0, 0>));
    }
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

int dispatch_decode(nrldpc_handle* h, const NrGraph& g, DecArgs& a, int inDtype, int computeDtype, cudaStream_t s)
{
    if (inDtype != NRLDPC_F32 && inDtype != NRLDPC_F64 && !(inDtype == NRLDPC_F16 && a.rm)) {
        nr_set_error("decode: bad input dtype (NRLDPC_F16 is accepted by nrldpc_decode_tb only)");
        return NRLDPC_ERR_ARG;
    }
    a.inF64 = (inDtype == NRLDPC_F64);
    a.inF16 = (inDtype == NRLDPC_F16);
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
    if (in_dtype != NRLDPC_F32 && in_dtype != NRLDPC_F64) {
        nr_set_error("decode: bad input dtype (NRLDPC_F16 is accepted by nrldpc_decode_tb only)");
        return NRLDPC_ERR_ARG;
    }
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

extern "C" int nrldpc_decode2(nrldpc_handle* h, int bg, int zc, int in_dtype, int compute_dtype, const void* llr,
                              int64_t num_cb, int64_t llr_stride, int in_cols, int max_iter, double alpha,
                              int stop_on_good_parity, int out_cols, int8_t* bits, void* beliefs, int32_t* iters,
                              nrldpc_stream stream)
{
    if (!h) { nr_set_error("decode2: null handle"); return NRLDPC_ERR_ARG; }
    NrGraph g;
    if (nr_build_graph(bg, zc, &g)) return NRLDPC_ERR_ARG;
    if (num_cb <= 0 || in_cols < 0 || in_cols > g.ncols - 2 || out_cols < 1 || out_cols > g.ncols || max_iter < 0 ||
        llr_stride < (int64_t)in_cols * zc || !(alpha == alpha)) {
        nr_set_error("decode2: bad arguments");
        return NRLDPC_ERR_ARG;
    }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    DecArgs a{};
    a.numCb = num_cb;
    a.numIter = max_iter;
    a.flags = NRLDPC_DEC_ALL_ROWS | (stop_on_good_parity ? NRLDPC_DEC_EARLY_STOP : 0);
    a.llr = llr;
    a.llrStride = llr_stride;
    a.inCols = in_cols;
    a.outCols = out_cols;
    a.bits = (signed char*)bits;
    a.bitsStride = (long long)out_cols * zc;
    a.beliefs = beliefs;
    a.iters = iters;
    a.numRows = g.P;          // every row: the closed form of skipped rows is specific to the standard rule
    a.trueMin2 = 1;
    a.alpha = alpha;
    return dispatch_decode(h, g, a, in_dtype, compute_dtype, (cudaStream_t)stream);
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
