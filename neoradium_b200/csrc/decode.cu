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
#include "decode_kernel.cuh"
#include "decode_launch.cuh"

namespace {

// shared memory a resident CTA costs on top of its dynamic allocation: 1 KB reserved by the system + the kernel's static variables
// (~1.6 KB).  Counting only the 1 KB let the planes of the tiered kernels fill the budget so exactly that the second CTA no longer
// fitted (BG1 Zc=288, all 46 rows: ONE resident CTA, 328 instead of ~700 G edge-updates/s).
constexpr size_t kCtaSmemOverhead = 1024 + 2048;
constexpr int NR_DEC_UNSUPPORTED = -1000;   // internal: this launch shape cannot take the requested input form (never leaves the library)

// last position (punctured frame) holding a non-zero LLR, max over the batch -> numRows for mode A.  A CTA walks whole blocks
// from their END in chunks of 1024 values (coalesced 4/8-byte loads, four per thread in flight) and leaves a block at the first
// chunk with a non-zero value or when it reaches the best position any block has reported so far: the pass reads the zero tail
// of the batch once (nothing at all when the last column is in use) instead of every value with a 64-bit division each.
template <typename TIn>
__global__ void __launch_bounds__(256) nr_last_nonzero_kernel(const TIn* __restrict__ llr, long long numCb, long long stride, int len, int* lastPos)
{
    constexpr int V = 4, CH = 256 * V;
    __shared__ int loSh;
    for (long long cb = blockIdx.x; cb < numCb; cb += gridDim.x) {
        const TIn* __restrict__ p = llr + cb * stride;
        if (threadIdx.x == 0) loSh = *reinterpret_cast<volatile int*>(lastPos) + 1;   // positions below the current best cannot raise it
        __syncthreads();
        const int lo = loSh;   // one value for the whole CTA: the loop below holds barriers
        for (int hi = len; hi > lo; hi -= CH) {
            int best = -1;
            TIn v[V];
#pragma unroll
            for (int k = 0; k < V; k++) {
                const int n = hi - 1 - (int)threadIdx.x - k * 256;
                v[k] = (n >= lo) ? p[n] : (TIn)0;
            }
#pragma unroll
            for (int k = 0; k < V; k++)
                if (v[k] != (TIn)0) best = max(best, hi - 1 - (int)threadIdx.x - k * 256);
            if (__syncthreads_or(best >= 0)) {
                for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(0xffffffffu, best, o));
                if ((threadIdx.x & 31) == 0 && best >= 0) atomicMax(lastPos, best);
                break;
            }
        }
        __syncthreads();   // the atomicMax of this block is issued before the next block reads the bound
    }
}

template <typename T>
int launch_decode(nrldpc_handle* h, const NrGraph& g, DecArgs& a, cudaStream_t s, bool allowMulti = true)
{
    const int Z = g.Z;
    a.cbPerCta = max(1, 384 / Z);
    if ((long long)a.cbPerCta > a.numCb) a.cbPerCta = (int)a.numCb;
    int nT = a.cbPerCta * Z;
    nT = (nT + 31) & ~31;
    const bool oneCb = (a.cbPerCta == 1 && nT == Z);
    int P2 = 1;
    while (P2 < Z) P2 <<= 1;
    size_t rBytes = (size_t)a.cbPerCta * g.ncore * Z * sizeof(T);
    const size_t rowBytes = (size_t)NPLANES * nT * sizeof(T);
    size_t miscBytes = ((size_t)((a.cbPerCta + 31) & ~31) + 32 + (size_t)a.cbPerCta * P2) * sizeof(uint32_t) + 16 +
                       (size_t)nT * (sizeof(MinSlot<T>) + sizeof(T));
    // static kernels: mbarriers, CRC factor table, XOR exchange (see the kernel's `extra` region)
    // statically scheduled kernels: one block per CTA (Zc a multiple of 32, >= 224) or, without early termination, several
    // blocks per CTA (every other lifting size; decode_kernel.cuh "MB")
    const bool multiStatic = !oneCb && allowMulti && sizeof(T) == 4 && !h->noStaticRows && !a.trueMin2 &&
                             !(a.flags & NRLDPC_DEC_EARLY_STOP) && Z >= 2 && !getenv("NRLDPC_NO_STATIC_MB");
    const bool staticRows = (oneCb && sizeof(T) == 4 && !h->noStaticRows && !a.trueMin2) || multiStatic;
    a.nAreas = multiStatic ? (nT + Z - 1) / Z : a.cbPerCta;   // padding threads: tid / Z can reach past cbPerCta by more than one (Zc < 16)
    if (staticRows) rBytes = (size_t)a.nAreas * (g.ncore + 1) * Z * sizeof(T);   // + the dummy row (decode_static.cuh)
    a.packWords = 0;
    if (staticRows && (a.flags & NRLDPC_DEC_EARLY_STOP))
        a.packWords = (((g.ncore * 2 + (a.numRows - 4) + 1) * (nT >> 5) + 3) & ~3)   // packed bits; +1 row: the funnel shift reads one word past the end
                      + ((a.numRows * (nT >> 5) + 3) & ~3) + 160 + 160;                 // syndrome words, task list, edge table (decode_kernel.cuh)
    if (staticRows) miscBytes += 16 + 16 + 64 * sizeof(uint32_t) + (size_t)a.packWords * sizeof(uint32_t);
    // target resident CTAs per SM (env NRLDPC_DEC_OCC overrides): two for the fp32 one-block-per-CTA kernel, whose
    // registers are capped at 80 and whose row state lives in Tensor Memory; one otherwise
    // (measured at 17 scheduled rows, profiles/r01_decode_per_lifting_size.json: two resident CTAs lift the generic fp32
    // kernels by 11-45 %, a third one helps only the 7-8 warp CTAs of Zc = 208 / 240)
    int occ = h->decOcc > 0 ? h->decOcc : (sizeof(T) == 4 ? ((!oneCb && !multiStatic && a.cbPerCta == 1 && nT <= 256) ? 3 : 2) : 1);
    occ = max(1, min(occ, 2048 / nT));
    if (sizeof(T) == 8) occ = 1;
    // static fp32 kernels: when the scheduled rows fit Tensor Memory only with ONE resident CTA (22-42 rows: low code rates,
    // every BG2 row), one all-TMEM CTA per SM beats two CTAs whose state spills to shared-memory planes / the L2 scratch
    // (measured: BG2 all rows 635 -> 773, 30 rows BG1 705 -> 775 G edge-updates/s); beyond 42 rows two spilling CTAs win
    // ... unless the rows beyond the 21 that fit 256 TMEM columns fit the shared-memory planes of two resident CTAs: the
    // "split" kernels keep both CTAs and know every row's tier at compile time
    // CTAs of at most 8 warps (Zc <= 256; no early termination): THREE per SM, each with 128 Tensor-Memory columns = 16 rows at two
    // warps per lane quadrant; further rows in shared-memory planes when they fit a third of the SM ("w8" instantiations)
    bool w8 = false, w8AllT = false, w8Tiered = false, w8Eligible = false;
    if (staticRows && sizeof(T) == 4 && nT <= 256 && !(a.flags & NRLDPC_DEC_EARLY_STOP) && h->decOcc <= 0 && !h->noTmem &&
        !getenv("NRLDPC_NO_W8")) {
        const size_t budget3 = min((size_t)h->smemPerSM / 3 - kCtaSmemOverhead, (size_t)h->maxSmemOptin);
        w8AllT = a.numRows <= 16;
        w8 = rBytes + miscBytes + (size_t)(w8AllT ? 0 : a.numRows - 16) * rowBytes <= budget3;
        if (w8) occ = 3;
        else w8Eligible = rBytes + miscBytes + 2 * rowBytes <= budget3;   // candidates of the "w8 tiered" route below
    }
    bool split = w8 && !w8AllT;
    if (!w8 && staticRows && h->decOcc <= 0 && occ == 2 && !h->noTmem && a.numRows > 21 && !getenv("NRLDPC_NO_SPLIT")) {
        const size_t budget2 = min((size_t)h->smemPerSM / 2 - kCtaSmemOverhead, (size_t)h->maxSmemOptin);
        split = rBytes + miscBytes + (size_t)(a.numRows - 21) * rowBytes <= budget2;
    }
    // low code rates on narrow CTAs: when the rows fit neither the three-CTA planes nor the two-CTA split, keep THREE CTAs per SM with
    // the rows beyond Tensor Memory and the planes in the L2 scratch ("w8 tiered") instead of two tiered CTAs / one all-TMEM CTA
    // (measured at 8 waves, G edge-updates/s: all 46 rows of BG1 Zc=256 572 -> 747, Zc=224 501 -> 689; 42 rows Zc=256 668 -> 752,
    // Zc=240 574 -> 674, Zc=208 507 -> 623; BG2 all rows Zc=240 513 -> 626.  Where the split fits it stays: 1016 vs 783 at 36 rows)
    if (w8Eligible && !split) {
        const char* wt = getenv("NRLDPC_W8_TIERED_FROM");
        const bool wouldBeOne = a.numRows * 3 * 4 > 256 && a.numRows * 3 * 4 <= 512;
        if (a.numRows >= (wt ? atoi(wt) : 43) || (wouldBeOne && !getenv("NRLDPC_NO_W8_TIERED_OCC1"))) {
            w8Tiered = true;
            occ = 3;
        }
    }
    if (!w8 && !split && staticRows && h->decOcc <= 0 && occ == 2 && a.numRows * 3 * 4 > 256 && a.numRows * 3 * 4 <= 512) occ = 1;
    // Tensor Memory rows (ONE_CB kernels): 512 columns per SM shared by the resident CTAs
    a.tmemRows = 0;
    a.tmemCols = 0;
    bool allT = w8 && w8AllT;
    if (w8) {
        a.tmemRows = a.numRows < 16 ? a.numRows : 16;
        a.tmemCols = 128;
    } else if (w8Tiered) {
        const int wpq = ((nT >> 5) + 3) >> 2;   // the kernel's run-time stride (ALLT = 0)
        a.tmemRows = min(a.numRows, 128 / (wpq * 4));
        a.tmemCols = 128;
    } else if ((oneCb || multiStatic) && !h->noTmem) {
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
    // (multi-block static kernels exist for all three state layouts: all-TMEM, split and -- low code rates -- tiered)
    // (measured, all 46 rows of BG1: +4..8 % over the generic kernel at Zc <= 192; CTAs of at most 8 warps -- Zc = 208, 240 -- are
    // better off with THREE resident generic CTAs: 604 / 675 vs 461 / 528 G edge-updates/s)
    if (multiStatic && !allT && !split && !w8Tiered && (nT <= 256 || getenv("NRLDPC_NO_STATIC_MB_TIERED"))) return launch_decode<T>(h, g, a, s, false);
    const int restRows = a.numRows - a.tmemRows;
    size_t budget = (size_t)h->smemPerSM / occ - kCtaSmemOverhead;
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
    if (staticRows && oneCb && a.rm && (a.tbBits || a.cbCrcOk || a.tbOk)) {
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
    if (staticRows && oneCb && a.rm && !a.softBuf && !a.inF64 && !h->noStage && (reinterpret_cast<uintptr_t>(a.llr) & 15) == 0) {
        const int Emax = a.E0 + ((a.nShort < a.C) ? a.fStep : 0);
        const int es = a.inSym ? 8 : (a.inF16 ? 2 : 4), epv = 16 / es;   // symbols: Emax / qm complex64 values
        const int elems = a.inSym ? Emax / a.qm : Emax;
        const size_t need = ((size_t)((elems + 2 * (epv - 1)) & ~(epv - 1)) * es + 15) & ~(size_t)15;
        const size_t used = rBytes + (size_t)smemRows * rowBytes + miscBytes;
        // half-precision streams are consumed from the staging buffer only by the no-repetition load (E <= Ncb - F): a launch
        // that holds a longer block would leave its copy unconsumed and the barrier phase out of step, so it does not stage
        const bool f16Wrap = (a.inF16 || a.inSym) && Emax > a.ncb - a.F;
        if (used + need <= budget && need <= (size_t)(1u << 19) && !f16Wrap) a.stageFloats = (int)(need / sizeof(float));
    }
    // symbol input is consumed by the staged load of the one-block static fp32 kernels only: the caller demaps first otherwise
    if (a.inSym && !(a.stageFloats > 0 && staticRows && oneCb && sizeof(T) == 4)) return NR_DEC_UNSUPPORTED;
    const size_t smem = rBytes + (size_t)smemRows * rowBytes + miscBytes + (size_t)a.stageFloats * sizeof(float);
    const long long numGroups = (a.numCb + a.cbPerCta - 1) / a.cbPerCta;
    int perSM = (int)((size_t)h->smemPerSM / (smem + kCtaSmemOverhead));
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
    a.workCounter = getenv("NRLDPC_NO_DYNQ") ? nullptr : h->workCounter;   // dynamic work queue (decode_kernel.cuh)
    a.esAuto = (a.workCounter && (a.flags & NRLDPC_DEC_EARLY_STOP) && (a.flags & NRLDPC_DEC_ES_AUTO)) ? h->workCounter + 4 : nullptr;
    NrDecGraph dg;
    build_dec_graph<T>(g, &dg, staticRows);
    auto launch = [&](auto kern) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<(unsigned)grid, nT, smem, s>>>(dg, a);
        return cudaGetLastError();
    };
    // the all-TMEM / split kernels exist with and without the early-termination code (its mere presence costs the row loop a
    // few %); a compile-time edge table (SpecTab) exists for the largest lifting size (no table operands in front of a row)
    const bool z384 = sizeof(T) == 4 && Z == 384 && !getenv("NRLDPC_NO_SPECZ");
    const bool noEs = sizeof(T) == 4 && !(a.flags & NRLDPC_DEC_EARLY_STOP) && !getenv("NRLDPC_ES_CODE");
    if (split && (a.tmemRows != (w8 ? 16 : 21) || a.smemRows != a.numRows - (w8 ? 16 : 21))) { nr_set_error("decode: internal error (split state layout)"); return NRLDPC_ERR_ARG; }
    if (staticRows) {
        const int allt = split ? 2 : (allT ? 1 : 0);
        const bool bg1 = g.P == NR_BG1_ROWS;
        // preference order: no early-termination code + compile-time table, no early-termination code, everything
        auto try_launch = [&](int esm, int zs) -> cudaError_t {
            if (w8 || w8Tiered) return bg1 ? nr_launch_static_bg1_w8(allt, oneCb ? 1 : 0, &dg, &a, (unsigned)grid, nT, smem, s)
                               : nr_launch_static_bg2_w8(allt, oneCb ? 1 : 0, &dg, &a, (unsigned)grid, nT, smem, s);
            if (multiStatic) return bg1 ? nr_launch_static_bg1_mb(allt, esm, zs, &dg, &a, (unsigned)grid, nT, smem, s)
                                        : nr_launch_static_bg2_mb(allt, esm, zs, &dg, &a, (unsigned)grid, nT, smem, s);
            if (esm) return bg1 ? nr_launch_static_bg1_es(allt, 1, zs, &dg, &a, (unsigned)grid, nT, smem, s)
                                : nr_launch_static_bg2_es(allt, 1, zs, &dg, &a, (unsigned)grid, nT, smem, s);
            return bg1 ? nr_launch_static_bg1(allt, 0, zs, &dg, &a, (unsigned)grid, nT, smem, s)
                       : nr_launch_static_bg2(allt, 0, zs, &dg, &a, (unsigned)grid, nT, smem, s);
        };
        cudaError_t e = cudaErrorNotSupported;
        if (noEs && allt != 0 && z384) e = try_launch(0, 384);
        if (e == cudaErrorNotSupported && noEs && (allt != 0 || !getenv("NRLDPC_TIERED_ES_CODE"))) e = try_launch(0, 0);
        if (e == cudaErrorNotSupported && multiStatic) e = try_launch(0, 0);   // tiered multi-block kernel
        if (e == cudaErrorNotSupported && z384) e = try_launch(1, 384);
        if (e == cudaErrorNotSupported) e = try_launch(1, 0);
        NR_CUDA_CHECK(e);
    } else if (oneCb) {
        NR_CUDA_CHECK(launch(nr_decode_kernel<T, true, 0, 0>));
    } else {
        NR_CUDA_CHECK(launch(nr_decode_kernel<T, false, 0, 0>));
    }
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
            const int blocks = (int)min((long long)h->numSMs * 8, (long long)num_cb);
            if (in_dtype == NRLDPC_F32)
                nr_last_nonzero_kernel<float><<<blocks, 256, 0, s>>>((const float*)llr, num_cb, llr_stride, len, d);
            else
                nr_last_nonzero_kernel<double><<<blocks, 256, 0, s>>>((const double*)llr, num_cb, llr_stride, len, d);
            NR_CUDA_CHECK(cudaGetLastError());
        }
        int last = -1;   // position, then column
        NR_CUDA_CHECK(cudaMemcpyAsync(&last, d, sizeof(int), cudaMemcpyDeviceToHost, s));
        NR_CUDA_CHECK(cudaStreamSynchronize(s));
        if (last >= 0) last /= zc;
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
    return nrldpc_decode2_offset(h, bg, zc, in_dtype, compute_dtype, llr, num_cb, llr_stride, in_cols, max_iter, alpha, 0.0,
                                 stop_on_good_parity, out_cols, bits, beliefs, iters, stream);
}

extern "C" int nrldpc_decode2_offset(nrldpc_handle* h, int bg, int zc, int in_dtype, int compute_dtype, const void* llr,
                                     int64_t num_cb, int64_t llr_stride, int in_cols, int max_iter, double alpha, double beta,
                                     int stop_on_good_parity, int out_cols, int8_t* bits, void* beliefs, int32_t* iters,
                                     nrldpc_stream stream)
{
    if (!h) { nr_set_error("decode2: null handle"); return NRLDPC_ERR_ARG; }
    NrGraph g;
    if (nr_build_graph(bg, zc, &g)) return NRLDPC_ERR_ARG;
    if (num_cb <= 0 || in_cols < 0 || in_cols > g.ncols - 2 || out_cols < 1 || out_cols > g.ncols || max_iter < 0 ||
        llr_stride < (int64_t)in_cols * zc || !(alpha == alpha) || !(beta >= 0.0)) {
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
    a.beta = beta;
    a.synRows = (stop_on_good_parity == 2) ? 1 : 0;   // 2: the reference's first-row-only stop test (ldpc.py:841-843, 1483-1485)
    return dispatch_decode(h, g, a, in_dtype, compute_dtype, (cudaStream_t)stream);
}

// nrldpc_decode_tb; noise_var > 0: `llr` holds complex64 symbols (llr_len / llr_stride still count LLRs), NR_DEC_UNSUPPORTED when
// this configuration has no fused symbol path
static int decode_tb_impl(nrldpc_handle* h, const nrldpc_tb_config* cfg, int in_dtype, int compute_dtype,
                          const void* llr, int64_t num_tb, int64_t llr_len, int64_t llr_stride,
                          void* soft_buffer, int num_iter, int flags, int8_t* tb_bits, int64_t tb_bits_stride,
                          uint8_t* cb_crc_ok, uint8_t* tb_crc_ok, int32_t* iters, nrldpc_stream stream, double noise_var)
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
    a.inSym = noise_var > 0.0 ? 1 : 0;
    a.invN0 = noise_var > 0.0 ? 1.0 / noise_var : 0.0;   // the demapper's own expression (linksim.cu)
    a.K = cfg->K; a.F = cfg->F; a.C = cfg->C; a.qm = cfg->qm; a.ncb = cfg->ncb;
    nr_tb_split(cfg, N, &a.E0, &a.nShort, &a.fStep, &a.k0);
    a.softBuf = soft_buffer;
    a.outCols = g.ksys;
    a.iters = iters;
    a.tbBits = (signed char*)tb_bits;
    a.tbBitsStride = tb_bits_stride;
    a.cbCrcOk = cb_crc_ok;
    const int Lk = cfg->K - cfg->F, per = cfg->C > 1 ? Lk - 24 : Lk;
    // transport-block CRC24A: combined by the decoder kernel itself (DecArgs::tbAcc), no second launch
    a.tbOk = tb_crc_ok;
    if (tb_crc_ok && cfg->C > 1) {
        const size_t accBytes = (size_t)num_tb * 2 * sizeof(unsigned int);
        if (accBytes > h->tbAccBytes) {
            if (h->tbAcc) NR_CUDA_CHECK(cudaFree(h->tbAcc));
            h->tbAcc = nullptr;
            h->tbAccBytes = 0;
            NR_CUDA_CHECK(cudaMalloc(&h->tbAcc, accBytes));
            h->tbAccBytes = accBytes;
            NR_CUDA_CHECK(cudaMemsetAsync(h->tbAcc, 0, accBytes, s));   // the kernels leave it zeroed
        }
        const unsigned long long key = ((unsigned long long)(unsigned)per << 32) | (unsigned)cfg->C;
        if (!h->tbFacDev || h->tbFacKey != key || (size_t)cfg->C > h->tbFacCap) {
            if ((size_t)cfg->C > h->tbFacCap) {
                if (h->tbFacDev) NR_CUDA_CHECK(cudaFree(h->tbFacDev));
                h->tbFacDev = nullptr;
                h->tbFacCap = 0;
                NR_CUDA_CHECK(cudaMalloc(&h->tbFacDev, (size_t)cfg->C * sizeof(unsigned int)));
                h->tbFacCap = (size_t)cfg->C;
            }
            unsigned int* host = (unsigned int*)malloc((size_t)cfg->C * sizeof(unsigned int));
            if (!host) { nr_set_error("decode_tb: out of host memory"); return NRLDPC_ERR_NOMEM; }
            const NrCrcPoly pa = nr_crc_poly(NRLDPC_CRC24A);
            uint32_t xp = 1, base = 2;   // x^per mod g by square and multiply
            for (int e = per; e; e >>= 1) {
                if (e & 1) xp = nr_gf_mulmod(xp, base, pa.poly, pa.len);
                base = nr_gf_mulmod(base, base, pa.poly, pa.len);
            }
            uint32_t f = 1;
            for (int r = cfg->C - 1; r >= 0; r--) {
                host[r] = f;
                f = nr_gf_mulmod(f, xp, pa.poly, pa.len);
            }
            cudaError_t ce = cudaMemcpyAsync(h->tbFacDev, host, (size_t)cfg->C * sizeof(unsigned int), cudaMemcpyHostToDevice, s);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);   // `host` is freed below
            free(host);
            NR_CUDA_CHECK(ce);
            h->tbFacKey = key;
        }
        a.tbAcc = (unsigned int*)h->tbAcc;
        a.tbFac = (const unsigned int*)h->tbFacDev;
    }
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
    return dispatch_decode(h, g, a, in_dtype, compute_dtype, s);
}

extern "C" int nrldpc_decode_tb(nrldpc_handle* h, const nrldpc_tb_config* cfg, int in_dtype, int compute_dtype,
                                const void* llr, int64_t num_tb, int64_t llr_len, int64_t llr_stride,
                                void* soft_buffer, int num_iter, int flags, int8_t* tb_bits, int64_t tb_bits_stride,
                                uint8_t* cb_crc_ok, uint8_t* tb_crc_ok, int32_t* iters, nrldpc_stream stream)
{
    return decode_tb_impl(h, cfg, in_dtype, compute_dtype, llr, num_tb, llr_len, llr_stride, soft_buffer, num_iter, flags, tb_bits,
                          tb_bits_stride, cb_crc_ok, tb_crc_ok, iters, stream, 0.0);
}

extern "C" int nrldpc_decode_tb_symbols(nrldpc_handle* h, const nrldpc_tb_config* cfg, const float* symbols, int64_t num_tb,
                                        int64_t num_sym, int64_t sym_stride, double noise_var, int num_iter, int flags,
                                        int8_t* tb_bits, int64_t tb_bits_stride, uint8_t* cb_crc_ok, uint8_t* tb_crc_ok,
                                        int32_t* iters, nrldpc_stream stream)
{
    if (!h || !cfg || !symbols) { nr_set_error("decode_tb_symbols: null argument"); return NRLDPC_ERR_ARG; }
    if (num_tb <= 0 || num_sym < 0 || sym_stride < num_sym || !(noise_var > 0.0) || cfg->qm < 1) {
        nr_set_error("decode_tb_symbols: bad arguments");
        return NRLDPC_ERR_ARG;
    }
    const int64_t llrLen = num_sym * cfg->qm, llrStride = sym_stride * cfg->qm;
    int rc = NR_DEC_UNSUPPORTED;
    if (!getenv("NRLDPC_NO_FUSED_DEMAP"))
        rc = decode_tb_impl(h, cfg, NRLDPC_F32, NRLDPC_F32, symbols, num_tb, llrLen, llrStride, nullptr, num_iter, flags, tb_bits,
                            tb_bits_stride, cb_crc_ok, tb_crc_ok, iters, stream, noise_var);
    if (rc != NR_DEC_UNSUPPORTED) return rc;
    // no fused form for this configuration (several blocks per CTA, repetition, no staging room): demap into a scratch buffer
    // of the handle, then the ordinary fused chain -- the same LLRs either way
    const size_t bytes = (size_t)num_tb * (size_t)llrLen * sizeof(float);
    if (bytes > h->symLlrBytes) {
        if (h->symLlr) NR_CUDA_CHECK(cudaFree(h->symLlr));
        h->symLlr = nullptr;
        h->symLlrBytes = 0;
        NR_CUDA_CHECK(cudaMalloc(&h->symLlr, bytes));
        h->symLlrBytes = bytes;
    }
    if (sym_stride == num_sym) {
        rc = nrldpc_demap_maxlog(h, cfg->qm, NRLDPC_F32, symbols, num_tb * num_sym, noise_var, NRLDPC_F32, h->symLlr, stream);
        if (rc) return rc;
    } else {
        for (int64_t t = 0; t < num_tb; t++) {
            rc = nrldpc_demap_maxlog(h, cfg->qm, NRLDPC_F32, symbols + 2 * t * sym_stride, num_sym, noise_var, NRLDPC_F32,
                                     (float*)h->symLlr + t * llrLen, stream);
            if (rc) return rc;
        }
    }
    return decode_tb_impl(h, cfg, NRLDPC_F32, NRLDPC_F32, h->symLlr, num_tb, llrLen, llrLen, nullptr, num_iter, flags, tb_bits,
                          tb_bits_stride, cb_crc_ok, tb_crc_ok, iters, stream, 0.0);
}

extern "C" int nrldpc_decode_tb_groups(nrldpc_handle* h, const nrldpc_tb_group* groups, int num_groups, int compute_dtype,
                                       int num_iter, int flags, nrldpc_stream stream)
{
    if (!h || !groups || num_groups <= 0) { nr_set_error("decode_tb_groups: bad argument"); return NRLDPC_ERR_ARG; }
    auto run = [&](nrldpc_handle* hh, const nrldpc_tb_group& g, cudaStream_t s) {
        return nrldpc_decode_tb(hh, &g.cfg, g.in_dtype, compute_dtype, g.llr, g.num_tb, g.llr_len, g.llr_stride, g.soft_buffer,
                                num_iter, flags, g.tb_bits, g.tb_bits_stride, g.cb_crc_ok, g.tb_crc_ok, g.iters, (nrldpc_stream)s);
    };
    cudaStream_t s = (cudaStream_t)stream;
    if (num_groups == 1) return run(h, groups[0], s);
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const int nsub = num_groups < 4 ? num_groups : 4;
    if (!h->subFork) NR_CUDA_CHECK(cudaEventCreateWithFlags(&h->subFork, cudaEventDisableTiming));
    for (int i = 0; i < nsub; i++) {
        if (!h->sub[i]) {
            int rc = nrldpc_create(h->device, &h->sub[i]);
            if (rc) return rc;
            NR_CUDA_CHECK(cudaStreamCreateWithFlags(&h->subStream[i], cudaStreamNonBlocking));
            NR_CUDA_CHECK(cudaEventCreateWithFlags(&h->subJoin[i], cudaEventDisableTiming));
        }
    }
    NR_CUDA_CHECK(cudaEventRecord(h->subFork, s));
    for (int i = 0; i < nsub; i++) NR_CUDA_CHECK(cudaStreamWaitEvent(h->subStream[i], h->subFork, 0));
    int rc = NRLDPC_OK;
    for (int i = 0; i < num_groups && rc == NRLDPC_OK; i++) rc = run(h->sub[i % nsub], groups[i], h->subStream[i % nsub]);
    for (int i = 0; i < nsub; i++) {   // join even after an error: the caller's stream must not run ahead of queued work
        cudaEventRecord(h->subJoin[i], h->subStream[i]);
        cudaStreamWaitEvent(s, h->subJoin[i], 0);
    }
    return rc;
}
