// On-device link around the codec: Gray-mapped QAM of TS 38.211 5.1, complex AWGN and the
// max-log LLR demapper, so that BLER sweeps and benchmarks never round-trip through the host.
//
// Mirrors Modem.modulate / Modem.getLLRsFromSymbols(useMax=True) of neoradium/modulation.py:127-204:
//   * constellation point of the qm-bit label b0..b(qm-1) (MSB first, modulation.py:64-74): per axis the recursion
//     a = 2^(q/2) - (1 - 2 b[qm-q]) a for q = 2, 4, .., qm-2 starting from 1, times (1 - 2 b0) [real] or (1 - 2 b1) [imag],
//     scaled by 1/sqrt(2 | 2 | 10 | 42 | 170 | 682); BPSK puts the same sign on both axes;
//   * max-log LLR of bit i: (min over points with b_i = 1 of |y - x|^2 - min over points with b_i = 0) / noiseVar, positive => 0.
//     The reference searches the full 2-D constellation; for these square constellations with independent per-axis Gray
//     labels the other axis cancels in the difference, so the search here is per axis (2^(qm/2) levels).  Same value
//     in exact arithmetic, a few ulp apart in floating point (the reference takes |.| by hypot and squares it again):
//     the parity tests state the tolerance.
// HBM-bound streaming kernels: a thread owns a group of symbols whose LLRs fill whole 16-byte stores.
#include <math.h>

#include "demap_device.cuh"
#include "nrldpc_internal.cuh"

namespace {

constexpr int LS_THREADS = 256;

// integer amplitudes (re, im) of the label `v` (qm bits, b0 = MSB), modulation.py:64-72
__device__ __forceinline__ void qam_point(uint32_t v, int qm, int& re, int& im)
{
    auto bit = [&](int i) { return (int)((v >> (qm - 1 - i)) & 1u); };
    re = 1;
    im = 1;
    for (int q = 2; q < qm; q += 2) {
        re = (1 << (q / 2)) - (1 - 2 * bit(qm - q)) * re;
        im = (1 << (q / 2)) - (1 - 2 * bit(qm + 1 - q)) * im;
    }
    re *= 1 - 2 * bit(0);
    im *= 1 - 2 * bit(qm > 1 ? 1 : 0);
}

template <typename T>
__global__ void __launch_bounds__(LS_THREADS)
    nr_modulate_kernel(const signed char* __restrict__ bits, long long numSym, int qm, T* __restrict__ out)
{
    const T scale = (T)qam_scale(qm);
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < numSym; s += (long long)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        for (int i = 0; i < qm; i++) v = (v << 1) | (uint32_t)(bits[s * qm + i] & 1);
        int re, im;
        qam_point(v, qm, re, im);
        out[2 * s] = scale * (T)re;
        out[2 * s + 1] = scale * (T)im;
    }
}

// per-axis max-log LLRs of one received coordinate y: llr[p] for the `half` label bits of the axis
template <typename T, int MAXH>
__device__ __forceinline__ void axis_llr(T y, int half, const T* __restrict__ levels, T invN0, T (&llr)[MAXH])
{
    T m0[MAXH], m1[MAXH];
#pragma unroll
    for (int p = 0; p < MAXH; p++) { m0[p] = (T)INFINITY; m1[p] = (T)INFINITY; }
    const int nl = 1 << half;
    for (int l = 0; l < nl; l++) {
        const T d = y - levels[l];
        const T d2 = d * d;
#pragma unroll
        for (int p = 0; p < MAXH; p++) {
            if (p < half) {
                if ((l >> (half - 1 - p)) & 1) m1[p] = fmin(m1[p], d2);
                else m0[p] = fmin(m0[p], d2);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < MAXH; p++) llr[p] = (m1[p] - m0[p]) * invN0;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(LS_THREADS)
    nr_demap_kernel(const TIn* __restrict__ sym, long long numSym, int qm, double noiseVar, TOut* __restrict__ llr)
{
    __shared__ double levels[32];
    const int half = qm >> 1;
    if (qm > 1 && (int)threadIdx.x < (1 << half)) levels[threadIdx.x] = qam_scale(qm) * (double)pam_level(threadIdx.x, half);
    __syncthreads();
    const double invN0 = 1.0 / noiseVar;
    for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < numSym; s += (long long)gridDim.x * blockDim.x) {
        const double yr = (double)sym[2 * s], yi = (double)sym[2 * s + 1];
        if (qm == 1) {
            // BPSK: points +-(1 + j)/sqrt(2): |y - x1|^2 - |y - x0|^2 = 4 a (yr + yi)
            const double a = qam_scale(1);
            const double d0 = (yr - a) * (yr - a) + (yi - a) * (yi - a), d1 = (yr + a) * (yr + a) + (yi + a) * (yi + a);
            llr[s] = (TOut)((d1 - d0) * invN0);
            continue;
        }
        double lr[5], li[5];
        axis_llr<double, 5>(yr, half, levels, invN0, lr);
        axis_llr<double, 5>(yi, half, levels, invN0, li);
        for (int p = 0; p < half; p++) {
            llr[s * qm + 2 * p] = (TOut)lr[p];
            llr[s * qm + 2 * p + 1] = (TOut)li[p];
        }
    }
}

// ---- counter-based RNG: Philox4x32-10 (Salmon et al., SC'11), written out; one call = four 32-bit words -----------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
// two independent N(0,1) values from two 32-bit words (Box-Muller; u1 in (0,1] keeps the tail to 6.7 sigma).  The noise
// only has to be Gaussian, not bit-reproducible against another implementation, so the logarithm and the sine / cosine
// are the SFU approximations (absolute error ~2^-21 on these ranges; checked by test_awgn_generator_statistics_*).
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b)
{
    const float u1 = (float)a * 2.3283064365386963e-10f + 1.1641532182693481e-10f;   // (a + 0.5) / 2^32
    const float ang = (float)b * 1.4629180792671596e-9f;                              // 2 pi b / 2^32
    const float r = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(ang, &sn, &cs);
    return make_float2(r * cs, r * sn);
}

// Fused transmit-channel-receive for workload generation: bits -> Gray QAM -> + CN(0, noiseVar) -> max-log LLR (fp32).
// Symbol n of the call has the absolute index a = offset + n and draws its noise from Philox counter a >> 1 under `seed`
// (words 0,1 for even a, words 2,3 for odd a: one Philox call serves two symbols): the result does not depend on the
// launch geometry, and a sweep sharded over GPUs / batches stays reproducible by passing the global symbol offset.
// A thread owns SPT (even) symbols that start at an even absolute index, so that its LLRs fill whole float4 stores.
template <int QM, int SPT>
__global__ void __launch_bounds__(LS_THREADS)
    nr_awgn_llr_kernel(const signed char* __restrict__ bits, long long numSym, float noiseVar, unsigned long long seed,
                       unsigned long long offset, float* __restrict__ llr)
{
    static_assert(SPT % 2 == 0, "pairs of symbols share one Philox call");
    constexpr int HALF = QM / 2;
    constexpr int NLL = QM * SPT;       // LLRs per thread
    __shared__ float levels[32];
    if (QM > 1 && (int)threadIdx.x < (1 << HALF)) levels[threadIdx.x] = (float)qam_scale(QM) * (float)pam_level(threadIdx.x, HALF);
    __syncthreads();
    const float scale = (float)qam_scale(QM);
    const float sigma = sqrtf(0.5f * noiseVar);
    const float invN0 = 1.0f / noiseVar;
    const long long o1 = (long long)(offset & 1ull);      // symbol n sits at shifted index n + o1 (same parity as a)
    const unsigned long long pair0 = offset >> 1;         // Philox counter of shifted indices 0, 1
    const long long numGroups = (numSym + o1 + SPT - 1) / SPT;
    const bool vec = ((reinterpret_cast<uintptr_t>(llr) & 15) == 0) && (NLL % 4 == 0) && ((o1 * QM) % 4 == 0);
    const bool bitsWord = (QM % 4 == 0) && ((reinterpret_cast<uintptr_t>(bits) & 3) == 0);
    for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < numGroups;
         gidx += (long long)gridDim.x * blockDim.x) {
        float o[NLL];
#pragma unroll
        for (int k2 = 0; k2 < SPT; k2 += 2) {
            const unsigned long long c = pair0 + (unsigned long long)((gidx * SPT + k2) >> 1);
            const uint4 rnd = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u),
                                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int k = k2 + h;
                const long long s = gidx * SPT + k - o1;
                if (s >= 0 && s < numSym) {
                    uint32_t v = 0;
                    if (bitsWord) {   // labels of QM = 4 | 8 bits as whole words: byte i of a word is bit i (MSB first)
#pragma unroll
                        for (int wd = 0; wd < QM / 4; wd++) {
                            const uint32_t w = reinterpret_cast<const uint32_t*>(bits)[s * (QM / 4) + wd];
                            v = (v << 4) | (((w & 0x01010101u) * 0x08040201u) >> 24);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < QM; i++) v = (v << 1) | (uint32_t)(bits[s * QM + i] & 1);
                    }
                    int re, im;
                    qam_point(v, QM, re, im);
                    const float2 nz = h ? box_muller(rnd.z, rnd.w) : box_muller(rnd.x, rnd.y);
                    const float yr = scale * (float)re + sigma * nz.x, yi = scale * (float)im + sigma * nz.y;
                    if (QM == 1) {
                        const float a = scale;
                        const float d0 = (yr - a) * (yr - a) + (yi - a) * (yi - a), d1 = (yr + a) * (yr + a) + (yi + a) * (yi + a);
                        o[k] = (d1 - d0) * invN0;
                    } else {
                        float lr[HALF > 0 ? HALF : 1], li[HALF > 0 ? HALF : 1];
                        axis_llr<float, (HALF > 0 ? HALF : 1)>(yr, HALF, levels, invN0, lr);
                        axis_llr<float, (HALF > 0 ? HALF : 1)>(yi, HALF, levels, invN0, li);
#pragma unroll
                        for (int p = 0; p < HALF; p++) {
                            o[k * QM + 2 * p] = lr[p];
                            o[k * QM + 2 * p + 1] = li[p];
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < QM; i++) o[k * QM + i] = 0.f;
                }
            }
        }
        const long long base = (gidx * SPT - o1) * QM;   // may be negative for the first group when the offset is odd
        if (vec && gidx * SPT - o1 >= 0 && (gidx + 1) * SPT - o1 <= numSym) {
#pragma unroll
            for (int q4 = 0; q4 < NLL / 4; q4++)
                reinterpret_cast<float4*>(llr + base)[q4] = make_float4(o[4 * q4], o[4 * q4 + 1], o[4 * q4 + 2], o[4 * q4 + 3]);
        } else {
            for (int i = 0; i < NLL; i++)
                if (base + i >= 0 && base + i < numSym * QM) llr[base + i] = o[i];
        }
    }
}

bool qm_ok(int qm) { return qm == 1 || qm == 2 || qm == 4 || qm == 6 || qm == 8 || qm == 10; }

}   // namespace

extern "C" int nrldpc_modulate(nrldpc_handle* h, int qm, const int8_t* bits, int64_t num_sym, int out_dtype, void* symbols,
                               nrldpc_stream stream)
{
    if (!h) { nr_set_error("modulate: null handle"); return NRLDPC_ERR_ARG; }
    if (!qm_ok(qm) || num_sym <= 0) { nr_set_error("modulate: bad arguments (qm=%d)", qm); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const int grid = (int)min((long long)((num_sym + LS_THREADS - 1) / LS_THREADS), (long long)h->numSMs * 16);
    if (out_dtype == NRLDPC_F64)
        nr_modulate_kernel<double><<<grid, LS_THREADS, 0, (cudaStream_t)stream>>>((const signed char*)bits, num_sym, qm, (double*)symbols);
    else if (out_dtype == NRLDPC_F32)
        nr_modulate_kernel<float><<<grid, LS_THREADS, 0, (cudaStream_t)stream>>>((const signed char*)bits, num_sym, qm, (float*)symbols);
    else { nr_set_error("modulate: bad dtype"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_demap_maxlog(nrldpc_handle* h, int qm, int in_dtype, const void* symbols, int64_t num_sym,
                                   double noise_var, int out_dtype, void* llr, nrldpc_stream stream)
{
    if (!h) { nr_set_error("demap: null handle"); return NRLDPC_ERR_ARG; }
    if (!qm_ok(qm) || num_sym <= 0 || !(noise_var > 0.0)) { nr_set_error("demap: bad arguments (qm=%d)", qm); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const int grid = (int)min((long long)((num_sym + LS_THREADS - 1) / LS_THREADS), (long long)h->numSMs * 16);
    cudaStream_t s = (cudaStream_t)stream;
    if (in_dtype == NRLDPC_F64 && out_dtype == NRLDPC_F64)
        nr_demap_kernel<double, double><<<grid, LS_THREADS, 0, s>>>((const double*)symbols, num_sym, qm, noise_var, (double*)llr);
    else if (in_dtype == NRLDPC_F64 && out_dtype == NRLDPC_F32)
        nr_demap_kernel<double, float><<<grid, LS_THREADS, 0, s>>>((const double*)symbols, num_sym, qm, noise_var, (float*)llr);
    else if (in_dtype == NRLDPC_F32 && out_dtype == NRLDPC_F64)
        nr_demap_kernel<float, double><<<grid, LS_THREADS, 0, s>>>((const float*)symbols, num_sym, qm, noise_var, (double*)llr);
    else if (in_dtype == NRLDPC_F32 && out_dtype == NRLDPC_F32)
        nr_demap_kernel<float, float><<<grid, LS_THREADS, 0, s>>>((const float*)symbols, num_sym, qm, noise_var, (float*)llr);
    else { nr_set_error("demap: bad dtype"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

extern "C" int nrldpc_awgn_llr(nrldpc_handle* h, int qm, const int8_t* bits, int64_t num_sym, double noise_var,
                               uint64_t seed, uint64_t offset, float* llr, nrldpc_stream stream)
{
    if (!h) { nr_set_error("awgn_llr: null handle"); return NRLDPC_ERR_ARG; }
    if (!qm_ok(qm) || num_sym <= 0 || !(noise_var > 0.0)) { nr_set_error("awgn_llr: bad arguments (qm=%d)", qm); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const signed char* b = (const signed char*)bits;
    const float nv = (float)noise_var;
#define NR_LAUNCH_AWGN(QM, SPT)                                                                               \
    {                                                                                                          \
        const long long groups = (num_sym + 1 + (SPT)-1) / (SPT);                                              \
        const int grid = (int)min((groups + LS_THREADS - 1) / LS_THREADS, (long long)h->numSMs * 16);          \
        nr_awgn_llr_kernel<QM, SPT><<<grid, LS_THREADS, 0, s>>>(b, num_sym, nv, seed, offset, llr);            \
    }
    switch (qm) {
        case 1: NR_LAUNCH_AWGN(1, 4) break;
        case 2: NR_LAUNCH_AWGN(2, 2) break;
        case 4: NR_LAUNCH_AWGN(4, 2) break;
        case 6: NR_LAUNCH_AWGN(6, 2) break;
        case 8: NR_LAUNCH_AWGN(8, 2) break;
        default: NR_LAUNCH_AWGN(10, 2) break;
    }
#undef NR_LAUNCH_AWGN
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

// =================================================================================================================
// Scrambling (TS 38.211 5.2.1 / 7.3.1.1): Gold sequence c(n) = x1(n + 1600) ^ x2(n + 1600) and its application to
// bits (XOR) and LLRs (sign flip).  Mirrors goldSequence (neoradium/utils.py:70-94) and PDSCH.scrambleBits /
// scrambleLLRs (neoradium/pdsch.py:603-616).
//
// The two 31-bit LFSRs advance 31 positions per "word step" (x1 <- A1 x1, x2 <- A2 x2 over GF(2), the same word
// recurrences the reference iterates); word step 51 holds positions 1581..1611, whose top 12 bits are c(0..11), and every
// later word holds the next 31 bits of c, LSB first.  The reference walks the words one after the other; here a thread
// jumps straight to its first word with precomputed powers A^(2^i) (host, constant bank) and then steps through a
// short run of words, so the sequence is produced in parallel.
// =================================================================================================================
namespace {

constexpr int GOLD_POW = 26;          // word steps up to 2^26 (2 Gbit of sequence)

struct GoldTables {
    uint32_t a1[GOLD_POW][31];   // column j of A1^(2^i): image of basis vector e_j
    uint32_t a2[GOLD_POW][31];
};
GoldTables g_goldHost;
bool g_goldReady = false;

inline uint32_t gold_step1(uint32_t x)
{
    x ^= (x >> 3);
    x ^= (x << 28) & 0x7FFFFFFFu;
    return x;
}
inline uint32_t gold_step2(uint32_t x)
{
    x ^= (x >> 3) ^ (x >> 2) ^ (x >> 1);
    x ^= ((x << 28) ^ (x << 29) ^ (x << 30)) & 0x7FFFFFFFu;
    return x;
}
inline uint32_t gold_apply_host(const uint32_t* cols, uint32_t x)
{
    uint32_t r = 0;
    for (int j = 0; j < 31; j++)
        if ((x >> j) & 1u) r ^= cols[j];
    return r;
}
void gold_build_tables()
{
    if (g_goldReady) return;
    for (int j = 0; j < 31; j++) {
        g_goldHost.a1[0][j] = gold_step1(1u << j);
        g_goldHost.a2[0][j] = gold_step2(1u << j);
    }
    for (int i = 1; i < GOLD_POW; i++)
        for (int j = 0; j < 31; j++) {   // A^(2^i) e_j = A^(2^(i-1)) (A^(2^(i-1)) e_j)
            g_goldHost.a1[i][j] = gold_apply_host(g_goldHost.a1[i - 1], g_goldHost.a1[i - 1][j]);
            g_goldHost.a2[i][j] = gold_apply_host(g_goldHost.a2[i - 1], g_goldHost.a2[i - 1][j]);
        }
    g_goldReady = true;
}

__device__ __forceinline__ uint32_t gold_apply(const uint32_t* __restrict__ cols, uint32_t x)
{
    uint32_t r = 0;
#pragma unroll
    for (int j = 0; j < 31; j++) r ^= ((x >> j) & 1u) ? cols[j] : 0u;
    return r;
}

// mode 0: out8[n] = c(n); mode 1: out8[n] = bits[n] ^ c(n); mode 2/3: out[n] = llr[n] * (1 - 2 c(n)) (fp32 / fp64)
//
// With q = n + 19, sequence bit n is bit q % 31 of LFSR word q / 31 (word 0 = word step 51, whose bits 19..30 are
// c(0..11)).  A CTA owns a tile of GOLD_TILE consecutive elements: its 256 threads first produce the (<= 2048) words the
// tile needs into shared memory -- thread t jumps to its first word with the precomputed powers A^(2^i) and steps
// through 8 words -- and then all threads sweep the tile with 16-byte coalesced accesses, taking their bits from a
// 62-bit window of two neighbouring words.  HBM-bound: 2 B (bits) / 8 B (fp32) / 16 B (fp64) per element.
constexpr int GOLD_WPT = 8;                      // words generated per thread per tile
constexpr int GOLD_TILE = 256 * 246;             // elements per tile: multiple of 16, needs <= 256 * GOLD_WPT - 1 words

__device__ __forceinline__ uint32_t spread4(uint32_t x)   // bits 0..3 -> bytes 0..3
{
    return ((x & 0xFu) * 0x00204081u) & 0x01010101u;
}

template <int MODE>
__global__ void __launch_bounds__(256)
    nr_gold_kernel(const GoldTables* __restrict__ tab, uint32_t x1w, uint32_t x2w, long long numBits, const void* __restrict__ in,
                   void* __restrict__ out)
{
    __shared__ uint32_t sA1[GOLD_POW * 31], sA2[GOLD_POW * 31];
    __shared__ uint32_t sW[256 * GOLD_WPT + 1];
    for (int i = threadIdx.x; i < GOLD_POW * 31; i += blockDim.x) {
        sA1[i] = (&tab->a1[0][0])[i];
        sA2[i] = (&tab->a2[0][0])[i];
    }
    if (threadIdx.x == 0) sW[256 * GOLD_WPT] = 0;
    __syncthreads();
    constexpr int V = (MODE == 2) ? 4 : (MODE == 3 ? 2 : 16);   // elements per 16-byte access
    const long long numTiles = (numBits + GOLD_TILE - 1) / GOLD_TILE;
    for (long long tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        const long long n0 = tile * GOLD_TILE;
        const long long wFirst = (n0 + 19) / 31;
        {   // this thread's words wFirst + tid * WPT .. + WPT - 1
            const long long w0 = wFirst + (long long)threadIdx.x * GOLD_WPT;
            uint32_t x1 = x1w, x2 = x2w;
            for (int i = 0; i < GOLD_POW; i++)
                if ((w0 >> i) & 1LL) {
                    x1 = gold_apply(sA1 + i * 31, x1);
                    x2 = gold_apply(sA2 + i * 31, x2);
                }
#pragma unroll
            for (int k = 0; k < GOLD_WPT; k++) {
                sW[threadIdx.x * GOLD_WPT + k] = x1 ^ x2;
                x1 ^= (x1 >> 3);
                x1 ^= (x1 << 28) & 0x7FFFFFFFu;
                x2 ^= (x2 >> 3) ^ (x2 >> 2) ^ (x2 >> 1);
                x2 ^= ((x2 << 28) ^ (x2 << 29) ^ (x2 << 30)) & 0x7FFFFFFFu;
            }
        }
        __syncthreads();
        const uint32_t qBase = (uint32_t)(n0 + 19 - wFirst * 31);   // local position of element n0 (< 31)
        const long long left = numBits - n0;
        const int cnt = (int)(left < GOLD_TILE ? left : GOLD_TILE);
        // V sequence bits starting at local element i (window of two 31-bit words; bit <= 30 leaves >= 32 bits)
        auto seq_bits = [&](int i) -> uint32_t {
            const uint32_t q = qBase + (uint32_t)i;
            const uint32_t w = q / 31u;
            const uint32_t b = q - w * 31u;
            const unsigned long long win = (unsigned long long)sW[w] | ((unsigned long long)sW[w + 1] << 31);
            return (uint32_t)(win >> b);
        };
        const bool aligned = ((reinterpret_cast<uintptr_t>(out) | (MODE ? reinterpret_cast<uintptr_t>(in) : 0)) & 15) == 0;
        if (aligned) {
            for (int i = threadIdx.x * V; i + V <= cnt; i += 256 * V) {
                const uint32_t c = seq_bits(i);
                const long long n = n0 + i;
                if (MODE == 0 || MODE == 1) {
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (MODE == 1) v = *reinterpret_cast<const uint4*>(reinterpret_cast<const signed char*>(in) + n);
                    v.x ^= spread4(c);
                    v.y ^= spread4(c >> 4);
                    v.z ^= spread4(c >> 8);
                    v.w ^= spread4(c >> 12);
                    *reinterpret_cast<uint4*>(reinterpret_cast<signed char*>(out) + n) = v;
                } else if (MODE == 2) {
                    uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(in) + n);
                    v.x ^= (c & 1u) << 31;
                    v.y ^= (c & 2u) << 30;
                    v.z ^= (c & 4u) << 29;
                    v.w ^= (c & 8u) << 28;
                    *reinterpret_cast<uint4*>(reinterpret_cast<float*>(out) + n) = v;
                } else {
                    uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const double*>(in) + n);
                    v.y ^= (c & 1u) << 31;
                    v.w ^= (c & 2u) << 30;
                    *reinterpret_cast<uint4*>(reinterpret_cast<double*>(out) + n) = v;
                }
            }
        }
        // scalar remainder (tail of the last tile, or everything when a pointer is not 16-byte aligned)
        for (int i = (aligned ? (cnt / V) * V : 0) + threadIdx.x; i < cnt; i += 256) {
            const uint32_t cb = seq_bits(i) & 1u;
            const long long n = n0 + i;
            if (MODE == 0) reinterpret_cast<signed char*>(out)[n] = (signed char)cb;
            if (MODE == 1) reinterpret_cast<signed char*>(out)[n] = (signed char)(reinterpret_cast<const signed char*>(in)[n] ^ (signed char)cb);
            if (MODE == 2) reinterpret_cast<uint32_t*>(out)[n] = reinterpret_cast<const uint32_t*>(in)[n] ^ (cb << 31);
            if (MODE == 3) reinterpret_cast<unsigned long long*>(out)[n] = reinterpret_cast<const unsigned long long*>(in)[n] ^ ((unsigned long long)cb << 63);
        }
        __syncthreads();   // sW is rewritten by the next tile
    }
}

int gold_launch(nrldpc_handle* h, int mode, uint32_t c_init, int64_t num_bits, const void* in, void* out, nrldpc_stream stream)
{
    if (!h) { nr_set_error("scramble: null handle"); return NRLDPC_ERR_ARG; }
    if (num_bits <= 0 || c_init >= 0x80000000u) { nr_set_error("scramble: bad arguments"); return NRLDPC_ERR_ARG; }
    const long long numWords = 1 + (num_bits > 12 ? (num_bits - 12 + 30) / 31 : 0);
    if (numWords >= (1LL << GOLD_POW)) { nr_set_error("scramble: sequence too long"); return NRLDPC_ERR_ARG; }
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    gold_build_tables();
    if (!h->goldTables) {
        NR_CUDA_CHECK(cudaMalloc(&h->goldTables, sizeof(GoldTables)));
        NR_CUDA_CHECK(cudaMemcpy(h->goldTables, &g_goldHost, sizeof(GoldTables), cudaMemcpyHostToDevice));
    }
    // LFSR words at word step 51 (utils.py:73-78: x1 is the pre-computed constant, x2 comes from cInit)
    uint32_t x1 = 0x42054D21u, x2 = c_init;
    for (int i = 0; i < 51; i++) x2 = gold_step2(x2);
    const long long tiles = (num_bits + GOLD_TILE - 1) / GOLD_TILE;
    const int grid = (int)max(1LL, min(tiles, (long long)h->numSMs * 4));
    const GoldTables* t = (const GoldTables*)h->goldTables;
    cudaStream_t s = (cudaStream_t)stream;
    switch (mode) {
        case 0: nr_gold_kernel<0><<<grid, 256, 0, s>>>(t, x1, x2, num_bits, in, out); break;
        case 1: nr_gold_kernel<1><<<grid, 256, 0, s>>>(t, x1, x2, num_bits, in, out); break;
        case 2: nr_gold_kernel<2><<<grid, 256, 0, s>>>(t, x1, x2, num_bits, in, out); break;
        default: nr_gold_kernel<3><<<grid, 256, 0, s>>>(t, x1, x2, num_bits, in, out); break;
    }
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}

}   // namespace

extern "C" int nrldpc_gold_sequence(nrldpc_handle* h, uint32_t c_init, int64_t num_bits, int8_t* out, nrldpc_stream stream)
{
    return gold_launch(h, 0, c_init, num_bits, nullptr, out, stream);
}

extern "C" int nrldpc_scramble_bits(nrldpc_handle* h, uint32_t c_init, const int8_t* bits, int64_t num_bits, int8_t* out,
                                    nrldpc_stream stream)
{
    return gold_launch(h, 1, c_init, num_bits, bits, out, stream);
}

extern "C" int nrldpc_scramble_llrs(nrldpc_handle* h, uint32_t c_init, int dtype, const void* llrs, int64_t num, void* out,
                                    nrldpc_stream stream)
{
    if (dtype != NRLDPC_F32 && dtype != NRLDPC_F64) { nr_set_error("scramble_llrs: bad dtype"); return NRLDPC_ERR_ARG; }
    return gold_launch(h, dtype == NRLDPC_F32 ? 2 : 3, c_init, num, llrs, out, stream);
}

// ---- payload bits (include/nrldpc.h, nrldpc_random_bits) ----------------------------------------------------------------
// one thread = one Philox call = 128 consecutive bits of the global stream, written as int8 0/1
__global__ void __launch_bounds__(256) nr_random_bits_kernel(unsigned long long seed, unsigned long long offset, signed char* out, long long n)
{
    const unsigned long long blk0 = offset >> 7;
    const long long numBlk = (long long)(((offset + (unsigned long long)n + 127ull) >> 7) - blk0);
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < numBlk; k += (long long)gridDim.x * blockDim.x) {
        const unsigned long long c = blk0 + (unsigned long long)k;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0x62697473u, 0u),
                                      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
        const long long j0 = (long long)(c << 7) - (long long)offset;   // index in `out` of the block's first bit
        if (j0 >= 0 && j0 + 128 <= n && ((reinterpret_cast<uintptr_t>(out + j0) & 15) == 0)) {
#pragma unroll
            for (int q = 0; q < 8; q++) {   // 16 bits -> 16 bytes
                const uint32_t h16 = (w[q >> 1] >> ((q & 1) * 16)) & 0xffffu;
                uint4 v;
                v.x = ((h16 >> 0) & 1u) | (((h16 >> 1) & 1u) << 8) | (((h16 >> 2) & 1u) << 16) | (((h16 >> 3) & 1u) << 24);
                v.y = ((h16 >> 4) & 1u) | (((h16 >> 5) & 1u) << 8) | (((h16 >> 6) & 1u) << 16) | (((h16 >> 7) & 1u) << 24);
                v.z = ((h16 >> 8) & 1u) | (((h16 >> 9) & 1u) << 8) | (((h16 >> 10) & 1u) << 16) | (((h16 >> 11) & 1u) << 24);
                v.w = ((h16 >> 12) & 1u) | (((h16 >> 13) & 1u) << 8) | (((h16 >> 14) & 1u) << 16) | (((h16 >> 15) & 1u) << 24);
                *reinterpret_cast<uint4*>(out + j0 + q * 16) = v;
            }
        } else {
            for (int b = 0; b < 128; b++) {
                const long long j = j0 + b;
                if (j >= 0 && j < n) out[j] = (signed char)((w[b >> 5] >> (b & 31)) & 1u);
            }
        }
    }
}

extern "C" int nrldpc_random_bits(nrldpc_handle* h, uint64_t seed, uint64_t offset, int8_t* out, int64_t n, nrldpc_stream stream)
{
    if (!h || !out || n < 0) { nr_set_error("random_bits: bad argument"); return NRLDPC_ERR_ARG; }
    if (n == 0) return NRLDPC_OK;
    NR_CUDA_CHECK(cudaSetDevice(h->device));
    const long long numBlk = (long long)(((offset + (uint64_t)n + 127ull) >> 7) - (offset >> 7));
    const int grid = (int)max(1LL, min((numBlk + 255) / 256, (long long)h->numSMs * 16));
    nr_random_bits_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, offset, (signed char*)out, n);
    NR_CUDA_CHECK(cudaGetLastError());
    return NRLDPC_OK;
}
