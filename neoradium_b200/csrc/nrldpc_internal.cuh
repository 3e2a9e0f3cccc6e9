// Internal shared declarations of libnrldpc (B200 / sm_100a).  Not part of the C-ABI (see include/nrldpc.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/nrldpc.h"

#define NR_MAX_ROWS 46
#define NR_MAX_EDGES 316
#define NR_MAX_Z 384

// Lifted base graph handed to kernels BY VALUE (__grid_constant__): it lands in the constant bank, so the
// warp-uniform, dynamically indexed reads of (column, shift) cost a ULDC and no shared-memory or L1 traffic.
struct __align__(16) NrGraph {
    int P;        // base-graph rows (46 | 42)
    int ncols;    // base-graph columns (68 | 52)
    int ksys;     // systematic columns (22 | 10)
    int ncore;    // ksys + 4: columns of degree > 1, the only ones shared between layers
    int Z;        // lifting size
    int pad[3];
    uint16_t rowEdge0[NR_MAX_ROWS + 2];   // prefix sums of row degrees
    uint32_t edge[NR_MAX_EDGES];          // lo16 = shift mod Z, hi16 = column index
};

struct nrldpc_handle {
    int device;
    int numSMs;
    int maxSmemOptin;
    void* scratch;          // decoder overflow state (rows that do not fit shared memory), grown on demand
    size_t scratchBytes;
    unsigned int* workCounter;  // device words: [0] dynamic scheduling counter, [1] last-non-zero-column scan, [4] / [5] NRLDPC_DEC_ES_AUTO hint / running minimum,
    void* symLlr;           // nrldpc_decode_tb_symbols without a fused form: LLR scratch, grown on demand
    size_t symLlrBytes;
    void* tmp;              // small per-call temporaries (per-code-block CRC partials), grown on demand
    size_t tmpBytes;
    void* tmp2;             // second temporary (multi-segment CRC accumulators; may be live together with `tmp`)
    size_t tmp2Bytes;
    void* crcFacDev;        // cached per-thread CRC factors of the fused decoder epilogue (decode.cu), key below
    unsigned long long crcFacKey;
    void* tbAcc;            // per-transport-block CRC24A accumulators of the fused decoder (decode.cu), zero between launches
    size_t tbAccBytes;
    void* tbFacDev;         // cached x^(per (C-1-r)) mod g24A of the fused decoder, key below
    unsigned long long tbFacKey;
    size_t tbFacCap;
    // nrldpc_decode_tb_groups: private sub-handles (own scratch / temporaries / work queue) on internal streams
    nrldpc_handle* sub[4];
    cudaStream_t subStream[4];
    cudaEvent_t subFork, subJoin[4];
    void* goldTables;       // device copy of the Gold-sequence jump tables (linksim.cu), created on first use
    int smemPerSM;
    int decOcc;             // target resident decoder CTAs per SM (0 = automatic), env NRLDPC_DEC_OCC
    int noStaticRows;       // env NRLDPC_NO_STATIC_ROWS=1: use the dynamic-row decoder kernel everywhere (A/B measurements)
    int noTmem;             // env NRLDPC_NO_TMEM=1: keep the decoder's row state out of Tensor Memory (A/B measurements)
    int noStage;            // env NRLDPC_NO_STAGE=1: no TMA staging of the rate-matched stream (A/B measurements)
};
int nr_reserve_tmp(nrldpc_handle* h, size_t bytes, void** out);
int nr_reserve_tmp2(nrldpc_handle* h, size_t bytes, void** out);
struct NrGraph;
// E_r split (getRateMatchedCbLens, ldpc.py:846-856) and k0 (ldpc.py:1145) of a transport-block configuration
int nr_tb_split(const nrldpc_tb_config* c, int N, int* E0, int* nShort, int* fStep, int* k0);
int nr_check_tb_config(const nrldpc_tb_config* cfg, const NrGraph& g, const char* who);

void nr_set_error(const char* fmt, ...);
int nr_build_graph(int bg, int zc, NrGraph* g);   // host: fills g from the TS 38.212 tables; 0 on success

#define NR_CUDA_CHECK(expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            nr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return NRLDPC_ERR_CUDA;                                                               \
        }                                                                                         \
    } while (0)

// ---- CRC helpers shared by crc.cu and the decoder epilogue -----------------------------------------------------
struct NrCrcPoly {
    uint32_t poly;   // generator without the leading 1
    int len;         // 6 | 11 | 16 | 24
};
__host__ __device__ inline NrCrcPoly nr_crc_poly(int id)
{
    // chancodebase.py:37-44 (TS 38.212 section 5.1)
    switch (id) {
        case NRLDPC_CRC6: return {0x21u, 6};
        case NRLDPC_CRC11: return {0x621u, 11};
        case NRLDPC_CRC16: return {0x1021u, 16};
        case NRLDPC_CRC24A: return {0x864CFBu, 24};
        case NRLDPC_CRC24B: return {0x800063u, 24};
        default: return {0xB2B117u, 24};   // 24C
    }
}

// Resident CTAs per SM of `kern` at this launch shape.  The grid-stride helper kernels are launched as ONE full wave: a grid
// of "SMs x 8" on a kernel that fits 5 CTAs per SM runs 1.6 waves, the second one 60 % full (ncu launch__waves_per_multiprocessor).
template <typename K>
static inline int nr_ctas_per_sm(K kern, int threads, size_t smem)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem) != cudaSuccess || n < 1) {
        (void)cudaGetLastError();
        n = 1;
    }
    return n;
}
