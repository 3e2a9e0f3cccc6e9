/* libnrldpc -- C-ABI of the B200-native (sm_100a) 5G NR LDPC hot path.
 *
 * Drop-in boundary for the channel-coding path of InterDigitalInc/NeoRadium v0.4.0 (neoradium/ldpc.py +
 * neoradium/chancodebase.py).  The reference is pure Python/NumPy and has no FFI of its own; the boundary is the
 * method surface of LdpcEncoder / LdpcDecoder / ChanCodeBase.  Each entry point below names the reference method
 * (file:line, relative to the reference root) it replaces.  neoradium_b200/ldpc.py binds these through ctypes and
 * mirrors the reference classes; INTEGRATION.md shows the stub a NeoRadium maintainer would add.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer in the current CUDA context of `device` (e.g. torch tensor.data_ptr());
 *     buffers are caller-owned, nothing is retained after the call is enqueued
 *   - bits are one int8 per bit (0/1), exactly the reference's array layout; LLRs are positive => bit 0
 *   - `stream` is a cudaStream_t (NULL = default stream); calls are asynchronous unless stated otherwise
 *   - return value: NRLDPC_OK or an error code; nrldpc_last_error() gives the message (thread local).  No exception
 *     crosses the ABI.  There is NO CPU fallback: without a usable GPU every compute call fails with NRLDPC_ERR_CUDA.
 *   - one handle per (device, stream); a handle is not thread-safe, distinct handles are independent
 */
#ifndef NRLDPC_H
#define NRLDPC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRLDPC_VERSION 100

typedef struct nrldpc_handle nrldpc_handle;
typedef void* nrldpc_stream; /* cudaStream_t */

enum {
    NRLDPC_OK = 0,
    NRLDPC_ERR_ARG = 1,    /* the reference raises ValueError / AssertionError for these */
    NRLDPC_ERR_CUDA = 2,   /* CUDA runtime error, no device, launch failure */
    NRLDPC_ERR_NOMEM = 3
};

/* generator polynomials of chancodebase.py:37-44 */
enum { NRLDPC_CRC6 = 0, NRLDPC_CRC11 = 1, NRLDPC_CRC16 = 2, NRLDPC_CRC24A = 3, NRLDPC_CRC24B = 4, NRLDPC_CRC24C = 5 };

/* element types of LLR / belief buffers.  NRLDPC_F16 (IEEE half) is an INPUT type of nrldpc_decode_tb only: the values
 * are widened exactly to the compute type on load, so the result equals that of the same values passed as float. */
enum { NRLDPC_F32 = 0, NRLDPC_F64 = 1, NRLDPC_F16 = 2 };

/* decoder flags */
enum {
    NRLDPC_DEC_EARLY_STOP = 1,  /* extension: stop a code block once all parity checks hold after a full iteration */
    NRLDPC_DEC_ALL_ROWS = 2,    /* disable the (exact) skipping of extension rows whose parity LLRs are all zero */
    NRLDPC_DEC_ES_AUTO = 4      /* with NRLDPC_DEC_EARLY_STOP: the first tested iteration follows the previous launch on this handle
                                 * (one less than the smallest iteration count any of its blocks needed; never below ES_FROM):
                                 * a syndrome test costs ~15 % of an iteration and is wasted before the first block converges */
};
/* with NRLDPC_DEC_EARLY_STOP: bits 8..15 of `flags` hold the first iteration (1-based) after which the syndrome is tested;
 * 0 or 1 = after every iteration.  A block then runs at least that many iterations (results per returned iteration count
 * are unchanged: they equal a fixed-iteration decode with that count). */
#define NRLDPC_DEC_ES_FROM(k) ((((k) < 0 ? 0 : (k) > 255 ? 255 : (k)) & 0xff) << 8)

int nrldpc_version(void);
const char* nrldpc_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Host-only table queries (usable without a GPU)
 * ---------------------------------------------------------------------------------------------------------------- */

/* LdpcBase.baseGraph, ldpc.py:775-789.  out[P*n] int16 row-major, -1 = no edge, else V % zc (the table keeps the
 * reference's verbatim 880 entry, ldpc.py:143).  set_index < 0 => derived from zc. */
int nrldpc_base_graph(int bg, int set_index, int zc, int16_t* out);
/* liftingSizeSets lookup, ldpc.py:657-666; returns iLS or -1 */
int nrldpc_lifting_set_index(int zc);
/* (P, n, k, number of base-graph edges) of ldpc.py:780 */
int nrldpc_graph_info(int bg, int* rows, int* cols, int* sys_cols, int* edges);

/* ------------------------------------------------------------------------------------------------------------------
 * Handle
 * ---------------------------------------------------------------------------------------------------------------- */
int nrldpc_create(int device, nrldpc_handle** out);
int nrldpc_destroy(nrldpc_handle* h);

/* ------------------------------------------------------------------------------------------------------------------
 * Unified-memory buffers -- HARQ state that stays on the device behind host-visible arrays.
 * HarqCW.encBuffer / HarqCW.decBuffer (harq.py:120-121, 145-178) and the [C, N] array that recoverRate hands to decode
 * (ldpc.py:1414-1418, harq.py:169-170) are host NumPy arrays in the reference and are passed back into the codec on the
 * next call.  Allocated here as CUDA managed memory they are valid DEVICE pointers for every entry point of this header
 * (the kernels update them at HBM speed, nothing crosses PCIe between calls) and at the same time ordinary host memory
 * for whoever reads them (pages migrate on a CPU access: a lazy, hardware-coherent host mirror).
 * ---------------------------------------------------------------------------------------------------------------- */
/* 1 if the device supports concurrent managed access (needed for the scheme above), 0 otherwise */
int nrldpc_managed_supported(nrldpc_handle* h);
/* *out = managed allocation of `bytes` bytes, populated on the device; zero != 0 clears it (asynchronously on `stream`) */
int nrldpc_managed_alloc(nrldpc_handle* h, uint64_t bytes, int zero, void** out, nrldpc_stream stream);
int nrldpc_managed_free(nrldpc_handle* h, void* p);
/* clear a managed block on the device (asynchronously on `stream`) */
int nrldpc_managed_clear(nrldpc_handle* h, void* p, uint64_t bytes, nrldpc_stream stream);
/* migrate the pages to the device (to_device != 0) or to the host ahead of use; purely a performance hint */
int nrldpc_managed_prefetch(nrldpc_handle* h, void* p, uint64_t bytes, int to_device, nrldpc_stream stream);

/* ------------------------------------------------------------------------------------------------------------------
 * CRC -- ChanCodeBase.getCrc / checkCrc / appendCrc, chancodebase.py:83-128, 132-157, 161-189
 * bits: [num_streams, len] int8 with row pitch `stride` (elements).  MSB-first long division, zero initial state.
 * ---------------------------------------------------------------------------------------------------------------- */
/* crc_bits [num_streams, c] int8 (may be NULL); rem [num_streams] uint32 remainder, MSB of the CRC in bit c-1 (may
 * be NULL) */
int nrldpc_crc(nrldpc_handle* h, const int8_t* bits, int64_t num_streams, int64_t len, int64_t stride, int poly,
               int8_t* crc_bits, uint32_t* rem, nrldpc_stream stream);
/* out [num_streams, len + c]: the input followed by its CRC */
int nrldpc_crc_attach(nrldpc_handle* h, const int8_t* bits, int64_t num_streams, int64_t len, int64_t stride,
                      int poly, int8_t* out, nrldpc_stream stream);
/* ok [num_streams] uint8: 1 where the remainder of the whole stream (data || crc) is zero */
int nrldpc_crc_check(nrldpc_handle* h, const int8_t* bits, int64_t num_streams, int64_t len, int64_t stride,
                     int poly, uint8_t* ok, nrldpc_stream stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Transport-block configuration shared by the TX and RX chains.  All fields are what LdpcBase.initialize
 * (ldpc.py:859-892), getRateMatchedCbLens (:846-856) and rateMatch/recoverRate (:1135-1145, :1365-1395) derive;
 * the host side (neoradium_b200/params.py) computes them with the reference's integer arithmetic.
 * ---------------------------------------------------------------------------------------------------------------- */
typedef struct nrldpc_tb_config {
    int32_t bg;   /* base graph 1 | 2 */
    int32_t zc;   /* lifting size Zc */
    int32_t K;    /* code block size 22 Zc | 10 Zc */
    int32_t F;    /* filler bits per code block */
    int32_t C;    /* code blocks per transport block */
    int32_t qm;   /* bits per modulation symbol */
    int32_t nl;   /* transmission layers */
    int32_t ncb;  /* circular buffer length INCLUDING fillers: N or min(N, nRef) */
    int32_t rv;   /* redundancy version 0..3 */
    int32_t reserved;
    int64_t G;    /* rate-matched bits per transport block the E_r split is derived from (sum E_r >= G) */
} nrldpc_tb_config;

/* ------------------------------------------------------------------------------------------------------------------
 * TX chain
 * ---------------------------------------------------------------------------------------------------------------- */
/* LdpcEncoder.doSegmentation, ldpc.py:1011-1030.  tb [num_tb, B] (pitch tb_stride) already carries its CRC24A.
 * out [num_tb*C, K]: zero pad at the end of the TB, CRC24B per code block when C > 1, F zero filler bits. */
int nrldpc_segment(nrldpc_handle* h, const nrldpc_tb_config* cfg, const int8_t* tb, int64_t num_tb, int64_t B,
                   int64_t tb_stride, int8_t* code_blocks, nrldpc_stream stream);

/* LdpcEncoder.encode, ldpc.py:1057-1090.  code_blocks [num_cb, K] -> coded [num_cb, N] (puncture != 0, the first
 * 2 Zc bits dropped) or [num_cb, N + 2 Zc]. */
int nrldpc_encode(nrldpc_handle* h, int bg, int zc, const int8_t* code_blocks, int64_t num_cb, int8_t* coded,
                  int puncture, nrldpc_stream stream);

/* LdpcEncoder.rateMatch, ldpc.py:1128-1159.  coded [num_tb*C, N] -> out [num_tb, sumE] (pitch out_stride):
 * bit selection from the filler-less circular buffer starting at k0(rv), then the bit interleaver. */
int nrldpc_rate_match(nrldpc_handle* h, const nrldpc_tb_config* cfg, const int8_t* coded, int64_t num_tb,
                      int8_t* out, int64_t out_stride, nrldpc_stream stream);

/* LdpcBase.isValidCodedBlock done right (all rows; the reference's ldpc.py:841-843 returns after the first row).
 * coded_full [num_cb, n*Zc] un-punctured bits; ok [num_cb] uint8. */
int nrldpc_parity_check(nrldpc_handle* h, int bg, int zc, const int8_t* coded_full, int64_t num_cb, uint8_t* ok,
                        nrldpc_stream stream);
/* the same over the first `rows` base-graph rows only (0 = all).  rows = 1 reproduces what the reference's
 * isValidCodedBlock actually tests (its loop returns inside the first row, ldpc.py:841-843): compatibility switch. */
int nrldpc_parity_check_rows(nrldpc_handle* h, int bg, int zc, const int8_t* coded_full, int64_t num_cb, int rows,
                             uint8_t* ok, nrldpc_stream stream);

/* ------------------------------------------------------------------------------------------------------------------
 * RX chain
 * ---------------------------------------------------------------------------------------------------------------- */
/* LdpcDecoder.recoverRate, ldpc.py:1365-1418.  llr [num_tb, llr_len] (pitch llr_stride; llr_len <= sumE, the tail is
 * zero-extended as ldpc.py:1402-1403 does) of element type `dtype`; soft_buffer NULL or [num_tb*C, ncb-F] in/out
 * (HARQ decBuffer: combined into, ldpc.py:1407-1412); out [num_tb*C, ncb] with F entries of 1e20 inserted after the
 * systematic part (:1415-1418).  All buffers share `dtype`; accumulation order = ascending stream position. */
int nrldpc_rate_recover(nrldpc_handle* h, const nrldpc_tb_config* cfg, int dtype, const void* llr, int64_t num_tb,
                        int64_t llr_len, int64_t llr_stride, void* soft_buffer, void* out, nrldpc_stream stream);

/* LdpcDecoder.decode, ldpc.py:1535-1581: layered normalised min-sum (alpha = 0.75), `num_iter` full iterations.
 *   llr      [num_cb, in_cols*Zc] of in_dtype (pitch llr_stride); in_cols is normally n-2 (66 | 50)
 *   compute_dtype  NRLDPC_F64 reproduces the reference's float64 arithmetic bit for bit; NRLDPC_F32 is the same
 *                  operation order evaluated in float32 (no FMA contraction)
 *   out_cols 22|10 (onlyInfoBits) ... n; bits [num_cb, out_cols*Zc] int8 (or NULL); beliefs same shape in
 *            compute_dtype (or NULL)
 *   iters    NULL or [num_cb] int32: iterations actually run (== num_iter unless NRLDPC_DEC_EARLY_STOP) */
int nrldpc_decode(nrldpc_handle* h, int bg, int zc, int in_dtype, int compute_dtype, const void* llr,
                  int64_t num_cb, int64_t llr_stride, int in_cols, int num_iter, int flags, int out_cols,
                  int8_t* bits, void* beliefs, int32_t* iters, nrldpc_stream stream);

/* LdpcDecoder.decode2, ldpc.py:1421-1492 (undocumented verification decoder): the same layered schedule -- the reference
 * walks the P*Zc lifted rows one by one, and the Zc rows of a base-graph row touch disjoint positions -- with the TRUE
 * second minimum (no "+100000" term), a caller-chosen `alpha`, and an optional stop after the first iteration whose hard
 * decisions satisfy every parity check.  (The reference's own stop test calls isValidCodedBlock, which looks at the first
 * base-graph row only, ldpc.py:841-843; here all rows are checked with stop_on_good_parity = 1, and
 * stop_on_good_parity = 2 reproduces the reference's first-row-only test: compatibility switch.)  Other arguments as
 * nrldpc_decode; all P rows are always scheduled. */
int nrldpc_decode2(nrldpc_handle* h, int bg, int zc, int in_dtype, int compute_dtype, const void* llr, int64_t num_cb,
                   int64_t llr_stride, int in_cols, int max_iter, double alpha, int stop_on_good_parity, int out_cols,
                   int8_t* bits, void* beliefs, int32_t* iters, nrldpc_stream stream);
/* the same with an OFFSET on top of the normalisation (extension; SURVEY 8f row 4 -- the reference has only the normalised
 * form): |message| = max(alpha * min - beta, 0), beta >= 0; beta = 0 is nrldpc_decode2. */
int nrldpc_decode2_offset(nrldpc_handle* h, int bg, int zc, int in_dtype, int compute_dtype, const void* llr, int64_t num_cb,
                          int64_t llr_stride, int in_cols, int max_iter, double alpha, double beta, int stop_on_good_parity,
                          int out_cols, int8_t* bits, void* beliefs, int32_t* iters, nrldpc_stream stream);

/* Fused RX chain: recoverRate -> decode -> checkCrcAndMerge -> checkCrc('24A'), i.e. HarqCW.decodeLLRs
 * (harq.py:165-173) / the documented usage ldpc.py:1234-1251, in ONE kernel per code-block group: rate recovery is
 * the decoder's load phase, the CRC runs on the decoder's hard decisions in shared memory.
 *   llr        [num_tb, llr_len] in_dtype, pitch llr_stride
 *   soft_buffer NULL or [num_tb*C, ncb-F] compute_dtype in/out
 *   tb_bits    [num_tb, C*per_cb] int8: merged transport block incl. its CRC24A (per_cb = K-F-24 if C>1 else K-F)
 *   cb_crc_ok  [num_tb*C] uint8 (CRC24B per code block; CRC24A when C == 1) ; tb_crc_ok [num_tb] uint8 (CRC24A
 *              over the merged block); iters [num_tb*C] int32.  Any output may be NULL. */
int nrldpc_decode_tb(nrldpc_handle* h, const nrldpc_tb_config* cfg, int in_dtype, int compute_dtype,
                     const void* llr, int64_t num_tb, int64_t llr_len, int64_t llr_stride, void* soft_buffer,
                     int num_iter, int flags, int8_t* tb_bits, int64_t tb_bits_stride, uint8_t* cb_crc_ok,
                     uint8_t* tb_crc_ok, int32_t* iters, nrldpc_stream stream);

/* The same chain fed with the EQUALISED SYMBOLS of the codewords instead of their LLRs: what PDSCH.getLLRsFromGrid hands to
 * Modem.getLLRsFromSymbols (pdsch.py:935-1000, modulation.py:159-204) in front of HarqCW.decodeLLRs (harq.py:165-173).
 *   symbols    [num_tb, num_sym] complex64 as (re, im) float pairs, pitch sym_stride symbols; num_sym * qm = the LLRs of a
 *              transport block (G'); noise_var > 0 is the demapper's noise variance
 * Max-log demapping happens inside the decoder's load phase (no LLR buffer in HBM, no demapper launch) for the lifting sizes
 * that run one code block per CTA without repetition; every other configuration demaps into a scratch buffer of the handle
 * first.  Either way the LLRs are, value for value, those of nrldpc_demap_maxlog(F32 -> F32) and the result equals
 * nrldpc_decode_tb(F32, F32) on them.  fp32 compute, no soft buffer (HARQ combining takes the LLR entry point). */
int nrldpc_decode_tb_symbols(nrldpc_handle* h, const nrldpc_tb_config* cfg, const float* symbols, int64_t num_tb,
                             int64_t num_sym, int64_t sym_stride, double noise_var, int num_iter, int flags,
                             int8_t* tb_bits, int64_t tb_bits_stride, uint8_t* cb_crc_ok, uint8_t* tb_crc_ok,
                             int32_t* iters, nrldpc_stream stream);

/* Mixed-configuration batch in ONE call (BASELINE configs[2]: a PDSCH slot whose codewords differ in base graph, lifting
 * size, modulation, layers, redundancy version -- the per-codeword loop of HarqProcess.decodeLLRs, harq.py:331-347, over
 * HarqCW.decodeLLRs, harq.py:165-173).  `groups` is a HOST array of descriptors, one per set of equally configured transport
 * blocks; every field has the meaning of the same-named nrldpc_decode_tb argument.  The groups are decoded CONCURRENTLY:
 * each one is a fused kernel on one of the handle's internal streams (forked from and joined back into `stream`), so a slot
 * costs the time of its largest group, not the sum, and small groups share the SMs.  Results equal num_groups separate
 * nrldpc_decode_tb calls. */
typedef struct nrldpc_tb_group {
    nrldpc_tb_config cfg;
    int32_t in_dtype;        /* NRLDPC_F32 | NRLDPC_F64 | NRLDPC_F16 */
    int32_t reserved;
    const void* llr;
    int64_t num_tb, llr_len, llr_stride;
    void* soft_buffer;       /* NULL or [num_tb*C, ncb-F] compute_dtype in/out: HarqCW.decBuffer on the device */
    int8_t* tb_bits;
    int64_t tb_bits_stride;
    uint8_t* cb_crc_ok;
    uint8_t* tb_crc_ok;
    int32_t* iters;
} nrldpc_tb_group;
int nrldpc_decode_tb_groups(nrldpc_handle* h, const nrldpc_tb_group* groups, int num_groups, int compute_dtype,
                            int num_iter, int flags, nrldpc_stream stream);

/* LdpcDecoder.checkCrcAndMerge, ldpc.py:1610-1619, for already decoded blocks.  decoded [num_tb*C, K] ->
 * tb_bits [num_tb, C*per_cb] and cb_crc_ok [num_tb*C]. */
int nrldpc_check_crc_and_merge(nrldpc_handle* h, const nrldpc_tb_config* cfg, const int8_t* decoded, int64_t num_tb,
                               int8_t* tb_bits, int64_t tb_bits_stride, uint8_t* cb_crc_ok, nrldpc_stream stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Link-level counters (the quantities harq.py:173 / the BLER notebooks accumulate on the host).
 * counters int64[8] += {code blocks, CB CRC failures, transport blocks, TB CRC failures, bit errors vs ref_bits,
 * sum of iterations, 0, 0}; reduce across GPUs with one NCCL all-reduce (neoradium_b200/dist.py).
 * ---------------------------------------------------------------------------------------------------------------- */
int nrldpc_accumulate_counters(nrldpc_handle* h, int64_t num_tb, int C, const uint8_t* cb_crc_ok,
                               const uint8_t* tb_crc_ok, const int32_t* iters, const int8_t* tb_bits,
                               const int8_t* ref_bits, int64_t bits_per_tb, int64_t bits_stride, int64_t* counters,
                               nrldpc_stream stream);
/* the same with separate row pitches for the decoded blocks and the reference payload (no padded copy of the payload) */
int nrldpc_accumulate_counters_ref(nrldpc_handle* h, int64_t num_tb, int C, const uint8_t* cb_crc_ok,
                                   const uint8_t* tb_crc_ok, const int32_t* iters, const int8_t* tb_bits, int64_t tb_stride,
                                   const int8_t* ref_bits, int64_t ref_stride, int64_t bits_per_tb, int64_t* counters,
                                   nrldpc_stream stream);

/* Payload generator of the on-device link simulator (random.bits, random.py:194, on the device): bit j of the call is bit
 * (offset + j) of the Philox4x32-10 stream under `seed` (128 bits per counter value), so a sweep draws the same payloads
 * whatever the batch size or the number of GPUs.  out [n] int8 (0/1). */
int nrldpc_random_bits(nrldpc_handle* h, uint64_t seed, uint64_t offset, int8_t* out, int64_t n, nrldpc_stream stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Link around the codec, on the device: the producers of the decoder's input in every NeoRadium
 * BLER loop.  qm in {1, 2, 4, 6, 8, 10}; symbols are interleaved (re, im) pairs of the given element type.
 * ---------------------------------------------------------------------------------------------------------------- */

/* Modem.modulate, modulation.py:127-157 (constellation of modulation.py:60-74): bits[num_sym*qm] -> symbols.
 * Bit-exact: the constellation values are scale * small integer. */
int nrldpc_modulate(nrldpc_handle* h, int qm, const int8_t* bits, int64_t num_sym, int out_dtype, void* symbols,
                    nrldpc_stream stream);

/* Modem.getLLRsFromSymbols(symbols, noiseVar, useMax=True), modulation.py:159-204: symbols -> llr[num_sym*qm],
 * positive => bit 0.  Per-axis search (equal to the reference's 2-D search in exact arithmetic; tolerance in tests). */
int nrldpc_demap_maxlog(nrldpc_handle* h, int qm, int in_dtype, const void* symbols, int64_t num_sym, double noise_var,
                        int out_dtype, void* llr, nrldpc_stream stream);

/* Fused modulate -> + CN(0, noise_var) -> max-log LLR in fp32 (what the BLER notebooks do with Modem.modulate,
 * random.awgn and getLLRsFromSymbols, PDSCH-BLER.ipynb raw lines 117-165).  Noise of symbol n comes from Philox4x32-10
 * counter (offset + n) under `seed`: independent of launch geometry and of how a sweep is sharded. */
int nrldpc_awgn_llr(nrldpc_handle* h, int qm, const int8_t* bits, int64_t num_sym, double noise_var, uint64_t seed,
                    uint64_t offset, float* llr, nrldpc_stream stream);

/* goldSequence(cInit, numBits), utils.py:70-94 (TS 38.211 5.2.1): out[n] = c(n), one int8 per bit.  Generated in
 * parallel by LFSR jump-ahead; c_init < 2^31. */
int nrldpc_gold_sequence(nrldpc_handle* h, uint32_t c_init, int64_t num_bits, int8_t* out, nrldpc_stream stream);

/* PDSCH.scrambleBits, pdsch.py:603-608: out = bits ^ c (in place allowed). */
int nrldpc_scramble_bits(nrldpc_handle* h, uint32_t c_init, const int8_t* bits, int64_t num_bits, int8_t* out,
                         nrldpc_stream stream);

/* PDSCH.scrambleLLRs, pdsch.py:611-616: out = llrs * (1 - 2 c), dtype NRLDPC_F32 | NRLDPC_F64 (in place allowed). */
int nrldpc_scramble_llrs(nrldpc_handle* h, uint32_t c_init, int dtype, const void* llrs, int64_t num, void* out,
                         nrldpc_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* NRLDPC_H */
