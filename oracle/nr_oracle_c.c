/* CPU ORACLE in C (test infrastructure, not product code): scalar restatement of NeoRadium's layered normalised
 * min-sum decoder, neoradium/ldpc.py:1535-1581, following the "a10 formula" of SURVEY.md section 8a line by line.
 *
 * It exists because the NumPy oracle (oracle/nr_oracle.py, pinned bit-for-bit against the unmodified reference) is
 * too slow for parity checks over thousands of code blocks.  tests/test_oracle_c.py pins THIS file against the NumPy
 * oracle (bit-identical beliefs in float64 and float32), so the chain of trust is
 *     MATLAB golden vectors / reference outputs  ->  nr_oracle.py  ->  nr_oracle_c.c.
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no fused multiply-add, every operation rounds once, as NumPy does)
 *
 * The graph is passed in flat form: row_deg[P], then per edge (row-major, ascending column) col[e] and shift[e]
 * (already reduced mod Z).  Beliefs are written for all n*Z positions.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DECODE_IMPL(NAME, T, ABS)                                                                                     \
    int NAME(const T *rx, long num_cb, int ncols_in, int P, int n, int Z, const int *row_deg, const int *col,       \
             const int *shift, int num_iter, T *belief)                                                              \
    {                                                                                                                 \
        int total_edges = 0;                                                                                          \
        for (int i = 0; i < P; i++) total_edges += row_deg[i];                                                        \
        T *msg = (T *)malloc(sizeof(T) * (size_t)total_edges * Z);                                                    \
        if (!msg) return -1;                                                                                          \
        for (long c = 0; c < num_cb; c++) {                                                                           \
            T *r = belief + (size_t)c * n * Z;                                                                        \
            const T *x = rx + (size_t)c * ncols_in * Z;                                                               \
            for (int i = 0; i < n * Z; i++) r[i] = (T)0;                                                              \
            for (int i = 0; i < ncols_in * Z; i++) {            /* clip, ldpc.py:1536; prepend 2Z zeros :1538 */       \
                T v = x[i];                                                                                           \
                if (v > (T)1e10) v = (T)1e10;                                                                         \
                if (v < (T)-1e10) v = (T)-1e10;                                                                       \
                r[2 * Z + i] = v;                                                                                     \
            }                                                                                                         \
            memset(msg, 0, sizeof(T) * (size_t)total_edges * Z);                                                      \
            for (int it = 0; it < num_iter; it++) {                                                                   \
                int e0 = 0;                                                                                           \
                for (int i = 0; i < P; i++) {                                                                         \
                    int d = row_deg[i];                                                                               \
                    for (int m = 0; m < Z; m++) {                                                                     \
                        T t[32];                                                                                      \
                        int pos[32];                                                                                  \
                        int jstar = 0, neg = 0;                                                                       \
                        T min1 = (T)0;                                                                                \
                        for (int j = 0; j < d; j++) {                                                                 \
                            int p = m + shift[e0 + j];                                                                \
                            if (p >= Z) p -= Z;                                                                       \
                            pos[j] = col[e0 + j] * Z + p;                                                             \
                            t[j] = r[pos[j]] - msg[(size_t)(e0 + j) * Z + m];                                         \
                            T a = ABS(t[j]);                                                                          \
                            if (t[j] < (T)0) neg ^= 1;                                                                \
                            if (j == 0 || a < min1) { min1 = a; jstar = j; }   /* first minimum, :1559 */             \
                        }                                                                                             \
                        T min2 = ABS(t[jstar] + (T)100000);                    /* the +100000 quirk, :1563 */         \
                        for (int j = 0; j < d; j++) {                                                                 \
                            if (j == jstar) continue;                                                                 \
                            T a = ABS(t[j]);                                                                          \
                            if (a < min2) min2 = a;                                                                   \
                        }                                                                                             \
                        for (int j = 0; j < d; j++) {                                                                 \
                            T mag = (j == jstar) ? min2 : min1;                                                       \
                            int s = (t[j] < (T)0) ^ neg;                                                              \
                            T nw = (s ? -mag : mag) * (T)0.75;                 /* :1567-1573 */                       \
                            msg[(size_t)(e0 + j) * Z + m] = nw;                                                       \
                            r[pos[j]] = t[j] + nw;                             /* :1576 */                            \
                        }                                                                                             \
                    }                                                                                                 \
                    e0 += d;                                                                                          \
                }                                                                                                     \
            }                                                                                                         \
        }                                                                                                             \
        free(msg);                                                                                                    \
        return 0;                                                                                                     \
    }

DECODE_IMPL(nr_oracle_decode_f64, double, fabs)
DECODE_IMPL(nr_oracle_decode_f32, float, fabsf)

/* CRC long division, chancodebase.py:120-128: MSB first, zero init, one value per bit.  poly without leading 1. */
uint32_t nr_oracle_crc(const int8_t *bits, long n, uint32_t poly, int c)
{
    uint32_t reg = 0, top = 1u << (c - 1), mask = (c == 32) ? 0xffffffffu : ((1u << c) - 1);
    for (long i = 0; i < n; i++) {
        uint32_t fb = ((reg & top) != 0) ^ (uint32_t)(bits[i] & 1);
        reg = (reg << 1) & mask;
        if (fb) reg ^= poly;
    }
    return reg;
}
