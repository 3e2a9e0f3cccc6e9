// QAM helpers shared by the link kernels (linksim.cu) and the decoder's symbol-input load phase (decode_kernel.cuh):
// Modem.getLLRsFromSymbols(useMax=True), neoradium/modulation.py:159-204, per axis (see linksim.cu for the derivation).
#pragma once
#include <math.h>
#include <stdint.h>

__host__ __device__ inline double qam_scale(int qm)
{
    switch (qm) {
        case 1: case 2: return 1.0 / sqrt(2.0);
        case 4: return 1.0 / sqrt(10.0);
        case 6: return 1.0 / sqrt(42.0);
        case 8: return 1.0 / sqrt(170.0);
        default: return 1.0 / sqrt(682.0);
    }
}

// amplitude of one axis from its `half` label bits (MSB = the sign bit b0 | b1, then outer .. inner)
__device__ __forceinline__ int pam_level(uint32_t lab, int half)
{
    int a = 1;
    for (int p = half - 1; p >= 1; p--) a = (1 << (half - p)) - (1 - 2 * (int)((lab >> (half - 1 - p)) & 1u)) * a;
    return (1 - 2 * (int)((lab >> (half - 1)) & 1u)) * a;
}

// max-log LLR of label bit p of ONE axis (received coordinate y): the same operations, value for value, as axis_llr of
// linksim.cu performs for that bit (its bits are independent of each other)
__device__ __forceinline__ double axis_llr_bit(double y, int half, const double* __restrict__ levels, double invN0, int p)
{
    double m0 = (double)INFINITY, m1 = (double)INFINITY;
    const int nl = 1 << half;
    for (int l = 0; l < nl; l++) {
        const double d = y - levels[l];
        const double d2 = d * d;
        if ((l >> (half - 1 - p)) & 1) m1 = fmin(m1, d2);
        else m0 = fmin(m0, d2);
    }
    return (m1 - m0) * invN0;
}
// LLR number b (0 .. qm-1) of the symbol (yr, yi): what nr_demap_kernel<float, float> writes to llr[s * qm + b]
__device__ __forceinline__ float demap_bit_f32(float yr, float yi, int b, int qm, const double* __restrict__ levels, double invN0)
{
    if (qm == 1) {
        const double a = qam_scale(1), r = (double)yr, i = (double)yi;
        const double d0 = (r - a) * (r - a) + (i - a) * (i - a), d1 = (r + a) * (r + a) + (i + a) * (i + a);
        return (float)((d1 - d0) * invN0);
    }
    return (float)axis_llr_bit((b & 1) ? (double)yi : (double)yr, qm >> 1, levels, invN0, b >> 1);
}
