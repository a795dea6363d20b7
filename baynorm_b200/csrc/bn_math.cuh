// bn_math.cuh -- FP64 logarithms for the likelihood kernels.
//
// The fit is bound by the FP64 pipe (64 lanes per SM on B200, one warp instruction every two
// cycles per SM sub-partition), so what matters is the number of FP64-pipe instructions per
// likelihood term.  libdevice's log() costs ~32 of them plus ~15 integer ones; the table-driven
// versions below cost 11 (log) and 15 (log1p) with one 16-byte shared-memory look-up:
//     x = 2^k z,  z in [0.6875, 1.375);   i = top 7 mantissa bits of (bits(x) - OFF)
//     r = fma(z, invc_i, -1),  |r| <= 2^-8;   log x = k ln2 + logc_i + (r - r^2/2 + ... - r^8/8)
// Absolute error <= ~1e-18 + 1 ulp of the result (tests/test_gpu_parity.py::test_device_log pins
// it against mpmath): the table entries are exact to 2^-63 (tools/gen_log_table.py), the only
// roundings that matter are the last two additions.
//
// Domain: finite, positive, normal doubles (every caller passes sizes, means and 1 + m/phi).
#pragma once
#include <stdint.h>

struct BnLogEntry {
    double invc, logc;
};

#define BN_LOG_TABLE_N 128
__device__ const BnLogEntry g_bn_log_table[BN_LOG_TABLE_N] = {
#include "bn_log_table.inc"
};

// copy the table into shared memory (call from all threads of the CTA, then __syncthreads())
__device__ __forceinline__ void bn_log_table_load(BnLogEntry *sh)
{
    for (int k = threadIdx.x; k < BN_LOG_TABLE_N; k += blockDim.x) sh[k] = g_bn_log_table[k];
}

#define BN_LOG_OFF_HI 0x3FE60000

// log1p(r) - r = r^2 * P(r) for |r| <= 2^-7.9: -1/2 + r/3 - r^2/4 + r^3/5 - r^4/6 + r^5/7 - r^6/8
__device__ __forceinline__ double bn_log_poly(double r)
{
    double p = -0.125;
    p = fma(p, r, 0.14285714285714285);
    p = fma(p, r, -0.16666666666666666);
    p = fma(p, r, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 0.33333333333333331);
    p = fma(p, r, -0.5);
    return p;
}

// reduction shared by log and log1p: returns r, sets kd (exponent as double) and the table entry
__device__ __forceinline__ void bn_log_reduce(double x, const BnLogEntry *__restrict__ tab, double &z, double &kd,
                                              int &k, BnLogEntry &e)
{
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int tmp = hi - BN_LOG_OFF_HI;          // top 32 bits of bits(x) - OFF (OFF's low word is 0)
    const int i = (tmp >> 13) & (BN_LOG_TABLE_N - 1); // bits 45..51 of the 64-bit difference
    k = tmp >> 20;                               // arithmetic shift: floor exponent
    z = __hiloint2double(hi - k * 1048576, lo);
    e = tab[i];
    // int -> double without the conversion unit: 2^52 + 2^31 + k, minus the magic constant
    kd = __hiloint2double(0x43300000, k ^ 0x80000000) - 4503601774854144.0;
}

__device__ __forceinline__ double bn_log(double x, const BnLogEntry *__restrict__ tab)
{
    double z, kd;
    int k;
    BnLogEntry e;
    bn_log_reduce(x, tab, z, kd, k, e);
    const double r = fma(z, e.invc, -1.0);
    const double w = fma(kd, 0.69314718055994529, e.logc);
    const double r2 = r * r;
    return w + fma(r2, bn_log_poly(r), r);
}

// log(1 + y) for y >= 0, relative accuracy kept for small y through the exact low part of 1 + y
__device__ __forceinline__ double bn_log1p_pos(double y, const BnLogEntry *__restrict__ tab)
{
    const bool big = y > 1.0;
    const double a = big ? y : 1.0, b = big ? 1.0 : y;
    const double u = a + b;
    const double c = b - (u - a); // Fast2Sum: u + c == 1 + y exactly
    double z, kd;
    int k;
    BnLogEntry e;
    bn_log_reduce(u, tab, z, kd, k, e);
    const double cs = c * __hiloint2double((1023 - k) * 1048576, 0); // c * 2^-k, exact
    double r = fma(z, e.invc, -1.0);
    r = fma(cs, e.invc, r);
    const double w = fma(kd, 0.69314718055994529, e.logc);
    const double r2 = r * r;
    return w + fma(r2, bn_log_poly(r), r);
}

// log(1 + y) for 0 <= y <= 2^-6 without the table: alternating series to y^9 (next term < 2^-57 y)
__device__ __forceinline__ double bn_log1p_small(double y)
{
    double p = 0.1111111111111111;
    p = fma(p, y, -0.125);
    p = fma(p, y, 0.14285714285714285);
    p = fma(p, y, -0.16666666666666666);
    p = fma(p, y, 0.2);
    p = fma(p, y, -0.25);
    p = fma(p, y, 0.33333333333333331);
    p = fma(p, y, -0.5);
    return fma(y * y, p, y);
}

// log(k!) for the integer counts of the likelihood constant sum_j lgamma(x_j + 1): a 2 KB table (L1-resident)
// up to 256, libdevice beyond
__device__ const double g_bn_lfact[257] = {
#include "bn_lfact_table.inc"
};
__device__ __forceinline__ double bn_lfact(double x)
{
    return x <= 256.0 ? __ldg(g_bn_lfact + (int)x) : lgamma(x + 1.0);
}

// lgamma(z) for z >= 64 by Stirling's series through 1/(1260 z^5) (next term 1/(1680 z^7) < 4e-16)
__device__ __forceinline__ double bn_lgamma_stirling(double z, const BnLogEntry *__restrict__ tab)
{
    const double lz = bn_log(z, tab);
    const double iz = 1.0 / z, iz2 = iz * iz;
    const double corr = iz * fma(iz2, fma(iz2, 7.9365079365079365e-4, -2.7777777777777778e-3), 8.3333333333333329e-2);
    return fma(z - 0.5, lz, -z) + (0.91893853320467278 + corr);
}
