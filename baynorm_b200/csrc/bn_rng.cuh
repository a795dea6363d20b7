// bn_rng.cuh -- counter-based Philox4x32-10 and the discrete samplers built on it.
//
// The reference draws from R's global Mersenne-Twister in loop order
// (src/bayNorm_main.cpp:1121, :907); those streams cannot be reproduced on a GPU, so the
// contract (BASELINE.json north_star) is distributional parity.  What this file guarantees
// instead: the value of draw (gene, cell, s) is a pure function of (seed, global gene index,
// global cell index, s), independent of tiling, sharding and GPU count.
#pragma once
#include <stdint.h>

struct Philox {
    uint32_t k0, k1;         // key  = seed
    uint32_t c0, c1, c2, c3; // counter: (element lo, element hi, sample/sub-stream, refill index)
    uint32_t out[4];
    int have;

    __device__ __forceinline__ void init(uint64_t seed, uint64_t element, uint32_t substream)
    {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        c0 = (uint32_t)element;
        c1 = (uint32_t)(element >> 32);
        c2 = substream;
        c3 = 0;
        have = 0;
    }
    __device__ __forceinline__ void refill()
    {
        uint32_t a0 = c0, a1 = c1, a2 = c2, a3 = c3, x0 = k0, x1 = k1;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, a0), lo0 = 0xD2511F53u * a0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, a2), lo1 = 0xCD9E8D57u * a2;
            const uint32_t n0 = hi1 ^ a1 ^ x0, n2 = hi0 ^ a3 ^ x1;
            a0 = n0;
            a1 = lo1;
            a2 = n2;
            a3 = lo0;
            x0 += 0x9E3779B9u;
            x1 += 0xBB67AE85u;
        }
        out[0] = a0;
        out[1] = a1;
        out[2] = a2;
        out[3] = a3;
        c3++;
        have = 4;
    }
    __device__ __forceinline__ uint32_t next()
    {
        if (have == 0) refill();
        have--;
        // select without dynamic register indexing
        return have == 3 ? out[0] : (have == 2 ? out[1] : (have == 1 ? out[2] : out[3]));
    }
    // uniform on (0,1), 32 random bits, never 0 or 1
    __device__ __forceinline__ double u32() { return ((double)next() + 0.5) * 2.3283064365386963e-10; }
    // uniform on (0,1) with 53 random bits
    __device__ __forceinline__ double u53()
    {
        const uint32_t a = next(), b = next();
        const uint64_t v = (((uint64_t)a << 32) | b) >> 11;
        return ((double)v + 0.5) * 1.1102230246251565e-16;
    }
};

// Philox4x32-10 with the key schedule precomputed on the host and passed as a kernel parameter:
// the 20 round keys then sit in the constant bank and cost no instructions (LOP3 takes them as a
// constant operand), 4 instructions per round instead of 6.
struct PhiloxKeys {
    uint32_t k[20];
};
static inline PhiloxKeys bn_philox_keys(uint64_t seed)
{
    PhiloxKeys K;
    uint32_t x0 = (uint32_t)seed, x1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) {
        K.k[2 * r] = x0;
        K.k[2 * r + 1] = x1;
        x0 += 0x9E3779B9u;
        x1 += 0xBB67AE85u;
    }
    return K;
}
#ifdef __CUDACC__
// The same round keys computed on the device from the seed (kernels that take the seed as a plain argument): the
// adds are warp-uniform, so they run on the uniform datapath once per kernel.
struct PhiloxKeysDev {
    uint32_t k[20];
    __device__ __forceinline__ explicit PhiloxKeysDev(uint64_t seed)
    {
        uint32_t x0 = (uint32_t)seed, x1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; r++) {
            k[2 * r] = x0;
            k[2 * r + 1] = x1;
            x0 += 0x9E3779B9u;
            x1 += 0xBB67AE85u;
        }
    }
    __device__ __forceinline__ void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t &o0, uint32_t &o1,
                                          uint32_t &o2, uint32_t &o3) const
    {
        uint32_t a0 = c0, a1 = c1, a2 = c2, a3 = c3;
#pragma unroll
        for (int r = 0; r < 10; r++) {
            const uint64_t m0 = (uint64_t)0xD2511F53u * a0, m1 = (uint64_t)0xCD9E8D57u * a2;
            const uint32_t n0 = (uint32_t)(m1 >> 32) ^ a1 ^ k[2 * r], n2 = (uint32_t)(m0 >> 32) ^ a3 ^ k[2 * r + 1];
            a1 = (uint32_t)m1;
            a3 = (uint32_t)m0;
            a0 = n0;
            a2 = n2;
        }
        o0 = a0;
        o1 = a1;
        o2 = a2;
        o3 = a3;
    }
};
__device__ __forceinline__ void bn_philox_block(const PhiloxKeys &K, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                uint32_t &o0, uint32_t &o1, uint32_t &o2, uint32_t &o3)
{
    uint32_t a0 = c0, a1 = c1, a2 = c2, a3 = c3;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t m0 = (uint64_t)0xD2511F53u * a0, m1 = (uint64_t)0xCD9E8D57u * a2;
        const uint32_t n0 = (uint32_t)(m1 >> 32) ^ a1 ^ K.k[2 * r], n2 = (uint32_t)(m0 >> 32) ^ a3 ^ K.k[2 * r + 1];
        a1 = (uint32_t)m1;
        a3 = (uint32_t)m0;
        a0 = n0;
        a2 = n2;
    }
    o0 = a0;
    o1 = a1;
    o2 = a2;
    o3 = a3;
}
#endif

// ---- continuous helpers ---------------------------------------------------------------
__device__ __forceinline__ double bn_rnorm(Philox &g)
{ // Box-Muller, one value per call (the sine branch is dropped)
    const double u = g.u53(), v = g.u32();
    return sqrt(-2.0 * log(u)) * cospi(2.0 * v);
}

// Gamma(shape a, scale 1), Marsaglia & Tsang (2000); a < 1 boosted by U^(1/a).
static __device__ double bn_rgamma(Philox &g, double a)
{
    double boost = 1.0;
    if (a < 1.0) {
        boost = pow(g.u53(), 1.0 / a);
        a += 1.0;
    }
    const double d = a - 1.0 / 3.0, c = rsqrt(9.0 * d);
    for (int it = 0; it < 64; it++) {
        double x, v;
        do {
            x = bn_rnorm(g);
            v = 1.0 + c * x;
        } while (v <= 0.0);
        v = v * v * v;
        const double u = g.u53();
        const double x2 = x * x;
        if (u < 1.0 - 0.0331 * x2 * x2) return d * v * boost;
        if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return d * v * boost;
    }
    return d * boost; // unreachable in practice (acceptance > 95 % per round)
}

// Poisson(lam): sequential inversion below 10, Hoermann's PTRS (1993) above.
static __device__ double bn_rpois(Philox &g, double lam)
{
    if (!(lam > 0.0)) return 0.0;
    if (lam < 10.0) {
        double p = exp(-lam), u = g.u53(), k = 0.0;
        while (u > p && k < 200.0) {
            u -= p;
            k += 1.0;
            p *= lam / k;
        }
        return k;
    }
    const double slam = sqrt(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0), ll = log(lam);
    for (int it = 0; it < 256; it++) {
        const double U = g.u53() - 0.5, V = g.u53(), us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * ll - lgamma(k + 1.0)) return k;
    }
    return floor(lam);
}

// ---- negative binomial, size r > 0, mean M >= 0 -------------------------------------------
// Small means (the vast majority of gene x cell entries: x = 0 and mu*beta << 1) are drawn
// by CDF inversion from ONE uniform, starting from P(0) = (r/(r+M))^r; large means fall back
// to the gamma-Poisson mixture the reference uses (rnbinom_mu == rpois(rgamma(size, mu/size))).
struct NBParams {
    double r, M, q, p0; // q = M/(r+M); p0 = P(0)
    bool inversion;
};
#define BN_NB_INVERSION_MAX_MEAN 12.0

__device__ __forceinline__ NBParams bn_nb_setup(double r, double M)
{
    NBParams P;
    P.r = r;
    P.M = M;
    P.q = M / (r + M);
    P.inversion = (M < BN_NB_INVERSION_MAX_MEAN);
    P.p0 = P.inversion ? exp(-r * log1p(M / r)) : 0.0;
    return P;
}

__device__ __forceinline__ double bn_nb_draw(const NBParams &P, Philox &g)
{
    if (!(P.M > 0.0)) return 0.0; // rgamma(shape, scale = 0) = 0 -> rpois(0) = 0
    if (P.inversion) {
        double u = g.u32(), p = P.p0, k = 0.0;
        // p(k+1) = p(k) * q * (r + k) / (k + 1)
        while (u > p && k < 4096.0) {
            u -= p;
            p *= P.q * (P.r + k) / (k + 1.0);
            k += 1.0;
        }
        return k;
    }
    const double lam = bn_rgamma(g, P.r) * (P.M / P.r);
    return bn_rpois(g, lam);
}

// ---- binomial(n, p) for DownSampling ----------------------------------------------------------
// Inversion from 0 for n*min(p,1-p) < 30 (counts are small and beta ~ 0.06), Hoermann's BTRS above.
static __device__ double bn_rbinom(Philox &g, double n, double pp)
{
    if (!(n > 0.0) || !(pp > 0.0)) return 0.0;
    if (pp >= 1.0) return n;
    const bool flip = pp > 0.5;
    const double p = flip ? 1.0 - pp : pp, q = 1.0 - p;
    double k;
    if (n * p < 30.0) {
        const double s = p / q;
        double f = exp(n * log1p(-p)), u = g.u53();
        k = 0.0;
        while (u > f && k < n) {
            u -= f;
            k += 1.0;
            f *= s * (n - k + 1.0) / k;
        }
    } else {
        const double spq = sqrt(n * p * q), b = 1.15 + 2.53 * spq, a = -0.0873 + 0.0248 * b + 0.01 * p;
        const double c = n * p + 0.5, vr = 0.92 - 4.2 / b, alpha = (2.83 + 5.1 / b) * spq;
        const double lpq = log(p / q), m = floor((n + 1.0) * p), h = lgamma(m + 1.0) + lgamma(n - m + 1.0);
        k = m;
        for (int it = 0; it < 256; it++) {
            const double U = g.u53() - 0.5, us = 0.5 - fabs(U);
            double V = g.u53();
            const double kk = floor((2.0 * a / us + b) * U + c);
            if (kk < 0.0 || kk > n) continue;
            if (us >= 0.07 && V <= vr) {
                k = kk;
                break;
            }
            V = log(V * alpha / (a / (us * us) + b));
            if (V <= h - lgamma(kk + 1.0) - lgamma(n - kk + 1.0) + (kk - m) * lpq) {
                k = kk;
                break;
            }
        }
    }
    return flip ? n - k : k;
}

// =================================================================================================
// Single-precision samplers for the 3-D posterior kernel.
//
// The 3-D output is 8 S bytes per gene x cell; at HBM speed a B200 has ~45 issue slots per sample,
// so the draws run on the FP32/integer pipes (128 lanes per SM) and only the rare slow-path
// acceptance tests use FP64.  Parity for draws is distributional (RNG streams differ from R's by
// construction): a relative error of ~1e-6 in a probability is far below what the per-element
// moment tests and the pooled PIT chi-square (3e6 draws) can resolve.
// =================================================================================================
__device__ __forceinline__ float bn_u01f(uint32_t w) { return ((float)(w >> 8) + 0.5f) * 5.9604644775390625e-8f; } // (0,1)

// NB(size r, mean M) by CDF inversion from 0 in float; one uniform per draw.
//   p(0) = (r/(r+M))^r,  p(k+1) = p(k) * q * (r+k)/(k+1),  q = M/(r+M)
struct NBLightF {
    float r, q, p0;
};
__device__ __forceinline__ NBLightF bn_nb_light_setup(float r, float M)
{
    NBLightF P;
    P.r = r;
    P.q = M / (r + M);
    P.p0 = __expf(-r * log1pf(M / r));
    return P;
}
__device__ __forceinline__ float bn_nb_light_draw(const NBLightF &P, uint32_t w, const float *__restrict__ inv_tab /*[64]: 1/(k+1)*/)
{
    float u = bn_u01f(w), p = P.p0, k = 0.0f;
#pragma unroll 1
    // A float CDF sums to 1 only within ~1e-7, so a uniform in the last 1e-7 could walk past the
    // whole support; once p has fallen below the resolution of u the crossing has happened.
    for (int it = 0; it < 64; it++) {
        if (u <= p || p < 1e-10f) return k;
        u -= p;
        p *= P.q * (P.r + k) * inv_tab[it];
        k += 1.0f;
    }
    // beyond 64 (reached with probability < 1e-9 for the means routed here): continue with divisions
    while (u > p && p >= 1e-10f && k < 4096.0f) {
        u -= p;
        p *= P.q * (P.r + k) / (k + 1.0f);
        k += 1.0f;
    }
    return k;
}

__device__ __forceinline__ float bn_rnormf(Philox &g)
{
    const float u1 = ((float)g.next() + 0.5f) * 2.3283064365386963e-10f, u2 = bn_u01f(g.next());
    return sqrtf(-2.0f * __logf(fminf(u1, 0.99999994f))) * cospif(2.0f * u2);
}

// Gamma(shape a, scale 1): Marsaglia & Tsang.  Squeeze and full test in float; the full test is
// repeated in double only when the float comparison is closer than its own rounding error bound
// (probability ~1e-4), so an FP64 logarithm almost never stalls the 31 other lanes of the warp.
//   log v = 3 log1p(cx),  1 - v = -(3t + 3t^2 + t^3)  (t = cx): no cancellation against 1
static __device__ float bn_rgammaf(Philox &g, float a)
{
    float boost = 1.0f;
    if (a < 1.0f) {
        boost = __powf(bn_u01f(g.next()), 1.0f / a);
        a += 1.0f;
    }
    const float d = a - 0.33333334f, c = rsqrtf(9.0f * d);
    for (int it = 0; it < 64; it++) {
        float x, t;
        do {
            x = bn_rnormf(g);
            t = c * x;
        } while (t <= -1.0f);
        const float v1 = 1.0f + t, v = v1 * v1 * v1;
        const float u = bn_u01f(g.next());
        const float x2 = x * x;
        if (u < 1.0f - 0.0331f * x2 * x2) return d * v * boost;
        const float gq = 3.0f * (log1pf(t) - t) - t * t * (3.0f + t);
        const float diff = logf(u) - (0.5f * x2 + d * gq);
        const float guard = 4e-6f + 2e-6f * (x2 + fabsf(x) * sqrtf(d));
        bool acc = diff < 0.0f;
        if (fabsf(diff) < guard) {
            const double td = (double)c * (double)x, vd = (1.0 + td) * (1.0 + td) * (1.0 + td);
            acc = log((double)u) < 0.5 * (double)x2 + (double)d * (1.0 - vd + log(vd));
        }
        if (acc) return d * v * boost;
    }
    return d * boost;
}

// log(k!) - [k log k - k + 0.5 log(2 pi k)]: the Stirling remainder, k >= 1
__device__ __forceinline__ float bn_stirling_remf(float k)
{
    const float ik = __frcp_rn(k), ik2 = ik * ik;
    return ik * (0.083333336f - ik2 * (0.0027777778f - ik2 * 0.0007936508f));
}

// Poisson(lam): inversion below 10, Hoermann's PTRS above.  The PTRS log test is evaluated in float in
// a cancellation-free form,
//   -lam + k log lam - log k! = lam (delta - (1 + delta) log1p(delta)) - 0.5 log(2 pi k) - rem(k),
//   delta = (k - lam)/lam,
// whose rounding error is ~1e-7 |k - lam|, and repeated in double only inside that error band.
static __device__ float bn_rpoisf(Philox &g, float lam)
{
    if (!(lam > 0.0f)) return 0.0f;
    if (lam < 10.0f) {
        float p = __expf(-lam), u = bn_u01f(g.next()), k = 0.0f;
        while (u > p && (p >= 1e-10f || k < lam) && k < 200.0f) {
            u -= p;
            k += 1.0f;
            p *= __fdividef(lam, k);
        }
        return k;
    }
    const float slam = sqrtf(lam), b = 0.931f + 2.53f * slam, a = -0.059f + 0.02483f * b;
    const float inv_alpha = 1.1239f + __fdividef(1.1328f, b - 3.4f), vr = 0.9277f - __fdividef(3.6224f, b - 2.0f);
    const float ilam = __frcp_rn(lam);
    for (int it = 0; it < 256; it++) {
        const float U = bn_u01f(g.next()) - 0.5f, V = bn_u01f(g.next()), us = 0.5f - fabsf(U);
        const float k = floorf((2.0f * a / us + b) * U + lam + 0.43f);
        if (us >= 0.07f && V <= vr) return k;
        if (k < 0.0f || (us < 0.013f && V > us)) continue;
        const float lhs = logf(V * inv_alpha / (a / (us * us) + b));
        float rhs, guard;
        if (k >= 8.0f) {
            const float dk = k - lam, delta = dk * ilam;
            rhs = lam * (delta - (1.0f + delta) * log1pf(delta)) - 0.5f * logf(6.2831855f * k) - bn_stirling_remf(k);
            guard = 1e-5f + 1e-6f * fabsf(dk) + 2e-7f * fabsf(rhs);
        } else { // far left tail of lam >= 10: tiny terms, plain form
            rhs = -lam + k * logf(lam) - lgammaf(k + 1.0f);
            guard = 1e-5f + 4e-7f * (lam + k * 3.0f);
        }
        const float diff = lhs - rhs;
        bool acc = diff <= 0.0f;
        if (fabsf(diff) < guard) {
            const double dl = lam, dk = k;
            acc = log((double)V) + log((double)inv_alpha) - log((double)a / ((double)us * us) + b) <= -dl + dk * log(dl) - lgamma(dk + 1.0);
        }
        if (acc) return k;
    }
    return floorf(lam);
}

// heavy NB draw: gamma-Poisson mixture, rnbinom_mu(size r, mu M) == rpois(rgamma(r) * M / r)
__device__ __forceinline__ double bn_nb_heavy_draw(float r, float M, Philox &g)
{
    if (M * (1.0f + M / r) > 3.0e4f * 3.0e4f || M > 1.0e5f) { // variance beyond float's integer grid: double path
        const double lam = bn_rgamma(g, (double)r) * ((double)M / (double)r);
        return bn_rpois(g, lam);
    }
    return (double)bn_rpoisf(g, bn_rgammaf(g, r) * (M / r));
}
