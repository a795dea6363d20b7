// bn_fit.cu -- per-gene maximum marginal likelihood of the binomial-thinned negative
// binomial: the whole of BB_Fun (R/PRIOR_FUNCTIONS.R:380-549, FIX_MU=TRUE) as ONE kernel
// launch instead of genes x evaluations .Call round trips into MarginalF_NB_1D
// (src/bayNorm_main.cpp:928-943).
//
// Likelihood of gene i at size phi (mean mu_i fixed, m_j = mu_i * beta_j):
//   l(phi) = sum_j log NB(x_ij; size phi, mean m_j)
//          = -phi * sum_{all cells j} log1p(m_j/phi)                          [dense part, C terms]
//            + sum_{x_ij > 0} { lgamma(x+phi) - lgamma(phi) + x*(log m_j - log(phi+m_j)) }
//            - sum_{x_ij > 0} lgamma(x+1)                                     [constant in phi]
// The x = 0 term phi*log(phi/(phi+m)) is what R's dnbinom_mu evaluates for zeros
// (`size * log(size/(size+mu))` or its log1p twin); written as -phi*log1p(m/phi) it is
// accurate in both regimes with no branch and no division per cell.
//
// Geometry: a group of TPG threads (one warp, or one 256-thread CTA for long genes) owns a
// gene; groups pull genes from an atomic queue (iteration counts vary 10x between genes).
// The optimiser state is replicated in every thread of the group (uniform control flow);
// each evaluation is a group-wide FP64 reduction with a fixed summation order, so a gene's
// result does not depend on which group ran it, on the grid size or on the GPU count.
#include <math.h>
#include <stdlib.h>

#include <cooperative_groups.h>

#include "bn_internal.cuh"
#include "bn_math.cuh"

// counts up to HMAX go through the tail table (one logarithm per table entry per evaluation, shared
// by all non-zeros of the gene), larger ones through Stirling's series
template <int TPG> struct FitCfg {
    static constexpr int HMAX = (TPG == 32) ? 64 : 256;
};

struct FitGene {
    const uint64_t *nz; // this gene's non-zeros, (cell << 32) | count
    int64_t nnz;
    const double *beta, *logbeta;
    int64_t C;
    double mu, lmu;
    double Kc;             // sum x*log(mu*beta_j) - sum lgamma(x+1): does not depend on phi
    double beta_max;
    const uint32_t *tail;  // shared memory: tail[k] = #{non-zeros with x > k}, k < HMAX
    int kmax;              // min(HMAX, largest count)
    int n_big;             // non-zeros with x > FIT_HMAX
};

// -l is evaluated as
//   l(phi) = -phi * sum_{all cells} log1p(m_j/phi)  -  sum_{nz} x_j log(phi + m_j)
//            + sum_{k < kmax} tail_k log(phi + k)  +  sum_{x_j > HMAX} [lgamma(x_j+phi) - lgamma(HMAX+phi)]  +  Kc
// using lgamma(x+phi) - lgamma(phi) = sum_{k<x} log(phi+k): all non-zeros of the gene share the
// same 64 logarithms, so a non-zero costs ONE log per evaluation and a cell costs one log1p.
template <int TPG>
__device__ double nb_loglik(const FitGene &g, double phi, const BnLogEntry *__restrict__ tab, double *scratch, int tid)
{
    const double r = g.mu / phi;
    // Dense part: sum_j log(1 + r beta_j) = log prod_j (1 + r beta_j).  Four independent running
    // products per thread, folded through ONE logarithm every FOLD factors: 2 FP64 instructions per
    // cell (fma, mul) instead of a 15-instruction log1p.  Each multiplication perturbs the logarithm
    // of the product by <= 2^-53 absolute, the same size as one rounding of the sum it replaces and
    // far below one ulp of l(phi); FOLD keeps the product inside the double range:
    // (1 + r beta_max)^16 < 2^1008 whenever r beta_max < 2^63.
    const double umax = fma(r, g.beta_max, 1.0);
    const int fold = umax < 9.0e18 ? 16 : (umax < 1.0e75 ? 4 : 1);
    double a0 = 0.0, a1 = 0.0;
    double p0 = 1.0, p1 = 1.0, p2 = 1.0, p3 = 1.0;
    int64_t j = tid;
    int cnt = 0;
    for (; j + 3 * TPG < g.C; j += 4 * TPG) {
        p0 *= fma(r, __ldg(g.beta + j), 1.0);
        p1 *= fma(r, __ldg(g.beta + j + TPG), 1.0);
        p2 *= fma(r, __ldg(g.beta + j + 2 * TPG), 1.0);
        p3 *= fma(r, __ldg(g.beta + j + 3 * TPG), 1.0);
        if (++cnt == fold) {
            a0 += bn_log(p0, tab) + bn_log(p1, tab);
            a1 += bn_log(p2, tab) + bn_log(p3, tab);
            p0 = p1 = p2 = p3 = 1.0;
            cnt = 0;
        }
    }
    for (; j < g.C; j += TPG) {
        p0 *= fma(r, __ldg(g.beta + j), 1.0);
        if (++cnt >= fold) {
            a0 += bn_log(p0, tab);
            p0 = 1.0;
            cnt = 0;
        }
    }
    a0 += bn_log(p0, tab) + bn_log(p1, tab);
    a1 += bn_log(p2, tab) + bn_log(p3, tab);
    const double a2 = 0.0, a3 = 0.0;
    double acc = -phi * ((a0 + a1) + (a2 + a3));
    constexpr int HMAX = FitCfg<TPG>::HMAX;
    double s0 = 0.0, s1 = 0.0;
    int64_t k = tid;
    // non-zeros: x_j log(phi + mu beta_j); the packed words and the beta gathers of four non-zeros
    // are in flight together (the gather address depends on the packed word)
    for (; k + 3 * TPG < g.nnz; k += 4 * TPG) {
        const uint64_t e0 = __ldg(g.nz + k), e1 = __ldg(g.nz + k + TPG), e2 = __ldg(g.nz + k + 2 * TPG),
                       e3 = __ldg(g.nz + k + 3 * TPG);
        const double b0 = __ldg(g.beta + (e0 >> 32)), b1 = __ldg(g.beta + (e1 >> 32)), b2 = __ldg(g.beta + (e2 >> 32)),
                     b3 = __ldg(g.beta + (e3 >> 32));
        s0 = fma(-(double)(uint32_t)e0, bn_log(fma(g.mu, b0, phi), tab), s0);
        s1 = fma(-(double)(uint32_t)e1, bn_log(fma(g.mu, b1, phi), tab), s1);
        s0 = fma(-(double)(uint32_t)e2, bn_log(fma(g.mu, b2, phi), tab), s0);
        s1 = fma(-(double)(uint32_t)e3, bn_log(fma(g.mu, b3, phi), tab), s1);
    }
    for (; k < g.nnz; k += TPG) {
        const uint64_t e0 = __ldg(g.nz + k);
        s0 = fma(-(double)(uint32_t)e0, bn_log(fma(g.mu, __ldg(g.beta + (e0 >> 32)), phi), tab), s0);
    }
    if (g.n_big != 0) { // rare: counts above the tail table, lgamma(x+phi) - lgamma(HMAX+phi) by Stirling
        const double lgref = bn_lgamma_stirling(phi + (double)HMAX, tab);
        for (k = tid; k < g.nnz; k += TPG) {
            const double x = (double)(uint32_t)__ldg(g.nz + k);
            if (x > (double)HMAX) s1 += bn_lgamma_stirling(x + phi, tab) - lgref;
        }
    }
    for (int q = tid; q < g.kmax; q += TPG) s1 = fma((double)g.tail[q], bn_log(phi + (double)q, tab), s1);
    return bn_group_sum<TPG>(acc + (s0 + s1), scratch) + g.Kc;
}

// d l / d phi = sum_j [ digamma(x+phi) - digamma(phi) + log(phi/(phi+m_j)) + (m_j - x)/(m_j + phi) ]
// (GradientFun_NB_1D, src/bayNorm_main.cpp:972-987)
__device__ __forceinline__ double bn_digamma(double x)
{
    double r = 0.0;
    while (x < 10.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    const double f = 1.0 / (x * x);
    const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}

template <int TPG> __device__ double nb_loglik_grad(const FitGene &g, double phi, double *scratch, int tid)
{
    const double r = g.mu / phi;
    double acc = 0.0;
    for (int64_t j = tid; j < g.C; j += TPG) {
        const double t = r * __ldg(g.beta + j); // m/phi
        acc += t / (1.0 + t) - log1p(t);        // log(phi/(phi+m)) + m/(m+phi)
    }
    const double dgphi = bn_digamma(phi);
    double sp = 0.0;
    for (int64_t k = tid; k < g.nnz; k += TPG) {
        const uint64_t e0 = __ldg(g.nz + k);
        const double x = (double)(uint32_t)e0;
        const double m = g.mu * __ldg(g.beta + (e0 >> 32));
        double dd;
        if (x <= 16.0) {
            dd = 0.0;
            for (double q = 0.0; q < x; q += 1.0) dd += 1.0 / (phi + q);
        } else
            dd = bn_digamma(x + phi) - dgphi;
        sp += dd - x / (m + phi);
    }
    return bn_group_sum<TPG>(acc + sp, scratch);
}

// ---------------------------------------------------------------------------------
// Solvers as resumable state machines: `next` is the point to evaluate, step(value) consumes
// -l(next) (or -dl/dphi when next_is_grad) and either asks for another point (true) or is
// finished (false).  One evaluation site in the kernel keeps the code small (the likelihood
// body is ~1.5k instructions) and every thread of the group steps the same state.
//
// SpgState is BB::spg restated for one bounded parameter (see oracle/baynorm_oracle.c:
// orc_spg for the line-by-line CPU twin and the provenance of every constant);
// control$maximize = TRUE, so the objective is func = -l.
// ---------------------------------------------------------------------------------
struct FitOpts {
    int method, maxfeval, maxit, M, use_gradient, evcap;
    double ftol, gtol, eps, xatol;
};

__device__ __forceinline__ double proj(double p, double lo, double hi) { return p < lo ? lo : (p > hi ? hi : p); }

struct SpgState {
    enum { INIT_F = 0, INIT_G = 1, LS = 2, GRAD = 3 };
    double lower, upper;
    double lastfv[16];
    double par, f, gr, lambda, pbest, fbest, fchg, pginfn;
    double d, gtd, fmaxv, alpha, pnew;
    int phase, iter, feval, lsflag, M;
    bool ls_first;
    // outputs / requests
    double next;
    bool next_is_grad;
    int conv;

    __device__ void init(double par0, double lo, double hi, const FitOpts &o)
    {
        lower = lo;
        upper = hi;
        M = o.M > 16 ? 16 : (o.M < 1 ? 1 : o.M);
#pragma unroll
        for (int i = 0; i < 16; i++) lastfv[i] = -1e99;
        iter = 0;
        feval = 0;
        lsflag = -1;
        fchg = INFINITY;
        lambda = 1.0;
        par = proj(par0, lo, hi);
        pbest = par;
        phase = INIT_F;
        next = par;
        next_is_grad = false;
        conv = 0;
    }
    __device__ __forceinline__ void ask_grad(double at, const FitOpts &o)
    {
        next_is_grad = o.use_gradient != 0;
        next = next_is_grad ? at : at + o.eps;
    }
    __device__ __forceinline__ bool finish()
    {
        if (lsflag <= 0) conv = 0; // iter >= maxit handled by the caller of finish()
        else conv = (lsflag == 1) ? 3 : (lsflag == 2) ? 2 : (lsflag == 3) ? 4 : 5;
        next = pbest;
        return false;
    }
    // start the next outer iteration or stop (the `while` header of spg)
    __device__ bool enter_iter(const FitOpts &o)
    {
        if (!(pginfn > o.gtol && iter <= o.maxit && fchg > o.ftol)) {
            const bool r = finish();
            if (lsflag <= 0 && iter >= o.maxit) conv = 1;
            return r;
        }
        iter++;
        d = proj(par - lambda * gr, lower, upper) - par;
        gtd = gr * d;
        if (isinf(gtd)) {
            lsflag = 4;
            return finish();
        }
        fmaxv = lastfv[0];
#pragma unroll
        for (int i = 1; i < 16; i++)
            if (i < M && lastfv[i] > fmaxv) fmaxv = lastfv[i];
        alpha = 1.0;
        pnew = par + alpha * d;
        phase = LS;
        ls_first = true;
        next = pnew;
        next_is_grad = false;
        return true;
    }
    __device__ bool step(double v, const FitOpts &o)
    {
        const double lmin = 1e-30, lmax = 1e30;
        switch (phase) {
        case INIT_F: {
            f = v;
            feval++;
            if (f != f || isinf(f)) {
                conv = 3;
                next = par;
                return false;
            }
            fbest = f;
            phase = INIT_G;
            ask_grad(par, o);
            return true;
        }
        case INIT_G: {
            gr = next_is_grad ? v : (v - f) / o.eps;
            if (gr != gr) {
                conv = 4;
                next = par;
                return false;
            }
            lastfv[0] = f;
            pginfn = fabs(proj(par - gr, lower, upper) - par);
            if (pginfn != 0.0) lambda = fmin(lmax, fmax(lmin, 1.0 / pginfn));
            return enter_iter(o);
        }
        case LS: {
            const double fnew = v;
            feval++;
            lsflag = 0;
            if (fnew != fnew) {
                lsflag = 1;
                return finish();
            }
            if (!ls_first && feval > o.maxfeval) {
                lsflag = 2;
                return finish();
            }
            ls_first = false;
            if (fnew > fmaxv + 1e-4 * alpha * gtd) { // nonmonotone Armijo test failed: backtrack
                if (alpha <= 0.1) alpha = alpha / 2;
                else {
                    double atemp = -(gtd * alpha * alpha) / (2 * (fnew - f - alpha * gtd));
                    if (atemp < 0.1 || atemp > 0.9 * alpha) atemp = alpha / 2;
                    alpha = atemp;
                }
                pnew = par + alpha * d;
                next = pnew;
                next_is_grad = false;
                return true;
            }
            fchg = fabs(f - fnew);
            f = fnew;
            const int slot = iter % M;
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (i == slot) lastfv[i] = f;
            phase = GRAD;
            ask_grad(pnew, o);
            return true;
        }
        default: { // GRAD
            const double gnew = next_is_grad ? v : (v - f) / o.eps;
            if (gnew != gnew) {
                lsflag = 3;
                return finish();
            }
            const double s = pnew - par, y = gnew - gr;
            par = pnew;
            gr = gnew;
            const double sts = s * s, yty = y * y;
            lambda = (sts == 0.0 || yty == 0.0) ? lmax : fmin(lmax, fmax(lmin, sqrt(sts / yty)));
            pginfn = fabs(proj(par - gr, lower, upper) - par);
            if (f < fbest) {
                fbest = f;
                pbest = pnew;
            }
            return enter_iter(o);
        }
        }
    }
};

// Bounded Brent (golden section + parabolic interpolation) on -l: the "argmax" solver the
// north star names; not what the reference runs, offered as opts.method = BN_FIT_BRENT.
struct BrentState {
    double a, b, x, w, v, fx, fw, fv, d, e, u;
    int it, conv;
    bool first;
    double next;

    __device__ void init(double lo, double hi)
    {
        a = lo;
        b = hi;
        d = e = 0.0;
        x = w = v = a + 0.3819660112501051 * (b - a);
        it = 0;
        conv = 1;
        first = true;
        next = x;
    }
    // choose the next trial point or stop
    __device__ bool propose(const FitOpts &o)
    {
        const double golden = 0.3819660112501051;
        if (it >= o.maxfeval) {
            next = x;
            return false;
        }
        it++;
        const double xm = 0.5 * (a + b), tol1 = 1.4901161193847656e-08 * fabs(x) + o.xatol / 3.0, tol2 = 2.0 * tol1;
        if (fabs(x - xm) <= tol2 - 0.5 * (b - a)) {
            conv = 0;
            next = x;
            return false;
        }
        bool golden_step = true;
        if (fabs(e) > tol1) {
            double r = (x - w) * (fx - fv), q = (x - v) * (fx - fw), pp = (x - v) * q - (x - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) pp = -pp;
            q = fabs(q);
            r = e;
            e = d;
            if (fabs(pp) < fabs(0.5 * q * r) && pp > q * (a - x) && pp < q * (b - x)) {
                d = pp / q;
                const double ut = x + d;
                if ((ut - a) < tol2 || (b - ut) < tol2) d = (xm - x >= 0) ? tol1 : -tol1;
                golden_step = false;
            }
        }
        if (golden_step) {
            e = (x >= xm) ? a - x : b - x;
            d = golden * e;
        }
        u = (fabs(d) >= tol1) ? x + d : x + ((d >= 0) ? tol1 : -tol1);
        next = u;
        return true;
    }
    __device__ bool step(double fu, const FitOpts &o)
    {
        if (first) {
            first = false;
            fx = fw = fv = fu;
            return propose(o);
        }
        if (fu <= fx) {
            if (u >= x) a = x;
            else b = x;
            v = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
        } else {
            if (u < x) a = u;
            else b = u;
            if (fu <= fw || w == x) { v = w; fv = fw; w = u; fw = fu; }
            else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
        }
        return propose(o);
    }
};

template <int TPG> __device__ __forceinline__ void group_sync()
{
    if (TPG == 32) __syncwarp();
    else __syncthreads();
}

template <int TPG, bool GRAD>
__global__ void __launch_bounds__(TPG == 32 ? 128 : TPG, TPG == 32 ? 8 : 4)
    k_fit_size(int64_t G, int64_t C, const int64_t *__restrict__ rowptr, const uint64_t *__restrict__ gm,
               const double *__restrict__ beta, const double *__restrict__ logbeta,
               const double *__restrict__ mu0, const double *__restrict__ size0, const double *__restrict__ params,
               FitOpts o, unsigned long long *__restrict__ queue, double *__restrict__ size_out,
               int32_t *__restrict__ conv_out, int32_t *__restrict__ nfe_out)
{
    const double lower = params[0], upper = params[1], beta_max = params[2]; // device-resident: no host round trip
    constexpr int NGROUP = (TPG == 32) ? 4 : 1;
    constexpr int HMAX = FitCfg<TPG>::HMAX;
    __shared__ BnLogEntry tab[BN_LOG_TABLE_N];
    __shared__ double scratch[8];
    __shared__ unsigned long long next_gene;
    __shared__ uint32_t hist_all[NGROUP][HMAX + 1];
    __shared__ int xmax_all[NGROUP];
    bn_log_table_load(tab);
    __syncthreads();
    const int tid = (TPG == 32) ? (threadIdx.x & 31) : threadIdx.x;
    const int grp = (TPG == 32) ? (threadIdx.x >> 5) : 0;
    uint32_t *hist = hist_all[grp];
    for (;;) {
        unsigned long long gq;
        if (TPG == 32) {
            if (tid == 0) gq = atomicAdd(queue, 1ULL);
            gq = __shfl_sync(0xffffffffu, gq, 0);
        } else {
            __syncthreads();
            if (tid == 0) next_gene = atomicAdd(queue, 1ULL);
            __syncthreads();
            gq = next_gene;
        }
        if (gq >= (unsigned long long)G) break;
        const int64_t gi = (int64_t)gq;
        FitGene g;
        const int64_t a = rowptr[gi];
        g.nz = gm + a;
        g.nnz = rowptr[gi + 1] - a;
        g.beta = beta;
        g.logbeta = logbeta;
        g.C = C;
        g.mu = mu0[gi];
        g.lmu = log(g.mu);
        g.beta_max = beta_max;
        g.tail = hist;
        double par = size0[gi], res = par;
        int conv = 0, nfe = 0;
        if (!(g.mu > 0.0)) {
            // flat likelihood (mean 0: an all-zero gene under the MME start): spg's first projected gradient is 0, it
            // returns the start.  A gene without stored entries but a caller-supplied mean > 0 is NOT flat
            // (l = -phi sum log1p(mu beta_j / phi)) and takes the normal path with an empty tail table.
            res = proj(par, lower, upper);
            if (o.method == BN_FIT_SPG) nfe = 2;
        } else {
            // per-gene constants: count histogram -> tail table, Kc, largest count
            for (int q = tid; q <= HMAX; q += TPG) hist[q] = 0;
            if (tid == 0) xmax_all[grp] = 0;
            group_sync<TPG>();
            double kk = 0.0;
            int xm = 0;
            for (int64_t k = tid; k < g.nnz; k += TPG) {
                const uint64_t e0 = __ldg(g.nz + k);
                const double x = (double)(uint32_t)e0;
                kk += x * (g.lmu + __ldg(logbeta + (e0 >> 32))) - bn_lfact(x);
                const int xi = x > (double)HMAX ? HMAX + 1 : (int)x;
                if (xi > 0) atomicAdd(&hist[xi - 1], 1u); // an explicitly stored zero has no lgamma term
                xm = max(xm, xi);
            }
            atomicMax(&xmax_all[grp], xm);
            g.Kc = bn_group_sum<TPG>(kk, scratch);
            group_sync<TPG>();
            if (tid == 0) { // suffix sum: tail[k] = #{x > k}
                uint32_t run = hist[HMAX];
                for (int q = HMAX - 1; q >= 0; q--) {
                    run += hist[q];
                    hist[q] = run;
                }
            }
            group_sync<TPG>();
            g.n_big = (int)(hist[HMAX]);
            g.kmax = min(HMAX, xmax_all[grp]);
            if (o.method == BN_FIT_BRENT) {
                BrentState st;
                st.init(lower, upper);
                bool more = true;
                while (more) {
                    const double val = -nb_loglik<TPG>(g, st.next, tab, scratch, tid);
                    nfe++;
                    more = st.step(val, o);
                }
                res = st.next;
                conv = st.conv;
            } else {
                SpgState st;
                st.init(par, lower, upper, o);
                bool more = true;
                while (more) {
                    double val;
                    if (GRAD && st.next_is_grad) val = -nb_loglik_grad<TPG>(g, st.next, scratch, tid);
                    else val = -nb_loglik<TPG>(g, st.next, tab, scratch, tid);
                    nfe++;
                    more = st.step(val, o);
                }
                res = st.next;
                conv = st.conv;
            }
            group_sync<TPG>(); // hist is reused by the next gene
        }
        if (tid == 0) {
            size_out[gi] = res;
            if (conv_out) conv_out[gi] = conv;
            if (nfe_out) nfe_out[gi] = nfe;
        }
    }
}

// ---------------------------------------------------------------------------------
// Version 2 of the 1-D fit: the cost of one evaluation is proportional to the gene's NON-ZEROS, not to the cells.
//
//   l(phi) = -phi F(mu/phi)  -  sum_nz x_j log(1 + r beta_j)  +  sum_{1 <= k < kmax} tail_k log1p(k/phi)  +  [x > HMAX]  +  Kc,
//   F(r) = sum_{all cells j} log(1 + r beta_j),   r = mu/phi
//
// (the expression of nb_loglik with log(phi + mu beta_j) = log phi + log(1 + r beta_j); the X log phi terms of the non-zero
// sum and of the tail sum cancel analytically instead of numerically).
//  * F depends on the gene only through the scalar r, so it is tabulated ONCE per call: piecewise Chebyshev interpolants of
//    degree 12 in t = log r on intervals of width 1/4 (F(e^t) is analytic in the strip |Im t| < pi: the interpolation error is
//    ~50^-13 relative, far below one rounding), node values by compensated summation over the C cells.  An evaluation costs
//    ~50 FP64 instructions in one thread instead of 3 per cell in every thread.
//  * The non-zeros of a gene are staged once per call as beta_j, sorted by count (descending).  prod_j u_j^x_j is then a
//    product of prefix products of the staged order (see nb_loglik2): 2 FP64 instructions per non-zero and trial size
//    whatever the count, against 14 for x log(phi + mu beta_j) through bn_log, and 8 bytes of coalesced traffic per non-zero
//    and pass against a packed word plus a gathered BETA sector (40 bytes).  Every multiplication perturbs the logarithm of
//    the product by <= 2^-53 absolute -- less than the rounding of the partial sums it replaces.
//  * spg's value at p and its forward-difference partner at p + eps share one pass over the entries (see k_fit2).
// Results are a pure function of the gene's data (fixed thread map and order); which genes share a launch does not matter.
// ---------------------------------------------------------------------------------
#define DT_T0 (-48.0) // table covers r in [e^-48, e^48]; the call builds only the intervals its box and means can reach
#define DT_NI 384     // intervals of width DT_H = 1/4 in log r
#define DT_N 13       // Chebyshev nodes / coefficients per interval
#define DT_STRIDE 16  // doubles per interval: DT_N coefficients (c0 halved), 1 / r_centre, padding
#define FIT2_HMAX 256 // tail table length (counts beyond: Stirling)
#define FIT2_CMAX 32  // counts up to here go through the prefix products, larger ones take one logarithm per entry

struct DenseMeta {
    int i_lo, i_hi;    // intervals built
    double r_lo, r_hi; // r inside [r_lo, r_hi] is covered
};

// range of r = mu/phi the solvers can ask for: phi in [lower, upper + eps], mu over the genes with mu > 0
__global__ void __launch_bounds__(1024) k_dt_range(int64_t G, const double *__restrict__ mu, const double *__restrict__ params,
                                                   double r_lo_user, double r_hi_user, DenseMeta *__restrict__ meta)
{
    __shared__ double smin[32], smax[32];
    double mn = INFINITY, mx = 0.0;
    for (int64_t g = threadIdx.x; g < G; g += blockDim.x) {
        const double v = mu ? mu[g] : 0.0;
        if (v > 0.0 && v < INFINITY) {
            mn = fmin(mn, v);
            mx = fmax(mx, v);
        }
    }
    mn = bn_warp_min(mn);
    mx = bn_warp_max(mx);
    if ((threadIdx.x & 31) == 0) {
        smin[threadIdx.x >> 5] = mn;
        smax[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; k++) {
            mn = fmin(mn, smin[k]);
            mx = fmax(mx, smax[k]);
        }
        DenseMeta m;
        m.i_lo = 1;
        m.i_hi = 0;
        m.r_lo = INFINITY;
        m.r_hi = 0.0;
        double rl, rh;
        if (params) { // the fit: box from device memory
            const double lower = params[0], upper = params[1];
            rl = mn / (upper * (1.0 + 1e-6) + 1e-6);
            rh = mx / lower;
        } else {
            rl = r_lo_user;
            rh = r_hi_user;
        }
        if (rh > 0.0 && rl <= rh && rl > 0.0 && rh < INFINITY) {
            int a = (int)floor((log(rl) - DT_T0) * 4.0) - 1, b = (int)floor((log(rh) - DT_T0) * 4.0) + 1;
            a = a < 0 ? 0 : a;
            b = b > DT_NI - 1 ? DT_NI - 1 : b;
            if (a <= b) {
                m.i_lo = a;
                m.i_hi = b;
                // strictly inside the built intervals: the index computed from a rounded log r may be one off at a boundary
                m.r_lo = exp(DT_T0 + 0.25 * a + 1e-3);
                m.r_hi = exp(DT_T0 + 0.25 * (b + 1) - 1e-3);
            }
        }
        *meta = m;
    }
}

// node values: part[(i * nsplit + s) * DT_N + k] = sum over the cells of split s of log1p(r_k beta_j), compensated
__global__ void __launch_bounds__(256) k_dt_nodes(int64_t C, const double *__restrict__ beta, const DenseMeta *__restrict__ meta,
                                                  int nsplit, double *__restrict__ part)
{
    const int i = blockIdx.x, s = blockIdx.y;
    if (i < meta->i_lo || i > meta->i_hi) return;
    __shared__ double sh[8][DT_N];
    const double rc = exp(DT_T0 + 0.25 * (i + 0.5));
    double rk[DT_N], acc[DT_N], cmp[DT_N];
#pragma unroll
    for (int k = 0; k < DT_N; k++) {
        rk[k] = rc * exp(0.125 * cospi((k + 0.5) / DT_N));
        acc[k] = 0.0;
        cmp[k] = 0.0;
    }
    const int64_t j0 = C * s / nsplit, j1 = C * (s + 1) / nsplit;
    for (int64_t j = j0 + threadIdx.x; j < j1; j += 256) {
        const double b = __ldg(beta + j);
#pragma unroll
        for (int k = 0; k < DT_N; k++) { // Kahan: the node values carry the accuracy of the whole table
            const double y = log1p(rk[k] * b) - cmp[k];
            const double t = acc[k] + y;
            cmp[k] = (t - acc[k]) - y;
            acc[k] = t;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < DT_N; k++) {
        const double v = bn_warp_sum(acc[k] - cmp[k]);
        if (lane == 0) sh[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < DT_N) {
        double v = 0.0;
        for (int w = 0; w < 8; w++) v += sh[w][threadIdx.x];
        part[((int64_t)i * nsplit + s) * DT_N + threadIdx.x] = v;
    }
}

// Chebyshev coefficients of every built interval: c_m = (2/N) sum_k f_k cos(m pi (k + 1/2) / N)
__global__ void __launch_bounds__(32) k_dt_coef(const DenseMeta *__restrict__ meta, int nsplit, const double *__restrict__ part,
                                                double *__restrict__ coef)
{
    const int i = blockIdx.x, lane = threadIdx.x;
    if (i < meta->i_lo || i > meta->i_hi) return;
    double f = 0.0, c = 0.0;
    if (lane < DT_N)
        for (int s = 0; s < nsplit; s++) {
            const double y = part[((int64_t)i * nsplit + s) * DT_N + lane] - c;
            const double t = f + y;
            c = (t - f) - y;
            f = t;
        }
    double cm = 0.0;
    for (int k = 0; k < DT_N; k++) {
        const double fk = __shfl_sync(0xffffffffu, f, k);
        cm += fk * cospi((double)lane * (k + 0.5) / DT_N);
    }
    cm *= 2.0 / DT_N;
    if (lane == 0) cm *= 0.5;
    if (lane < DT_N) coef[i * DT_STRIDE + lane] = cm;
    if (lane == DT_N) coef[i * DT_STRIDE + DT_N] = 1.0 / exp(DT_T0 + 0.25 * (i + 0.5));
}

// F(r) from the table (r inside [meta.r_lo, meta.r_hi])
__device__ __forceinline__ double dt_eval(double r, const double *__restrict__ coef, const DenseMeta *__restrict__ meta,
                                          const BnLogEntry *__restrict__ tab)
{
    int i = (int)floor((bn_log(r, tab) - DT_T0) * 4.0);
    i = i < meta->i_lo ? meta->i_lo : (i > meta->i_hi ? meta->i_hi : i);
    const double *__restrict__ c = coef + i * DT_STRIDE;
    // local coordinate from the ratio to the interval centre: log of a number in [0.88, 1.14] keeps full relative accuracy
    const double x = 8.0 * bn_log(r * __ldg(c + DT_N), tab), x2 = 2.0 * x;
    double b1 = 0.0, b2 = 0.0;
#pragma unroll
    for (int k = DT_N - 1; k >= 1; k--) {
        const double b0 = fma(x2, b1, __ldg(c + k) - b2);
        b2 = b1;
        b1 = b0;
    }
    return fma(x, b1, __ldg(c) - b2);
}

// per-gene record written by the staging kernel
struct GeneMeta2 {
    double Kbase; // sum x log beta_j - sum lgamma(x + 1)
    double X;     // sum x
    int kmax;     // min(FIT2_HMAX, largest count)
    int n_big;    // stored entries with x > FIT2_HMAX
};

#define FIT2_RW 8   // entries (four pairs) a thread of the CTA kernel copies per round
#define FIT2_NST 4  // rounds of its shared-memory ring
#define FIT2_RING_BYTES (FIT2_NST * FIT2_RW * 256 * 8)
#define FIT2_NK (FIT2_HMAX + 2) // sort keys: 0 (explicitly stored zero: dropped), 1..HMAX, HMAX + 1 (larger counts)

// Staging, CTA per gene (atomic queue): a counting sort of the gene's stored entries by count, DESCENDING -- entries with
// x > HMAX first (positions [0, tail[HMAX])), then count c at [tail[c], tail[c - 1]), tail[c] = #{x > c} -- plus the tail
// table, Kbase and X.  gs[rowptr[g] + position] = beta_j (for x > HMAX: the position of the entry in gm, as raw bits).
// Every warp owns a contiguous eighth of the gene and keeps the order of its entries, so the staged order is one fixed
// function of the gene's data (cells ascending inside a count group).
__global__ void __launch_bounds__(256) k_fit2_stage(int64_t G, const int64_t *__restrict__ rowptr, const uint64_t *__restrict__ gm,
                                                    const double *__restrict__ beta, const double *__restrict__ logbeta,
                                                    unsigned long long *__restrict__ queue, double *__restrict__ gs,
                                                    uint8_t *__restrict__ gx, GeneMeta2 *__restrict__ gmeta,
                                                    uint32_t *__restrict__ tail_all)
{
    __shared__ uint32_t hist[FIT2_HMAX + 1]; // hist[q]: #{x == q + 1} (q = HMAX: larger), then tail[q] = #{x > q}
    __shared__ uint32_t wh[8][FIT2_NK];      // per-warp counts per key, then the warp's next free position per key
    __shared__ double red8[8], red8b[8];
    __shared__ int xmax_sh;
    __shared__ unsigned long long next_gene;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned lt = (1u << lane) - 1u;
    for (;;) {
        __syncthreads();
        if (tid == 0) next_gene = atomicAdd(queue, 1ULL);
        for (int q = tid; q < 8 * FIT2_NK; q += 256) (&wh[0][0])[q] = 0;
        if (tid == 0) xmax_sh = 0;
        __syncthreads();
        if (next_gene >= (unsigned long long)G) break;
        const int64_t gi = (int64_t)next_gene, a = rowptr[gi], nnz = rowptr[gi + 1] - a;
        const uint64_t *__restrict__ nz = gm + a;
        const int64_t c0 = nnz * warp / 8, c1 = nnz * (warp + 1) / 8;
        double kk = 0.0, xs = 0.0;
        int xm = 0;
        // four batches of 32 entries in flight per warp (packed word, then the gathered log BETA): one batch at a time left the
        // kernel waiting on memory 27 cycles per issued instruction
        for (int64_t base = c0; base < c1; base += 128) {
            uint64_t e[4];
            double lb[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t k = base + 32 * u + lane;
                e[u] = k < c1 ? __ldg(nz + k) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) lb[u] = (base + 32 * u + lane < c1) ? __ldg(logbeta + (e[u] >> 32)) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                int key = -1;
                if (base + 32 * u + lane < c1) {
                    const uint32_t xi = (uint32_t)e[u];
                    const double x = (double)xi;
                    kk += x * lb[u] - bn_lfact(x);
                    xs += x;
                    key = xi > FIT2_HMAX ? FIT2_HMAX + 1 : (int)xi;
                    xm = max(xm, key);
                }
                const unsigned same = __match_any_sync(0xffffffffu, key);
                if (key > 0 && (same & lt) == 0) wh[warp][key] += (uint32_t)__popc(same);
                __syncwarp();
            }
        }
        atomicMax(&xmax_sh, xm);
        const double Kbase = bn_group_sum<256>(kk, red8);
        const double X = bn_group_sum<256>(xs, red8b);
        __syncthreads();
        for (int key = 1 + tid; key < FIT2_NK; key += 256) {
            uint32_t t = 0;
            for (int w = 0; w < 8; w++) t += wh[w][key];
            hist[key - 1] = t;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t r = hist[FIT2_HMAX];
            GeneMeta2 gmv;
            gmv.Kbase = Kbase;
            gmv.X = X;
            gmv.n_big = (int)r;
            gmv.kmax = min(FIT2_HMAX, xmax_sh);
            gmeta[gi] = gmv;
            for (int q = FIT2_HMAX - 1; q >= 0; q--) { // suffix sum: tail[k] = #{x > k}
                r += hist[q];
                hist[q] = r;
            }
        }
        __syncthreads();
        uint32_t *__restrict__ tl = tail_all + gi * (int64_t)(FIT2_HMAX + 1);
        for (int q = tid; q <= FIT2_HMAX; q += 256) tl[q] = hist[q];
        for (int key = 1 + tid; key < FIT2_NK; key += 256) {
            uint32_t pos = key > FIT2_HMAX ? 0u : hist[key]; // first position of the key's group
            for (int w = 0; w < 8; w++) {
                const uint32_t n = wh[w][key];
                wh[w][key] = pos;
                pos += n;
            }
        }
        __syncthreads();
        double *__restrict__ out = gs + a;
        for (int64_t base = c0; base < c1; base += 128) {
            uint64_t e[4];
            double bv[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t k = base + 32 * u + lane;
                e[u] = k < c1 ? __ldg(nz + k) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) bv[u] = (base + 32 * u + lane < c1) ? __ldg(beta + (e[u] >> 32)) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int64_t k = base + 32 * u + lane;
                int key = -1;
                if (k < c1) {
                    const uint32_t xi = (uint32_t)e[u];
                    key = xi > FIT2_HMAX ? FIT2_HMAX + 1 : (int)xi;
                }
                const unsigned same = __match_any_sync(0xffffffffu, key);
                const uint32_t pos = key > 0 ? wh[warp][key] : 0u;
                __syncwarp();
                if (key > 0) {
                    if ((same & lt) == 0) wh[warp][key] = pos + (uint32_t)__popc(same);
                    const uint32_t at = pos + (uint32_t)__popc(same & lt);
                    out[at] = key <= FIT2_HMAX ? bv[u] : __longlong_as_double((long long)(a + k));
                    gx[a + at] = (uint8_t)(key <= FIT2_HMAX ? key - 1 : 0); // count - 1 (1..256 in a byte)
                }
                __syncwarp();
            }
        }
    }
}

struct FitGene2 {
    const double *gs;     // staged entries of the gene
    const uint8_t *gx;    // their counts - 1 (entries up to FIT2_HMAX)
    double *ring;         // CTA kernel: FIT2_RING_BYTES of shared memory
    const uint32_t *tail; // shared memory: tail[k] = #{x > k}, k <= FIT2_HMAX
    double mu, Kc;
    int kmax, n_big;
};

// rare paths of nb_loglik2 as real calls: the evaluation body stays small enough for the instruction cache
__device__ __noinline__ double bn_log_far(double x, const BnLogEntry *__restrict__ tab) { return bn_log(x, tab); }

// entries [tid, nb) step TPG of the gene's staged order (counts beyond the tail table): -x log u plus Stirling's series for
// lgamma(x + phi) - lgamma(HMAX + phi) - (x - HMAX) log phi
__device__ __noinline__ double fit2_big(const double *__restrict__ gs, int nb, int tid, int tpg, const uint64_t *__restrict__ gm,
                                        const double *__restrict__ beta, double r, double phi, const BnLogEntry *__restrict__ tab)
{
    const double lgref = bn_lgamma_stirling(phi + (double)FIT2_HMAX, tab), lphi = bn_log(phi, tab);
    double s = 0.0;
    for (int k = tid; k < nb; k += tpg) {
        const uint64_t e0 = __ldg(gm + __double_as_longlong(__ldg(gs + k)));
        const double x = (double)(uint32_t)e0, b = __ldg(beta + (e0 >> 32));
        s = fma(-x, bn_log(fma(r, b, 1.0), tab), s);
        s += (bn_lgamma_stirling(x + phi, tab) - lgref) - (x - (double)FIT2_HMAX) * lphi;
    }
    return s;
}

// the dense part summed directly (an r the table was not built for: cannot happen for the box the table was built from)
__device__ __noinline__ double fit2_dense_direct(const double *__restrict__ beta, int64_t C, int tid, int tpg, double r,
                                                 const BnLogEntry *__restrict__ tab)
{
    double d = 0.0;
    for (int64_t j = tid; j < C; j += tpg) d += bn_log1p_pos(r * __ldg(beta + j), tab);
    return d;
}

// -l at NP trial sizes in ONE pass over the gene's staged entries.
//
// sum_nz x_j log u_j (u_j = 1 + r beta_j) without a logarithm per entry: with the entries sorted by count, descending,
//   prod_j u_j^x_j = prod_{c >= 1} S_c,   S_c = prod_{x_j >= c} u_j  (a prefix of the staged order),
// so a thread walks its entries once, multiplies each u into the running prefix product S (fma + mul per entry and trial
// size, whatever the count) and multiplies S into T each time it passes from count c to c - 1.  Exponent budgets keep S and T
// inside the double range: before either could overflow -- and across runs of absent counts, where crossings would
// otherwise cost a multiplication each -- S is folded as acc += c log S (its entries still have c crossings ahead), T as
// acc += log T.  Every multiplication perturbs the logarithm of its product by <= 2^-53 absolute.  Counts above FIT2_CMAX
// (few entries per count and thread: the crossings would cost more than they save) take one logarithm per entry.
template <int TPG, int NP>
__device__ __forceinline__ void nb_loglik2(const FitGene2 &g, const double *phi, double *val, const uint64_t *__restrict__ gm,
                                           const double *__restrict__ beta, int64_t C, double beta_max, const double *__restrict__ coef,
                                           const DenseMeta *__restrict__ meta, const BnLogEntry *__restrict__ tab, double *scratch, int tid)
{
    double r[NP], S[NP], T[NP], acc[NP];
    double rmax = 0.0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
        r[p] = g.mu / phi[p];
        rmax = fmax(rmax, r[p]);
        S[p] = 1.0;
        T[p] = 1.0;
        acc[p] = 0.0;
    }
    int ub = ((__double2hiint(fma(rmax, beta_max, 1.0)) >> 20) & 0x7ff) - 1022; // u < 2^ub for every cell and trial size
    ub = ub > 1000 ? 1000 : ub;
    const uint32_t *__restrict__ tail = g.tail;
    const int n = (int)tail[0]; // staged entries (a gene has fewer than 2^31: one per cell at most)
    int eS = 0, eT = 0, level = g.kmax < FIT2_CMAX ? g.kmax : FIT2_CMAX;
    int nextb = level >= 1 ? (int)tail[level - 1] : n; // first position whose count is below `level`
#define FIT2_FOLD_S()                                                                  \
    {                                                                                  \
        _Pragma("unroll") for (int p = 0; p < NP; p++)                                 \
        {                                                                              \
            acc[p] = fma((double)level, bn_log_far(S[p], tab), acc[p]);                \
            S[p] = 1.0;                                                                \
        }                                                                              \
        eS = 0;                                                                        \
    }
#define FIT2_ENTRY(kk_, b_)                                                            \
    if ((kk_) < n) {                                                                   \
        while ((kk_) >= nextb) {                                                       \
            if (eS) {                                                                  \
                const bool gap = level >= 2 && (int)tail[level - 2] == nextb;          \
                if (gap || eT + eS > 1000) {                                           \
                    FIT2_FOLD_S()                                                      \
                    if (eT > 500) {                                                    \
                        _Pragma("unroll") for (int p = 0; p < NP; p++)                 \
                        {                                                              \
                            acc[p] += bn_log_far(T[p], tab);                           \
                            T[p] = 1.0;                                                \
                        }                                                              \
                        eT = 0;                                                        \
                    }                                                                  \
                } else {                                                               \
                    _Pragma("unroll") for (int p = 0; p < NP; p++) T[p] *= S[p];       \
                    eT += eS;                                                          \
                }                                                                      \
            }                                                                          \
            level--;                                                                   \
            nextb = level >= 1 ? (int)tail[level - 1] : n;                             \
        }                                                                              \
        if (eS + ub > 1000) FIT2_FOLD_S()                                              \
        _Pragma("unroll") for (int p = 0; p < NP; p++) S[p] *= fma(r[p], (b_), 1.0);   \
        eS += ub;                                                                      \
    }
    const double *__restrict__ gs = g.gs;
    const int nm = g.kmax > FIT2_CMAX ? (int)tail[FIT2_CMAX] : g.n_big; // start of the prefix-product section
    // A round whose entries all carry the current count and whose factors fit the exponent budget -- nearly every round of a
    // long gene -- needs none of the per-entry tests: RW fused multiply-adds and a product tree per trial size.
    if (TPG == 32) {
        for (int k = nm + tid; k < n; k += 4 * TPG) {
            const int k1 = k + TPG, k2 = k + 2 * TPG, k3 = k + 3 * TPG;
            const double b0 = __ldg(gs + k), b1 = k1 < n ? __ldg(gs + k1) : 0.0, b2 = k2 < n ? __ldg(gs + k2) : 0.0,
                         b3 = k3 < n ? __ldg(gs + k3) : 0.0;
            if (k3 < nextb && eS + 4 * ub <= 1000) {
#pragma unroll
                for (int p = 0; p < NP; p++)
                    S[p] *= (fma(r[p], b0, 1.0) * fma(r[p], b1, 1.0)) * (fma(r[p], b2, 1.0) * fma(r[p], b3, 1.0));
                eS += 4 * ub;
            } else {
                FIT2_ENTRY(k, b0)
                FIT2_ENTRY(k1, b1)
                FIT2_ENTRY(k2, b2)
                FIT2_ENTRY(k3, b3)
            }
        }
    } else {
        // Long genes stream their entries from L2 / HBM (config 5: 830 KB per gene and pass).  Every thread keeps FIT2_NST - 1
        // rounds of four entry PAIRS in flight through its own column of a shared-memory ring (cp.async, 16 bytes each; a
        // thread only ever reads what it copied itself, so cp.async.wait_group is the only synchronisation).  With the loads
        // of a round issued only after the previous round was consumed, every round exposed one DRAM latency (long-scoreboard
        // stall 12 cycles per issued instruction); 8-byte copies were bound by the load/store unit's queue instead.
        // Pairs start at an even absolute position (16-byte alignment): an odd first entry goes to thread 0 ahead of its pairs.
        const int nm2 = nm + (int)((reinterpret_cast<uintptr_t>(gs + nm) >> 3) & 1);
        if (nm2 > nm && tid == 0) FIT2_ENTRY(nm, __ldg(gs + nm))
        double2 *rg = reinterpret_cast<double2 *>(g.ring) + tid; // pair i of stage s at rg[(s * 4 + i) * TPG]
        const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(rg);
        const double2 *src = reinterpret_cast<const double2 *>(gs + nm2) + tid;
        int kis = nm2 + 2 * tid, sti = 0, stc = 0; // first entry of the thread's first pair in the next round to issue
#define FIT2_ISSUE()                                                                                                     \
    {                                                                                                                    \
        const uint32_t dst_ = ring0 + (uint32_t)sti * (4 * TPG * 16);                                                    \
        if (kis + 6 * TPG + 1 < n) {                                                                                     \
            _Pragma("unroll") for (int i = 0; i < 4; i++)                                                                \
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_ + i * (TPG * 16)), "l"(src + i * TPG) : "memory"); \
        } else if (kis < n) {                                                                                            \
            _Pragma("unroll") for (int i = 0; i < 4; i++)                                                                \
            {                                                                                                            \
                const int e0_ = kis + 2 * i * TPG;                                                                       \
                if (e0_ + 1 < n)                                                                                         \
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_ + i * (TPG * 16)), "l"(src + i * TPG) : "memory"); \
                else if (e0_ < n)                                                                                        \
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_ + i * (TPG * 16)), "l"(src + i * TPG) : "memory"); \
            }                                                                                                            \
        }                                                                                                                \
        asm volatile("cp.async.commit_group;" ::: "memory");                                                             \
        src += 4 * TPG;                                                                                                  \
        kis += 8 * TPG;                                                                                                  \
        sti = sti + 1 == FIT2_NST ? 0 : sti + 1;                                                                         \
    }
#pragma unroll
        for (int q = 0; q < FIT2_NST - 1; q++) FIT2_ISSUE()
        for (int k = nm2 + 2 * tid; k < n; k += 8 * TPG) {
            FIT2_ISSUE()
            asm volatile("cp.async.wait_group %0;" ::"n"(FIT2_NST - 1) : "memory");
            const double2 *st = rg + stc * (4 * TPG);
            stc = stc + 1 == FIT2_NST ? 0 : stc + 1;
            const double2 v0 = st[0], v1 = st[TPG], v2 = st[2 * TPG], v3 = st[3 * TPG];
            const int kl = k + 6 * TPG + 1; // the round's last entry
            if (kl < nextb && eS + 8 * ub <= 1000) {
#pragma unroll
                for (int p = 0; p < NP; p++)
                    S[p] *= ((fma(r[p], v0.x, 1.0) * fma(r[p], v0.y, 1.0)) * (fma(r[p], v1.x, 1.0) * fma(r[p], v1.y, 1.0))) *
                            ((fma(r[p], v2.x, 1.0) * fma(r[p], v2.y, 1.0)) * (fma(r[p], v3.x, 1.0) * fma(r[p], v3.y, 1.0)));
                eS += 8 * ub;
            } else {
                const int e1 = k + 2 * TPG, e2 = k + 4 * TPG, e3 = k + 6 * TPG;
                FIT2_ENTRY(k, v0.x)
                FIT2_ENTRY(k + 1, v0.y)
                FIT2_ENTRY(e1, v1.x)
                FIT2_ENTRY(e1 + 1, v1.y)
                FIT2_ENTRY(e2, v2.x)
                FIT2_ENTRY(e2 + 1, v2.y)
                FIT2_ENTRY(e3, v3.x)
                FIT2_ENTRY(e3 + 1, v3.y)
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#undef FIT2_ISSUE
    }
#undef FIT2_ENTRY
#undef FIT2_FOLD_S
    double s[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
        if (eS) acc[p] = fma((double)level, bn_log_far(S[p], tab), acc[p]);
        if (eT) acc[p] += bn_log_far(T[p], tab);
        s[p] = -acc[p];
    }
    if (g.kmax > FIT2_CMAX) {
        const uint8_t *__restrict__ gx = g.gx;
        int k = g.n_big + tid;
        for (; k + TPG < nm; k += 2 * TPG) {
            const double b0 = __ldg(gs + k), b1 = __ldg(gs + k + TPG);
            const double x0 = (double)((int)__ldg(gx + k) + 1), x1 = (double)((int)__ldg(gx + k + TPG) + 1);
#pragma unroll
            for (int p = 0; p < NP; p++) {
                s[p] = fma(-x0, bn_log(fma(r[p], b0, 1.0), tab), s[p]);
                s[p] = fma(-x1, bn_log(fma(r[p], b1, 1.0), tab), s[p]);
            }
        }
        if (k < nm) {
            const double b0 = __ldg(gs + k), x0 = (double)((int)__ldg(gx + k) + 1);
#pragma unroll
            for (int p = 0; p < NP; p++) s[p] = fma(-x0, bn_log(fma(r[p], b0, 1.0), tab), s[p]);
        }
        if (g.n_big) {
#pragma unroll
            for (int p = 0; p < NP; p++) s[p] += fit2_big(gs, g.n_big, tid, TPG, gm, beta, r[p], phi[p], tab);
        }
    }
    if (g.kmax > 1) {
        double iphi[NP];
#pragma unroll
        for (int p = 0; p < NP; p++) iphi[p] = 1.0 / phi[p];
        for (int q = 1 + tid; q < g.kmax; q += TPG) {
            const double t = (double)tail[q], qd = (double)q;
#pragma unroll
            for (int p = 0; p < NP; p++) s[p] = fma(t, bn_log1p_pos(qd * iphi[p], tab), s[p]);
        }
    }
    // the dense part: every cell, through the table (thread p serves trial size p)
    {
        const bool mine = tid < NP;
        double rr = r[0], ph = phi[0];
#pragma unroll
        for (int p = 1; p < NP; p++)
            if (tid == p) {
                rr = r[p];
                ph = phi[p];
            }
        bool covered = true;
#pragma unroll
        for (int p = 0; p < NP; p++) covered = covered && r[p] >= meta->r_lo && r[p] <= meta->r_hi;
        if (covered) {
            if (mine) {
                const double d = -ph * dt_eval(rr, coef, meta, tab);
#pragma unroll
                for (int p = 0; p < NP; p++)
                    if (tid == p) s[p] += d;
            }
        } else {
#pragma unroll
            for (int p = 0; p < NP; p++) s[p] = fma(-phi[p], fit2_dense_direct(beta, C, tid, TPG, r[p], tab), s[p]);
        }
    }
#pragma unroll
    for (int p = 0; p < NP; p++) s[p] = bn_warp_sum(s[p]);
    if (TPG != 32) {
        const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
        __syncthreads(); // scratch may still be read by the previous call
        if (l == 0) {
#pragma unroll
            for (int p = 0; p < NP; p++) scratch[p * (TPG / 32) + w] = s[p];
        }
        __syncthreads();
#pragma unroll
        for (int p = 0; p < NP; p++) {
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < TPG / 32; k++) t += scratch[p * (TPG / 32) + k];
            s[p] = t;
        }
    }
#pragma unroll
    for (int p = 0; p < NP; p++) val[p] = -(s[p] + g.Kc);
}

// gene lists of the two geometries: warp per gene below FIT2_SPLIT stored entries, CTA per gene from there on
#define FIT2_SPLIT 4096
// One CTA.  Both lists come out longest gene first (128 length classes, a quarter of an octave each; the order inside a class
// is left to the atomics -- it only decides which group fits a gene): the queue then schedules longest-processing-time
// first, so the phase does not end on a 1.3 M-entry gene that happened to sit late in the list (gene blocks of 4125 genes:
// fit 41 - 109 ms over 8 ranks before this).
__global__ void __launch_bounds__(1024) k_fit2_lists(int64_t G, const int64_t *__restrict__ rowptr, int split, int32_t *__restrict__ small,
                                                     int32_t *__restrict__ big, unsigned int *__restrict__ counts)
{
    __shared__ unsigned int hist[2][128], start[2][128];
    for (int q = threadIdx.x; q < 256; q += 1024) (&hist[0][0])[q] = 0;
    __syncthreads();
    auto cls = [](int64_t n) -> int { // descending: class 0 holds the longest genes
        if (n < 4) return 127;
        const int e = 63 - __clzll((long long)n);
        const int c = 4 * e + (int)((n >> (e - 2)) & 3);
        return c > 127 ? 0 : 127 - c;
    };
    for (int64_t g = threadIdx.x; g < G; g += 1024) {
        const int64_t n = rowptr[g + 1] - rowptr[g];
        atomicAdd(&hist[n >= split ? 1 : 0][cls(n)], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 2) {
        unsigned int acc = 0;
        for (int c = 0; c < 128; c++) {
            start[threadIdx.x][c] = acc;
            acc += hist[threadIdx.x][c];
        }
        counts[threadIdx.x] = acc;
    }
    __syncthreads();
    for (int64_t g = threadIdx.x; g < G; g += 1024) {
        const int64_t n = rowptr[g + 1] - rowptr[g];
        const int w = n >= split ? 1 : 0;
        (w ? big : small)[atomicAdd(&start[w][cls(n)], 1u)] = (int32_t)g;
    }
}

// spg asks for -l(p) and then, unless the line search rejects p, for -l(p + eps) (its forward-difference gradient): both are
// evaluated in the same pass over the gene (the entries are read once, the second value costs two more FP64 instructions per
// entry) and the second is handed over only if it is what the solver asks for next -- the sequence of values the solver sees,
// and the count of evaluations it consumed, are those of one-at-a-time evaluation.
template <int TPG, int METHOD>
__global__ void __launch_bounds__(TPG == 32 ? 128 : TPG, TPG == 32 ? 6 : 3)
    k_fit2(const int32_t *__restrict__ list, const unsigned int *__restrict__ count, int64_t C, const int64_t *__restrict__ rowptr,
           const uint64_t *__restrict__ gm, const double *__restrict__ gs, const uint8_t *__restrict__ gx,
           const GeneMeta2 *__restrict__ gmeta, const uint32_t *__restrict__ tail_all, const double *__restrict__ beta,
           const double *__restrict__ mu0,
           const double *__restrict__ size0, const double *__restrict__ params, const double *__restrict__ coef,
           const DenseMeta *__restrict__ meta, FitOpts o, unsigned long long *__restrict__ queue, double *__restrict__ size_out,
           int32_t *__restrict__ conv_out, int32_t *__restrict__ nfe_out)
{
    extern __shared__ __align__(16) double fit2_ring[]; // CTA kernel only (FIT2_RING_BYTES)
    constexpr int GPC = TPG == 32 ? 4 : 1; // genes in flight per CTA
    const double lower = params[0], upper = params[1], beta_max = params[2];
    __shared__ BnLogEntry tab[BN_LOG_TABLE_N];
    __shared__ double scratch[16];
    __shared__ uint32_t tail_sh[GPC][FIT2_HMAX + 1];
    __shared__ unsigned long long next_gene;
    bn_log_table_load(tab);
    __syncthreads();
    const int tid = (TPG == 32) ? (threadIdx.x & 31) : threadIdx.x;
    uint32_t *tl = tail_sh[TPG == 32 ? (threadIdx.x >> 5) : 0];
    const unsigned long long n = *count;
    for (;;) {
        unsigned long long gq;
        if (TPG == 32) {
            if (tid == 0) gq = atomicAdd(queue, 1ULL);
            gq = __shfl_sync(0xffffffffu, gq, 0);
        } else {
            __syncthreads();
            if (tid == 0) next_gene = atomicAdd(queue, 1ULL);
            __syncthreads();
            gq = next_gene;
        }
        if (gq >= n) break;
        const int64_t gi = (int64_t)list[gq];
        const int64_t a = rowptr[gi];
        const GeneMeta2 gmv = gmeta[gi];
        for (int q = tid; q <= FIT2_HMAX; q += TPG) tl[q] = __ldg(tail_all + gi * (int64_t)(FIT2_HMAX + 1) + q);
        if (TPG == 32) __syncwarp();
        else __syncthreads();
        FitGene2 g;
        g.gs = gs + a;
        g.gx = gx + a;
        g.ring = TPG == 32 ? nullptr : fit2_ring;
        g.tail = tl;
        g.mu = mu0[gi];
        g.kmax = gmv.kmax;
        g.n_big = gmv.n_big;
        const double par = size0[gi];
        double res = par;
        int conv = 0, nfe = 0;
        if (!(g.mu > 0.0)) {
            // flat likelihood (mean 0): spg's first projected gradient is 0, it returns the projected start
            res = proj(par, lower, upper);
            if (METHOD == BN_FIT_SPG) nfe = 2;
        } else {
            g.Kc = fma(gmv.X, log(g.mu), gmv.Kbase);
            if (METHOD == BN_FIT_BRENT) {
                BrentState st;
                st.init(lower, upper);
                bool more = true;
                while (more) {
                    double at[1] = {st.next}, v[1];
                    nb_loglik2<TPG, 1>(g, at, v, gm, beta, C, beta_max, coef, meta, tab, scratch, tid);
                    nfe++;
                    more = st.step(v[0], o);
                }
                res = st.next;
                conv = st.conv;
            } else {
                SpgState st;
                st.init(par, lower, upper, o);
                bool more = true;
                while (more) {
                    double at[2] = {st.next, st.next + o.eps}, v[2];
                    nb_loglik2<TPG, 2>(g, at, v, gm, beta, C, beta_max, coef, meta, tab, scratch, tid);
                    nfe++;
                    more = st.step(v[0], o);
                    if (more && !st.next_is_grad && st.next == at[1] && (st.phase == SpgState::INIT_G || st.phase == SpgState::GRAD)) {
                        nfe++;
                        more = st.step(v[1], o);
                    }
                }
                res = st.next;
                conv = st.conv;
            }
        }
        if (tid == 0) {
            size_out[gi] = res;
            if (conv_out) conv_out[gi] = conv;
            if (nfe_out) nfe_out[gi] = nfe;
        }
        if (TPG == 32) __syncwarp(); // the next gene overwrites this warp's tail table
    }
}

// ---------------------------------------------------------------------------------
// Multi-slot variant for long genes (C >= 8192): one 256-thread CTA evaluates FIT_NS genes
// ("slots") per sweep over the cells, each at its own trial size.  Every beta_j fetched from
// L1/L2 feeds FIT_NS likelihoods, which is what bounds the single-slot kernel at 1e5+ cells (each
// evaluation streams the whole beta vector; 33k genes x 26 evaluations x 0.8 MB = 680 GB of L2
// traffic for config 3).  Slots are independent state machines in shared memory: a slot that
// converges pops the next gene from the queue while the others continue, so no slot idles until the
// queue is empty.  After a sweep, thread s steps the solver of slot s (the FIT_NS state machines
// run as lanes of one warp).  A gene's result is a pure function of its own data: summation order,
// fold points and solver path do not depend on which genes share the CTA.
// ---------------------------------------------------------------------------------
#define FIT_NS 4
#define FIT_EVCAP 128      // default number of evaluations after which a gene is handed to the cluster kernel (bn_fit_opts.park_after)
#define FIT_CL 8           // CTAs per cluster (portable maximum)
#define FIT_CL_THREADS 512

// A gene suspended by the sweep kernel: everything needed to resume its solver elsewhere.
struct Straggler {
    int64_t gene;
    int nfe, kmax, n_big, pad;
    double Kc, mu, next;
    SpgState spg;
    BrentState brent;
    uint32_t tail[FitCfg<256>::HMAX + 1];
};

struct FitSlot {
    const uint64_t *nz;
    int64_t nnz, gene;
    double mu, Kc, next;
    int kmax, n_big, nfe, state; // state: 0 idle, 1 fresh (needs prologue), 2 running
    SpgState spg;
    BrentState brent;
    uint32_t tail[FitCfg<256>::HMAX + 1];
};

template <int MINB, int UNR, int NS>
__global__ void __launch_bounds__(256, MINB)
    k_fit_size_ms(int64_t G, int64_t C, const int64_t *__restrict__ rowptr, const uint64_t *__restrict__ gm,
                  const double *__restrict__ beta, const double *__restrict__ logbeta, const double *__restrict__ mu0,
                  const double *__restrict__ size0, const double *__restrict__ params, FitOpts o,
                  unsigned long long *__restrict__ queue, double *__restrict__ size_out, int32_t *__restrict__ conv_out,
                  int32_t *__restrict__ nfe_out, Straggler *__restrict__ strag, unsigned int *__restrict__ n_strag)
{
    const double lower = params[0], upper = params[1], beta_max = params[2]; // device-resident: no host round trip
    constexpr int HMAX = FitCfg<256>::HMAX;
    __shared__ BnLogEntry tab[BN_LOG_TABLE_N];
    __shared__ FitSlot slot[NS];
    __shared__ double scratch[NS][8];
    __shared__ double red8[8];
    __shared__ int xmax_sh, n_running, queue_dry; // queue_dry: stop polling the global queue once it is empty
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bn_log_table_load(tab);
    if (tid < NS) slot[tid].state = 0;
    if (tid == 0) queue_dry = 0;
    __syncthreads();
    for (;;) {
        // ---- refill idle slots (thread 0); all-zero genes are answered on the spot -------------
        if (tid == 0) {
            int running = 0;
            for (int s = 0; s < NS; s++) {
                FitSlot &sl = slot[s];
                while (sl.state == 0 && !queue_dry) {
                    const unsigned long long gq = atomicAdd(queue, 1ULL);
                    if (gq >= (unsigned long long)G) {
                        queue_dry = 1;
                        break;
                    }
                    const int64_t gi = (int64_t)gq, a = rowptr[gi], n = rowptr[gi + 1] - a;
                    const double mu = mu0[gi];
                    if (!(mu > 0.0)) { // flat likelihood (mean 0): spg returns the projected start
                        size_out[gi] = proj(size0[gi], lower, upper);
                        if (conv_out) conv_out[gi] = 0;
                        if (nfe_out) nfe_out[gi] = (o.method == BN_FIT_SPG) ? 2 : 0;
                        continue;
                    }
                    sl.nz = gm + a;
                    sl.nnz = n;
                    sl.gene = gi;
                    sl.mu = mu;
                    sl.nfe = 0;
                    sl.state = 1;
                }
                running += (sl.state != 0);
            }
            n_running = running;
        }
        __syncthreads();
        if (n_running == 0) break;
        // ---- prologue of fresh slots: count histogram -> tail table, Kc, largest count ---------
        for (int s = 0; s < NS; s++) {
            FitSlot &sl = slot[s];
            if (sl.state != 1) continue; // uniform
            for (int q = tid; q <= HMAX; q += 256) sl.tail[q] = 0;
            if (tid == 0) xmax_sh = 0;
            __syncthreads();
            const double lmu = log(sl.mu);
            double kk = 0.0;
            int xm = 0;
            for (int64_t k = tid; k < sl.nnz; k += 256) {
                const uint64_t e0 = __ldg(sl.nz + k);
                const double x = (double)(uint32_t)e0;
                kk += x * (lmu + __ldg(logbeta + (e0 >> 32))) - bn_lfact(x);
                const int xi = x > (double)HMAX ? HMAX + 1 : (int)x;
                if (xi > 0) atomicAdd(&sl.tail[xi - 1], 1u); // an explicitly stored zero has no lgamma term
                xm = max(xm, xi);
            }
            atomicMax(&xmax_sh, xm);
            const double Kc = bn_group_sum<256>(kk, red8);
            __syncthreads();
            if (tid == 0) {
                uint32_t run = sl.tail[HMAX];
                sl.n_big = (int)run;
                for (int q = HMAX - 1; q >= 0; q--) {
                    run += sl.tail[q];
                    sl.tail[q] = run;
                }
                sl.Kc = Kc;
                sl.kmax = min(HMAX, xmax_sh);
                if (o.method == BN_FIT_BRENT) {
                    sl.brent.init(lower, upper);
                    sl.next = sl.brent.next;
                } else {
                    sl.spg.init(size0[sl.gene], lower, upper, o);
                    sl.next = sl.spg.next;
                }
                sl.state = 2;
            }
            __syncthreads();
        }
        // ---- one sweep: l(next) of every running slot ------------------------------------------
        double r[NS], pa[NS], pb[NS], acc[NS];
        bool any_slow = false;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            r[s] = (slot[s].state == 2) ? slot[s].mu / slot[s].next : 0.0;
            pa[s] = 1.0;
            pb[s] = 1.0;
            acc[s] = 0.0;
            any_slow |= !(fma(r[s], beta_max, 1.0) < 2.0e9); // (1 + r beta_max)^32 must stay finite (remainder loop: < 32 factors)
        }
        if (!any_slow) {
            // blocks of 16 x 512 cells: 32 factors per slot between folds, no branch inside the block
            int64_t j = tid;
            for (; j + 15 * 512 + 256 < C; j += 16 * 512) {
#pragma unroll 4
                for (int it = 0; it < 16; it++) {
                    const double b0 = __ldg(beta + j + it * 512), b1 = __ldg(beta + j + it * 512 + 256);
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        pa[s] *= fma(r[s], b0, 1.0);
                        pb[s] *= fma(r[s], b1, 1.0);
                    }
                }
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    acc[s] += bn_log(pa[s], tab) + bn_log(pb[s], tab);
                    pa[s] = 1.0;
                    pb[s] = 1.0;
                }
            }
            for (; j < C; j += 256) { // fewer than 32 factors left per thread
                const double b0 = __ldg(beta + j);
#pragma unroll
                for (int s = 0; s < NS; s++) pa[s] *= fma(r[s], b0, 1.0);
            }
        } else { // absurd mu/size ratios: one logarithm per cell keeps everything finite
            for (int64_t j = tid; j < C; j += 256) {
                const double b0 = __ldg(beta + j);
#pragma unroll
                for (int s = 0; s < NS; s++) acc[s] += bn_log(fma(r[s], b0, 1.0), tab);
            }
        }
#pragma unroll
        for (int s = 0; s < NS; s++) acc[s] += bn_log(pa[s], tab) + bn_log(pb[s], tab);
#pragma unroll 1
        for (int s = 0; s < NS; s++) {
            const FitSlot &sl = slot[s];
            if (sl.state != 2) continue; // uniform
            const double phi = sl.next, mu = sl.mu;
            double t = -phi * acc[s];
            double s0 = 0.0, s1 = 0.0;
            const uint64_t *__restrict__ nz = sl.nz;
            const int64_t nnz = sl.nnz;
            int64_t k = tid;
            for (; k + (UNR - 1) * 256 < nnz; k += UNR * 256) {
                uint64_t e[UNR];
                double b[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) e[u] = __ldg(nz + k + u * 256);
#pragma unroll
                for (int u = 0; u < UNR; u++) b[u] = __ldg(beta + (e[u] >> 32));
#pragma unroll
                for (int u = 0; u < UNR; u += 2) {
                    s0 = fma(-(double)(uint32_t)e[u], bn_log(fma(mu, b[u], phi), tab), s0);
                    s1 = fma(-(double)(uint32_t)e[u + 1], bn_log(fma(mu, b[u + 1], phi), tab), s1);
                }
            }
            for (; k < nnz; k += 256) {
                const uint64_t e0 = __ldg(nz + k);
                s0 = fma(-(double)(uint32_t)e0, bn_log(fma(mu, __ldg(beta + (e0 >> 32)), phi), tab), s0);
            }
            if (sl.n_big != 0) {
                const double lgref = bn_lgamma_stirling(phi + (double)HMAX, tab);
                for (k = tid; k < nnz; k += 256) {
                    const double x = (double)(uint32_t)__ldg(nz + k);
                    if (x > (double)HMAX) s1 += bn_lgamma_stirling(x + phi, tab) - lgref;
                }
            }
            for (int q = tid; q < sl.kmax; q += 256) s1 = fma((double)sl.tail[q], bn_log(phi + (double)q, tab), s1);
            t += s0 + s1;
            t = bn_warp_sum(t);
            if (lane == 0) scratch[s][warp] = t;
        }
        __syncthreads();
        // ---- thread s steps the solver of slot s ------------------------------------------------
        if (tid < NS && slot[tid].state == 2) {
            FitSlot &sl = slot[tid];
            double tot = 0.0;
#pragma unroll
            for (int w = 0; w < 8; w++) tot += scratch[tid][w];
            const double val = -(tot + sl.Kc);
            sl.nfe++;
            bool more;
            int conv;
            if (o.method == BN_FIT_BRENT) {
                more = sl.brent.step(val, o);
                sl.next = sl.brent.next;
                conv = sl.brent.conv;
            } else {
                more = sl.spg.step(val, o);
                sl.next = sl.spg.next;
                conv = sl.spg.conv;
            }
            if (!more) {
                size_out[sl.gene] = sl.next;
                if (conv_out) conv_out[sl.gene] = conv;
                if (nfe_out) nfe_out[sl.gene] = sl.nfe;
                sl.state = 0;
            } else if (strag && sl.nfe >= o.evcap) {
                // a straggler (BB::spg may run to maxit = 1500 iterations): one CTA would hold the tail of
                // the whole launch for milliseconds, so park the solver state for the cluster kernel
                Straggler &g = strag[atomicAdd(n_strag, 1u)];
                g.gene = sl.gene;
                g.nfe = sl.nfe;
                g.kmax = sl.kmax;
                g.n_big = sl.n_big;
                g.Kc = sl.Kc;
                g.mu = sl.mu;
                g.next = sl.next;
                g.spg = sl.spg;
                g.brent = sl.brent;
                for (int q = 0; q <= HMAX; q++) g.tail[q] = sl.tail[q];
                sl.state = 0;
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------
// Stragglers: one thread-block cluster (8 CTAs x 512 threads on neighbouring SMs) per gene.
// Every evaluation is split over the 4096 threads of the cluster; the 8 CTA partial sums travel
// to the leader CTA through distributed shared memory, the leader steps the solver and the next
// trial size is read back by all CTAs the same way: two cluster barriers per evaluation instead
// of a 256-thread CTA walking all cells alone.  Results stay deterministic (fixed cell -> thread
// map and summation order; which genes are parked depends only on their own evaluation count).
// ---------------------------------------------------------------------------------
__global__ void __cluster_dims__(FIT_CL, 1, 1) __launch_bounds__(FIT_CL_THREADS, 1)
    k_fit_cluster(int64_t C, const int64_t *__restrict__ rowptr, const uint64_t *__restrict__ gm,
                  const double *__restrict__ beta, FitOpts o, const Straggler *__restrict__ strag,
                  const unsigned int *__restrict__ n_strag, double *__restrict__ size_out, int32_t *__restrict__ conv_out,
                  int32_t *__restrict__ nfe_out)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    constexpr int HMAX = FitCfg<256>::HMAX;
    constexpr int NT = FIT_CL * FIT_CL_THREADS;
    __shared__ BnLogEntry tab[BN_LOG_TABLE_N];
    __shared__ Straggler st;             // leader's copy is the live solver state
    __shared__ double part[FIT_CL];      // leader: partial sums of the 8 CTAs
    __shared__ double warp_part[FIT_CL_THREADS / 32];
    __shared__ double next_phi;          // leader: trial size of the coming evaluation (< 0: gene finished)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned rank = cluster.block_rank();
    const int64_t t = (int64_t)rank * FIT_CL_THREADS + tid;
    double *lead_part = cluster.map_shared_rank(part, 0);
    const double *lead_next = cluster.map_shared_rank(&next_phi, 0);
    bn_log_table_load(tab);
    const unsigned n = *n_strag;
    const unsigned ncluster = gridDim.x / FIT_CL;
    for (unsigned w = blockIdx.x / FIT_CL; w < n; w += ncluster) {
        // every CTA keeps a read-only copy of the gene's constants; the leader's copy also holds the solver
        {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(strag + w);
            uint32_t *dst = reinterpret_cast<uint32_t *>(&st);
            for (int q = tid; q < (int)(sizeof(Straggler) / 4); q += FIT_CL_THREADS) dst[q] = src[q];
        }
        __syncthreads();
        if (rank == 0 && tid == 0) next_phi = st.next;
        const int64_t a = rowptr[st.gene];
        const uint64_t *__restrict__ nz = gm + a;
        const int64_t nnz = rowptr[st.gene + 1] - a;
        const double mu = st.mu;
        cluster.sync();
        for (;;) {
            const double phi = *lead_next;
            if (!(phi > 0.0)) break; // finished (sizes are positive: the box has lower > 0)
            const double r = mu / phi;
            double acc = 0.0, pa = 1.0, pb = 1.0;
            if (fma(r, 1.0, 1.0) < 2.0e9) { // (1 + r)^32 stays finite: the remainder loop multiplies up to 31 factors
                int64_t j = t;
                for (; j + 15 * 2 * NT + NT < C; j += 16 * 2 * NT) {
#pragma unroll 4
                    for (int it = 0; it < 16; it++) {
                        pa *= fma(r, __ldg(beta + j + it * 2 * NT), 1.0);
                        pb *= fma(r, __ldg(beta + j + it * 2 * NT + NT), 1.0);
                    }
                    acc += bn_log(pa, tab) + bn_log(pb, tab);
                    pa = 1.0;
                    pb = 1.0;
                }
                for (; j < C; j += NT) pa *= fma(r, __ldg(beta + j), 1.0); // < 32 factors
            } else {
                for (int64_t j = t; j < C; j += NT) acc += bn_log(fma(r, __ldg(beta + j), 1.0), tab);
            }
            acc += bn_log(pa, tab) + bn_log(pb, tab);
            double s0 = 0.0, s1 = 0.0;
            for (int64_t k = t; k < nnz; k += NT) {
                const uint64_t e0 = __ldg(nz + k);
                const double x = (double)(uint32_t)e0;
                s0 = fma(-x, bn_log(fma(mu, __ldg(beta + (e0 >> 32)), phi), tab), s0);
                if (x > (double)HMAX) s1 += bn_lgamma_stirling(x + phi, tab) - bn_lgamma_stirling(phi + (double)HMAX, tab);
            }
            if (t < st.kmax) s1 = fma((double)st.tail[t], bn_log(phi + (double)t, tab), s1);
            double tot = bn_warp_sum(-phi * acc + (s0 + s1));
            if (lane == 0) warp_part[warp] = tot;
            __syncthreads();
            if (tid == 0) {
                double c = 0.0;
#pragma unroll
                for (int q = 0; q < FIT_CL_THREADS / 32; q++) c += warp_part[q];
                lead_part[rank] = c; // distributed shared memory store into the leader CTA
            }
            cluster.sync();
            if (rank == 0 && tid == 0) {
                double f = 0.0;
#pragma unroll
                for (int q = 0; q < FIT_CL; q++) f += part[q];
                const double val = -(f + st.Kc);
                st.nfe++;
                bool more;
                int conv;
                if (o.method == BN_FIT_BRENT) {
                    more = st.brent.step(val, o);
                    st.next = st.brent.next;
                    conv = st.brent.conv;
                } else {
                    more = st.spg.step(val, o);
                    st.next = st.spg.next;
                    conv = st.spg.conv;
                }
                if (!more) {
                    size_out[st.gene] = st.next;
                    if (conv_out) conv_out[st.gene] = conv;
                    if (nfe_out) nfe_out[st.gene] = st.nfe;
                    next_phi = -1.0;
                } else
                    next_phi = st.next;
            }
            cluster.sync();
        }
        cluster.sync(); // nobody may still be reading the leader's next_phi when the next gene is loaded
    }
}

// ---------------------------------------------------------------------------------
// FIX_MU = FALSE (R/PRIOR_FUNCTIONS.R:414-419, :457-470): BB::spg over par = (size, mu) in the box
// [SIZE_lower, SIZE_upper] x [MU_lower, MU_upper], objective MarginalF_NB_2D
// (src/bayNorm_main.cpp:948-965), gradient by forward differences (eps = 1e-7 per coordinate) or
// GradientFun_NB_2D (:993-1013) when GR = TRUE.  Spg2State is orc_spg (oracle/baynorm_oracle.c) for
// n = 2 as a resumable state machine; one 256-thread CTA works on one gene at a time.
// The likelihood is the same expression as in the 1-D fit with mu a variable: the only mu-dependent
// constant is Kc = X log mu + K0, X = sum x, K0 = sum x log beta_j - sum lgamma(x+1).
// ---------------------------------------------------------------------------------
struct Spg2State {
    enum { INIT_F = 0, INIT_G = 1, LS = 2, GRAD = 3 };
    double lo[2], hi[2], par[2], g[2], pnew[2], gnew[2], d[2], pbest[2];
    double lastfv[16];
    double f, fbest, fchg, pginfn, lambda, gtd, fmaxv, alpha;
    int phase, gi, iter, feval, lsflag, M;
    bool ls_first;
    // request: evaluate at next[]; kind 0 = -l, 1 = -dl/dsize, 2 = -dl/dmu
    double next[2];
    int kind, conv;

    __device__ void init(const double *par0, const double *lower, const double *upper, const FitOpts &o)
    {
        M = o.M > 16 ? 16 : (o.M < 1 ? 1 : o.M);
#pragma unroll
        for (int i = 0; i < 16; i++) lastfv[i] = -1e99;
        for (int i = 0; i < 2; i++) {
            lo[i] = lower[i];
            hi[i] = upper[i];
            par[i] = proj(par0[i], lo[i], hi[i]);
            pbest[i] = par[i];
            next[i] = par[i];
        }
        iter = feval = 0;
        lsflag = -1;
        fchg = INFINITY;
        lambda = 1.0;
        phase = INIT_F;
        kind = 0;
        conv = 0;
    }
    // ask for gradient component gi at point p (held in `at`)
    __device__ __forceinline__ void ask_grad(const double *at, const FitOpts &o)
    {
        next[0] = at[0];
        next[1] = at[1];
        if (o.use_gradient) kind = 1 + gi;
        else {
            kind = 0;
            next[gi] = at[gi] + o.eps;
        }
    }
    __device__ bool finish()
    {
        if (lsflag <= 0) conv = 0;
        else conv = (lsflag == 1) ? 3 : (lsflag == 2) ? 2 : (lsflag == 3) ? 4 : 5;
        next[0] = pbest[0];
        next[1] = pbest[1];
        return false;
    }
    __device__ void proj_grad_norm()
    {
        pginfn = 0.0;
        for (int i = 0; i < 2; i++) pginfn = fmax(pginfn, fabs(proj(par[i] - g[i], lo[i], hi[i]) - par[i]));
    }
    __device__ bool enter_iter(const FitOpts &o)
    {
        if (!(pginfn > o.gtol && iter <= o.maxit && fchg > o.ftol)) {
            const bool r = finish();
            if (lsflag <= 0 && iter >= o.maxit) conv = 1;
            return r;
        }
        iter++;
        gtd = 0.0;
        for (int i = 0; i < 2; i++) {
            d[i] = proj(par[i] - lambda * g[i], lo[i], hi[i]) - par[i];
            gtd += g[i] * d[i];
        }
        if (isinf(gtd)) {
            lsflag = 4;
            return finish();
        }
        fmaxv = lastfv[0];
#pragma unroll
        for (int i = 1; i < 16; i++)
            if (i < M && lastfv[i] > fmaxv) fmaxv = lastfv[i];
        alpha = 1.0;
        for (int i = 0; i < 2; i++) pnew[i] = next[i] = par[i] + alpha * d[i];
        phase = LS;
        ls_first = true;
        kind = 0;
        return true;
    }
    __device__ bool step(double v, const FitOpts &o)
    {
        const double lmin = 1e-30, lmax = 1e30;
        switch (phase) {
        case INIT_F:
            f = v;
            feval++;
            if (f != f || isinf(f)) {
                conv = 3;
                next[0] = par[0];
                next[1] = par[1];
                return false;
            }
            fbest = f;
            phase = INIT_G;
            gi = 0;
            ask_grad(par, o);
            return true;
        case INIT_G:
            g[gi] = o.use_gradient ? v : (v - f) / o.eps;
            if (++gi < 2) {
                ask_grad(par, o);
                return true;
            }
            if (g[0] != g[0] || g[1] != g[1]) {
                conv = 4;
                next[0] = par[0];
                next[1] = par[1];
                return false;
            }
            lastfv[0] = f;
            proj_grad_norm();
            if (pginfn != 0.0) lambda = fmin(lmax, fmax(lmin, 1.0 / pginfn));
            return enter_iter(o);
        case LS: {
            const double fnew = v;
            feval++;
            lsflag = 0;
            if (fnew != fnew) {
                lsflag = 1;
                return finish();
            }
            if (!ls_first && feval > o.maxfeval) {
                lsflag = 2;
                return finish();
            }
            ls_first = false;
            if (fnew > fmaxv + 1e-4 * alpha * gtd) { // nonmonotone Armijo test failed: backtrack
                if (alpha <= 0.1) alpha = alpha / 2;
                else {
                    double atemp = -(gtd * alpha * alpha) / (2 * (fnew - f - alpha * gtd));
                    if (atemp < 0.1 || atemp > 0.9 * alpha) atemp = alpha / 2;
                    alpha = atemp;
                }
                for (int i = 0; i < 2; i++) pnew[i] = next[i] = par[i] + alpha * d[i];
                kind = 0;
                return true;
            }
            fchg = fabs(f - fnew);
            f = fnew;
            const int slot = iter % M;
#pragma unroll
            for (int i = 0; i < 16; i++)
                if (i == slot) lastfv[i] = f;
            phase = GRAD;
            gi = 0;
            ask_grad(pnew, o);
            return true;
        }
        default: { // GRAD
            gnew[gi] = o.use_gradient ? v : (v - f) / o.eps;
            if (++gi < 2) {
                ask_grad(pnew, o);
                return true;
            }
            if (gnew[0] != gnew[0] || gnew[1] != gnew[1]) {
                lsflag = 3;
                return finish();
            }
            double sts = 0.0, yty = 0.0;
            for (int i = 0; i < 2; i++) {
                const double sv = pnew[i] - par[i], yv = gnew[i] - g[i];
                sts += sv * sv;
                yty += yv * yv;
                par[i] = pnew[i];
                g[i] = gnew[i];
            }
            lambda = (sts == 0.0 || yty == 0.0) ? lmax : fmin(lmax, fmax(lmin, sqrt(sts / yty)));
            proj_grad_norm();
            if (f < fbest) {
                fbest = f;
                pbest[0] = pnew[0];
                pbest[1] = pnew[1];
            }
            return enter_iter(o);
        }
        }
    }
};

// d l / d mu = sum_j size (x_j - mu beta_j) / (mu (mu beta_j + size))       (GradientFun_NB_2D, :1004)
template <int TPG> __device__ double nb_loglik_grad_mu(const FitGene &g, double phi, double *scratch, int tid)
{
    const double mu = g.mu;
    double acc = 0.0;
    for (int64_t j = tid; j < g.C; j += TPG) {
        const double mb = mu * __ldg(g.beta + j);
        acc += (0.0 * phi - mb * phi) / (mu * (mb + phi)); // the x = 0 term of every cell
    }
    for (int64_t k = tid; k < g.nnz; k += TPG) {
        const uint64_t e0 = __ldg(g.nz + k);
        const double x = (double)(uint32_t)e0;
        const double mb = mu * __ldg(g.beta + (e0 >> 32));
        acc += (x * phi) / (mu * (mb + phi)); // (x phi - mb phi) minus the x = 0 term already counted
    }
    return bn_group_sum<TPG>(acc, scratch);
}

template <bool GRAD>
__global__ void __launch_bounds__(256, 4)
    k_fit_2d(int64_t G, int64_t C, const int64_t *__restrict__ rowptr, const uint64_t *__restrict__ gm,
             const double *__restrict__ beta, const double *__restrict__ logbeta, const double *__restrict__ mu0,
             const double *__restrict__ size0, const double *__restrict__ box /* lo_size, lo_mu, hi_size, hi_mu, beta_max */,
             FitOpts o, unsigned long long *__restrict__ queue, double *__restrict__ size_out, double *__restrict__ mu_out,
             int32_t *__restrict__ conv_out, int32_t *__restrict__ nfe_out)
{
    constexpr int TPG = 256;
    constexpr int HMAX = FitCfg<TPG>::HMAX;
    __shared__ BnLogEntry tab[BN_LOG_TABLE_N];
    __shared__ double scratch[8];
    __shared__ unsigned long long next_gene;
    __shared__ uint32_t hist[HMAX + 1];
    __shared__ int xmax_sh;
    bn_log_table_load(tab);
    __syncthreads();
    const int tid = threadIdx.x;
    const double lower[2] = {box[0], box[1]}, upper[2] = {box[2], box[3]};
    for (;;) {
        __syncthreads();
        if (tid == 0) next_gene = atomicAdd(queue, 1ULL);
        __syncthreads();
        const unsigned long long gq = next_gene;
        if (gq >= (unsigned long long)G) break;
        const int64_t gi = (int64_t)gq;
        FitGene g;
        const int64_t a = rowptr[gi];
        g.nz = gm + a;
        g.nnz = rowptr[gi + 1] - a;
        g.beta = beta;
        g.logbeta = logbeta;
        g.C = C;
        g.beta_max = box[4];
        g.tail = hist;
        // per-gene constants: count histogram -> tail table, X = sum x, K0 = sum x log beta - sum lgamma(x+1)
        for (int q = tid; q <= HMAX; q += TPG) hist[q] = 0;
        if (tid == 0) xmax_sh = 0;
        __syncthreads();
        double k0 = 0.0, xs = 0.0;
        int xm = 0;
        for (int64_t k = tid; k < g.nnz; k += TPG) {
            const uint64_t e0 = __ldg(g.nz + k);
            const double x = (double)(uint32_t)e0;
            if (x > 0.0) { // explicit zeros carry no term
                k0 += x * __ldg(logbeta + (e0 >> 32)) - bn_lfact(x);
                xs += x;
                const int xi = x > (double)HMAX ? HMAX + 1 : (int)x;
                atomicAdd(&hist[xi - 1], 1u);
                xm = max(xm, xi);
            }
        }
        atomicMax(&xmax_sh, xm);
        const double K0 = bn_group_sum<TPG>(k0, scratch);
        const double X = bn_group_sum<TPG>(xs, scratch);
        __syncthreads();
        if (tid == 0) { // suffix sum: tail[k] = #{x > k}
            uint32_t run = hist[HMAX];
            for (int q = HMAX - 1; q >= 0; q--) {
                run += hist[q];
                hist[q] = run;
            }
        }
        __syncthreads();
        g.n_big = (int)(hist[HMAX]);
        g.kmax = min(HMAX, xmax_sh);
        Spg2State st;
        const double par0[2] = {size0[gi], mu0[gi]};
        st.init(par0, lower, upper, o);
        int nfe = 0;
        bool more = true;
        while (more) {
            const double phi = st.next[0];
            g.mu = st.next[1];
            double val;
            if (GRAD && st.kind == 1) val = -nb_loglik_grad<TPG>(g, phi, scratch, tid);
            else if (GRAD && st.kind == 2) val = -nb_loglik_grad_mu<TPG>(g, phi, scratch, tid);
            else {
                // X log mu: dnbinom_mu(x > 0, mu = 0) = 0 gives l = -inf, which spg's line search backs away from
                g.Kc = (X > 0.0 ? X * log(g.mu) : 0.0) + K0;
                val = -nb_loglik<TPG>(g, phi, tab, scratch, tid);
            }
            nfe++;
            more = st.step(val, o);
        }
        __syncthreads(); // hist is reused by the next gene
        if (tid == 0) {
            size_out[gi] = st.next[0];
            mu_out[gi] = st.next[1];
            if (conv_out) conv_out[gi] = st.conv;
            if (nfe_out) nfe_out[gi] = nfe;
        }
    }
}

__global__ void k_fit_box(double lo_size, double lo_mu, double hi_size, double hi_mu, double *__restrict__ box)
{
    box[0] = lo_size;
    box[1] = lo_mu;
    box[2] = hi_size;
    box[3] = hi_mu;
}

// largest beta (decides between the series and the table form of log1p per evaluation)
__global__ void __launch_bounds__(1024) k_max_vec(int64_t n, const double *__restrict__ v, double *__restrict__ out)
{
    __shared__ double sh[32];
    double m = 0.0;
    for (int64_t k = threadIdx.x; k < n; k += blockDim.x) m = fmax(m, v[k]);
    m = bn_warp_max(m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; k++) m = fmax(m, sh[k]);
        out[0] = m;
    }
}

__global__ void k_log_vec(int64_t n, const double *__restrict__ in, double *__restrict__ out)
{
    const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (k < n) out[k] = log(in[k]);
}

extern "C" void bn_fit_opts_default(bn_fit_opts *o)
{
    if (!o) return;
    o->method = BN_FIT_SPG;
    o->maxfeval = 500;
    o->maxit = 1500;
    o->M = 10;
    o->ftol = 1e-10;
    o->gtol = 1e-5;
    o->eps = 1e-7;
    o->xatol = 1e-10;
    o->use_gradient = 0;
    o->park_after = 0;
}

// params[0..1] = box: either the host scalars, or (from_mm) derived on the device from
// mm = {min, -max} of the MME sizes: lower = min, upper = ceiling(max)  (R/PRIOR_FUNCTIONS.R:277-279)
__global__ void k_fit_params(double lower, double upper, const double *__restrict__ mm, double *__restrict__ params)
{
    params[0] = mm ? mm[0] : lower;
    params[1] = mm ? ceil(-mm[1]) : upper;
}

// device-pointer core shared by bn_fit_size and bn_prior.  The box is (lower, upper) or, when d_mm
// is given, read from device memory so that bn_prior never waits for the host between stages.
int bn_fit_size_device(bn_ctx *ctx, bn_mat *m, const double *d_beta, const double *d_mu0, const double *d_size0,
                       double lower, double upper, const double *d_mm, const bn_fit_opts *opts, double *d_size_out,
                       int32_t *d_conv, int32_t *d_nfe)
{
    if (!m->is_count)
        return bn_fail(ctx, BN_ERR_NONCOUNT, "the likelihood fit needs non-negative integer counts (dnbinom_mu is 0 elsewhere)");
    if (!d_mm && (!(lower > 0.0) || !(upper >= lower))) return bn_fail(ctx, BN_ERR_INVALID, "bad size bounds [%g, %g]", lower, upper);
    bn_fit_opts od;
    bn_fit_opts_default(&od);
    if (opts) od = *opts;
    FitOpts o;
    o.method = od.method;
    o.maxfeval = od.maxfeval;
    o.maxit = od.maxit;
    o.M = od.M;
    o.use_gradient = od.use_gradient;
    o.ftol = od.ftol;
    o.gtol = od.gtol;
    o.eps = od.eps;
    o.xatol = od.xatol;
    o.evcap = od.park_after > 0 ? od.park_after : FIT_EVCAP;
    BN_TRY(bn_mat_ensure_csr(ctx, m));
    DevBuf<double> lb;
    BN_TRY(lb.alloc(ctx, (size_t)m->C));
    BN_LAUNCH(ctx, k_log_vec, (unsigned)((m->C + 255) / 256), 256, 0, m->C, d_beta, lb.p);
    DevBuf<unsigned long long> q;
    BN_TRY(q.alloc(ctx, 1));
    BN_CUDA(ctx, cudaMemsetAsync(q.p, 0, 8, ctx->stream));
    DevBuf<double> params;
    BN_TRY(params.alloc(ctx, 4));
    BN_LAUNCH(ctx, k_fit_params, 1, 1, 0, lower, upper, d_mm, params.p);
    BN_LAUNCH(ctx, k_max_vec, 1, 1024, 0, m->C, d_beta, params.p + 2);
    const int sms = ctx->prop.multiProcessorCount;
    const bool grad = o.use_gradient != 0;
    *ctx->fit_parked = 0;
    ctx->last_fit[2] = o.evcap;
    ctx->last_fit[3] = grad ? 1 : 0;
#define FIT_ARGS m->G, m->C, m->rowptr, m->gm, d_beta, lb.p, d_mu0, d_size0, params.p, o, q.p, d_size_out, d_conv, d_nfe
    const bool sparse_mat = (double)m->nnz <= 0.2 * (double)m->G * (double)m->C;
    if (!grad && (ctx->tune[BN_TUNE_FIT_VARIANT] == 5 || (ctx->tune[BN_TUNE_FIT_VARIANT] == 0 && sparse_mat))) {
        // version 2 (see k_fit2): dense part from the per-call table, non-zeros staged by count.  Measured against the
        // first-generation kernels: 33k x 100k at 8 % non-zeros 37 -> 14.5 ms, 33k x 1.3M 598 -> 205 ms; on matrices with
        // >= 20 % non-zeros and counts in the hundreds (config 2) most entries take the one-logarithm path and the
        // first-generation sweep kernel is faster (7.5 against 9.8 ms): those keep it.
        ctx->last_fit[0] = 5;
        const int64_t G = m->G, nnz = m->nnz;
        struct { double *p; } gs, dt;
        struct { uint8_t *p; } gx;
        struct { uint32_t *p; } tails;
        struct { GeneMeta2 *p; } gmeta;
        struct { int32_t *p; } lists;
        constexpr int NSPLIT = 6;
        const size_t dt_doubles = (size_t)DT_NI * DT_STRIDE + (size_t)DT_NI * NSPLIT * DT_N + 8;
        BN_TRY(bn_ctx_scratch(ctx, 14, (size_t)std::max<int64_t>(nnz, 1) * 8, (void **)&gs.p));
        BN_TRY(bn_ctx_scratch(ctx, 19, (size_t)std::max<int64_t>(nnz, 1), (void **)&gx.p));
        BN_TRY(bn_ctx_scratch(ctx, 15, (size_t)G * (FIT2_HMAX + 1) * 4, (void **)&tails.p));
        BN_TRY(bn_ctx_scratch(ctx, 16, (size_t)G * sizeof(GeneMeta2), (void **)&gmeta.p));
        BN_TRY(bn_ctx_scratch(ctx, 17, (size_t)G * 2 * 4 + 64, (void **)&lists.p));
        BN_TRY(bn_ctx_scratch(ctx, 18, dt_doubles * 8, (void **)&dt.p));
        double *coef = dt.p, *part = dt.p + (size_t)DT_NI * DT_STRIDE;
        DenseMeta *meta = reinterpret_cast<DenseMeta *>(part + (size_t)DT_NI * NSPLIT * DT_N);
        DevBuf<unsigned long long> q2; // queues: staging, warp-per-gene fit, CTA-per-gene fit
        DevBuf<unsigned int> cnt2;
        BN_TRY(q2.alloc(ctx, 4));
        BN_TRY(cnt2.alloc(ctx, 2));
        BN_CUDA(ctx, cudaMemsetAsync(q2.p, 0, 32, ctx->stream));
        BN_CUDA(ctx, cudaMemsetAsync(cnt2.p, 0, 8, ctx->stream));
        BN_LAUNCH(ctx, k_dt_range, 1, 1024, 0, G, d_mu0, params.p, 0.0, 0.0, meta);
        BN_LAUNCH(ctx, k_dt_nodes, dim3(DT_NI, NSPLIT), 256, 0, m->C, d_beta, meta, NSPLIT, part);
        BN_LAUNCH(ctx, k_dt_coef, DT_NI, 32, 0, meta, NSPLIT, part, coef);
        BN_LAUNCH(ctx, k_fit2_stage, (int)std::min<int64_t>(G, (int64_t)sms * 8), 256, 0, G, m->rowptr, m->gm, d_beta, lb.p, q2.p, gs.p, gx.p, gmeta.p,
                  tails.p);
        int32_t *l_small = lists.p, *l_big = lists.p + G;
        const int split = ctx->tune[BN_TUNE_FIT_SPLIT] > 0 ? ctx->tune[BN_TUNE_FIT_SPLIT] : FIT2_SPLIT;
        BN_LAUNCH(ctx, k_fit2_lists, 1, 1024, 0, G, m->rowptr, split, l_small, l_big, cnt2.p);
#define FIT2_ARGS(L, K, Q) L, cnt2.p + K, m->C, m->rowptr, m->gm, gs.p, gx.p, gmeta.p, tails.p, d_beta, d_mu0, d_size0, params.p, coef, meta, o, q2.p + Q, \
    d_size_out, d_conv, d_nfe
        // the long genes first: their tail is what ends the phase
        const int grid_big = (int)std::min<int64_t>(G, (int64_t)sms * 3), grid_small = (int)std::min<int64_t>((G + 3) / 4, (int64_t)sms * 6);
        if (!ctx->fit2_attr_set) { // function attributes are per device: once per context
            BN_CUDA(ctx, cudaFuncSetAttribute(k_fit2<256, BN_FIT_BRENT>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIT2_RING_BYTES));
            BN_CUDA(ctx, cudaFuncSetAttribute(k_fit2<256, BN_FIT_SPG>, cudaFuncAttributeMaxDynamicSharedMemorySize, FIT2_RING_BYTES));
            ctx->fit2_attr_set = true;
        }
        if (o.method == BN_FIT_BRENT) {
            BN_LAUNCH(ctx, (k_fit2<256, BN_FIT_BRENT>), grid_big, 256, FIT2_RING_BYTES, FIT2_ARGS(l_big, 1, 2));
            BN_LAUNCH(ctx, (k_fit2<32, BN_FIT_BRENT>), grid_small, 128, 0, FIT2_ARGS(l_small, 0, 1));
        } else {
            BN_LAUNCH(ctx, (k_fit2<256, BN_FIT_SPG>), grid_big, 256, FIT2_RING_BYTES, FIT2_ARGS(l_big, 1, 2));
            BN_LAUNCH(ctx, (k_fit2<32, BN_FIT_SPG>), grid_small, 128, 0, FIT2_ARGS(l_small, 0, 1));
        }
#undef FIT2_ARGS
    } else if (m->C >= 8192) {
        // long genes: a whole CTA per gene keeps every gene's evaluations short and the tail small
        const int grid = (int)std::min<int64_t>(m->G, (int64_t)sms * 12);
        const int force = ctx->tune[BN_TUNE_FIT_VARIANT]; // A/B switch of profiling runs (bn_debug_tune)
        ctx->last_fit[0] = 1;
        if (grad) BN_LAUNCH(ctx, (k_fit_size<256, true>), grid, 256, 0, FIT_ARGS);
        else if (force == 1) BN_LAUNCH(ctx, (k_fit_size<256, false>), grid, 256, 0, FIT_ARGS);
        else {
            // sweep kernel, then the cluster kernel for the genes it parked (usually none or a handful)
            DevBuf<Straggler> strag;
            DevBuf<unsigned int> nstrag;
            BN_TRY(strag.alloc(ctx, (size_t)m->G));
            BN_TRY(nstrag.alloc(ctx, 1));
            BN_CUDA(ctx, cudaMemsetAsync(nstrag.p, 0, 4, ctx->stream));
            // Sparse single-cell matrices: 4 genes per sweep at 4 resident CTAs per SM (64 registers; measured 5-9 % faster
            // than 3 x 78 registers, 2 x 112 is 25 % slower).  Dense matrices (>= 20 % non-zero), where the per-gene
            // non-zero loop dominates and BETA re-use across genes matters less: 2 genes per sweep at 6 CTAs per SM
            // (40 registers) is 3.5 % faster; 3 x 5 and 6 x 3 are slower on both.
            const bool dense = force == 2 || ((force == 0 || force == 4) && (double)m->nnz > 0.2 * (double)m->G * (double)m->C);
            ctx->last_fit[0] = dense ? 2 : 3;
            if (dense)
                BN_LAUNCH(ctx, (k_fit_size_ms<6, 4, 2>), (int)std::min<int64_t>((m->G + 1) / 2, (int64_t)sms * 6), 256, 0, FIT_ARGS,
                          strag.p, nstrag.p);
            else
                BN_LAUNCH(ctx, (k_fit_size_ms<4, 4, 4>), (int)std::min<int64_t>((m->G + 3) / 4, (int64_t)sms * 4), 256, 0, FIT_ARGS,
                          strag.p, nstrag.p);
            const int nclusters = std::max(1, sms / FIT_CL - 2);
            BN_LAUNCH(ctx, k_fit_cluster, nclusters * FIT_CL, FIT_CL_THREADS, 0, m->C, m->rowptr, m->gm, d_beta, o, strag.p,
                      nstrag.p, d_size_out, d_conv, d_nfe);
            BN_CUDA(ctx, cudaMemcpyAsync(ctx->fit_parked, nstrag.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        }
    } else {
        ctx->last_fit[0] = 0;
        const int grid = (int)std::min<int64_t>((m->G + 3) / 4, (int64_t)sms * 24);
        if (grad) BN_LAUNCH(ctx, (k_fit_size<32, true>), grid, 128, 0, FIT_ARGS);
        else BN_LAUNCH(ctx, (k_fit_size<32, false>), grid, 128, 0, FIT_ARGS);
    }
#undef FIT_ARGS
    return BN_OK;
}

extern "C" int bn_fit_size(bn_ctx *ctx, bn_mat *m, const double *BETA_vec, const double *INITIAL_MU_vec,
                           const double *INITIAL_SIZE_vec, double SIZE_lower, double SIZE_upper, const bn_fit_opts *opts,
                           double *size_out, int32_t *conv, int32_t *nfeval)
{
    if (!ctx || !m || !BETA_vec || !INITIAL_MU_vec || !INITIAL_SIZE_vec || !size_out)
        return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_fit_size: NULL argument") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevIn<double> b, mu0, s0;
    BN_TRY(b.bind(ctx, BETA_vec, (size_t)m->C));
    BN_TRY(mu0.bind(ctx, INITIAL_MU_vec, (size_t)m->G));
    BN_TRY(s0.bind(ctx, INITIAL_SIZE_vec, (size_t)m->G));
    DevOut<double> so;
    DevOut<int32_t> co, no;
    BN_TRY(so.bind(ctx, size_out, (size_t)m->G));
    BN_TRY(co.bind(ctx, conv, (size_t)m->G));
    BN_TRY(no.bind(ctx, nfeval, (size_t)m->G));
    BN_TRY(bn_fit_size_device(ctx, m, b.p, mu0.p, s0.p, SIZE_lower, SIZE_upper, nullptr, opts, so.p, co.p, no.p));
    BN_TRY(so.commit(ctx));
    BN_TRY(co.commit(ctx));
    BN_TRY(no.commit(ctx));
    return bn_sync(ctx);
}

// ---------------------------------------------------------------------------------
// single-gene entry points on a dense count vector (the registered .Call routines
// MarginalF_NB_1D/2D and GradientFun_NB_1D/2D).  One CTA, fixed-order reduction.
// ---------------------------------------------------------------------------------
// which: 0 = log-likelihood, 1 = d/dsize, 2 = d/dmu
__global__ void __launch_bounds__(1024) k_dense_gene(int which, double size, double mu, const double *__restrict__ x,
                                                     const double *__restrict__ beta, int64_t n, double *__restrict__ out)
{
    __shared__ double sh[32];
    double acc = 0.0;
    const double lgs = lgamma(size), dgs = (which == 1) ? bn_digamma(size) : 0.0;
    for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double xv = x[j], m = beta[j] * mu;
        double t;
        if (which == 0) {
            if (xv < 0.0 || xv != floor(xv)) t = -INFINITY; // R_D_nonint_check / x < 0: density 0
            else if (xv == 0.0) t = -size * log1p(m / size);
            else
                t = lgamma(xv + size) - lgs - lgamma(xv + 1.0) - size * log1p(m / size) + xv * (log(m) - log(size + m));
        } else if (which == 1) {
            t = bn_digamma(xv + size) - dgs + log(size / (size + m)) + (m - xv) / (m + size);
        } else {
            t = (xv * size - m * size) / (mu * (m + size));
        }
        acc += t;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    acc = bn_warp_sum(acc);
    if (lane == 0) sh[w] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) s += sh[k];
        out[0] = s;
    }
}

static int dense_gene(bn_ctx *ctx, int which, double size, double mu, const double *x, const double *beta, int64_t n,
                      double *out)
{
    if (!ctx || !x || !beta || !out || n <= 0) return ctx ? bn_fail(ctx, BN_ERR_INVALID, "likelihood entry: bad arguments") : BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevIn<double> dx, db;
    BN_TRY(dx.bind(ctx, x, (size_t)n));
    BN_TRY(db.bind(ctx, beta, (size_t)n));
    DevOut<double> o;
    BN_TRY(o.bind(ctx, out, 1));
    BN_LAUNCH(ctx, k_dense_gene, 1, 1024, 0, which, size, mu, dx.p, db.p, n, o.p);
    BN_TRY(o.commit(ctx));
    return bn_sync(ctx);
}

extern "C" int bn_marginal_nb_1d(bn_ctx *ctx, double SIZE, double MU, const double *m_observed, const double *BETA,
                                 int64_t n, double *out)
{
    return dense_gene(ctx, 0, SIZE, MU, m_observed, BETA, n, out);
}
extern "C" int bn_marginal_nb_2d(bn_ctx *ctx, const double SIZE_MU[2], const double *m_observed, const double *BETA,
                                 int64_t n, double *out)
{
    if (!SIZE_MU) return BN_ERR_INVALID;
    return dense_gene(ctx, 0, SIZE_MU[0], SIZE_MU[1], m_observed, BETA, n, out);
}
extern "C" int bn_gradient_nb_1d(bn_ctx *ctx, double SIZE, double MU, const double *m_observed, const double *BETA,
                                 int64_t n, double *out)
{
    return dense_gene(ctx, 1, SIZE, MU, m_observed, BETA, n, out);
}
extern "C" int bn_gradient_nb_2d(bn_ctx *ctx, const double SIZE_MU[2], const double *m_observed, const double *BETA,
                                 int64_t n, double out[2])
{
    if (!SIZE_MU || !out) return BN_ERR_INVALID;
    BN_TRY(dense_gene(ctx, 1, SIZE_MU[0], SIZE_MU[1], m_observed, BETA, n, out));
    // out may be host or device; element 1 written by a second launch
    return dense_gene(ctx, 2, SIZE_MU[0], SIZE_MU[1], m_observed, BETA, n, out + 1);
}

// ---------------------------------------------------------------------------------
// debug hook: the device logarithms on a vector, so tests can pin them against mpmath.
// which: 0 = bn_log, 1 = bn_log1p_pos, 2 = bn_log1p_small, 3 = bn_lgamma_stirling
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_debug_log(int which, int64_t n, const double *__restrict__ x, double *__restrict__ out)
{
    __shared__ BnLogEntry tab[BN_LOG_TABLE_N];
    bn_log_table_load(tab);
    __syncthreads();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += gridDim.x * (int64_t)blockDim.x) {
        const double v = x[k];
        out[k] = which == 0 ? bn_log(v, tab) : which == 1 ? bn_log1p_pos(v, tab) : which == 2 ? bn_log1p_small(v) : bn_lgamma_stirling(v, tab);
    }
}

extern "C" int bn_debug_log(bn_ctx *ctx, int which, const double *x, int64_t n, double *out)
{
    if (!ctx || !x || !out || n <= 0 || which < 0 || which > 3) return BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevIn<double> dx;
    BN_TRY(dx.bind(ctx, x, (size_t)n));
    DevOut<double> o;
    BN_TRY(o.bind(ctx, out, (size_t)n));
    BN_LAUNCH(ctx, k_debug_log, (unsigned)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, which, n, dx.p, o.p);
    BN_TRY(o.commit(ctx));
    return bn_sync(ctx);
}

// BB_Fun with FIX_MU = FALSE: one call fits (size, mu) of every gene.
extern "C" int bn_fit_size_mu(bn_ctx *ctx, bn_mat *m, const double *BETA_vec, const double *INITIAL_MU_vec,
                              const double *INITIAL_SIZE_vec, double SIZE_lower, double SIZE_upper, double MU_lower,
                              double MU_upper, const bn_fit_opts *opts, double *size_out, double *mu_out, int32_t *conv,
                              int32_t *nfeval)
{
    if (!ctx || !m || !BETA_vec || !INITIAL_MU_vec || !INITIAL_SIZE_vec || !size_out || !mu_out)
        return ctx ? bn_fail(ctx, BN_ERR_INVALID, "bn_fit_size_mu: NULL argument") : BN_ERR_INVALID;
    if (!m->is_count)
        return bn_fail(ctx, BN_ERR_NONCOUNT, "the likelihood fit needs non-negative integer counts (dnbinom_mu is 0 elsewhere)");
    if (!(SIZE_lower > 0.0) || !(SIZE_upper >= SIZE_lower) || !(MU_lower >= 0.0) || !(MU_upper >= MU_lower))
        return bn_fail(ctx, BN_ERR_INVALID, "bad bounds: size [%g, %g], mu [%g, %g]", SIZE_lower, SIZE_upper, MU_lower, MU_upper);
    cudaSetDevice(ctx->device);
    bn_fit_opts od;
    bn_fit_opts_default(&od);
    if (opts) od = *opts;
    if (od.method != BN_FIT_SPG) return bn_fail(ctx, BN_ERR_UNSUPPORTED, "the 2-D fit runs BB::spg only");
    FitOpts o;
    o.method = od.method;
    o.maxfeval = od.maxfeval;
    o.maxit = od.maxit;
    o.M = od.M;
    o.use_gradient = od.use_gradient;
    o.ftol = od.ftol;
    o.gtol = od.gtol;
    o.eps = od.eps;
    o.xatol = od.xatol;
    o.evcap = FIT_EVCAP;
    DevIn<double> b, mu0, s0;
    BN_TRY(b.bind(ctx, BETA_vec, (size_t)m->C));
    BN_TRY(mu0.bind(ctx, INITIAL_MU_vec, (size_t)m->G));
    BN_TRY(s0.bind(ctx, INITIAL_SIZE_vec, (size_t)m->G));
    DevOut<double> so, mo;
    DevOut<int32_t> co, no;
    BN_TRY(so.bind(ctx, size_out, (size_t)m->G));
    BN_TRY(mo.bind(ctx, mu_out, (size_t)m->G));
    BN_TRY(co.bind(ctx, conv, (size_t)m->G));
    BN_TRY(no.bind(ctx, nfeval, (size_t)m->G));
    BN_TRY(bn_mat_ensure_csr(ctx, m));
    DevBuf<double> lb, box;
    BN_TRY(lb.alloc(ctx, (size_t)m->C));
    BN_TRY(box.alloc(ctx, 8));
    BN_LAUNCH(ctx, k_log_vec, (unsigned)((m->C + 255) / 256), 256, 0, m->C, b.p, lb.p);
    BN_LAUNCH(ctx, k_fit_box, 1, 1, 0, SIZE_lower, MU_lower, SIZE_upper, MU_upper, box.p);
    BN_LAUNCH(ctx, k_max_vec, 1, 1024, 0, m->C, b.p, box.p + 4);
    DevBuf<unsigned long long> q;
    BN_TRY(q.alloc(ctx, 1));
    BN_CUDA(ctx, cudaMemsetAsync(q.p, 0, 8, ctx->stream));
    const int grid = (int)std::min<int64_t>(m->G, (int64_t)ctx->prop.multiProcessorCount * 4);
    ctx->last_fit[0] = 4;
    *ctx->fit_parked = 0;
    ctx->last_fit[2] = 0;
    ctx->last_fit[3] = o.use_gradient ? 1 : 0;
    if (o.use_gradient)
        BN_LAUNCH(ctx, k_fit_2d<true>, grid, 256, 0, m->G, m->C, m->rowptr, m->gm, b.p, lb.p, mu0.p, s0.p, box.p, o, q.p, so.p, mo.p,
                  co.p, no.p);
    else
        BN_LAUNCH(ctx, k_fit_2d<false>, grid, 256, 0, m->G, m->C, m->rowptr, m->gm, b.p, lb.p, mu0.p, s0.p, box.p, o, q.p, so.p, mo.p,
                  co.p, no.p);
    BN_TRY(so.commit(ctx));
    BN_TRY(mo.commit(ctx));
    BN_TRY(co.commit(ctx));
    BN_TRY(no.commit(ctx));
    return bn_sync(ctx);
}

// debug hook: the dense-part table F(r) = sum_j log(1 + r BETA_j) built for [r_lo, r_hi] and evaluated at r[0..n)
__global__ void __launch_bounds__(256) k_dt_debug(int64_t n, const double *__restrict__ r, const double *__restrict__ coef,
                                                  const DenseMeta *__restrict__ meta, double *__restrict__ out)
{
    __shared__ BnLogEntry tab[BN_LOG_TABLE_N];
    bn_log_table_load(tab);
    __syncthreads();
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += gridDim.x * (int64_t)blockDim.x) {
        const double v = r[k];
        out[k] = (v >= meta->r_lo && v <= meta->r_hi) ? dt_eval(v, coef, meta, tab) : nan("");
    }
}

extern "C" int bn_debug_dense_table(bn_ctx *ctx, const double *BETA, int64_t C, double r_lo, double r_hi, const double *r, int64_t n,
                                    double *out)
{
    if (!ctx || !BETA || !r || !out || C <= 0 || n <= 0) return BN_ERR_INVALID;
    cudaSetDevice(ctx->device);
    DevIn<double> b, dr;
    BN_TRY(b.bind(ctx, BETA, (size_t)C));
    BN_TRY(dr.bind(ctx, r, (size_t)n));
    DevOut<double> o;
    BN_TRY(o.bind(ctx, out, (size_t)n));
    constexpr int NSPLIT = 6;
    DevBuf<double> dt;
    BN_TRY(dt.alloc(ctx, (size_t)DT_NI * DT_STRIDE + (size_t)DT_NI * NSPLIT * DT_N + 8));
    double *coef = dt.p, *part = dt.p + (size_t)DT_NI * DT_STRIDE;
    DenseMeta *meta = reinterpret_cast<DenseMeta *>(part + (size_t)DT_NI * NSPLIT * DT_N);
    BN_LAUNCH(ctx, k_dt_range, 1, 1024, 0, (int64_t)0, (const double *)nullptr, (const double *)nullptr, r_lo, r_hi, meta);
    BN_LAUNCH(ctx, k_dt_nodes, dim3(DT_NI, NSPLIT), 256, 0, C, b.p, meta, NSPLIT, part);
    BN_LAUNCH(ctx, k_dt_coef, DT_NI, 32, 0, meta, NSPLIT, part, coef);
    BN_LAUNCH(ctx, k_dt_debug, 64, 256, 0, n, dr.p, coef, meta, o.p);
    BN_TRY(o.commit(ctx));
    return bn_sync(ctx);
}

extern "C" int bn_debug_last_fit(const bn_ctx *ctx, int64_t out[4])
{
    if (!ctx || !out) return BN_ERR_INVALID;
    out[0] = ctx->last_fit[0];
    out[1] = (int64_t)*ctx->fit_parked;
    out[2] = ctx->last_fit[2];
    out[3] = ctx->last_fit[3];
    return BN_OK;
}

extern "C" int bn_debug_tune(bn_ctx *ctx, int knob, int value)
{
    if (!ctx || knob < 0 || knob >= BN_TUNE_COUNT) return BN_ERR_INVALID;
    ctx->tune[knob] = value;
    return BN_OK;
}
