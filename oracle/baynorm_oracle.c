/*
 * baynorm_oracle.c -- CPU restatement of bayNorm's data-parallel hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it,
 * and only as the checker / the timed CPU baseline.  The shipped library
 * (baynorm_b200/csrc) never links, includes or calls anything in oracle/.
 *
 * PARITY STATUS: "parity unpinned" at the third-party boundary.  The reference
 * (WT215/bayNorm v1.5.14) cannot be compiled here: src/bayNorm_main.cpp:2-7
 * includes Rcpp / RcppArmadillo / Rmath, none of which exist in this image, and
 * its tests (tests/testthat/test-bayNorm.r:9-29) hold no numeric golden vector.
 * Closed-form rows (beta, MME, mean, mode, OLS adjust) are fully specified by the
 * reference's own source and are restated operation-for-operation below.  The
 * likelihood goes through R nmath::dnbinom_mu and the optimiser through BB::spg;
 * neither is vendored nor version-pinned (DESCRIPTION:18-21).  Their published
 * algorithms are restated from the upstream sources as the author recalls them
 * (R 3.5 - 4.3 nmath dnbinom.c / dbinom.c / bd0.c / stirlerr.c; BB 2019.10-1
 * spg.R) and every such function is tagged [recall].  tests/test_oracle.py pins
 * what can be pinned: dnbinom_mu against scipy/mpmath, spg against a bounded
 * Brent argmax, closed forms against an independent numpy implementation.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -march=native, to
 * mirror a stock R build on x86-64: no FMA contraction, IEEE double).
 *
 * All matrices are dgCMatrix-style CSC: p[C+1] (int64 here), i[nnz] int32 sorted
 * within a column, x[nnz] double.  Dense matrices are column-major (R layout).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

#ifndef M_LN_SQRT_2PI
#define M_LN_SQRT_2PI 0.918938533204672741780329736406 /* log(sqrt(2*pi)) */
#endif
#ifndef M_LN_2PI
#define M_LN_2PI 1.837877066409345483560659472811 /* log(2*pi) */
#endif

/* ------------------------------------------------------------------------- */
/* R nmath restatements  [recall]                                             */
/* ------------------------------------------------------------------------- */

/* stirlerr(n) = log(n!) - log( sqrt(2*pi*n)*(n/e)^n )   (nmath/stirlerr.c, R < 4.4)
 * Upstream keeps a table sferr_halves[] for n <= 15 with 2n integer; the table
 * entries are the same mathematical quantity, so the lgamma expression is used
 * for those arguments too (deviation: last-bit only). */
static double orc_stirlerr(double n)
{
    const double S0 = 0.083333333333333333333;        /* 1/12 */
    const double S1 = 0.00277777777777777777778;      /* 1/360 */
    const double S2 = 0.00079365079365079365079365;   /* 1/1260 */
    const double S3 = 0.000595238095238095238095238;  /* 1/1680 */
    const double S4 = 0.0008417508417508417508417508; /* 1/1188 */
    double nn;
    if (n <= 15.0)
        return lgamma(n + 1.0) - (n + 0.5) * log(n) + n - M_LN_SQRT_2PI;
    nn = n * n;
    if (n > 500) return (S0 - S1 / nn) / n;
    if (n > 80) return (S0 - (S1 - S2 / nn) / nn) / n;
    if (n > 35) return (S0 - (S1 - (S2 - S3 / nn) / nn) / nn) / n;
    return (S0 - (S1 - (S2 - (S3 - S4 / nn) / nn) / nn) / nn) / n;
}

/* bd0(x, np) = x log(x/np) + np - x, evaluated stably  (nmath/bd0.c) */
static double orc_bd0(double x, double np)
{
    double ej, s, s1, v;
    int j;
    if (!isfinite(x) || !isfinite(np) || np == 0.0) return NAN;
    if (fabs(x - np) < 0.1 * (x + np)) {
        v = (x - np) / (x + np);
        s = (x - np) * v;
        if (fabs(s) < DBL_MIN) return s;
        ej = 2 * x * v;
        v = v * v;
        for (j = 1; j < 1000; j++) {
            ej *= v;
            s1 = s + ej / ((j << 1) + 1);
            if (s1 == s) return s1;
            s = s1;
        }
    }
    return x * log(x / np) + np - x;
}

/* dbinom_raw(x, n, p, q, log=TRUE)  (nmath/dbinom.c) */
static double orc_dbinom_raw_log(double x, double n, double p, double q)
{
    double lf, lc;
    if (p == 0) return (x == 0) ? 0.0 : -INFINITY;
    if (q == 0) return (x == n) ? 0.0 : -INFINITY;
    if (x == 0) {
        if (n == 0) return 0.0;
        lc = (p < 0.1) ? -orc_bd0(n, n * q) - n * p : n * log(q);
        return lc;
    }
    if (x == n) {
        lc = (q < 0.1) ? -orc_bd0(n, n * p) - n * q : n * log(p);
        return lc;
    }
    if (x < 0 || x > n) return -INFINITY;
    lc = orc_stirlerr(n) - orc_stirlerr(x) - orc_stirlerr(n - x) - orc_bd0(x, n * p) - orc_bd0(n - x, n * q);
    lf = M_LN_2PI + log(x) + log1p(-x / n);
    return lc - 0.5 * lf;
}

/* dnbinom_mu(x, size, mu, log=TRUE)  (nmath/dnbinom.c); the call the reference
 * makes at src/bayNorm_main.cpp:939 and :960. */
ORC_API double orc_dnbinom_mu_log(double x, double size, double mu)
{
    if (isnan(x) || isnan(size) || isnan(mu)) return x + size + mu;
    if (mu < 0 || size < 0) return NAN;
    if (fabs(x - nearbyint(x)) > 1e-7 * fmax(1.0, fabs(x))) return -INFINITY; /* R_D_nonint_check */
    if (x < 0 || !isfinite(x)) return -INFINITY;
    if (x == 0 && size == 0) return 0.0;
    x = nearbyint(x);
    if (!isfinite(size)) { /* Poisson limit: dpois_raw(x, mu, log) */
        if (mu == 0) return (x == 0) ? 0.0 : -INFINITY;
        if (x == 0) return -mu;
        return -0.5 * (M_LN_2PI + log(x)) - orc_stirlerr(x) - orc_bd0(x, mu);
    }
    if (x == 0)
        return size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu)));
    if (x < 1e-10 * size) {
        double p = (size < mu ? log(size / (1 + size / mu)) : log(mu / (1 + mu / size)));
        return x * p - mu - lgamma(x + 1) + log1p(x * (x - 1) / (2 * size));
    } else {
        double p = size / (size + x);
        double ans = orc_dbinom_raw_log(size, x + size, size / (size + mu), mu / (size + mu));
        return log(p) + ans;
    }
}

/* Same density written directly with lgamma; used to cross-check the saddle-point
 * form above.  Not a reference code path. */
ORC_API double orc_dnbinom_mu_log_direct(double x, double size, double mu)
{
    if (x == 0) return size * (size < mu ? log(size / (size + mu)) : log1p(-mu / (size + mu)));
    return lgamma(x + size) - lgamma(size) - lgamma(x + 1.0) + size * log(size / (size + mu)) +
           x * log(mu / (size + mu));
}

/* digamma: recurrence up to x >= 6 then asymptotic series.  R uses Amos' dpsifn;
 * any digamma accurate to ~1e-15 is the same function.  [recall: n/a] */
ORC_API double orc_digamma(double x)
{
    double r = 0.0, f, t;
    if (!(x > 0)) return NAN;
    while (x < 10.0) {
        r -= 1.0 / x;
        x += 1.0;
    }
    f = 1.0 / (x * x);
    t = f * (-1.0 / 12.0 +
             f * (1.0 / 120.0 +
                  f * (-1.0 / 252.0 +
                       f * (1.0 / 240.0 + f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
    return r + log(x) - 0.5 / x + t;
}

/* ------------------------------------------------------------------------- */
/* Likelihood and gradients  (src/bayNorm_main.cpp:928-1013)                  */
/* ------------------------------------------------------------------------- */

/* MarginalF_NB_1D  src/bayNorm_main.cpp:928-943.  temp_vec_2(i) then Rcpp sugar
 * sum(): sequential accumulation in cell order. */
ORC_API double orc_marginal_nb_1d(double SIZE, double MU, const double *m_observed, const double *BETA, int64_t n)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += orc_dnbinom_mu_log(m_observed[i], SIZE, BETA[i] * MU);
    return s;
}

/* MarginalF_NB_2D  src/bayNorm_main.cpp:948-965  (SIZE_MU = {size, mu}) */
ORC_API double orc_marginal_nb_2d(const double *SIZE_MU, const double *m_observed, const double *BETA, int64_t n)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; i++) s += orc_dnbinom_mu_log(m_observed[i], SIZE_MU[0], SIZE_MU[1] * BETA[i]);
    return s;
}

/* GradientFun_NB_1D  src/bayNorm_main.cpp:972-987 */
ORC_API double orc_gradient_nb_1d(double SIZE, double MU, const double *m_observed, const double *BETA, int64_t n)
{
    double s = 0.0;
    for (int64_t i = 0; i < n; i++)
        s += orc_digamma(m_observed[i] + SIZE) - orc_digamma(SIZE) + log(SIZE / (SIZE + MU * BETA[i])) +
             (BETA[i] * MU - m_observed[i]) / (BETA[i] * MU + SIZE);
    return s;
}

/* GradientFun_NB_2D  src/bayNorm_main.cpp:993-1013;  out = {d/dsize, d/dmu} */
ORC_API void orc_gradient_nb_2d(const double *SIZE_MU, const double *m_observed, const double *BETA, int64_t n,
                                double *out)
{
    double ss = 0.0, sm = 0.0;
    const double sz = SIZE_MU[0], mu = SIZE_MU[1];
    for (int64_t i = 0; i < n; i++) {
        sm += (m_observed[i] * sz - mu * BETA[i] * sz) / (mu * (mu * BETA[i] + sz));
        ss += orc_digamma(m_observed[i] + sz) - orc_digamma(sz) + log(sz / (sz + mu * BETA[i])) +
              (BETA[i] * mu - m_observed[i]) / (BETA[i] * mu + sz);
    }
    out[0] = ss;
    out[1] = sm;
}

/* ------------------------------------------------------------------------- */
/* beta (library size)  R/bayNorm.r:164-170                                   */
/* ------------------------------------------------------------------------- */

ORC_API void orc_colsums(int64_t C, const int64_t *p, const double *x, double *out)
{
    for (int64_t j = 0; j < C; j++) {
        double s = 0.0;
        for (int64_t k = p[j]; k < p[j + 1]; k++) s += x[k];
        out[j] = s;
    }
}

/* xx/mean(xx)*mean_beta, then the two clamps in the reference's order. Returns 0,
 * or 1 if the range check at R/bayNorm.r:185-187 would stop(). */
ORC_API int orc_beta_from_colsums(int64_t C, const double *xx, double mean_beta, double *beta)
{
    double tot = 0.0, mean, minpos = INFINITY, maxlt1 = -INFINITY, lo = INFINITY, hi = -INFINITY;
    for (int64_t j = 0; j < C; j++) tot += xx[j];
    mean = tot / (double)C;
    for (int64_t j = 0; j < C; j++) beta[j] = xx[j] / mean * mean_beta;
    for (int64_t j = 0; j < C; j++)
        if (beta[j] > 0 && beta[j] < minpos) minpos = beta[j];
    for (int64_t j = 0; j < C; j++)
        if (beta[j] <= 0) beta[j] = minpos;
    for (int64_t j = 0; j < C; j++)
        if (beta[j] < 1 && beta[j] > maxlt1) maxlt1 = beta[j];
    for (int64_t j = 0; j < C; j++)
        if (beta[j] >= 1) beta[j] = maxlt1;
    for (int64_t j = 0; j < C; j++) {
        if (beta[j] < lo) lo = beta[j];
        if (beta[j] > hi) hi = beta[j];
    }
    return (lo <= 0 || hi >= 1 || isnan(lo) || isnan(hi)) ? 1 : 0;
}

/* ------------------------------------------------------------------------- */
/* MME  src/bayNorm_main.cpp:1303-1394 + R/PRIOR_FUNCTIONS.R:163, :253        */
/* ------------------------------------------------------------------------- */

/* beta may be NULL (EstPrior on data that is already normalised).  With beta the
 * normalisation x/beta_j of R/PRIOR_FUNCTIONS.R:253 is applied on the fly; R
 * materialises normcount_N first, the values are the same doubles.
 * Outputs: MU[G], SIZE[G] (NaN where v <= m), v[G].                           */
ORC_API void orc_mme(int64_t G, int64_t C, const int64_t *p, const int32_t *ri, const double *x, const double *beta,
                     double *MU, double *SIZE, double *V)
{
    double *nz = (double *)calloc((size_t)G, sizeof(double));
    const double n = (double)C;
    for (int64_t g = 0; g < G; g++) MU[g] = 0.0;
    /* rowMeansFast :1303-1320 : iterate CSC in storage order, scatter-add */
    for (int64_t j = 0; j < C; j++)
        for (int64_t k = p[j]; k < p[j + 1]; k++) MU[ri[k]] += beta ? x[k] / beta[j] : x[k];
    for (int64_t g = 0; g < G; g++) MU[g] /= (double)C;
    /* rowVarsFast :1323-1342 */
    for (int64_t g = 0; g < G; g++) V[g] = 0.0;
    for (int64_t j = 0; j < C; j++)
        for (int64_t k = p[j]; k < p[j + 1]; k++) {
            const double val = beta ? x[k] / beta[j] : x[k];
            const int32_t r = ri[k];
            V[r] += (val - MU[r]) * (val - MU[r]);
            nz[r] += 1;
        }
    for (int64_t g = 0; g < G; g++) {
        V[g] += ((double)C - nz[g]) * (MU[g] * MU[g]);
        V[g] /= (double)(C - 1);
    }
    /* EstPrior_sprcpp :1385-1386 then NaN rule R/PRIOR_FUNCTIONS.R:163 */
    for (int64_t g = 0; g < G; g++) {
        V[g] = (n - 1) / n * V[g];
        SIZE[g] = pow(MU[g], 2) / (V[g] - MU[g]);
        if (V[g] <= MU[g]) SIZE[g] = NAN;
    }
    free(nz);
}

/* ------------------------------------------------------------------------- */
/* BB::spg restated  [recall: BB 2019.10-1, R/spg.R]                          */
/* ------------------------------------------------------------------------- */

typedef double (*orc_obj_fn)(const double *par, void *user);

typedef struct {
    int M;          /* 10 */
    int maxit;      /* 1500 */
    double ftol;    /* 1e-10 */
    double gtol;    /* 1e-5 */
    int maxfeval;   /* reference passes 500 (R/PRIOR_FUNCTIONS.R:451-454) */
    double eps;     /* 1e-7 */
    int method;     /* 3 */
} orc_spg_ctrl;

static void orc_project(int n, double *p, const double *lo, const double *hi)
{
    for (int i = 0; i < n; i++) {
        if (p[i] < lo[i]) p[i] = lo[i];
        if (p[i] > hi[i]) p[i] = hi[i];
    }
}

/* Maximises fn (control$maximize=TRUE: func = -fn).  n <= 2.  Returns the
 * convergence code of spg (0 ok, 1 maxit, 2 maxfeval, 3/4/5 failures); par is
 * overwritten with the result ($par == pbest).  counts[0]=feval (line-search +
 * initial), counts[1]=extra evaluations spent on finite differences, counts[2]=iter. */
static int orc_spg(int n, double *par, orc_obj_fn fn, void *user, const double *lower, const double *upper,
                   const orc_spg_ctrl *ct, int *counts)
{
    enum { NMAX = 2, MMAX = 64 };
    const double lmin = 1e-30, lmax = 1e30;
    double lastfv[MMAX], pbest[NMAX], g[NMAX], pg[NMAX], d[NMAX], pnew[NMAX], gnew[NMAX], dx[NMAX];
    double f, fbest, fchg = INFINITY, pginfn, lambda = 1.0;
    int iter = 0, feval = 0, fdeval = 0, lsflag = -1, M = ct->M, conv = 0;
    if (M > MMAX) M = MMAX;
    for (int i = 0; i < M; i++) lastfv[i] = -1e99;
#define FUNC(P) (-(fn((P), user)))
#define GRAD(P, F, OUT)                                   \
    do {                                                  \
        for (int i_ = 0; i_ < n; i_++) {                  \
            for (int k_ = 0; k_ < n; k_++) dx[k_] = (P)[k_]; \
            dx[i_] = dx[i_] + ct->eps;                    \
            (OUT)[i_] = (FUNC(dx) - (F)) / ct->eps;       \
            fdeval++;                                     \
        }                                                 \
    } while (0)

    orc_project(n, par, lower, upper);
    for (int i = 0; i < n; i++) pbest[i] = par[i];
    f = FUNC(par);
    feval++;
    if (isnan(f) || isinf(f)) {
        if (counts) { counts[0] = feval; counts[1] = fdeval; counts[2] = iter; }
        return 3;
    }
    fbest = f;
    GRAD(par, f, g);
    for (int i = 0; i < n; i++)
        if (isnan(g[i])) {
            if (counts) { counts[0] = feval; counts[1] = fdeval; counts[2] = iter; }
            return 4;
        }
    lastfv[0] = f;
    for (int i = 0; i < n; i++) pg[i] = par[i] - g[i];
    orc_project(n, pg, lower, upper);
    pginfn = 0.0;
    for (int i = 0; i < n; i++) {
        pg[i] -= par[i];
        if (fabs(pg[i]) > pginfn) pginfn = fabs(pg[i]);
    }
    if (pginfn != 0) lambda = fmin(lmax, fmax(lmin, 1.0 / pginfn));

    while (pginfn > ct->gtol && iter <= ct->maxit && fchg > ct->ftol) {
        double gtd = 0.0, fmax_, alpha, fnew, sts = 0, yty = 0, sty = 0;
        iter++;
        for (int i = 0; i < n; i++) d[i] = par[i] - lambda * g[i];
        orc_project(n, d, lower, upper);
        for (int i = 0; i < n; i++) {
            d[i] -= par[i];
            gtd += g[i] * d[i];
        }
        if (isinf(gtd)) { lsflag = 4; break; }
        /* nmls: nonmonotone line search, safeguarded quadratic interpolation */
        fmax_ = lastfv[0];
        for (int i = 1; i < M; i++)
            if (lastfv[i] > fmax_) fmax_ = lastfv[i];
        alpha = 1.0;
        for (int i = 0; i < n; i++) pnew[i] = par[i] + alpha * d[i];
        fnew = FUNC(pnew);
        feval++;
        lsflag = 0;
        if (isnan(fnew)) { lsflag = 1; break; }
        while (fnew > fmax_ + 1e-4 * alpha * gtd) {
            if (alpha <= 0.1) alpha = alpha / 2;
            else {
                double atemp = -(gtd * alpha * alpha) / (2 * (fnew - f - alpha * gtd));
                if (atemp < 0.1 || atemp > 0.9 * alpha) atemp = alpha / 2;
                alpha = atemp;
            }
            for (int i = 0; i < n; i++) pnew[i] = par[i] + alpha * d[i];
            fnew = FUNC(pnew);
            feval++;
            if (isnan(fnew)) { lsflag = 1; break; }
            if (feval > ct->maxfeval) { lsflag = 2; break; }
        }
        if (lsflag != 0) break;
        fchg = fabs(f - fnew);
        f = fnew;
        lastfv[iter % M] = f;
        GRAD(pnew, f, gnew);
        {
            int bad = 0;
            for (int i = 0; i < n; i++)
                if (isnan(gnew[i])) bad = 1;
            if (bad) { lsflag = 3; break; }
        }
        for (int i = 0; i < n; i++) {
            const double s = pnew[i] - par[i], y = gnew[i] - g[i];
            sts += s * s;
            yty += y * y;
            sty += s * y;
            par[i] = pnew[i];
            g[i] = gnew[i];
        }
        if (ct->method == 1) lambda = (sts == 0 || sty < 0) ? lmax : fmin(lmax, fmax(lmin, sts / sty));
        else if (ct->method == 2) lambda = (sty < 0 || yty == 0) ? lmax : fmin(lmax, fmax(lmin, sty / yty));
        else lambda = (sts == 0 || yty == 0) ? lmax : fmin(lmax, fmax(lmin, sqrt(sts / yty)));
        for (int i = 0; i < n; i++) pg[i] = par[i] - g[i];
        orc_project(n, pg, lower, upper);
        pginfn = 0.0;
        for (int i = 0; i < n; i++) {
            pg[i] -= par[i];
            if (fabs(pg[i]) > pginfn) pginfn = fabs(pg[i]);
        }
        if (f < fbest) {
            fbest = f;
            for (int i = 0; i < n; i++) pbest[i] = pnew[i];
        }
    }
    if (lsflag <= 0) {
        conv = 0;
        if (iter >= ct->maxit) conv = 1;
    } else {
        conv = (lsflag == 1) ? 3 : (lsflag == 2) ? 2 : (lsflag == 3) ? 4 : 5;
    }
    for (int i = 0; i < n; i++) par[i] = pbest[i];
    if (counts) { counts[0] = feval; counts[1] = fdeval; counts[2] = iter; }
#undef FUNC
#undef GRAD
    return conv;
}

typedef struct {
    const double *m_observed;
    const double *beta;
    int64_t n;
    double mu;
} orc_gene_ctx;

static double orc_obj_1d(const double *par, void *user)
{
    const orc_gene_ctx *c = (const orc_gene_ctx *)user;
    return orc_marginal_nb_1d(par[0], c->mu, c->m_observed, c->beta, c->n);
}
static double orc_obj_2d(const double *par, void *user)
{
    const orc_gene_ctx *c = (const orc_gene_ctx *)user;
    return orc_marginal_nb_2d(par, c->m_observed, c->beta, c->n);
}

/* One gene, dense row: exactly the spg call at R/PRIOR_FUNCTIONS.R:446-456. */
ORC_API int orc_spg_gene_1d(double size0, double mu, const double *m_observed, const double *beta, int64_t n,
                            double lower, double upper, int maxfeval, double *size_out, int *counts)
{
    orc_spg_ctrl ct = {10, 1500, 1e-10, 1e-5, maxfeval, 1e-7, 3};
    orc_gene_ctx c = {m_observed, beta, n, mu};
    double par[1] = {size0};
    int rc = orc_spg(1, par, orc_obj_1d, &c, &lower, &upper, &ct, counts);
    *size_out = par[0];
    return rc;
}

ORC_API int orc_spg_gene_2d(const double *par0, const double *m_observed, const double *beta, int64_t n,
                            const double *lower, const double *upper, int maxfeval, double *par_out, int *counts)
{
    orc_spg_ctrl ct = {10, 1500, 1e-10, 1e-5, maxfeval, 1e-7, 3};
    orc_gene_ctx c = {m_observed, beta, n, 0.0};
    double par[2] = {par0[0], par0[1]};
    int rc = orc_spg(2, par, orc_obj_2d, &c, lower, upper, &ct, counts);
    par_out[0] = par[0];
    par_out[1] = par[1];
    return rc;
}

/* CSC -> gene-major (CSR) transpose; stands in for `Data[Geneind, ]`
 * (R/PRIOR_FUNCTIONS.R:442).  rp[G+1], cj[nnz], cx[nnz]. */
static void orc_csc_to_csr(int64_t G, int64_t C, const int64_t *p, const int32_t *ri, const double *x, int64_t *rp,
                           int32_t *cj, double *cx)
{
    int64_t *cur;
    memset(rp, 0, sizeof(int64_t) * (size_t)(G + 1));
    for (int64_t k = 0; k < p[C]; k++) rp[ri[k] + 1]++;
    for (int64_t g = 0; g < G; g++) rp[g + 1] += rp[g];
    cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)G);
    memcpy(cur, rp, sizeof(int64_t) * (size_t)G);
    for (int64_t j = 0; j < C; j++)
        for (int64_t k = p[j]; k < p[j + 1]; k++) {
            const int64_t q = cur[ri[k]]++;
            cj[q] = (int32_t)j;
            cx[q] = x[k];
        }
    free(cur);
}

/* BB_Fun, FIX_MU=TRUE  (R/PRIOR_FUNCTIONS.R:380-549).  Genes [g0, g1) only, so a
 * bounded sample can be timed.  nthreads stands in for the SOCK cluster's
 * NCores (the reference's only parallel step).  counts_out may be NULL or
 * int[3*(g1-g0)].  Row extraction cost of `Data[g, ]` on a dgCMatrix, the R
 * interpreter and .Call marshalling are NOT reproduced: optimistic stand-in. */
ORC_API int orc_bb_fun_1d(int64_t G, int64_t C, const int64_t *p, const int32_t *ri, const double *x,
                          const double *beta, const double *mu0, const double *size0, double lower, double upper,
                          int maxfeval, int64_t g0, int64_t g1, int nthreads, double *size_out, int *conv_out,
                          int *counts_out)
{
    const int64_t nnz = p[C];
    int64_t *rp = (int64_t *)malloc(sizeof(int64_t) * (size_t)(G + 1));
    int32_t *cj = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
    double *cx = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
    if (!rp || !cj || !cx) return -1;
    orc_csc_to_csr(G, C, p, ri, x, rp, cj, cx);
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
    {
        double *row = (double *)malloc(sizeof(double) * (size_t)C);
#pragma omp for schedule(dynamic, 1)
        for (int64_t g = g0; g < g1; g++) {
            int counts[3];
            memset(row, 0, sizeof(double) * (size_t)C);
            for (int64_t k = rp[g]; k < rp[g + 1]; k++) row[cj[k]] = cx[k];
            int rc = orc_spg_gene_1d(size0[g], mu0[g], row, beta, C, lower, upper, maxfeval, &size_out[g - g0],
                                     counts);
            if (conv_out) conv_out[g - g0] = rc;
            if (counts_out) memcpy(&counts_out[3 * (g - g0)], counts, sizeof(counts));
        }
        free(row);
    }
    free(rp);
    free(cj);
    free(cx);
    return 0;
}

/* Golden-section + parabolic (Brent) bounded maximiser of the 1-D likelihood:
 * the "argmax truth" the survey asks to report beside the spg stopping point.
 * Not a reference code path. */
ORC_API double orc_brent_gene_1d(double mu, const double *m_observed, const double *beta, int64_t n, double lower,
                                 double upper, double xatol, int maxiter, double *fopt)
{
    const double golden = 0.3819660112501051;
    double a = lower, b = upper, x, w, v, fx, fw, fv, d = 0.0, e = 0.0;
    x = w = v = a + golden * (b - a);
    fx = fw = fv = -orc_marginal_nb_1d(x, mu, m_observed, beta, n);
    for (int it = 0; it < maxiter; it++) {
        const double xm = 0.5 * (a + b), tol1 = 1.4901161193847656e-08 * fabs(x) + xatol / 3.0, tol2 = 2.0 * tol1;
        double u, fu;
        if (fabs(x - xm) <= tol2 - 0.5 * (b - a)) break;
        int golden_step = 1;
        if (fabs(e) > tol1) {
            double r = (x - w) * (fx - fv), q = (x - v) * (fx - fw), pp = (x - v) * q - (x - w) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) pp = -pp;
            q = fabs(q);
            r = e;
            e = d;
            if (fabs(pp) < fabs(0.5 * q * r) && pp > q * (a - x) && pp < q * (b - x)) {
                d = pp / q;
                u = x + d;
                if ((u - a) < tol2 || (b - u) < tol2) d = (xm - x >= 0) ? tol1 : -tol1;
                golden_step = 0;
            }
        }
        if (golden_step) {
            e = (x >= xm) ? a - x : b - x;
            d = golden * e;
        }
        u = (fabs(d) >= tol1) ? x + d : x + ((d >= 0) ? tol1 : -tol1);
        fu = -orc_marginal_nb_1d(u, mu, m_observed, beta, n);
        if (fu <= fx) {
            if (u >= x) a = x; else b = x;
            v = w; fv = fw; w = x; fw = fx; x = u; fx = fu;
        } else {
            if (u < x) a = u; else b = u;
            if (fu <= fw || w == x) { v = w; fv = fw; w = u; fw = fu; }
            else if (fu <= fv || v == x || v == w) { v = u; fv = fu; }
        }
    }
    if (fopt) *fopt = -fx;
    return x;
}

/* ------------------------------------------------------------------------- */
/* AdjustSIZE_fun  R/sup_funs.R:27-33  (lm() = least squares; centred sums)   */
/* ------------------------------------------------------------------------- */
ORC_API int orc_adjust_size(int64_t G, const double *BB_SIZE, const double *MME_SIZE, double *out, double *coef)
{
    double mx = -INFINITY, mn = INFINITY, sx = 0, sy = 0, sxx = 0, sxy = 0, a, b;
    int64_t n = 0;
    for (int64_t g = 0; g < G; g++) {
        if (BB_SIZE[g] > mx) mx = BB_SIZE[g];
        if (BB_SIZE[g] < mn) mn = BB_SIZE[g];
    }
    for (int64_t g = 0; g < G; g++)
        if (BB_SIZE[g] < mx && BB_SIZE[g] > mn) {
            sx += log(MME_SIZE[g]);
            sy += log(BB_SIZE[g]);
            n++;
        }
    if (n > 0) {
        sx /= (double)n;
        sy /= (double)n;
    }
    for (int64_t g = 0; g < G; g++)
        if (BB_SIZE[g] < mx && BB_SIZE[g] > mn) {
            const double dxv = log(MME_SIZE[g]) - sx, dyv = log(BB_SIZE[g]) - sy;
            sxx += dxv * dxv;
            sxy += dxv * dyv;
        }
    b = sxy / sxx;
    a = sy - b * sx;
    if (coef) { coef[0] = a; coef[1] = b; }
    for (int64_t g = 0; g < G; g++) out[g] = exp(a + b * log(MME_SIZE[g]));
    return (int)n;
}

/* ------------------------------------------------------------------------- */
/* Posterior mean / mode  src/bayNorm_main.cpp:1144-1299                      */
/* ------------------------------------------------------------------------- */

/* Main_mean_NB_Bay :1144-1219; out is column-major G x C.  Loop order as the
 * reference (cell outer, gene inner); M(j,i) look-ups become a dense column. */
ORC_API void orc_post_mean(int64_t G, int64_t C, const int64_t *p, const int32_t *ri, const double *x,
                           const double *Beta, const double *size, const double *M_ave, double *out)
{
    double *col = (double *)malloc(sizeof(double) * (size_t)G);
    for (int64_t i = 0; i < C; i++) {
        memset(col, 0, sizeof(double) * (size_t)G);
        for (int64_t k = p[i]; k < p[i + 1]; k++) col[ri[k]] = x[k];
        for (int64_t j = 0; j < G; j++) {
            /* :1196 */
            double tempmu = (col[j] + size[j]) * (M_ave[j] - M_ave[j] * Beta[i]) / (size[j] + M_ave[j] * Beta[i]) + col[j];
            out[j + G * i] = tempmu;
        }
    }
    free(col);
}

/* Main_mode_NB_Bay :1226-1299 */
ORC_API void orc_post_mode(int64_t G, int64_t C, const int64_t *p, const int32_t *ri, const double *x,
                           const double *Beta, const double *size, const double *M_ave, double *out)
{
    double *col = (double *)malloc(sizeof(double) * (size_t)G);
    for (int64_t i = 0; i < C; i++) {
        memset(col, 0, sizeof(double) * (size_t)G);
        for (int64_t k = p[i]; k < p[i + 1]; k++) col[ri[k]] = x[k];
        for (int64_t j = 0; j < G; j++) {
            double r = 0.0;
            if ((size[j] + col[j]) <= 1) {
                r = 0; /* :1275-1276 */
            } else if ((size[j] + col[j]) > 1) {
                double meann = (col[j] + size[j]) * (M_ave[j] - M_ave[j] * Beta[i]) / (size[j] + M_ave[j] * Beta[i]); /* :1278 */
                r = floor(meann / (size[j] + col[j]) * ((size[j] + col[j]) - 1)) + col[j]; /* :1281 */
            }
            out[j + G * i] = r; /* NaN size+x: neither branch fires upstream and the cell keeps arma::mat's fill; 0 here */
        }
    }
    free(col);
}

/* Parameters of the 3-D draw (Main_NB_Bay :1118-1121): NB(size = size_j + x,
 * mu = tempmu).  Written out so the Python side can draw with numpy and compute
 * exact pmfs with scipy.  out_size/out_mu are column-major G x C. */
ORC_API void orc_post_sample_params(int64_t G, int64_t C, const int64_t *p, const int32_t *ri, const double *x,
                                    const double *Beta, const double *size, const double *M_ave, double *out_size,
                                    double *out_mu)
{
    double *col = (double *)malloc(sizeof(double) * (size_t)G);
    for (int64_t i = 0; i < C; i++) {
        memset(col, 0, sizeof(double) * (size_t)G);
        for (int64_t k = p[i]; k < p[i + 1]; k++) col[ri[k]] = x[k];
        for (int64_t j = 0; j < G; j++) {
            out_mu[j + G * i] = (col[j] + size[j]) * (M_ave[j] - M_ave[j] * Beta[i]) / (size[j] + M_ave[j] * Beta[i]);
            out_size[j + G * i] = size[j] + col[j];
        }
    }
    free(col);
}

/* ------------------------------------------------------------------------- */
/* A small self-contained RNG so 3-D draws and DownSampling can be timed in C  */
/* (the reference uses R's Mersenne-Twister + rgamma/rpois/rbinom; streams are */
/* not reproducible, parity is distributional -- SURVEY.md 8(c)).              */
/* ------------------------------------------------------------------------- */
typedef struct { uint64_t s[4]; } orc_rng;
static inline uint64_t orc_rotl(uint64_t v, int k) { return (v << k) | (v >> (64 - k)); }
static inline uint64_t orc_next(orc_rng *r)
{ /* xoshiro256++ */
    uint64_t *s = r->s, res = orc_rotl(s[0] + s[3], 23) + s[0], t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = orc_rotl(s[3], 45);
    return res;
}
static void orc_seed(orc_rng *r, uint64_t seed)
{ /* splitmix64 */
    for (int i = 0; i < 4; i++) {
        uint64_t z = (seed += 0x9e3779b97f4a7c15ULL);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        r->s[i] = z ^ (z >> 31);
    }
}
static inline double orc_unif(orc_rng *r) { return ((orc_next(r) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static double orc_norm(orc_rng *r)
{ /* polar Box-Muller, one value per call */
    double u, v, s;
    do {
        u = 2 * orc_unif(r) - 1;
        v = 2 * orc_unif(r) - 1;
        s = u * u + v * v;
    } while (s >= 1 || s == 0);
    return u * sqrt(-2 * log(s) / s);
}
static double orc_gamma(orc_rng *r, double a, double scale)
{ /* Marsaglia-Tsang; rgamma(a, scale) with R's scale==0 -> 0 convention */
    if (scale == 0 || a == 0) return 0.0;
    if (a < 1) return orc_gamma(r, a + 1, scale) * pow(orc_unif(r), 1.0 / a);
    {
        const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9 * d);
        for (;;) {
            double x, v, u;
            do {
                x = orc_norm(r);
                v = 1 + c * x;
            } while (v <= 0);
            v = v * v * v;
            u = orc_unif(r);
            if (u < 1 - 0.0331 * x * x * x * x) return d * v * scale;
            if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) return d * v * scale;
        }
    }
}
static double orc_pois(orc_rng *r, double lam)
{
    if (lam <= 0) return 0.0;
    if (lam < 30) { /* inversion by sequential search */
        double pk = exp(-lam), u = orc_unif(r), k = 0;
        while (u > pk) {
            u -= pk;
            k += 1;
            pk *= lam / k;
            if (k > 1000) break;
        }
        return k;
    }
    { /* PTRS, Hoermann 1993 */
        const double slam = sqrt(lam), b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b,
                     inv_alpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2), ll = log(lam);
        for (;;) {
            double U = orc_unif(r) - 0.5, V = orc_unif(r), us = 0.5 - fabs(U), k = floor((2 * a / us + b) * U + lam + 0.43);
            if (us >= 0.07 && V <= vr) return k;
            if (k < 0 || (us < 0.013 && V > us)) continue;
            if (log(V) + log(inv_alpha) - log(a / (us * us) + b) <= -lam + k * ll - lgamma(k + 1)) return k;
        }
    }
}

/* Main_NB_Bay :1063-1139.  out[G][C][S] column-major cube: j + G*(i + C*q).
 * rnbinom_mu(S, size, mu) == rpois(rgamma(size, mu/size))  [recall: Rcpp sugar]. */
ORC_API void orc_post_sample(int64_t G, int64_t C, const int64_t *p, const int32_t *ri, const double *x,
                             const double *Beta, const double *size, const double *M_ave, int S, uint64_t seed,
                             double *out)
{
    orc_rng rng;
    double *col = (double *)malloc(sizeof(double) * (size_t)G);
    orc_seed(&rng, seed);
    for (int64_t i = 0; i < C; i++) {
        memset(col, 0, sizeof(double) * (size_t)G);
        for (int64_t k = p[i]; k < p[i + 1]; k++) col[ri[k]] = x[k];
        for (int64_t j = 0; j < G; j++) {
            const double tempmu = (col[j] + size[j]) * (M_ave[j] - M_ave[j] * Beta[i]) / (size[j] + M_ave[j] * Beta[i]);
            const double sz = size[j] + col[j];
            for (int q = 0; q < S; q++)
                out[j + G * (i + C * (int64_t)q)] = orc_pois(&rng, orc_gamma(&rng, sz, tempmu / sz)) + col[j];
        }
    }
    free(col);
}

/* DownSampling src/bayNorm_main.cpp:892-915 : dense in/out, rows outer. */
ORC_API void orc_downsample(int64_t G, int64_t C, const double *Data, const double *BETA_vec, uint64_t seed,
                            double *out)
{
    orc_rng rng;
    orc_seed(&rng, seed);
    for (int64_t i = 0; i < G; i++)
        for (int64_t j = 0; j < C; j++) {
            const double n = Data[i + G * j], pr = BETA_vec[j];
            double k = 0;
            if (n > 0) {
                if (n < 64) {
                    for (int t = 0; t < (int)n; t++) k += (orc_unif(&rng) < pr);
                } else { /* inversion from 0 in log space is fine for the small p used here */
                    double q = 1 - pr, pk = pow(q, n), u = orc_unif(&rng), ratio = pr / q;
                    while (u > pk && k < n) {
                        u -= pk;
                        k += 1;
                        pk *= ratio * (n - k + 1) / k;
                    }
                }
            }
            out[i + G * j] = k;
        }
}

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
