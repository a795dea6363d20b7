/*
 * baynorm_b200_shim.c -- the R side of the drop-in boundary: a plain-C `.Call` shim that forwards
 * the reference's registered routines (src/RcppExports.cpp:211-228) and six new batched entries
 * to libbaynorm_b200.so.  NOT compiled or executed in this repository (no R toolchain in the
 * image); it is the binding a maintainer adds to the package's src/ directory.  Build:
 *     PKG_CPPFLAGS = -I<repo>/include        PKG_LIBS = -L<repo>/baynorm_b200 -lbaynorm_b200
 *
 * Conventions: the library never calls the R API; the shim allocates every result with
 * allocVector/allocMatrix and hands REAL() pointers down; non-zero status codes become Rf_error
 * after device temporaries were released (the library frees its own scratch before returning).
 * The Philox seed is drawn from R's RNG inside GetRNGstate/PutRNGstate, so set.seed() keeps two runs
 * identical (tests/testthat/test-bayNorm.r:9-17).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>

#include "baynorm_b200.h"

static bn_ctx *g_ctx = NULL;

static bn_ctx *ctx(void)
{
    if (!g_ctx) {
        int rc = bn_ctx_create(0, &g_ctx);
        if (rc != BN_OK) Rf_error("baynorm_b200: %s", bn_last_error(NULL));
    }
    return g_ctx;
}

static void check(int rc)
{
    if (rc != BN_OK) Rf_error("baynorm_b200 (%d): %s", rc, bn_last_error(g_ctx));
}

/* dgCMatrix (S4: @i, @p, @x, @Dim) -> device matrix wrapped in an external pointer with a finalizer */
static void mat_finalizer(SEXP p)
{
    bn_mat *m = (bn_mat *)R_ExternalPtrAddr(p);
    if (m) bn_mat_free(m);
    R_ClearExternalPtr(p);
}

static bn_mat *upload(SEXP Data)
{
    SEXP i = R_do_slot(Data, Rf_install("i")), p = R_do_slot(Data, Rf_install("p")),
         x = R_do_slot(Data, Rf_install("x")), dim = R_do_slot(Data, Rf_install("Dim"));
    bn_mat *m = NULL;
    check(bn_mat_from_csc(ctx(), INTEGER(dim)[0], INTEGER(dim)[1], INTEGER(p), 0, INTEGER(i), REAL(x), &m));
    return m;
}

SEXP bnb_upload(SEXP Data)
{
    SEXP ext = PROTECT(R_MakeExternalPtr(upload(Data), R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ext, mat_finalizer, TRUE);
    UNPROTECT(1);
    return ext;
}

static uint64_t seed_from_R(void)
{
    GetRNGstate();
    uint64_t hi = (uint64_t)(unif_rand() * 4294967296.0), lo = (uint64_t)(unif_rand() * 4294967296.0);
    PutRNGstate();
    return (hi << 32) | lo;
}

/* ---- the reference's registered routines, same names and arity ------------------------------- */

/* Main_mean_NB_Bay / Main_mode_NB_Bay / Main_NB_Bay (Data, BETA_vec, size, mu, S, thres) */
static SEXP posterior(SEXP Data, SEXP BETA_vec, SEXP size, SEXP mu, SEXP S, int kind)
{
    if (Rf_isNull(mu)) Rf_error("mu must be supplied");
    bn_mat *m = upload(Data);
    int64_t info[4];
    bn_mat_info(m, info);
    const int64_t G = info[0], C = info[1];
    SEXP out;
    int rc;
    if (kind == 2) {
        const int s = Rf_asInteger(S);
        out = PROTECT(Rf_alloc3DArray(REALSXP, (int)G, (int)C, s));
        rc = bn_post_sample(ctx(), m, REAL(BETA_vec), REAL(size), REAL(mu), 0, C, s, seed_from_R(), REAL(out));
    } else {
        out = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
        rc = (kind == 0 ? bn_post_mean : bn_post_mode)(ctx(), m, REAL(BETA_vec), REAL(size), REAL(mu), 0, C, REAL(out));
    }
    bn_mat_free(m);
    check(rc);
    UNPROTECT(1);
    return out;
}
SEXP _bayNorm_Main_mean_NB_Bay(SEXP D, SEXP B, SEXP s, SEXP m, SEXP S, SEXP t) { return posterior(D, B, s, m, S, 0); }
SEXP _bayNorm_Main_mode_NB_Bay(SEXP D, SEXP B, SEXP s, SEXP m, SEXP S, SEXP t) { return posterior(D, B, s, m, S, 1); }
SEXP _bayNorm_Main_NB_Bay(SEXP D, SEXP B, SEXP s, SEXP m, SEXP S, SEXP t) { return posterior(D, B, s, m, S, 2); }
/* Main_mean_NB_spBay: R wrapper calls Main_mean_NB_Bay then as(., "dgCMatrix") (src/bayNorm_main.cpp:1598) */

SEXP _bayNorm_DownSampling(SEXP Data, SEXP BETA_vec)
{
    SEXP dim = Rf_getAttrib(Data, R_DimSymbol);
    const int G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, G, C));
    check(bn_downsample(ctx(), G, C, REAL(Data), REAL(BETA_vec), seed_from_R(), REAL(out)));
    UNPROTECT(1);
    return out;
}

SEXP _bayNorm_MarginalF_NB_1D(SEXP SIZE, SEXP MU, SEXP m_observed, SEXP BETA)
{
    double out;
    check(bn_marginal_nb_1d(ctx(), Rf_asReal(SIZE), Rf_asReal(MU), REAL(m_observed), REAL(BETA), XLENGTH(m_observed), &out));
    return Rf_ScalarReal(out);
}
SEXP _bayNorm_MarginalF_NB_2D(SEXP SIZE_MU, SEXP m_observed, SEXP BETA)
{
    double out;
    check(bn_marginal_nb_2d(ctx(), REAL(SIZE_MU), REAL(m_observed), REAL(BETA), XLENGTH(m_observed), &out));
    return Rf_ScalarReal(out);
}
SEXP _bayNorm_GradientFun_NB_1D(SEXP SIZE, SEXP MU, SEXP m_observed, SEXP BETA)
{
    double out;
    check(bn_gradient_nb_1d(ctx(), Rf_asReal(SIZE), Rf_asReal(MU), REAL(m_observed), REAL(BETA), XLENGTH(m_observed), &out));
    return Rf_ScalarReal(out);
}
SEXP _bayNorm_GradientFun_NB_2D(SEXP SIZE_MU, SEXP m_observed, SEXP BETA)
{
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 2));
    check(bn_gradient_nb_2d(ctx(), REAL(SIZE_MU), REAL(m_observed), REAL(BETA), XLENGTH(m_observed), REAL(out)));
    UNPROTECT(1);
    return out;
}

SEXP _bayNorm_rowMeansFast(SEXP x)
{
    bn_mat *m = upload(x);
    int64_t info[4];
    bn_mat_info(m, info);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, info[0]));
    int rc = bn_row_means(ctx(), m, REAL(out));
    bn_mat_free(m);
    check(rc);
    UNPROTECT(1);
    return out;
}
SEXP _bayNorm_rowVarsFast(SEXP x, SEXP means)
{
    bn_mat *m = upload(x);
    int64_t info[4];
    bn_mat_info(m, info);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, info[0]));
    int rc = bn_row_vars(ctx(), m, REAL(means), REAL(out));
    bn_mat_free(m);
    check(rc);
    UNPROTECT(1);
    return out;
}

static SEXP mme_list(double *MU, double *SIZE, double *V, SEXP a, SEXP b, SEXP c)
{
    SEXP L = PROTECT(Rf_allocVector(VECSXP, 3)), nm = PROTECT(Rf_allocVector(STRSXP, 3));
    SET_VECTOR_ELT(L, 0, a);
    SET_VECTOR_ELT(L, 1, b);
    SET_VECTOR_ELT(L, 2, c);
    SET_STRING_ELT(nm, 0, Rf_mkChar("MU"));
    SET_STRING_ELT(nm, 1, Rf_mkChar("SIZE"));
    SET_STRING_ELT(nm, 2, Rf_mkChar("v"));
    Rf_setAttrib(L, R_NamesSymbol, nm);
    UNPROTECT(2);
    return L;
}
SEXP _bayNorm_EstPrior_sprcpp(SEXP Data)
{
    bn_mat *m = upload(Data);
    int64_t info[4];
    bn_mat_info(m, info);
    SEXP a = PROTECT(Rf_allocVector(REALSXP, info[0])), b = PROTECT(Rf_allocVector(REALSXP, info[0])),
         c = PROTECT(Rf_allocVector(REALSXP, info[0]));
    int rc = bn_mme(ctx(), m, NULL, REAL(a), REAL(b), REAL(c)); /* NaN where v <= m: EstPrior() applies the same rule at R/PRIOR_FUNCTIONS.R:163 */
    bn_mat_free(m);
    check(rc);
    SEXP L = mme_list(REAL(a), REAL(b), REAL(c), a, b, c);
    UNPROTECT(3);
    return L;
}
SEXP _bayNorm_EstPrior_rcpp(SEXP Data)
{
    SEXP dim = Rf_getAttrib(Data, R_DimSymbol);
    const int G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    SEXP a = PROTECT(Rf_allocVector(REALSXP, G)), b = PROTECT(Rf_allocVector(REALSXP, G)), c = PROTECT(Rf_allocVector(REALSXP, G));
    check(bn_mme_dense(ctx(), G, C, REAL(Data), REAL(a), REAL(b), REAL(c)));
    SEXP L = mme_list(REAL(a), REAL(b), REAL(c), a, b, c);
    UNPROTECT(3);
    return L;
}

/* ---- new batched entries: the R-level loops that sit on the hot path --------------------------- */

/* BB_Fun(FIX_MU = TRUE): whole loop of R/PRIOR_FUNCTIONS.R:435-477 in one call */
SEXP bnb_fit_size(SEXP Data, SEXP BETA_vec, SEXP mu0, SEXP size0, SEXP lower, SEXP upper, SEXP GR)
{
    bn_mat *m = upload(Data);
    int64_t info[4];
    bn_mat_info(m, info);
    bn_fit_opts o;
    bn_fit_opts_default(&o);
    o.use_gradient = Rf_asLogical(GR) == TRUE;
    SEXP out = PROTECT(Rf_allocVector(REALSXP, info[0]));
    int rc = bn_fit_size(ctx(), m, REAL(BETA_vec), REAL(mu0), REAL(size0), Rf_asReal(lower), Rf_asReal(upper), &o,
                         REAL(out), NULL, NULL);
    bn_mat_free(m);
    check(rc);
    UNPROTECT(1);
    return out;
}

/* Prior_fun(FIX_MU = TRUE): list(MME_MU, MME_SIZE, BB_SIZE, MME_SIZE_adjust); R assembles the data.frame */
SEXP bnb_prior(SEXP Data, SEXP BETA_vec, SEXP BB_SIZE, SEXP GR)
{
    bn_mat *m = upload(Data);
    int64_t info[4];
    bn_mat_info(m, info);
    const int bb = Rf_asLogical(BB_SIZE) == TRUE;
    SEXP L = PROTECT(Rf_allocVector(VECSXP, 4));
    for (int k = 0; k < 4; k++) SET_VECTOR_ELT(L, k, Rf_allocVector(REALSXP, (k < 2 || bb) ? info[0] : 0));
    bn_prior_out po = {0};
    po.MME_MU = REAL(VECTOR_ELT(L, 0));
    po.MME_SIZE = REAL(VECTOR_ELT(L, 1));
    po.BB_SIZE = bb ? REAL(VECTOR_ELT(L, 2)) : NULL;
    po.MME_SIZE_adjust = bb ? REAL(VECTOR_ELT(L, 3)) : NULL;
    bn_fit_opts o;
    bn_fit_opts_default(&o);
    o.use_gradient = Rf_asLogical(GR) == TRUE;
    int rc = bn_prior(ctx(), m, REAL(BETA_vec), bb, &o, &po);
    bn_mat_free(m);
    check(rc);
    UNPROTECT(1);
    return L;
}

/* BB_Fun(FIX_MU = FALSE): G x 2 matrix c('size', 'mu') of R/PRIOR_FUNCTIONS.R:457-470, :543-546 */
SEXP bnb_fit_size_mu(SEXP Data, SEXP BETA_vec, SEXP mu0, SEXP size0, SEXP lower, SEXP upper, SEXP GR)
{
    bn_mat *m = upload(Data);
    int64_t info[4];
    bn_mat_info(m, info);
    bn_fit_opts o;
    bn_fit_opts_default(&o);
    o.use_gradient = Rf_asLogical(GR) == TRUE;
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, info[0], 2)); /* column 1 = size, column 2 = mu */
    /* lower = c(SIZE_lower, MU_lower), upper = c(SIZE_upper, MU_upper) as in the parallel = TRUE branch (:414-419) */
    int rc = bn_fit_size_mu(ctx(), m, REAL(BETA_vec), REAL(mu0), REAL(size0), REAL(lower)[0], REAL(upper)[0], REAL(lower)[1],
                            REAL(upper)[1], &o, REAL(out), REAL(out) + info[0], NULL, NULL);
    bn_mat_free(m);
    check(rc);
    UNPROTECT(1);
    return out;
}

/* BetaFun (R/PRIOR_FUNCTIONS.R:30-71): list(BETA, selected) -- selected is a logical vector over the genes */
SEXP bnb_beta_fun(SEXP Data, SEXP MeanBETA)
{
    bn_mat *m = upload(Data);
    int64_t info[4];
    bn_mat_info(m, info);
    SEXP L = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(L, 0, Rf_allocVector(REALSXP, info[1]));
    SET_VECTOR_ELT(L, 1, Rf_allocVector(LGLSXP, info[0]));
    int rc = bn_beta_fun(ctx(), m, Rf_asReal(MeanBETA), REAL(VECTOR_ELT(L, 0)), (int32_t *)LOGICAL(VECTOR_ELT(L, 1)), NULL);
    bn_mat_free(m);
    check(rc);
    UNPROTECT(1);
    return L;
}

/* SyntheticControl (R/synthetic_controls.r:57-70): list(N_c, lambda) */
SEXP bnb_synthetic_control(SEXP Data, SEXP BETA_vec)
{
    bn_mat *m = upload(Data);
    int64_t info[4];
    bn_mat_info(m, info);
    GetRNGstate();
    const uint64_t seed = ((uint64_t)(unif_rand() * 4294967296.0) << 32) | (uint64_t)(unif_rand() * 4294967296.0);
    PutRNGstate();
    SEXP L = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(L, 0, Rf_allocMatrix(REALSXP, info[0], info[1]));
    SET_VECTOR_ELT(L, 1, Rf_allocVector(REALSXP, info[0]));
    int rc = bn_synthetic_control(ctx(), m, REAL(BETA_vec), seed, REAL(VECTOR_ELT(L, 1)), REAL(VECTOR_ELT(L, 0)));
    bn_mat_free(m);
    check(rc);
    UNPROTECT(1);
    return L;
}

/* default BETA_vec of R/bayNorm.r:164-170 (+ the range check of :185-187) */
SEXP bnb_beta_default(SEXP Data)
{
    bn_mat *m = upload(Data);
    int64_t info[4];
    bn_mat_info(m, info);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, info[1]));
    int rc = bn_beta_default(ctx(), m, 0.06, REAL(out));
    bn_mat_free(m);
    if (rc == BN_ERR_BETA_RANGE) Rf_error("The range of BETA must be within (0,1).");
    check(rc);
    UNPROTECT(1);
    return out;
}

static const R_CallMethodDef CallEntries[] = {
    {"_bayNorm_DownSampling", (DL_FUNC)&_bayNorm_DownSampling, 2},
    {"_bayNorm_MarginalF_NB_1D", (DL_FUNC)&_bayNorm_MarginalF_NB_1D, 4},
    {"_bayNorm_MarginalF_NB_2D", (DL_FUNC)&_bayNorm_MarginalF_NB_2D, 3},
    {"_bayNorm_GradientFun_NB_1D", (DL_FUNC)&_bayNorm_GradientFun_NB_1D, 4},
    {"_bayNorm_GradientFun_NB_2D", (DL_FUNC)&_bayNorm_GradientFun_NB_2D, 3},
    {"_bayNorm_Main_NB_Bay", (DL_FUNC)&_bayNorm_Main_NB_Bay, 6},
    {"_bayNorm_Main_mean_NB_Bay", (DL_FUNC)&_bayNorm_Main_mean_NB_Bay, 6},
    {"_bayNorm_Main_mode_NB_Bay", (DL_FUNC)&_bayNorm_Main_mode_NB_Bay, 6},
    {"_bayNorm_rowMeansFast", (DL_FUNC)&_bayNorm_rowMeansFast, 1},
    {"_bayNorm_rowVarsFast", (DL_FUNC)&_bayNorm_rowVarsFast, 2},
    {"_bayNorm_EstPrior_sprcpp", (DL_FUNC)&_bayNorm_EstPrior_sprcpp, 1},
    {"_bayNorm_EstPrior_rcpp", (DL_FUNC)&_bayNorm_EstPrior_rcpp, 1},
    {"bnb_upload", (DL_FUNC)&bnb_upload, 1},
    {"bnb_fit_size", (DL_FUNC)&bnb_fit_size, 7},
    {"bnb_prior", (DL_FUNC)&bnb_prior, 4},
    {"bnb_beta_default", (DL_FUNC)&bnb_beta_default, 1},
    {"bnb_fit_size_mu", (DL_FUNC)&bnb_fit_size_mu, 7},
    {"bnb_beta_fun", (DL_FUNC)&bnb_beta_fun, 2},
    {"bnb_synthetic_control", (DL_FUNC)&bnb_synthetic_control, 2},
    {NULL, NULL, 0}};

void R_init_bayNorm(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

void R_unload_bayNorm(DllInfo *dll)
{
    if (g_ctx) bn_ctx_destroy(g_ctx);
    g_ctx = NULL;
}
