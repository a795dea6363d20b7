/*
 * baynorm_b200_shim.c -- the R side of the drop-in boundary: a plain-C `.Call` shim that forwards
 * ALL 15 routines the reference registers (src/RcppExports.cpp:211-228, same names and arities, so the
 * unchanged R/RcppExports.R stubs resolve with R_useDynamicSymbols(dll, FALSE)) plus the batched entries
 * to libbaynorm_b200.so.  It replaces src/RcppExports.cpp + src/bayNorm_main.cpp in the package's src/:
 *     PKG_CPPFLAGS = -I<repo>/include        PKG_LIBS = -L<repo>/baynorm_b200 -lbaynorm_b200
 * There is no R toolchain in this repository's image: tests/test_abi.py type-checks this file against
 * include/baynorm_b200.h and EXECUTES it against a small stand-in for the R C API (tests/rmock/) on the GPU box.
 *
 * Conventions
 *  - The library never calls the R API.  The shim coerces every argument the way Rcpp's input_parameter<> did
 *    (integer <-> double, length checks against @Dim), allocates every result with allocVector/allocMatrix and
 *    hands REAL() pointers down.
 *  - A device matrix lives in an external pointer with a finalizer from the moment it exists, so an R allocation
 *    failure or Rf_error longjmp between upload and return cannot leak HBM; the normal path releases it eagerly.
 *  - Non-zero status codes become Rf_error AFTER the device matrix was released.
 *  - The Philox seed is drawn from R's RNG inside GetRNGstate/PutRNGstate, so set.seed() keeps two runs identical
 *    (tests/testthat/test-bayNorm.r:9-17).
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

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

/* Several GPUs from this R session (bnb_set_devices(n), called by the patched Prior_fun / bayNorm when
 * options(bayNorm.b200.devices = n) is set): the stand-in for the reference's SOCK cluster (R/PRIOR_FUNCTIONS.R:426-427).
 * The prior and the dense posterior routines then run over gene blocks on n devices (bn_group_*). */
static bn_group *g_group = NULL;

static void check_group(int rc)
{
    if (rc == BN_ERR_BETA_RANGE) Rf_error("The range of BETA must be within (0,1).");
    if (rc != BN_OK) Rf_error("baynorm_b200 (%d): %s", rc, bn_group_last_error(g_group));
}

SEXP bnb_set_devices(SEXP n)
{
    const int want = Rf_asInteger(n); /* NA_INTEGER is INT_MIN: rejected below */
    if (want < 1) Rf_error("the number of devices must be a positive integer");
    if (g_group && bn_group_size(g_group) != want) {
        bn_group_destroy(g_group);
        g_group = NULL;
    }
    if (want > 1 && !g_group) {
        int rc = bn_group_create(want, NULL, &g_group);
        if (rc != BN_OK) Rf_error("baynorm_b200 (%d): %s", rc, bn_group_last_error(NULL));
    }
    return Rf_ScalarInteger(g_group ? bn_group_size(g_group) : 1);
}

static void gmat_finalizer(SEXP p)
{
    bn_gmat *m = (bn_gmat *)R_ExternalPtrAddr(p);
    if (m) bn_gmat_free(m);
    R_ClearExternalPtr(p);
}

/* dgCMatrix -> external pointer holding the gene-sharded matrix of the current group.  Caller PROTECTs the result. */
static SEXP upload_group(SEXP Data, int64_t *G_out, int64_t *C_out)
{
    SEXP i = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("i")), INTSXP));
    SEXP p = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("p")), INTSXP));
    SEXP x = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("x")), REALSXP));
    SEXP dim = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("Dim")), INTSXP));
    if (XLENGTH(dim) != 2) Rf_error("Data@Dim must have length 2");
    const int G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    if (XLENGTH(p) != (R_xlen_t)C + 1) Rf_error("Data@p has length %ld, expected %d", (long)XLENGTH(p), C + 1);
    if (XLENGTH(i) != XLENGTH(x) || XLENGTH(i) < INTEGER(p)[C]) Rf_error("Data@i / Data@x shorter than Data@p[ncol + 1]");
    SEXP ext = PROTECT(R_MakeExternalPtr(NULL, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ext, gmat_finalizer, TRUE);
    bn_gmat *m = NULL;
    check_group(bn_group_mat_from_csc(g_group, G, C, INTEGER(p), 0, INTEGER(i), REAL(x), &m));
    R_SetExternalPtrAddr(ext, m);
    *G_out = G;
    *C_out = C;
    UNPROTECT(5);
    return ext;
}

static void check(int rc)
{
    if (rc == BN_ERR_BETA_RANGE) Rf_error("The range of BETA must be within (0,1)."); /* R/bayNorm.r:185-187 */
    if (rc != BN_OK) Rf_error("baynorm_b200 (%d): %s", rc, bn_last_error(g_ctx));
}

/* ---- argument coercion (what Rcpp::traits::input_parameter<NumericVector> did) ----------------------- */
/* returns a REALSXP (possibly x itself); the caller PROTECTs it */
static SEXP as_real(SEXP x, const char *what, R_xlen_t want_len)
{
    if (Rf_isNull(x)) Rf_error("%s must be supplied", what);
    SEXP r = Rf_isReal(x) ? x : Rf_coerceVector(x, REALSXP);
    if (want_len >= 0 && XLENGTH(r) != want_len) {
        PROTECT(r);
        UNPROTECT(1);
        Rf_error("%s has length %ld, expected %ld", what, (long)XLENGTH(r), (long)want_len);
    }
    return r;
}

/* ---- device matrices: always owned by an external pointer ---------------------------------------------- */
static void mat_finalizer(SEXP p)
{
    bn_mat *m = (bn_mat *)R_ExternalPtrAddr(p);
    if (m) bn_mat_free(m);
    R_ClearExternalPtr(p);
}

/* dgCMatrix (S4: @i, @p, @x, @Dim) -> external pointer holding the device matrix.  Caller PROTECTs the result. */
SEXP bnb_upload(SEXP Data)
{
    SEXP i = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("i")), INTSXP));
    SEXP p = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("p")), INTSXP));
    SEXP x = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("x")), REALSXP));
    SEXP dim = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("Dim")), INTSXP));
    if (XLENGTH(dim) != 2) Rf_error("Data@Dim must have length 2");
    const int G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    if (XLENGTH(p) != (R_xlen_t)C + 1) Rf_error("Data@p has length %ld, expected %d", (long)XLENGTH(p), C + 1);
    if (XLENGTH(i) != XLENGTH(x) || XLENGTH(i) < INTEGER(p)[C]) Rf_error("Data@i / Data@x shorter than Data@p[ncol + 1]");
    SEXP ext = PROTECT(R_MakeExternalPtr(NULL, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ext, mat_finalizer, TRUE);
    bn_mat *m = NULL;
    int rc = bn_mat_from_csc(ctx(), G, C, INTEGER(p), 0, INTEGER(i), REAL(x), &m);
    check(rc);
    R_SetExternalPtrAddr(ext, m);
    UNPROTECT(5);
    return ext;
}
static bn_mat *mat_of(SEXP ext) { return (bn_mat *)R_ExternalPtrAddr(ext); }
static void release(SEXP ext) { mat_finalizer(ext); } /* eager free on the normal path; the finalizer then finds NULL */

static uint64_t seed_from_R(void)
{
    GetRNGstate();
    uint64_t hi = (uint64_t)(unif_rand() * 4294967296.0), lo = (uint64_t)(unif_rand() * 4294967296.0);
    PutRNGstate();
    return (hi << 32) | lo;
}

/* a fresh dgCMatrix with the given slots (Matrix must be loaded: bayNorm imports it, DESCRIPTION:19-23) */
static SEXP new_dgCMatrix(int nrow, int ncol, SEXP i, SEXP p, SEXP x)
{
    SEXP cls = PROTECT(R_do_MAKE_CLASS("dgCMatrix"));
    SEXP obj = PROTECT(R_do_new_object(cls));
    SEXP dim = PROTECT(Rf_allocVector(INTSXP, 2));
    INTEGER(dim)[0] = nrow;
    INTEGER(dim)[1] = ncol;
    R_do_slot_assign(obj, Rf_install("i"), i);
    R_do_slot_assign(obj, Rf_install("p"), p);
    R_do_slot_assign(obj, Rf_install("x"), x);
    R_do_slot_assign(obj, Rf_install("Dim"), dim);
    UNPROTECT(3);
    return obj;
}

/* ---- the reference's registered routines, same names and arity ------------------------------- */

/* Main_mean_NB_Bay / Main_mode_NB_Bay / Main_NB_Bay (Data, BETA_vec, size, mu, S, thres): src/bayNorm_main.cpp:1063-1299.
 * thres is unused upstream as well (:1067, :1148, :1230). */
static SEXP posterior(SEXP Data, SEXP BETA_vec, SEXP size, SEXP mu, SEXP S, int kind)
{
    if (g_group) { /* gene blocks over the devices of the group */
        int64_t G, C;
        SEXP gext = PROTECT(upload_group(Data, &G, &C));
        SEXP b = PROTECT(as_real(BETA_vec, "BETA_vec", C)), sz = PROTECT(as_real(size, "size", G)), m = PROTECT(as_real(mu, "mu", G));
        const int s = kind == 2 ? Rf_asInteger(S) : 1;
        if (s <= 0) Rf_error("S must be a positive integer");
        SEXP out = PROTECT(kind == 2 ? Rf_alloc3DArray(REALSXP, (int)G, (int)C, s) : Rf_allocMatrix(REALSXP, (int)G, (int)C));
        int rc = bn_group_post(g_group, (bn_gmat *)R_ExternalPtrAddr(gext), kind, REAL(b), REAL(sz), REAL(m), 0, C, s,
                               kind == 2 ? seed_from_R() : 0, REAL(out));
        gmat_finalizer(gext);
        check_group(rc);
        UNPROTECT(5);
        return out;
    }
    SEXP ext = PROTECT(bnb_upload(Data));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    const int64_t G = info[0], C = info[1];
    SEXP b = PROTECT(as_real(BETA_vec, "BETA_vec", C)), sz = PROTECT(as_real(size, "size", G)), m = PROTECT(as_real(mu, "mu", G));
    SEXP out;
    int rc;
    if (kind == 2) {
        const int s = Rf_asInteger(S);
        if (s <= 0) Rf_error("S must be a positive integer");
        out = PROTECT(Rf_alloc3DArray(REALSXP, (int)G, (int)C, s));
        rc = bn_post_sample(ctx(), mat_of(ext), REAL(b), REAL(sz), REAL(m), 0, C, s, seed_from_R(), REAL(out));
    } else {
        out = PROTECT(Rf_allocMatrix(REALSXP, (int)G, (int)C));
        rc = (kind == 0 ? bn_post_mean : bn_post_mode)(ctx(), mat_of(ext), REAL(b), REAL(sz), REAL(m), 0, C, REAL(out));
    }
    release(ext);
    check(rc);
    UNPROTECT(5);
    return out;
}
SEXP _bayNorm_Main_mean_NB_Bay(SEXP D, SEXP B, SEXP s, SEXP m, SEXP S, SEXP t) { return posterior(D, B, s, m, S, 0); }
SEXP _bayNorm_Main_mode_NB_Bay(SEXP D, SEXP B, SEXP s, SEXP m, SEXP S, SEXP t) { return posterior(D, B, s, m, S, 1); }
SEXP _bayNorm_Main_NB_Bay(SEXP D, SEXP B, SEXP s, SEXP m, SEXP S, SEXP t) { return posterior(D, B, s, m, S, 2); }

/* Main_mean_NB_spBay (src/bayNorm_main.cpp:1519-1601): the posterior mean converted to a dgCMatrix (:1598).  The
 * conversion keeps every non-zero value, exactly like arma::sp_mat(dense): the mean is positive wherever mu > 0, so
 * the result is as dense as the matrix -- the library returns it in CSC form without a dense G x C host array
 * (bn_post_mean_csc counts, then fills). */
SEXP _bayNorm_Main_mean_NB_spBay(SEXP Data, SEXP BETA_vec, SEXP size, SEXP mu, SEXP S, SEXP thres)
{
    SEXP ext = PROTECT(bnb_upload(Data));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    const int64_t G = info[0], C = info[1];
    SEXP b = PROTECT(as_real(BETA_vec, "BETA_vec", C)), sz = PROTECT(as_real(size, "size", G)), m = PROTECT(as_real(mu, "mu", G));
    SEXP p = PROTECT(Rf_allocVector(INTSXP, C + 1));
    bn_csc_result *res = NULL;
    int rc = bn_post_mean_csc(ctx(), mat_of(ext), REAL(b), REAL(sz), REAL(m), 0, C, &res);
    release(ext);
    check(rc);
    int64_t rinfo[3];
    bn_csc_result_info(res, rinfo);
    const int64_t nnz = rinfo[2];
    if (nnz > 2147483647LL) {
        bn_csc_result_free(res);
        Rf_error("posterior mean has more than 2^31 - 1 non-zeros: a dgCMatrix cannot hold it");
    }
    /* R allocations can fail: park the device result in an external pointer first */
    SEXP rext = PROTECT(R_MakeExternalPtr(res, R_NilValue, R_NilValue));
    SEXP i = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)nnz)), x = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)nnz));
    rc = bn_csc_result_fetch(ctx(), res, NULL, INTEGER(p), INTEGER(i), REAL(x));
    bn_csc_result_free(res);
    R_ClearExternalPtr(rext);
    check(rc);
    SEXP obj = new_dgCMatrix((int)G, (int)C, i, p, x);
    UNPROTECT(8);
    return obj;
}

/* DownSampling(Data, BETA_vec)  src/bayNorm_main.cpp:892-915: dense (integer or double) matrix in, double matrix out */
SEXP _bayNorm_DownSampling(SEXP Data, SEXP BETA_vec)
{
    const int G = Rf_nrows(Data), C = Rf_ncols(Data);
    SEXP d = PROTECT(as_real(Data, "Data", (R_xlen_t)G * C)), b = PROTECT(as_real(BETA_vec, "BETA_vec", C));
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, G, C));
    check(bn_downsample(ctx(), G, C, REAL(d), REAL(b), seed_from_R(), REAL(out)));
    UNPROTECT(3);
    return out;
}

SEXP _bayNorm_MarginalF_NB_1D(SEXP SIZE, SEXP MU, SEXP m_observed, SEXP BETA)
{
    SEXP mo = PROTECT(as_real(m_observed, "m_observed", -1)), b = PROTECT(as_real(BETA, "BETA", XLENGTH(m_observed)));
    double out;
    check(bn_marginal_nb_1d(ctx(), Rf_asReal(SIZE), Rf_asReal(MU), REAL(mo), REAL(b), XLENGTH(mo), &out));
    UNPROTECT(2);
    return Rf_ScalarReal(out);
}
SEXP _bayNorm_MarginalF_NB_2D(SEXP SIZE_MU, SEXP m_observed, SEXP BETA)
{
    SEXP sm = PROTECT(as_real(SIZE_MU, "SIZE_MU", 2)), mo = PROTECT(as_real(m_observed, "m_observed", -1)),
         b = PROTECT(as_real(BETA, "BETA", XLENGTH(m_observed)));
    double out;
    check(bn_marginal_nb_2d(ctx(), REAL(sm), REAL(mo), REAL(b), XLENGTH(mo), &out));
    UNPROTECT(3);
    return Rf_ScalarReal(out);
}
SEXP _bayNorm_GradientFun_NB_1D(SEXP SIZE, SEXP MU, SEXP m_observed, SEXP BETA)
{
    SEXP mo = PROTECT(as_real(m_observed, "m_observed", -1)), b = PROTECT(as_real(BETA, "BETA", XLENGTH(m_observed)));
    double out;
    check(bn_gradient_nb_1d(ctx(), Rf_asReal(SIZE), Rf_asReal(MU), REAL(mo), REAL(b), XLENGTH(mo), &out));
    UNPROTECT(2);
    return Rf_ScalarReal(out);
}
SEXP _bayNorm_GradientFun_NB_2D(SEXP SIZE_MU, SEXP m_observed, SEXP BETA)
{
    SEXP sm = PROTECT(as_real(SIZE_MU, "SIZE_MU", 2)), mo = PROTECT(as_real(m_observed, "m_observed", -1)),
         b = PROTECT(as_real(BETA, "BETA", XLENGTH(m_observed)));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, 2));
    check(bn_gradient_nb_2d(ctx(), REAL(sm), REAL(mo), REAL(b), XLENGTH(mo), REAL(out)));
    UNPROTECT(4);
    return out;
}

SEXP _bayNorm_rowMeansFast(SEXP x)
{
    SEXP ext = PROTECT(bnb_upload(x));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, info[0]));
    int rc = bn_row_means(ctx(), mat_of(ext), REAL(out));
    release(ext);
    check(rc);
    UNPROTECT(2);
    return out;
}
SEXP _bayNorm_rowVarsFast(SEXP x, SEXP means)
{
    SEXP ext = PROTECT(bnb_upload(x));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    SEXP mv = PROTECT(as_real(means, "means", info[0]));
    SEXP out = PROTECT(Rf_allocVector(REALSXP, info[0]));
    int rc = bn_row_vars(ctx(), mat_of(ext), REAL(mv), REAL(out));
    release(ext);
    check(rc);
    UNPROTECT(3);
    return out;
}

static SEXP mme_list(SEXP a, SEXP b, SEXP c)
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
/* EstPrior_sprcpp (src/bayNorm_main.cpp:1370-1394): list(MU, SIZE, v) with SIZE the RAW quotient m^2/(v - m); the NaN
 * rule lives in R's EstPrior (R/PRIOR_FUNCTIONS.R:163), exactly as for the dense twin below. */
SEXP _bayNorm_EstPrior_sprcpp(SEXP Data)
{
    SEXP ext = PROTECT(bnb_upload(Data));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    SEXP a = PROTECT(Rf_allocVector(REALSXP, info[0])), b = PROTECT(Rf_allocVector(REALSXP, info[0])),
         c = PROTECT(Rf_allocVector(REALSXP, info[0]));
    int rc = bn_mme_raw(ctx(), mat_of(ext), NULL, REAL(a), REAL(b), REAL(c));
    release(ext);
    check(rc);
    SEXP L = mme_list(a, b, c);
    UNPROTECT(4);
    return L;
}
SEXP _bayNorm_EstPrior_rcpp(SEXP Data)
{
    const int G = Rf_nrows(Data), C = Rf_ncols(Data);
    SEXP d = PROTECT(as_real(Data, "Data", (R_xlen_t)G * C));
    SEXP a = PROTECT(Rf_allocVector(REALSXP, G)), b = PROTECT(Rf_allocVector(REALSXP, G)), c = PROTECT(Rf_allocVector(REALSXP, G));
    check(bn_mme_dense(ctx(), G, C, REAL(d), REAL(a), REAL(b), REAL(c)));
    SEXP L = mme_list(a, b, c);
    UNPROTECT(4);
    return L;
}

/* t_sp (src/bayNorm_main.cpp:1472-1477): transpose of a sparse matrix.  Off the hot path (the library builds its own
 * gene-major twin on the device); a host counting transpose keeps the exported function working. */
SEXP _bayNorm_t_sp(SEXP Data)
{
    SEXP i = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("i")), INTSXP));
    SEXP p = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("p")), INTSXP));
    SEXP x = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("x")), REALSXP));
    SEXP dim = PROTECT(Rf_coerceVector(R_do_slot(Data, Rf_install("Dim")), INTSXP));
    const int G = INTEGER(dim)[0], C = INTEGER(dim)[1];
    const int nnz = INTEGER(p)[C];
    SEXP tp = PROTECT(Rf_allocVector(INTSXP, (R_xlen_t)G + 1)), ti = PROTECT(Rf_allocVector(INTSXP, nnz)),
         tx = PROTECT(Rf_allocVector(REALSXP, nnz));
    int *rp = INTEGER(tp);
    memset(rp, 0, sizeof(int) * ((size_t)G + 1));
    for (int k = 0; k < nnz; k++) rp[INTEGER(i)[k] + 1]++;
    for (int g = 0; g < G; g++) rp[g + 1] += rp[g];
    SEXP cur = PROTECT(Rf_allocVector(INTSXP, G > 0 ? G : 1));
    memcpy(INTEGER(cur), rp, sizeof(int) * (size_t)G);
    for (int j = 0; j < C; j++)
        for (int k = INTEGER(p)[j]; k < INTEGER(p)[j + 1]; k++) {
            const int q = INTEGER(cur)[INTEGER(i)[k]]++;
            INTEGER(ti)[q] = j;
            REAL(tx)[q] = REAL(x)[k];
        }
    SEXP obj = new_dgCMatrix(C, G, ti, tp, tx);
    UNPROTECT(8);
    return obj;
}

/* asMatrix(rp, cp, z, nrows, ncols) (src/bayNorm_main.cpp:1497-1513): IntegerMatrix with mat(rp[i], cp[i]) = z[i] */
SEXP _bayNorm_asMatrix(SEXP rp, SEXP cp, SEXP z, SEXP nrows, SEXP ncols)
{
    const int nr = Rf_asInteger(nrows), nc = Rf_asInteger(ncols);
    if (nr < 0 || nc < 0) Rf_error("negative extents to matrix");
    SEXP r = PROTECT(as_real(rp, "rp", -1)), c = PROTECT(as_real(cp, "cp", XLENGTH(rp))), v = PROTECT(as_real(z, "z", XLENGTH(rp)));
    SEXP mat = PROTECT(Rf_allocMatrix(INTSXP, nr, nc));
    memset(INTEGER(mat), 0, sizeof(int) * (size_t)nr * (size_t)nc);
    const R_xlen_t k = XLENGTH(v);
    for (R_xlen_t t = 0; t < k; t++) {
        const double ri = REAL(r)[t], ci = REAL(c)[t];
        if (!(ri >= 0 && ri < nr && ci >= 0 && ci < nc)) Rf_error("index out of bounds");
        INTEGER(mat)[(R_xlen_t)ri + (R_xlen_t)nr * (R_xlen_t)ci] = (int)REAL(v)[t];
    }
    UNPROTECT(4);
    return mat;
}

/* ---- new batched entries: the R-level loops that sit on the hot path --------------------------- */

/* BB_Fun(FIX_MU = TRUE): whole loop of R/PRIOR_FUNCTIONS.R:435-477 in one call */
SEXP bnb_fit_size(SEXP Data, SEXP BETA_vec, SEXP mu0, SEXP size0, SEXP lower, SEXP upper, SEXP GR)
{
    SEXP ext = PROTECT(bnb_upload(Data));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    SEXP b = PROTECT(as_real(BETA_vec, "BETA_vec", info[1])), m0 = PROTECT(as_real(mu0, "INITIAL_MU_vec", info[0])),
         s0 = PROTECT(as_real(size0, "INITIAL_SIZE_vec", info[0]));
    bn_fit_opts o;
    bn_fit_opts_default(&o);
    o.use_gradient = Rf_asLogical(GR) == TRUE;
    SEXP out = PROTECT(Rf_allocVector(REALSXP, info[0]));
    int rc = bn_fit_size(ctx(), mat_of(ext), REAL(b), REAL(m0), REAL(s0), Rf_asReal(lower), Rf_asReal(upper), &o, REAL(out), NULL, NULL);
    release(ext);
    check(rc);
    UNPROTECT(5);
    return out;
}

/* Prior_fun(FIX_MU = TRUE): list(MME_MU, MME_SIZE, BB_SIZE, MME_SIZE_adjust); R assembles the data.frame */
SEXP bnb_prior(SEXP Data, SEXP BETA_vec, SEXP BB_SIZE, SEXP GR)
{
    if (g_group) { /* gene blocks over the devices of the group */
        int64_t G, C;
        SEXP gext = PROTECT(upload_group(Data, &G, &C));
        SEXP b = PROTECT(as_real(BETA_vec, "BETA_vec", C));
        const int bb = Rf_asLogical(BB_SIZE) == TRUE;
        SEXP L = PROTECT(Rf_allocVector(VECSXP, 4));
        for (int k = 0; k < 4; k++) SET_VECTOR_ELT(L, k, Rf_allocVector(REALSXP, (k < 2 || bb) ? G : 0));
        bn_prior_out po;
        memset(&po, 0, sizeof(po));
        po.MME_MU = REAL(VECTOR_ELT(L, 0));
        po.MME_SIZE = REAL(VECTOR_ELT(L, 1));
        po.BB_SIZE = bb ? REAL(VECTOR_ELT(L, 2)) : NULL;
        po.MME_SIZE_adjust = bb ? REAL(VECTOR_ELT(L, 3)) : NULL;
        bn_fit_opts o;
        bn_fit_opts_default(&o);
        o.use_gradient = Rf_asLogical(GR) == TRUE;
        int rc = bn_group_prior(g_group, (bn_gmat *)R_ExternalPtrAddr(gext), REAL(b), bb, &o, &po);
        gmat_finalizer(gext);
        check_group(rc);
        UNPROTECT(3);
        return L;
    }
    SEXP ext = PROTECT(bnb_upload(Data));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    SEXP b = PROTECT(as_real(BETA_vec, "BETA_vec", info[1]));
    const int bb = Rf_asLogical(BB_SIZE) == TRUE;
    SEXP L = PROTECT(Rf_allocVector(VECSXP, 4));
    for (int k = 0; k < 4; k++) SET_VECTOR_ELT(L, k, Rf_allocVector(REALSXP, (k < 2 || bb) ? info[0] : 0));
    bn_prior_out po;
    memset(&po, 0, sizeof(po));
    po.MME_MU = REAL(VECTOR_ELT(L, 0));
    po.MME_SIZE = REAL(VECTOR_ELT(L, 1));
    po.BB_SIZE = bb ? REAL(VECTOR_ELT(L, 2)) : NULL;
    po.MME_SIZE_adjust = bb ? REAL(VECTOR_ELT(L, 3)) : NULL;
    bn_fit_opts o;
    bn_fit_opts_default(&o);
    o.use_gradient = Rf_asLogical(GR) == TRUE;
    int rc = bn_prior(ctx(), mat_of(ext), REAL(b), bb, &o, &po);
    release(ext);
    check(rc);
    UNPROTECT(3);
    return L;
}

/* BB_Fun(FIX_MU = FALSE): G x 2 matrix c('size', 'mu') of R/PRIOR_FUNCTIONS.R:457-470, :543-546.
 * lower = c(SIZE_lower, MU_lower), upper = c(SIZE_upper, MU_upper): the box of the parallel = TRUE branch (:414-419);
 * the R wrapper passes rep(SIZE_lower, 2) / rep(SIZE_upper, 2) for parallel = FALSE, which is what :525-526 hands spg */
SEXP bnb_fit_size_mu(SEXP Data, SEXP BETA_vec, SEXP mu0, SEXP size0, SEXP lower, SEXP upper, SEXP GR)
{
    SEXP ext = PROTECT(bnb_upload(Data));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    SEXP b = PROTECT(as_real(BETA_vec, "BETA_vec", info[1])), m0 = PROTECT(as_real(mu0, "INITIAL_MU_vec", info[0])),
         s0 = PROTECT(as_real(size0, "INITIAL_SIZE_vec", info[0])), lo = PROTECT(as_real(lower, "lower", 2)),
         hi = PROTECT(as_real(upper, "upper", 2));
    bn_fit_opts o;
    bn_fit_opts_default(&o);
    o.use_gradient = Rf_asLogical(GR) == TRUE;
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, info[0], 2)); /* column 1 = size, column 2 = mu */
    int rc = bn_fit_size_mu(ctx(), mat_of(ext), REAL(b), REAL(m0), REAL(s0), REAL(lo)[0], REAL(hi)[0], REAL(lo)[1], REAL(hi)[1], &o,
                            REAL(out), REAL(out) + info[0], NULL, NULL);
    release(ext);
    check(rc);
    UNPROTECT(7);
    return out;
}

/* BetaFun (R/PRIOR_FUNCTIONS.R:30-71): list(BETA, selected) -- selected is a logical vector over the genes */
SEXP bnb_beta_fun(SEXP Data, SEXP MeanBETA)
{
    SEXP ext = PROTECT(bnb_upload(Data));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    SEXP L = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(L, 0, Rf_allocVector(REALSXP, info[1]));
    SET_VECTOR_ELT(L, 1, Rf_allocVector(LGLSXP, info[0]));
    int rc = bn_beta_fun(ctx(), mat_of(ext), Rf_asReal(MeanBETA), REAL(VECTOR_ELT(L, 0)), (int32_t *)LOGICAL(VECTOR_ELT(L, 1)), NULL);
    release(ext);
    check(rc);
    UNPROTECT(2);
    return L;
}

/* SyntheticControl (R/synthetic_controls.r:57-70): list(N_c, lambda) */
SEXP bnb_synthetic_control(SEXP Data, SEXP BETA_vec)
{
    SEXP ext = PROTECT(bnb_upload(Data));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    SEXP b = PROTECT(as_real(BETA_vec, "BETA_vec", info[1]));
    const uint64_t seed = seed_from_R();
    SEXP L = PROTECT(Rf_allocVector(VECSXP, 2));
    SET_VECTOR_ELT(L, 0, Rf_allocMatrix(REALSXP, info[0], info[1]));
    SET_VECTOR_ELT(L, 1, Rf_allocVector(REALSXP, info[0]));
    int rc = bn_synthetic_control(ctx(), mat_of(ext), REAL(b), seed, REAL(VECTOR_ELT(L, 1)), REAL(VECTOR_ELT(L, 0)));
    release(ext);
    check(rc);
    UNPROTECT(3);
    return L;
}

/* default BETA_vec of R/bayNorm.r:164-170 (+ the range check of :185-187) */
SEXP bnb_beta_default(SEXP Data)
{
    SEXP ext = PROTECT(bnb_upload(Data));
    int64_t info[4];
    bn_mat_info(mat_of(ext), info);
    SEXP out = PROTECT(Rf_allocVector(REALSXP, info[1]));
    int rc = bn_beta_default(ctx(), mat_of(ext), 0.06, REAL(out));
    release(ext);
    check(rc);
    UNPROTECT(2);
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
    {"_bayNorm_t_sp", (DL_FUNC)&_bayNorm_t_sp, 1},
    {"_bayNorm_asMatrix", (DL_FUNC)&_bayNorm_asMatrix, 5},
    {"_bayNorm_Main_mean_NB_spBay", (DL_FUNC)&_bayNorm_Main_mean_NB_spBay, 6},
    {"bnb_upload", (DL_FUNC)&bnb_upload, 1},
    {"bnb_fit_size", (DL_FUNC)&bnb_fit_size, 7},
    {"bnb_prior", (DL_FUNC)&bnb_prior, 4},
    {"bnb_beta_default", (DL_FUNC)&bnb_beta_default, 1},
    {"bnb_fit_size_mu", (DL_FUNC)&bnb_fit_size_mu, 7},
    {"bnb_beta_fun", (DL_FUNC)&bnb_beta_fun, 2},
    {"bnb_synthetic_control", (DL_FUNC)&bnb_synthetic_control, 2},
    {"bnb_set_devices", (DL_FUNC)&bnb_set_devices, 1},
    {NULL, NULL, 0}};

void R_init_bayNorm(DllInfo *dll)
{
    R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
    R_useDynamicSymbols(dll, FALSE);
}

void R_unload_bayNorm(DllInfo *dll)
{
    if (g_group) bn_group_destroy(g_group);
    g_group = NULL;
    if (g_ctx) bn_ctx_destroy(g_ctx);
    g_ctx = NULL;
}
