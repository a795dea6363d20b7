/*
 * baynorm_b200.h -- C ABI of the B200-native bayNorm hot path.
 *
 * One shared library (libbaynorm_b200.so, sm_100a only) replaces the native layer
 * of WT215/bayNorm (src/bayNorm_main.cpp, reached today through the 15 .Call
 * routines registered at src/RcppExports.cpp:211-228) plus the R-level per-gene
 * optimiser loop (BB_Fun, R/PRIOR_FUNCTIONS.R:380-549) and the prior orchestration
 * (Prior_fun, R/PRIOR_FUNCTIONS.R:241-311).  Every entry point below names the
 * reference interface it stands in for.  INTEGRATION.md shows the .Call shim a
 * maintainer adds on the R side.
 *
 * Conventions
 *  - extern "C", plain pointers and sizes; no R, Rcpp, torch or CUDA types.
 *  - Every data pointer may be HOST or DEVICE memory; the library inspects the
 *    pointer (cudaPointerGetAttributes) and stages host buffers itself.  Pinned
 *    host memory is copied at PCIe rate, pageable memory more slowly.
 *  - Dense matrices are column-major (R layout): element (gene i, cell j) of a
 *    G x C matrix is at [i + G*j].  3-D output is the reference's arma::cube
 *    layout dim=c(G,C,S): [i + G*(j + C*s)]  (src/bayNorm_main.cpp:1091,1123).
 *  - Sparse input is dgCMatrix-style CSC: p[C+1] column pointers (int32 as in
 *    R's @p, or int64 for > 2^31-1 non-zeros), i[nnz] 0-based int32 row indices
 *    sorted within each column, x[nnz] double  (R/sup_funs.R:99-106).
 *  - All entry points return 0 on success or a BN_ERR_* code; bn_last_error(ctx)
 *    gives a message.  No exceptions, no longjmp, no R API calls.  There is no
 *    CPU fallback: without a usable sm_100 device bn_ctx_create fails.
 *  - A context is driven by one host thread at a time.  One context = one GPU;
 *    multi-GPU runs use one process (or one context) per GPU, each holding a
 *    gene block of the matrix, and exchange the few cross-shard quantities through
 *    the all-reduce hook (bn_ctx_set_allreduce).
 *  - The library never keeps a host pointer after a call returns.
 */
#ifndef BAYNORM_B200_H
#define BAYNORM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define BN_API __attribute__((visibility("default")))
#else
#define BN_API
#endif

#define BN_VERSION 100 /* 0.1.0 */

/* ---- status codes --------------------------------------------------------- */
enum {
    BN_OK = 0,
    BN_ERR_INVALID = 1,   /* bad argument (NULL, negative size, unsorted rows, ...) */
    BN_ERR_CUDA = 2,      /* CUDA runtime error; message has the cudaError string */
    BN_ERR_NO_DEVICE = 3, /* no sm_100 GPU: there is no CPU fallback */
    BN_ERR_OOM = 4,       /* device or host allocation failed */
    BN_ERR_BETA_RANGE = 5,/* "The range of BETA must be within (0,1)." R/bayNorm.r:185-187 */
    BN_ERR_NONCOUNT = 6,  /* likelihood fit needs non-negative integer counts (dnbinom_mu) */
    BN_ERR_COLLECTIVE = 7,/* the all-reduce hook returned non-zero */
    BN_ERR_UNSUPPORTED = 8
};

typedef struct bn_ctx bn_ctx; /* one GPU: stream, workspace, last error */
typedef struct bn_mat bn_mat; /* device-resident gene x cell matrix (CSC + gene-major CSR built on demand) */

/* ---- context --------------------------------------------------------------- */
BN_API int bn_version(void);
BN_API int bn_ctx_create(int device, bn_ctx **out);
BN_API void bn_ctx_destroy(bn_ctx *ctx);
BN_API const char *bn_last_error(const bn_ctx *ctx); /* ctx may be NULL: last creation error */
BN_API int bn_ctx_synchronize(bn_ctx *ctx);
/* Opaque cudaStream_t the context launches on (so callers can time it with events). */
BN_API void *bn_ctx_stream(bn_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
BN_API int64_t bn_ctx_launch_count(const bn_ctx *ctx);
/* Device properties the bench reports: out[0]=SM count, [1]=clock kHz, [2]=cc major*10+minor, [3]=total bytes */
BN_API int bn_ctx_device_info(bn_ctx *ctx, int64_t out[4]);

/* Cross-shard all-reduce hook.  dtype: 0 = float64.  op: 0 = sum, 1 = min, 2 = max.
 * buf is DEVICE memory of `count` elements, reduced in place across all shards; the
 * work must be ordered after everything already enqueued on `stream` (cudaStream_t)
 * and complete (or be stream-ordered) before the hook returns.  With world == 1 (the
 * default) the hook is never called.  Replaces nothing in the reference (its only
 * parallelism is a SOCK cluster, R/PRIOR_FUNCTIONS.R:426-427); this is the one
 * exchange step of the gene-sharded path. */
typedef int (*bn_allreduce_fn)(void *user, void *buf, int64_t count, int dtype, int op, void *stream);
BN_API int bn_ctx_set_allreduce(bn_ctx *ctx, bn_allreduce_fn fn, void *user, int rank, int world);

/* ---- matrices ---------------------------------------------------------------- */
/* dgCMatrix in.  p_is_int64: 0 -> p is int32[C+1] (R's @p), 1 -> int64[C+1].
 * Stands in for `as<arma::sp_mat>(Data)` in every Main_* wrapper
 * (src/RcppExports.cpp:76-90) -- but the copy lands in HBM once and is reused. */
BN_API int bn_mat_from_csc(bn_ctx *ctx, int64_t G, int64_t C, const void *p, int p_is_int64, const int32_t *i,
                           const double *x, bn_mat **out);
/* Dense column-major double G x C in (what Check_input turns into a dgCMatrix at
 * R/sup_funs.R:105); explicit zeros are dropped on the device. */
BN_API int bn_mat_from_dense(bn_ctx *ctx, int64_t G, int64_t C, const double *dense, bn_mat **out);
BN_API void bn_mat_free(bn_mat *m);
/* out[0]=G, [1]=C, [2]=nnz, [3]=1 if all values are non-negative integers */
BN_API int bn_mat_info(const bn_mat *m, int64_t out[4]);
/* Position of this block inside the whole matrix (gene-sharded / column-blocked runs):
 * gene0 = global index of local gene 0, cell0 likewise.  Only the RNG counters use it, so
 * that sample (gene, cell, s) does not depend on how the matrix was split. */
BN_API int bn_mat_set_origin(bn_mat *m, int64_t gene0, int64_t cell0);
/* Copy the CSC arrays back (p as int64).  Any pointer may be NULL. */
BN_API int bn_mat_export_csc(bn_ctx *ctx, const bn_mat *m, int64_t *p, int32_t *i, double *x);
/* New matrix holding columns idx[0..n) of m (Data[, which(Conditions == level)],
 * R/bayNorm.r:265-270). idx is host or device int64. */
BN_API int bn_mat_select_cols(bn_ctx *ctx, const bn_mat *m, const int64_t *idx, int64_t n, bn_mat **out);
/* Synthetic NB counts x_ij ~ NB(size phi_i, mean mu_i*beta_j) generated on the device
 * straight into CSC (SURVEY.md 8(d) generator; bench/test input, not a reference API). */
BN_API int bn_mat_synth_nb(bn_ctx *ctx, int64_t G, int64_t C, const double *mu, const double *size,
                           const double *beta, uint64_t seed, int64_t gene0, bn_mat **out);

/* ---- beta / library size  (R/bayNorm.r:164-170) ------------------------------ */
/* colsums[C] = Matrix::colSums(Data) of this shard (no collective). */
BN_API int bn_colsums(bn_ctx *ctx, const bn_mat *m, double *colsums);
/* beta[C] = colsums/mean(colsums)*mean_beta with the two clamps, then the (0,1) range
 * check (BN_ERR_BETA_RANGE).  Runs the all-reduce hook on the column sums when world > 1. */
BN_API int bn_beta_default(bn_ctx *ctx, const bn_mat *m, double mean_beta, double *beta);

/* ---- method-of-moments prior  (EstPrior_sprcpp, src/bayNorm_main.cpp:1370-1394;
 *      rowMeansFast :1303, rowVarsFast :1323; NaN rule R/PRIOR_FUNCTIONS.R:163) ---- */
/* beta may be NULL (data already normalised, as EstPrior_sprcpp is called upstream) or
 * beta[C]: the x/beta normalisation of R/PRIOR_FUNCTIONS.R:253 is then fused in.
 * MU[G], SIZE[G] (NaN where v <= m, i.e. after the rule at :163), V[G]; any may be NULL. */
BN_API int bn_mme(bn_ctx *ctx, bn_mat *m, const double *beta, double *MU, double *SIZE, double *V);
/* As bn_mme, but SIZE is the raw quotient m^2 / (v - m) that EstPrior_sprcpp itself returns (src/bayNorm_main.cpp:1386);
 * the NaN rule is applied by R's EstPrior afterwards (R/PRIOR_FUNCTIONS.R:163). */
BN_API int bn_mme_raw(bn_ctx *ctx, bn_mat *m, const double *beta, double *MU, double *SIZE, double *V);
/* rowMeansFast / rowVarsFast as separate calls (registered .Call routines). */
BN_API int bn_row_means(bn_ctx *ctx, bn_mat *m, double *means);
BN_API int bn_row_vars(bn_ctx *ctx, bn_mat *m, const double *means, double *vars);
/* EstPrior_rcpp (dense twin, src/bayNorm_main.cpp:1423-1449). */
BN_API int bn_mme_dense(bn_ctx *ctx, int64_t G, int64_t C, const double *dense, double *MU, double *SIZE, double *V);

/* ---- likelihood / gradient for one gene (registered .Call routines) ------------ */
/* MarginalF_NB_1D(SIZE, MU, m_observed, BETA)  src/bayNorm_main.cpp:928-943 */
BN_API int bn_marginal_nb_1d(bn_ctx *ctx, double SIZE, double MU, const double *m_observed, const double *BETA,
                             int64_t n, double *out);
/* MarginalF_NB_2D(SIZE_MU, m_observed, BETA)  :948-965 */
BN_API int bn_marginal_nb_2d(bn_ctx *ctx, const double SIZE_MU[2], const double *m_observed, const double *BETA,
                             int64_t n, double *out);
/* GradientFun_NB_1D  :972-987 */
BN_API int bn_gradient_nb_1d(bn_ctx *ctx, double SIZE, double MU, const double *m_observed, const double *BETA,
                             int64_t n, double *out);
/* GradientFun_NB_2D  :993-1013; out[2] = {d/dsize, d/dmu} */
BN_API int bn_gradient_nb_2d(bn_ctx *ctx, const double SIZE_MU[2], const double *m_observed, const double *BETA,
                             int64_t n, double out[2]);

/* ---- per-gene maximum marginal likelihood (BB_Fun, R/PRIOR_FUNCTIONS.R:380-549) - */
enum { BN_FIT_SPG = 0, BN_FIT_BRENT = 1 };
typedef struct {
    int32_t method;    /* BN_FIT_SPG: BB::spg restated (reference-faithful, default);
                          BN_FIT_BRENT: bounded golden-section/parabolic argmax */
    int32_t maxfeval;  /* 500  (R/PRIOR_FUNCTIONS.R:451-454) */
    int32_t maxit;     /* 1500 (BB default) */
    int32_t M;         /* 10   (BB default, nonmonotone window) */
    double ftol;       /* 1e-10 */
    double gtol;       /* 1e-5 */
    double eps;        /* 1e-7 forward-difference step */
    double xatol;      /* 1e-10 (Brent only) */
    int32_t use_gradient; /* GR=TRUE: analytic GradientFun_NB_1D instead of differences */
    int32_t park_after;   /* long genes (C >= 8192): evaluations after which a gene still running is handed from the
                             sweep kernel to the cluster kernel; 0 = default (128).  Scheduling only: the solver state
                             travels with the gene */
} bn_fit_opts;
BN_API void bn_fit_opts_default(bn_fit_opts *o);
/* FIX_MU=TRUE fit of `size` for every gene of the shard: start INITIAL_SIZE_vec[G], mean
 * INITIAL_MU_vec[G] fixed, box [SIZE_lower, SIZE_upper].  Outputs (any may be NULL):
 * size_out[G] (BB_opt$par), conv[G] (spg convergence code), nfeval[G] (likelihood
 * evaluations incl. finite differences).  `parallel`/`NCores` have no equivalent. */
BN_API int bn_fit_size(bn_ctx *ctx, bn_mat *m, const double *BETA_vec, const double *INITIAL_MU_vec,
                       const double *INITIAL_SIZE_vec, double SIZE_lower, double SIZE_upper, const bn_fit_opts *opts,
                       double *size_out, int32_t *conv, int32_t *nfeval);
/* FIX_MU=FALSE (R/PRIOR_FUNCTIONS.R:414-419, :457-470): BB::spg over par = (size, mu) for every gene,
 * start (INITIAL_SIZE_vec, INITIAL_MU_vec), box [SIZE_lower, SIZE_upper] x [MU_lower, MU_upper]
 * (the bounds of the reference's parallel = TRUE branch), objective MarginalF_NB_2D, gradient by
 * forward differences or GradientFun_NB_2D (opts->use_gradient).  size_out[G], mu_out[G] = BB_opt$par. */
BN_API int bn_fit_size_mu(bn_ctx *ctx, bn_mat *m, const double *BETA_vec, const double *INITIAL_MU_vec,
                          const double *INITIAL_SIZE_vec, double SIZE_lower, double SIZE_upper, double MU_lower,
                          double MU_upper, const bn_fit_opts *opts, double *size_out, double *mu_out, int32_t *conv,
                          int32_t *nfeval);
/* min / max over the non-NaN entries of a vector, across all gene shards (through the all-reduce hook). */
BN_API int bn_vec_range(bn_ctx *ctx, const double *v, int64_t n, double out[2]);

/* ---- AdjustSIZE_fun  (R/sup_funs.R:27-33) --------------------------------------- */
/* out[G] = exp(a + b*log(MME_SIZE)), (a,b) = OLS of log BB_SIZE on log MME_SIZE over genes
 * strictly inside (min BB, max BB).  coef (may be NULL) receives {a, b, n_fit}.  Uses the
 * all-reduce hook for min/max and the five sums when world > 1. */
BN_API int bn_adjust_size(bn_ctx *ctx, int64_t G, const double *BB_SIZE, const double *MME_SIZE, double *out,
                          double coef[3]);

/* ---- Prior_fun  (R/PRIOR_FUNCTIONS.R:241-311), FIX_MU=TRUE ------------------------ */
typedef struct {
    double *MME_MU;          /* [G] */
    double *MME_SIZE;        /* [G] NaN already replaced by the global min (:259) */
    double *BB_SIZE;         /* [G] NULL unless BB_SIZE requested */
    double *MME_SIZE_adjust; /* [G] NULL unless BB_SIZE requested */
    int32_t *conv;           /* [G] optional */
    int32_t *nfeval;         /* [G] optional */
    double size_lower, size_upper; /* bounds handed to the fit (:277-279) */
    double coef[3];          /* a, b, n_fit of AdjustSIZE_fun */
    double seconds[4];       /* device time: MME, fit, adjust, total */
} bn_prior_out;
BN_API int bn_prior(bn_ctx *ctx, bn_mat *m, const double *BETA_vec, int BB_SIZE, const bn_fit_opts *opts,
                    bn_prior_out *out);

/* ---- posterior  (Main_*_NB_Bay, src/bayNorm_main.cpp:1063-1299, :1519-1601) ------- */
/* Columns [col0, col0+ncol) of the matrix; out is column-major G x ncol.  S/thres of the
 * reference signature are unused by mean/mode upstream and have no parameter here. */
BN_API int bn_post_mean(bn_ctx *ctx, const bn_mat *m, const double *BETA_vec, const double *size, const double *mu,
                        int64_t col0, int64_t ncol, double *out);
BN_API int bn_post_mode(bn_ctx *ctx, const bn_mat *m, const double *BETA_vec, const double *size, const double *mu,
                        int64_t col0, int64_t ncol, double *out);
/* S draws x + NB(size+x, tempmu) per gene x cell: out[i + G*(j + ncol*s)], j local to the
 * block.  Pure function of (inputs, seed, matrix origin): Philox4x32-10 keyed by seed with
 * the (global gene, global cell, s) triple as counter.  The R shim draws `seed` from R's RNG
 * inside GetRNGstate/PutRNGstate so set.seed() keeps runs reproducible
 * (tests/testthat/test-bayNorm.r:9-17). */
BN_API int bn_post_sample(bn_ctx *ctx, const bn_mat *m, const double *BETA_vec, const double *size, const double *mu,
                          int64_t col0, int64_t ncol, int S, uint64_t seed, double *out);

/* Sparse result (Main_mean_NB_spBay, src/bayNorm_main.cpp:1519-1601; out.sparse = TRUE, R/bayNorm.r:140-150): the
 * posterior mean of columns [col0, col0+ncol) as a device-resident CSC matrix holding every non-zero value -- what
 * `arma::sp_mat(dense)` keeps at :1598 -- without a dense G x C array on the host.  The handle owns device memory. */
typedef struct bn_csc_result bn_csc_result;
BN_API int bn_post_mean_csc(bn_ctx *ctx, const bn_mat *m, const double *BETA_vec, const double *size, const double *mu,
                            int64_t col0, int64_t ncol, bn_csc_result **out);
/* out[0] = G, [1] = ncol, [2] = nnz */
BN_API int bn_csc_result_info(const bn_csc_result *r, int64_t out[3]);
/* Copy the arrays out (host or device destinations; any pointer may be NULL): p64[ncol+1] and/or p32[ncol+1] (R's @p;
 * fails with BN_ERR_UNSUPPORTED when nnz > 2^31 - 1), i[nnz] 0-based rows, x[nnz]. */
BN_API int bn_csc_result_fetch(bn_ctx *ctx, const bn_csc_result *r, int64_t *p64, int32_t *p32, int32_t *i, double *x);
/* The result as a new matrix handle (stays on the device: e.g. the posterior mean fed to another stage). */
BN_API int bn_csc_result_to_mat(bn_ctx *ctx, bn_csc_result *r, bn_mat **out);
BN_API void bn_csc_result_free(bn_csc_result *r);

/* Column-block iterator: results larger than device or host memory (configs 3-5: 26 GB - 6.9 TB) stream through
 * caller-owned host tiles.  kind: 0 mean, 1 mode, 2 S draws.  The library picks the tile width (tile_cols = 0: about
 * 256 MB of result per tile), computes tile t+1 while tile t is copied to the host on a second stream, and calls
 * sink(user, col0, ncol, tile) from the calling thread once a tile is complete; `tile` is G x ncol (x S, planes
 * G*ncol apart) column-major pinned host memory owned by the library, valid until sink returns (write it to disk,
 * copy it into an R array, ...).  A non-zero return from sink stops the iteration (BN_ERR_INVALID). */
typedef int (*bn_tile_sink)(void *user, int64_t col0, int64_t ncol, const double *tile);
BN_API int bn_post_stream(bn_ctx *ctx, const bn_mat *m, int kind, const double *BETA_vec, const double *size,
                          const double *mu, int S, uint64_t seed, int64_t tile_cols, bn_tile_sink sink, void *user);

/* ---- DownSampling  (src/bayNorm_main.cpp:892-915) ---------------------------------- */
/* Dense column-major in/out; out[i,j] ~ Binomial(Data[i,j], BETA_vec[j]). */
BN_API int bn_downsample(bn_ctx *ctx, int64_t G, int64_t C, const double *Data, const double *BETA_vec,
                         uint64_t seed, double *out);

/* ---- BetaFun  (R/PRIOR_FUNCTIONS.R:30-71) ------------------------------------------- */
/* Capture efficiencies from the library sizes of a robust gene subset: genes whose largest
 * log(normalised count + 1) lies below median + 3 MAD (over all cells) and that are zero in
 * fewer than 35 % of the cells.  BETA[C]; selected[G] (0/1, may be NULL);
 * stats[4*G] = {median | mad | max | zero fraction} per gene (may be NULL; median/mad/max are NaN
 * for genes that fail the zero-fraction test: they are never evaluated). */
BN_API int bn_beta_fun(bn_ctx *ctx, bn_mat *m, double MeanBETA, double *BETA, int32_t *selected, double *stats);

/* ---- SyntheticControl  (R/synthetic_controls.r:57-70) ------------------------------ */
/* lambda[G] = rowMeans(t(t(Data)/colSums(Data)) * mean(colSums(Data)/BETA_vec));
 * N_c[G x C, column-major] ~ DownSampling(rpois(lambda_i), BETA_j).  N_c may be NULL (lambda only).
 * Fails with BN_ERR_INVALID when a cell has no counts (the reference returns NaN everywhere). */
BN_API int bn_synthetic_control(bn_ctx *ctx, bn_mat *m, const double *BETA_vec, uint64_t seed, double *lambda,
                                double *N_c);

/* ---- one process, several GPUs  (SURVEY.md 8(b) "Threading", 8(e)) ------------------------ */
/* A context per device whose gene-sharded collectives run over NCCL (ncclCommInitAll; libnccl.so.2 is loaded with dlopen on
 * the first call, never linked), plus group versions of the entry points an R session calls.  Stands where the reference
 * starts a SOCK cluster over genes (R/PRIOR_FUNCTIONS.R:426-427, parallel = TRUE): there every worker receives the whole
 * matrix; here device k holds the gene block [bounds[k], bounds[k+1]).  devices = NULL: 0 .. ndev-1.  The group entry points
 * take HOST pointers and run one host thread per device; results are those of the one-device path (tested bit for bit). */
typedef struct bn_group bn_group;
typedef struct bn_gmat bn_gmat; /* a matrix split into gene blocks, block k resident on device k */
BN_API int bn_group_create(int ndev, const int *devices, bn_group **out);
BN_API void bn_group_destroy(bn_group *g);
BN_API const char *bn_group_last_error(const bn_group *g); /* g may be NULL: last creation error of this thread */
BN_API int bn_group_size(const bn_group *g);
BN_API bn_ctx *bn_group_ctx(bn_group *g, int k);
/* fn(user, k, context k) on one host thread per device, all at once (whatever fn calls may use the collectives); returns
 * the largest status code. */
typedef int (*bn_group_fn)(void *user, int k, bn_ctx *ctx);
BN_API int bn_group_run(bn_group *g, bn_group_fn fn, void *user);
/* dgCMatrix in (host arrays, see bn_mat_from_csc); contiguous gene blocks balanced on C + 4 nnz_gene. */
BN_API int bn_group_mat_from_csc(bn_group *g, int64_t G, int64_t C, const void *p, int p_is_int64, const int32_t *i,
                                 const double *x, bn_gmat **out);
BN_API void bn_gmat_free(bn_gmat *gm);
BN_API int bn_gmat_bounds(const bn_gmat *gm, int64_t *bounds /* [ndev + 1] */);
BN_API bn_mat *bn_gmat_shard(bn_gmat *gm, int k);
/* bn_beta_default / bn_prior / bn_post_{mean,mode,sample} over the group: per-gene arrays are full length G, `out` of
 * bn_group_post is the full column-major G x ncol (x S) host array (kind: 0 mean, 1 mode, 2 S draws). */
BN_API int bn_group_beta_default(bn_group *g, bn_gmat *gm, double mean_beta, double *beta);
BN_API int bn_group_prior(bn_group *g, bn_gmat *gm, const double *BETA_vec, int BB_SIZE, const bn_fit_opts *opts,
                          bn_prior_out *out);
BN_API int bn_group_post(bn_group *g, bn_gmat *gm, int kind, const double *BETA_vec, const double *size, const double *mu,
                         int64_t col0, int64_t ncol, int S, uint64_t seed, double *out);

/* ---- benchmarking helpers ---------------------------------------------------------- */
/* Forget the gene-major twin of the matrix so the next per-gene stage rebuilds it (bench.py
 * times the build inside every step). */
BN_API int bn_mat_drop_gene_major(bn_mat *m);
/* Measured FP64 FMA rate of the device in flop/s (roofline denominator of the fit kernel). */
BN_API int bn_debug_fp64_peak(bn_ctx *ctx, double *flops_per_s);
/* Measured HBM write-only bandwidth in GB/s over a `bytes`-sized buffer, written as one stream
 * (planes = 1) or in the dim = c(G, C, S) pattern of the 3-D output (planes = S; a warp writes
 * run x 256 contiguous bytes to a plane before it moves to the next). */
BN_API int bn_debug_write_peak(bn_ctx *ctx, int64_t bytes, int planes, int run, double *gb_per_s);

/* ---- debug ------------------------------------------------------------------------- */
/* What the last likelihood fit of this context launched (valid after the call returned):
 * out[0] = kernel variant (0 warp per gene, 1 CTA per gene, 2 sweep kernel 2 genes x 6 CTAs/SM,
 *          3 sweep kernel 4 genes x 4 CTAs/SM, 4 the 2-D fit, 5 version 2: k_fit2), out[1] = genes the sweep kernel parked for the
 *          cluster kernel, out[2] = park_after in effect, out[3] = 1 if the analytic gradient was used. */
BN_API int bn_debug_last_fit(const bn_ctx *ctx, int64_t out[4]);
/* The likelihood fit's table of the dense part F(r) = sum_j log(1 + r BETA_j) (all C cells), built for r in [r_lo, r_hi] and
 * evaluated at r[0..n): out[k] = F(r[k]), NaN outside the range (tests pin it against extended-precision sums). */
BN_API int bn_debug_dense_table(bn_ctx *ctx, const double *BETA, int64_t C, double r_lo, double r_hi, const double *r, int64_t n,
                                double *out);
/* A/B switches for profiling runs (the product never reads the environment).  value 0 restores the default.
 *   BN_TUNE_FIT_VARIANT   1..3: force that first-generation fit kernel for long genes, 4: the first-generation kernels chosen by
 *                         shape, 5: version 2 (the table + staged-entries kernels).  Default 0: version 2 for matrices with
 *                         at most 20 % non-zeros, first generation otherwise (see bn_debug_last_fit)
 *   BN_TUNE_SAMPLE_NCS    columns walked per CTA by the 3-D sampler
 *   BN_TUNE_SAMPLE_SERIAL 1: the 3-D sampler queues every kernel on the context's stream (no side streams)
 *   BN_TUNE_FIT_SPLIT     version 2 of the fit: genes with at least this many stored entries get a CTA, the others a warp
 *   BN_TUNE_SAMPLE_CHUNKS the 3-D sampler splits a call into at least this many column chunks (default: by side-buffer size only) */
enum { BN_TUNE_FIT_VARIANT = 0, BN_TUNE_SAMPLE_NCS = 1, BN_TUNE_POST_NCS = 2, BN_TUNE_SAMPLE_SERIAL = 3, BN_TUNE_FIT_SPLIT = 4, BN_TUNE_SAMPLE_CHUNKS = 5,
       BN_TUNE_COUNT = 8 };
BN_API int bn_debug_tune(bn_ctx *ctx, int knob, int value);
/* One raw Philox4x32-10 block (counter ctr[4], key[2]) -> out[4]; lets tests pin the
 * generator against the published known-answer vectors. */
/* The device logarithms of the likelihood kernels on a vector (which: 0 log, 1 log1p for y >= 0,
 * 2 log1p series for y <= 2^-6, 3 Stirling lgamma for z >= 64). */
BN_API int bn_debug_log(bn_ctx *ctx, int which, const double *x, int64_t n, double *out);
BN_API int bn_debug_philox(bn_ctx *ctx, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

#ifdef __cplusplus
}
#endif
#endif /* BAYNORM_B200_H */
