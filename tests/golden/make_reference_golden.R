#!/usr/bin/env Rscript
# make_reference_golden.R -- produce reference-held golden values for the hot path with REAL R, bayNorm and BB.
#
# This repository's oracle (oracle/) restates the reference path from the sources because R does not exist in the build
# image ("parity unpinned", DESIGN.md section 2).  Anyone with R can pin it:
#
#     R -e 'install.packages("BB"); BiocManager::install("bayNorm")'      # or R CMD INSTALL /path/to/WT215/bayNorm
#     Rscript tests/golden/make_reference_golden.R                        # writes tests/golden/reference_fixture.txt
#     python -m pytest tests/test_oracle.py -k reference_golden           # oracle vs the file (CPU only)
#     python -m pytest tests/test_gpu_parity.py -m gpu -k reference_golden  # CUDA path vs the file
#
# Everything runs on data(EXAMPLE_DATA_list) (tests/testthat/test-bayNorm.r:9-29 uses the same object).  The output is a
# plain text file, one block per value:   "## <name> <nrow> <ncol>"  then nrow*ncol numbers (column-major, 17 digits).
suppressPackageStartupMessages({
    library(bayNorm)
    library(BB)
    library(Matrix)
})
args <- commandArgs(trailingOnly = TRUE)
out <- if (length(args) >= 1) args[1] else file.path("tests", "golden", "reference_fixture.txt")
con <- file(out, "w")
put <- function(name, x) {
    x <- as.matrix(x)
    storage.mode(x) <- "double"
    writeLines(sprintf("## %s %d %d", name, nrow(x), ncol(x)), con)
    writeLines(paste(formatC(as.vector(x), digits = 17, format = "g"), collapse = " "), con)
}
data("EXAMPLE_DATA_list")
X <- EXAMPLE_DATA_list$inputdata
beta <- EXAMPLE_DATA_list$inputbeta
size <- EXAMPLE_DATA_list$size
mu <- EXAMPLE_DATA_list$mu
writeLines(sprintf("# R %s, bayNorm %s, BB %s, Matrix %s", R.version.string, packageVersion("bayNorm"), packageVersion("BB"),
                   packageVersion("Matrix")), con)
put("inputdata", X); put("inputbeta", beta); put("size", size); put("mu", mu)

# registered routines on the fixture (src/bayNorm_main.cpp)
Xs <- as(as.matrix(X), "dgCMatrix")
put("Main_mean_NB_Bay", bayNorm:::Main_mean_NB_Bay(Xs, beta, size, mu, 20, 10000000))
put("Main_mode_NB_Bay", bayNorm:::Main_mode_NB_Bay(Xs, beta, size, mu, 20, 10000000))
ll <- sapply(seq_len(nrow(X)), function(g) sapply(c(0.05, 0.37, 1, 2.5, 11, 140), function(s)
    bayNorm:::MarginalF_NB_1D(s, mu[g], X[g, ], beta)))
put("MarginalF_NB_1D_sizes_x_genes", ll)
gr <- sapply(seq_len(nrow(X)), function(g) sapply(c(0.05, 0.37, 1, 2.5, 11, 140), function(s)
    bayNorm:::GradientFun_NB_1D(s, mu[g], X[g, ], beta)))
put("GradientFun_NB_1D_sizes_x_genes", gr)
ll2 <- sapply(seq_len(nrow(X)), function(g) bayNorm:::MarginalF_NB_2D(c(size[g], mu[g]), X[g, ], beta))
put("MarginalF_NB_2D", ll2)
gr2 <- sapply(seq_len(nrow(X)), function(g) bayNorm:::GradientFun_NB_2D(c(size[g], mu[g]), X[g, ], beta))
put("GradientFun_NB_2D", gr2)

# default BETA (R/bayNorm.r:164-170)
xx <- Matrix::colSums(Xs)
b0 <- xx / mean(xx) * 0.06
b0[b0 <= 0] <- min(b0[b0 > 0]); b0[b0 >= 1] <- max(b0[b0 < 1])
put("default_BETA", b0)

# EstPrior / Prior_fun / BB_Fun (R/PRIOR_FUNCTIONS.R)
ep <- bayNorm:::EstPrior_sprcpp(t(t(Xs) / beta))
put("EstPrior_sprcpp_MU", ep$MU); put("EstPrior_sprcpp_SIZE", ep$SIZE); put("EstPrior_sprcpp_v", ep$v)
pr <- Prior_fun(Data = Xs, BETA_vec = beta, parallel = FALSE, NCores = 1, FIX_MU = TRUE, GR = FALSE, BB_SIZE = TRUE, verbose = FALSE)
put("Prior_MME_MU", pr$MME_prior$MME_MU); put("Prior_MME_SIZE", pr$MME_prior$MME_SIZE)
put("Prior_BB_SIZE", pr$BB_prior[, 1]); put("Prior_MME_SIZE_adjust", pr$MME_SIZE_adjust)
lo <- min(pr$MME_prior$MME_SIZE); hi <- ceiling(max(pr$MME_prior$MME_SIZE))
put("Prior_box", c(lo, hi))
bb1 <- BB_Fun(Xs, beta, pr$MME_prior$MME_MU, pr$MME_prior$MME_SIZE, MU_lower = min(pr$MME_prior$MME_MU),
              MU_upper = max(pr$MME_prior$MME_MU), SIZE_lower = lo, SIZE_upper = hi, parallel = FALSE, FIX_MU = TRUE, GR = FALSE)
put("BB_Fun_FIX_MU_TRUE", bb1)
bb1g <- BB_Fun(Xs, beta, pr$MME_prior$MME_MU, pr$MME_prior$MME_SIZE, MU_lower = min(pr$MME_prior$MME_MU),
               MU_upper = max(pr$MME_prior$MME_MU), SIZE_lower = lo, SIZE_upper = hi, parallel = FALSE, FIX_MU = TRUE, GR = TRUE)
put("BB_Fun_FIX_MU_TRUE_GR", bb1g)
bb2 <- BB_Fun(Xs, beta, pr$MME_prior$MME_MU, pr$MME_prior$MME_SIZE, MU_lower = min(pr$MME_prior$MME_MU),
              MU_upper = max(pr$MME_prior$MME_MU), SIZE_lower = lo, SIZE_upper = hi, parallel = FALSE, FIX_MU = FALSE, GR = FALSE)
put("BB_Fun_FIX_MU_FALSE_serial", bb2)
# one gene through spg by hand, with the counts BB reports: pins the optimiser restatement (iterations, evaluations)
g <- 1
o <- BB::spg(par = pr$MME_prior$MME_SIZE[g], fn = bayNorm:::MarginalF_NB_1D, MU = pr$MME_prior$MME_MU[g], m_observed = X[g, ],
             BETA = beta, control = list(maximize = TRUE, trace = FALSE, maxfeval = 500), lower = lo, upper = hi)
put("spg_gene1_par_value_feval_iter_conv", c(o$par, o$value, o$feval, o$iter, o$convergence))

# BetaFun and AdjustSIZE_fun
bf <- BetaFun(Data = Xs, MeanBETA = 0.06)
put("BetaFun_BETA", bf$BETA)
put("AdjustSIZE_fun", AdjustSIZE_fun(pr$BB_prior[, 1], pr$MME_prior$MME_MU, pr$MME_prior$MME_SIZE))

# the whole thing (BASELINE config 1): bayNorm, GG, mean_version = TRUE, reference BETA_vec
bn <- bayNorm(Data = X, BETA_vec = beta, mean_version = TRUE, parallel = FALSE, NCores = 1, verbose = FALSE)
put("bayNorm_mean_Bay_out", as.matrix(bn$Bay_out))
put("bayNorm_mean_MME_SIZE_adjust", bn$PRIORS$MME_SIZE_adjust)
bm <- bayNorm(Data = X, BETA_vec = beta, mode_version = TRUE, parallel = FALSE, NCores = 1, verbose = FALSE)
put("bayNorm_mode_Bay_out", as.matrix(bm$Bay_out))
# 3-D draws: moments over many draws (distributional parity only: the RNG streams differ by construction)
set.seed(1)
a3 <- bayNorm:::Main_NB_Bay(Xs, beta, size, mu, 2000, 10000000)
put("Main_NB_Bay_mean_over_2000", apply(a3, c(1, 2), mean)); put("Main_NB_Bay_var_over_2000", apply(a3, c(1, 2), var))
close(con)
cat("wrote", out, "\n")
