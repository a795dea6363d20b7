## R-side changes a maintainer makes so that BB_Fun / Prior_fun / bayNorm call the batched entries.
## Signatures, defaults and return structures are unchanged (R/PRIOR_FUNCTIONS.R:380-384, :241-245,
## R/bayNorm.r:120-129).  `NCores` is accepted and ignored (the fit is one kernel launch per device); `parallel = TRUE`
## spreads the genes over options(bayNorm.b200.devices = n) GPUs of this session (default 1) instead of a SOCK cluster.
## Not executed in this repository (no R in the image).

BB_Fun <- function(Data, BETA_vec, INITIAL_MU_vec, INITIAL_SIZE_vec,
                   MU_lower = 0.01, MU_upper = 500, SIZE_lower = 0.01, SIZE_upper = 30,
                   parallel = FALSE, NCores = 5, FIX_MU = TRUE, GR = FALSE) {
    Data <- Check_input(Data)
    if (!FIX_MU) {
        BB_parmat <- .Call("bnb_fit_size_mu", Data, as.numeric(BETA_vec), as.numeric(INITIAL_MU_vec),
                           as.numeric(INITIAL_SIZE_vec), c(SIZE_lower, MU_lower), c(SIZE_upper, MU_upper), GR,
                           PACKAGE = "bayNorm")
        rownames(BB_parmat) <- rownames(Data)
        colnames(BB_parmat) <- c("size", "mu")
        return(BB_parmat)
    }
    BB_parmat <- .Call("bnb_fit_size", Data, as.numeric(BETA_vec), as.numeric(INITIAL_MU_vec),
                       as.numeric(INITIAL_SIZE_vec), SIZE_lower, SIZE_upper, GR, PACKAGE = "bayNorm")
    names(BB_parmat) <- rownames(Data)
    BB_parmat
}

Prior_fun <- function(Data, BETA_vec, parallel = TRUE, NCores = 5, FIX_MU = TRUE, GR = FALSE,
                      BB_SIZE = TRUE, verbose = TRUE) {
    Data <- Check_input(Data)
    .Call("bnb_set_devices", if (parallel) as.integer(getOption("bayNorm.b200.devices", 1L)) else 1L, PACKAGE = "bayNorm")
    r <- .Call("bnb_prior", Data, as.numeric(BETA_vec), BB_SIZE && FIX_MU, GR, PACKAGE = "bayNorm")
    MME_prior <- data.frame(MME_MU = r[[1]], MME_SIZE = r[[2]], row.names = rownames(Data))
    if (!BB_SIZE) return(list(MME_prior = MME_prior, BETA_vec = BETA_vec))
    if (!FIX_MU) {   # R/PRIOR_FUNCTIONS.R:271-296 with the 2-D fit: the MME step above, then BB_Fun and AdjustSIZE_fun as upstream
        BB_prior <- BB_Fun(Data, BETA_vec, INITIAL_MU_vec = MME_prior$MME_MU, INITIAL_SIZE_vec = MME_prior$MME_SIZE,
                           MU_lower = min(MME_prior$MME_MU), MU_upper = max(MME_prior$MME_MU),
                           SIZE_lower = min(MME_prior$MME_SIZE), SIZE_upper = ceiling(max(MME_prior$MME_SIZE)),
                           FIX_MU = FALSE, GR = GR)
        MME_SIZE_adjust <- AdjustSIZE_fun(BB_prior[, 1], MME_prior$MME_MU, MME_prior$MME_SIZE)
        names(MME_SIZE_adjust) <- rownames(Data)
        return(list(MME_prior = MME_prior, BB_prior = BB_prior, MME_SIZE_adjust = MME_SIZE_adjust, BETA_vec = BETA_vec))
    }
    BB_prior <- cbind(BB_SIZE = r[[3]], MME_MU = r[[1]])
    rownames(BB_prior) <- rownames(Data)
    MME_SIZE_adjust <- r[[4]]
    names(MME_SIZE_adjust) <- rownames(Data)
    if (verbose) message("Prior estimation has completed!")
    list(MME_prior = MME_prior, BB_prior = BB_prior, MME_SIZE_adjust = MME_SIZE_adjust, BETA_vec = BETA_vec)
}

## inside bayNorm(), replacing R/bayNorm.r:164-170:
##     if (is.null(BETA_vec)) BETA_vec <- .Call("bnb_beta_default", Data, PACKAGE = "bayNorm")
## Main_mean_NB_spBay keeps its name:
Main_mean_NB_spBay <- function(Data, BETA_vec, size, mu, S, thres)
    methods::as(Main_mean_NB_Bay(Data, BETA_vec, size, mu, S, thres), "dgCMatrix")

BetaFun <- function(Data, MeanBETA) {
    Data <- Check_input(Data)
    r <- .Call("bnb_beta_fun", Data, MeanBETA, PACKAGE = "bayNorm")
    BETA <- r[[1]]
    names(BETA) <- colnames(Data)
    list(BETA = BETA, Selected_genes = rownames(Data)[r[[2]]])
}

SyntheticControl <- function(Data, BETA_vec) {
    Data <- Check_input(Data)
    r <- .Call("bnb_synthetic_control", Data, as.numeric(BETA_vec), PACKAGE = "bayNorm")
    list(N_c = r[[1]], beta_c = BETA_vec, lambda = r[[2]])
}
