"""CPU oracle for bayNorm's hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``bench.py``'s CPU-baseline legs and ``__graft_entry__.smoke()``
may import this module.  The product (``baynorm_b200``) never does.

Two layers:

* ``libbaynorm_oracle.so`` (``oracle/baynorm_oracle.c``): plain-C restatement of the
  reference functions, operation for operation, each citing the reference file:line.
* this file: ctypes bindings, the R-level orchestration restated
  (``Prior_fun`` R/PRIOR_FUNCTIONS.R:241-311, ``bayNorm`` R/bayNorm.r:120-363,
  ``bayNorm_sup`` :457-617), and an *independent* numpy version of every closed form
  (``np_*``) used to cross-check the C code before it is trusted.

PARITY STATUS: "parity unpinned" -- the reference cannot run in this image (no R,
Rcpp, Armadillo, Rmath, BB) and its tests hold no numeric golden vectors; see the
header of baynorm_oracle.c and DESIGN.md.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

_D = np.float64
_pd = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_pi32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_pi64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> Path:
    """Compile the C oracle with oracle/Makefile (gcc only; no reference sources)."""
    so = _HERE / "libbaynorm_oracle.so"
    src = _HERE / "baynorm_oracle.c"
    if force or (not so.exists()) or (src.exists() and so.stat().st_mtime < src.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = build()
    L = C.CDLL(str(so))
    L.orc_dnbinom_mu_log.restype = C.c_double
    L.orc_dnbinom_mu_log.argtypes = [C.c_double] * 3
    L.orc_dnbinom_mu_log_direct.restype = C.c_double
    L.orc_dnbinom_mu_log_direct.argtypes = [C.c_double] * 3
    L.orc_digamma.restype = C.c_double
    L.orc_digamma.argtypes = [C.c_double]
    L.orc_marginal_nb_1d.restype = C.c_double
    L.orc_marginal_nb_1d.argtypes = [C.c_double, C.c_double, _pd, _pd, C.c_int64]
    L.orc_marginal_nb_2d.restype = C.c_double
    L.orc_marginal_nb_2d.argtypes = [_pd, _pd, _pd, C.c_int64]
    L.orc_gradient_nb_1d.restype = C.c_double
    L.orc_gradient_nb_1d.argtypes = [C.c_double, C.c_double, _pd, _pd, C.c_int64]
    L.orc_gradient_nb_2d.restype = None
    L.orc_gradient_nb_2d.argtypes = [_pd, _pd, _pd, C.c_int64, _pd]
    L.orc_colsums.restype = None
    L.orc_colsums.argtypes = [C.c_int64, _pi64, _pd, _pd]
    L.orc_beta_from_colsums.restype = C.c_int
    L.orc_beta_from_colsums.argtypes = [C.c_int64, _pd, C.c_double, _pd]
    L.orc_mme.restype = None
    L.orc_mme.argtypes = [C.c_int64, C.c_int64, _pi64, _pi32, _pd, C.c_void_p, _pd, _pd, _pd]
    L.orc_spg_gene_1d.restype = C.c_int
    L.orc_spg_gene_1d.argtypes = [C.c_double, C.c_double, _pd, _pd, C.c_int64, C.c_double, C.c_double, C.c_int,
                                  C.POINTER(C.c_double), C.POINTER(C.c_int)]
    L.orc_spg_gene_2d.restype = C.c_int
    L.orc_spg_gene_2d.argtypes = [_pd, _pd, _pd, C.c_int64, _pd, _pd, C.c_int, _pd, C.POINTER(C.c_int)]
    L.orc_bb_fun_1d.restype = C.c_int
    L.orc_bb_fun_1d.argtypes = [C.c_int64, C.c_int64, _pi64, _pi32, _pd, _pd, _pd, _pd, C.c_double, C.c_double,
                                C.c_int, C.c_int64, C.c_int64, C.c_int, _pd, _pi32, _pi32]
    L.orc_brent_gene_1d.restype = C.c_double
    L.orc_brent_gene_1d.argtypes = [C.c_double, _pd, _pd, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_int,
                                    C.POINTER(C.c_double)]
    L.orc_adjust_size.restype = C.c_int
    L.orc_adjust_size.argtypes = [C.c_int64, _pd, _pd, _pd, _pd]
    for name in ("orc_post_mean", "orc_post_mode"):
        f = getattr(L, name)
        f.restype = None
        f.argtypes = [C.c_int64, C.c_int64, _pi64, _pi32, _pd, _pd, _pd, _pd, _pd]
    L.orc_post_sample_params.restype = None
    L.orc_post_sample_params.argtypes = [C.c_int64, C.c_int64, _pi64, _pi32, _pd, _pd, _pd, _pd, _pd, _pd]
    L.orc_post_sample.restype = None
    L.orc_post_sample.argtypes = [C.c_int64, C.c_int64, _pi64, _pi32, _pd, _pd, _pd, _pd, C.c_int, C.c_uint64, _pd]
    L.orc_downsample.restype = None
    L.orc_downsample.argtypes = [C.c_int64, C.c_int64, _pd, _pd, C.c_uint64, _pd]
    L.orc_max_threads.restype = C.c_int
    L.orc_max_threads.argtypes = []
    _LIB = L
    return L


# --------------------------------------------------------------------------------------
# CSC container (dgCMatrix stand-in: @p, @i, @x, @Dim)
# --------------------------------------------------------------------------------------
class CSC:
    """Minimal dgCMatrix: p int64[C+1], i int32[nnz] (sorted within column), x float64[nnz]."""

    def __init__(self, G, Cn, p, i, x):
        self.G = int(G)
        self.C = int(Cn)
        self.p = np.ascontiguousarray(p, dtype=np.int64)
        self.i = np.ascontiguousarray(i, dtype=np.int32)
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        assert self.p.shape == (self.C + 1,)
        assert self.i.shape == self.x.shape == (int(self.p[-1]),)

    @property
    def nnz(self):
        return int(self.p[-1])

    @classmethod
    def from_dense(cls, A):
        """as(as.matrix(Data), 'dgCMatrix')  (R/sup_funs.R:105): explicit zeros dropped."""
        A = np.asarray(A, dtype=np.float64)
        G, Cn = A.shape
        mask = A.T != 0            # (C, G): column-major walk
        cols, rows = np.nonzero(mask)
        p = np.zeros(Cn + 1, dtype=np.int64)
        np.add.at(p, cols + 1, 1)
        p = np.cumsum(p)
        return cls(G, Cn, p, rows.astype(np.int32), A.T[mask])

    def to_dense(self):
        A = np.zeros((self.G, self.C), dtype=np.float64)
        cols = np.repeat(np.arange(self.C), np.diff(self.p))
        A[self.i, cols] = self.x
        return A

    def select_cols(self, idx):
        """Data[, idx]"""
        idx = np.asarray(idx, dtype=np.int64)
        lens = (self.p[idx + 1] - self.p[idx]).astype(np.int64)
        p = np.zeros(len(idx) + 1, dtype=np.int64)
        p[1:] = np.cumsum(lens)
        take = np.concatenate([np.arange(self.p[j], self.p[j + 1]) for j in idx]) if len(idx) else np.zeros(0, np.int64)
        return CSC(self.G, len(idx), p, self.i[take], self.x[take])

    def select_rows(self, g0, g1):
        """Data[g0:g1, ] (0-based, half-open) -- a gene shard."""
        keep = (self.i >= g0) & (self.i < g1)
        cols = np.repeat(np.arange(self.C), np.diff(self.p))
        p = np.zeros(self.C + 1, dtype=np.int64)
        np.add.at(p, cols[keep] + 1, 1)
        p = np.cumsum(p)
        return CSC(g1 - g0, self.C, p, self.i[keep] - g0, self.x[keep])


# --------------------------------------------------------------------------------------
# thin wrappers over the C restatement
# --------------------------------------------------------------------------------------
def dnbinom_mu_log(x, size, mu):
    f = np.vectorize(lambda a, b, c: lib().orc_dnbinom_mu_log(float(a), float(b), float(c)), otypes=[_D])
    return f(x, size, mu)


def dnbinom_mu_log_direct(x, size, mu):
    f = np.vectorize(lambda a, b, c: lib().orc_dnbinom_mu_log_direct(float(a), float(b), float(c)), otypes=[_D])
    return f(x, size, mu)


def digamma(x):
    return np.vectorize(lambda a: lib().orc_digamma(float(a)), otypes=[_D])(x)


def MarginalF_NB_1D(SIZE, MU, m_observed, BETA):
    m = np.ascontiguousarray(m_observed, _D)
    b = np.ascontiguousarray(BETA, _D)
    return lib().orc_marginal_nb_1d(float(SIZE), float(MU), m, b, m.size)


def MarginalF_NB_2D(SIZE_MU, m_observed, BETA):
    m = np.ascontiguousarray(m_observed, _D)
    b = np.ascontiguousarray(BETA, _D)
    return lib().orc_marginal_nb_2d(np.ascontiguousarray(SIZE_MU, _D), m, b, m.size)


def GradientFun_NB_1D(SIZE, MU, m_observed, BETA):
    m = np.ascontiguousarray(m_observed, _D)
    b = np.ascontiguousarray(BETA, _D)
    return lib().orc_gradient_nb_1d(float(SIZE), float(MU), m, b, m.size)


def GradientFun_NB_2D(SIZE_MU, m_observed, BETA):
    m = np.ascontiguousarray(m_observed, _D)
    b = np.ascontiguousarray(BETA, _D)
    out = np.zeros(2, _D)
    lib().orc_gradient_nb_2d(np.ascontiguousarray(SIZE_MU, _D), m, b, m.size, out)
    return out


def colsums(M: CSC):
    out = np.zeros(M.C, _D)
    lib().orc_colsums(M.C, M.p, M.x, out)
    return out


def beta_default(M: CSC, mean_beta=0.06):
    """R/bayNorm.r:164-170.  Returns (beta, bad) -- bad=True where :185-187 would stop()."""
    xx = colsums(M)
    beta = np.zeros(M.C, _D)
    bad = lib().orc_beta_from_colsums(M.C, xx, float(mean_beta), beta)
    return beta, bool(bad)


def EstPrior(M: CSC, beta=None):
    """EstPrior (R/PRIOR_FUNCTIONS.R:105-175) on x/beta (beta=None: data already normalised).
    Returns dict(MU, SIZE (NaN where v<=m), v)."""
    MU = np.zeros(M.G, _D)
    SIZE = np.zeros(M.G, _D)
    V = np.zeros(M.G, _D)
    bp = None if beta is None else np.ascontiguousarray(beta, _D).ctypes.data_as(C.c_void_p)
    lib().orc_mme(M.G, M.C, M.p, M.i, M.x, bp, MU, SIZE, V)
    return {"MU": MU, "SIZE": SIZE, "v": V}


def spg_gene_1d(size0, mu, m_observed, beta, lower, upper, maxfeval=500):
    m = np.ascontiguousarray(m_observed, _D)
    b = np.ascontiguousarray(beta, _D)
    out = C.c_double()
    counts = (C.c_int * 3)()
    rc = lib().orc_spg_gene_1d(float(size0), float(mu), m, b, m.size, float(lower), float(upper), int(maxfeval),
                               C.byref(out), counts)
    return out.value, rc, tuple(counts)


def spg_gene_2d(par0, m_observed, beta, lower, upper, maxfeval=500):
    m = np.ascontiguousarray(m_observed, _D)
    b = np.ascontiguousarray(beta, _D)
    out = np.zeros(2, _D)
    counts = (C.c_int * 3)()
    rc = lib().orc_spg_gene_2d(np.ascontiguousarray(par0, _D), m, b, m.size, np.ascontiguousarray(lower, _D),
                               np.ascontiguousarray(upper, _D), int(maxfeval), out, counts)
    return out, rc, tuple(counts)


def brent_gene_1d(mu, m_observed, beta, lower, upper, xatol=1e-12, maxiter=500):
    m = np.ascontiguousarray(m_observed, _D)
    b = np.ascontiguousarray(beta, _D)
    f = C.c_double()
    x = lib().orc_brent_gene_1d(float(mu), m, b, m.size, float(lower), float(upper), float(xatol), int(maxiter),
                                C.byref(f))
    return x, f.value


def BB_Fun(M: CSC, BETA_vec, INITIAL_MU_vec, INITIAL_SIZE_vec, SIZE_lower, SIZE_upper, maxfeval=500, g0=0, g1=None,
           nthreads=1, want_counts=False):
    """BB_Fun, FIX_MU=TRUE (R/PRIOR_FUNCTIONS.R:380-549) for genes [g0, g1)."""
    g1 = M.G if g1 is None else g1
    n = g1 - g0
    out = np.zeros(n, _D)
    conv = np.zeros(n, np.int32)
    counts = np.zeros(3 * n, np.int32)
    rc = lib().orc_bb_fun_1d(M.G, M.C, M.p, M.i, M.x, np.ascontiguousarray(BETA_vec, _D),
                             np.ascontiguousarray(INITIAL_MU_vec, _D), np.ascontiguousarray(INITIAL_SIZE_vec, _D),
                             float(SIZE_lower), float(SIZE_upper), int(maxfeval), g0, g1, int(nthreads), out, conv,
                             counts)
    if rc != 0:
        raise MemoryError("orc_bb_fun_1d")
    if want_counts:
        return out, conv, counts.reshape(n, 3)
    return out


def BB_Fun_2d(A, BETA_vec, INITIAL_MU_vec, INITIAL_SIZE_vec, SIZE_lower, SIZE_upper, MU_lower, MU_upper, maxfeval=500):
    """R/PRIOR_FUNCTIONS.R:457-470 (FIX_MU=FALSE, parallel=TRUE bounds :414-419) on a dense matrix, gene by gene
    through the restated spg: returns the G x 2 (size, mu) matrix, the convergence codes and the evaluation counts."""
    A = np.asarray(A, _D)
    G = A.shape[0]
    par = np.zeros((G, 2), _D)
    conv = np.zeros(G, np.int32)
    cnt = np.zeros((G, 3), np.int32)
    lo, hi = np.array([SIZE_lower, MU_lower], _D), np.array([SIZE_upper, MU_upper], _D)
    for g in range(G):
        p, rc, c = spg_gene_2d([INITIAL_SIZE_vec[g], INITIAL_MU_vec[g]], A[g], BETA_vec, lo, hi, maxfeval)
        par[g], conv[g], cnt[g] = p, rc, c
    return par, conv, cnt


def AdjustSIZE_fun(BB_SIZE, MME_MU, MME_SIZE):
    """R/sup_funs.R:27-33 (MME_MU is accepted and unused, as upstream)."""
    bb = np.ascontiguousarray(BB_SIZE, _D)
    ms = np.ascontiguousarray(MME_SIZE, _D)
    out = np.zeros_like(bb)
    coef = np.zeros(2, _D)
    lib().orc_adjust_size(bb.size, bb, ms, out, coef)
    return out


def _post(fn, M: CSC, BETA_vec, size, mu):
    out = np.zeros(M.G * M.C, _D)
    getattr(lib(), fn)(M.G, M.C, M.p, M.i, M.x, np.ascontiguousarray(BETA_vec, _D), np.ascontiguousarray(size, _D),
                       np.ascontiguousarray(mu, _D), out)
    return out.reshape(M.C, M.G).T        # column-major G x C -> numpy (G, C) view


def Main_mean_NB_Bay(M, BETA_vec, size, mu, S=20, thres=0):
    return _post("orc_post_mean", M, BETA_vec, size, mu)


def Main_mode_NB_Bay(M, BETA_vec, size, mu, S=20, thres=0):
    return _post("orc_post_mode", M, BETA_vec, size, mu)


def post_sample_params(M, BETA_vec, size, mu):
    """(nb_size, nb_mu) of the draw x + NB(size=phi+x, mu=tempmu) for every gene x cell."""
    a = np.zeros(M.G * M.C, _D)
    b = np.zeros(M.G * M.C, _D)
    lib().orc_post_sample_params(M.G, M.C, M.p, M.i, M.x, np.ascontiguousarray(BETA_vec, _D),
                                 np.ascontiguousarray(size, _D), np.ascontiguousarray(mu, _D), a, b)
    return a.reshape(M.C, M.G).T, b.reshape(M.C, M.G).T


def Main_NB_Bay(M, BETA_vec, size, mu, S=20, thres=0, seed=1):
    """3-D draws, array (G, C, S) like the reference's dim=c(G,C,S) cube."""
    out = np.zeros(M.G * M.C * S, _D)
    lib().orc_post_sample(M.G, M.C, M.p, M.i, M.x, np.ascontiguousarray(BETA_vec, _D),
                          np.ascontiguousarray(size, _D), np.ascontiguousarray(mu, _D), int(S), int(seed), out)
    return out.reshape(S, M.C, M.G).transpose(2, 1, 0)


def DownSampling(Data, BETA_vec, seed=1):
    A = np.asarray(Data, _D)
    G, Cn = A.shape
    flat = np.ascontiguousarray(A.T).ravel()           # column-major
    out = np.zeros_like(flat)
    lib().orc_downsample(G, Cn, flat, np.ascontiguousarray(BETA_vec, _D), int(seed), out)
    return out.reshape(Cn, G).T


def BetaFun(Data, MeanBETA):
    """R/PRIOR_FUNCTIONS.R:30-71 restated with numpy on a dense matrix (the reference's own formulation:
    dense `apply` loops): returns BETA, the selected row indices and the per-gene statistics."""
    A = np.asarray(Data, _D)
    xx = A.sum(axis=0)
    norm = (A / xx[None, :]) * xx.mean()                       # :39
    L = np.log(norm + 1.0)
    med = np.median(L, axis=1)                                   # :44-46
    mad = 1.4826 * np.median(np.abs(L - med[:, None]), axis=1)   # :47-49 stats::mad
    mx = L.max(axis=1)                                           # :51
    dropout = (A == 0).mean(axis=1)                              # :53-55
    sel = np.nonzero((mx < med + 3 * mad) & (dropout < 0.35))[0]  # :52, :56
    t = A[sel, :].sum(axis=0)                                    # :59
    beta = t / t.mean() * MeanBETA                               # :60
    if (beta >= 1).any():                                        # :61-63
        beta[beta >= 1] = beta[beta < 1].max()
    if (beta <= 0).any():                                        # :64-66
        beta[beta <= 0] = beta[beta > 0].min()
    return {"BETA": beta, "Selected_genes": sel, "stats": {"median": med, "mad": mad, "max": mx, "dropout": dropout}}


def SyntheticControl(Data, BETA_vec, seed=1):
    """R/synthetic_controls.r:57-70 restated with numpy (dense input):
    Mean_depth = mean(colSums/BETA); lambda = rowMeans(t(t(Data)/colSums) * Mean_depth);
    N_c = DownSampling(rpois(G*C, lambda), BETA)."""
    A = np.asarray(Data, _D)
    b = np.asarray(BETA_vec, _D)
    cs = A.sum(axis=0)
    mean_depth = np.mean(cs / b)
    lam = ((A / cs[None, :]) * mean_depth).mean(axis=1)
    rng = np.random.default_rng(seed)
    N = rng.poisson(np.repeat(lam[:, None], A.shape[1], axis=1)).astype(_D)
    return {"N_c": rng.binomial(N.astype(np.int64), b[None, :]).astype(_D), "beta_c": b, "lambda": lam}


# --------------------------------------------------------------------------------------
# R-level orchestration restated
# --------------------------------------------------------------------------------------
def Prior_fun(M: CSC, BETA_vec, BB_SIZE=True, maxfeval=500, nthreads=1):
    """R/PRIOR_FUNCTIONS.R:241-311 (FIX_MU=TRUE, GR=FALSE -- the defaults)."""
    BETA_vec = np.ascontiguousarray(BETA_vec, _D)
    pri = EstPrior(M, BETA_vec)                               # :253-256
    M_ave_ori = pri["MU"]
    size_est = pri["SIZE"].copy()
    nan = np.isnan(size_est)
    size_est[nan] = np.min(size_est[~nan])                    # :259
    out = {"MME_prior": {"MME_MU": M_ave_ori, "MME_SIZE": size_est}, "BETA_vec": BETA_vec}
    if BB_SIZE:
        lower = float(np.min(size_est))                       # :277
        upper = float(math.ceil(np.max(size_est)))            # :278-279
        bb = BB_Fun(M, BETA_vec, M_ave_ori, size_est, lower, upper, maxfeval=maxfeval, nthreads=nthreads)
        out["BB_prior"] = {"BB_SIZE": bb, "MME_MU": M_ave_ori}
        out["MME_SIZE_adjust"] = AdjustSIZE_fun(bb, M_ave_ori, size_est)   # :294-296
    return out


def _myfunc(mode_version, mean_version):
    if mode_version and mean_version:
        raise ValueError("Only one of mode_version and mean_version should be specified to be TRUE")
    if mode_version:
        return Main_mode_NB_Bay
    if mean_version:
        return Main_mean_NB_Bay
    return Main_NB_Bay


def bayNorm(M: CSC, BETA_vec=None, Conditions=None, Prior_type=None, mode_version=False, mean_version=False, S=20,
            BB_SIZE=True, seed=1, nthreads=1):
    """R/bayNorm.r:120-363 (UMI_sffl=NULL, FIX_MU=TRUE, GR=FALSE)."""
    fn = _myfunc(mode_version, mean_version)
    kw = {"seed": seed} if fn is Main_NB_Bay else {}
    input_params = {"BETA_vec": BETA_vec, "Conditions": Conditions, "UMI_sffl": None, "Prior_type": Prior_type,
                    "FIX_MU": True, "BB_SIZE": BB_SIZE, "GR": False}
    if BETA_vec is None:
        BETA_vec, _ = beta_default(M)                          # :164-170
    BETA_vec = np.ascontiguousarray(BETA_vec, _D)
    if BETA_vec.min() <= 0 or BETA_vec.max() >= 1:
        raise ValueError("The range of BETA must be within (0,1).")
    if M.C != BETA_vec.size:
        raise ValueError("The number of cells(columns) in Data is not consistent with BETA_vec.")
    if Conditions is None:
        PRIORS = Prior_fun(M, BETA_vec, BB_SIZE=BB_SIZE, nthreads=nthreads)
        MU_input = PRIORS["MME_prior"]["MME_MU"]
        SIZE_input = PRIORS["MME_SIZE_adjust"] if BB_SIZE else PRIORS["MME_prior"]["MME_SIZE"]
        Bay_out = fn(M, BETA_vec, SIZE_input, MU_input, S=S, **kw)
        return {"Bay_out": Bay_out, "PRIORS": PRIORS, "input_params": input_params}
    Conditions = np.asarray(Conditions)
    if Conditions.size != M.C:
        raise ValueError("Number of columns in expression matrix must match length of conditions vector!")
    if Prior_type is None:
        Prior_type = "LL"
    _, first = np.unique(Conditions, return_index=True)
    Levels = Conditions[np.sort(first)]                        # unique() keeps first-seen order
    idx = [np.nonzero(Conditions == lv)[0] for lv in Levels]
    DataList = [M.select_cols(ix) for ix in idx]
    BETAList = [BETA_vec[ix] for ix in idx]
    if Prior_type == "LL":
        PRIORS_LIST = [Prior_fun(d, b, BB_SIZE=BB_SIZE, nthreads=nthreads) for d, b in zip(DataList, BETAList)]
    else:
        allidx = np.concatenate(idx)
        tmp = Prior_fun(M.select_cols(allidx), BETA_vec[allidx], BB_SIZE=BB_SIZE, nthreads=nthreads)
        PRIORS_LIST = [tmp for _ in Levels]
    outs = []
    for d, b, P in zip(DataList, BETAList, PRIORS_LIST):
        MU_input = P["MME_prior"]["MME_MU"]
        SIZE_input = P["MME_SIZE_adjust"] if BB_SIZE else P["MME_prior"]["MME_SIZE"]
        outs.append(fn(d, b, SIZE_input, MU_input, S=S, **kw))
    names = ["Group %s" % lv for lv in Levels]
    return {"Bay_out_list": dict(zip(names, outs)), "PRIORS_LIST": dict(zip(names, PRIORS_LIST)),
            "input_params": input_params}


# --------------------------------------------------------------------------------------
# Independent numpy restatements of the closed forms (cross-checks for the C code)
# --------------------------------------------------------------------------------------
def np_beta_default(A, mean_beta=0.06):
    xx = A.sum(axis=0)
    b = xx / xx.mean() * mean_beta
    b[b <= 0] = b[b > 0].min()
    b[b >= 1] = b[b < 1].max()
    return b


def np_mme(A, beta=None):
    N = A / beta[None, :] if beta is not None else A
    n = A.shape[1]
    m = N.sum(axis=1) / n
    var = ((N - m[:, None]) ** 2).sum(axis=1) / (n - 1)
    v = (n - 1) / n * var
    with np.errstate(divide="ignore", invalid="ignore"):
        size = m ** 2 / (v - m)
    size[v <= m] = np.nan
    return m, size, v


def np_post_mean(A, beta, size, mu):
    x = A
    s = size[:, None]
    m = mu[:, None]
    b = beta[None, :]
    return (x + s) * (m - m * b) / (s + m * b) + x


def np_post_mode(A, beta, size, mu):
    x = A
    s = size[:, None]
    m = mu[:, None]
    b = beta[None, :]
    meann = (x + s) * (m - m * b) / (s + m * b)
    r = s + x
    with np.errstate(invalid="ignore", divide="ignore"):
        out = np.floor(meann / r * (r - 1)) + x
    return np.where(r <= 1, 0.0, out)


def np_adjust_size(bb, mme_size):
    fit = (bb < bb.max()) & (bb > bb.min())
    X = np.stack([np.ones(fit.sum()), np.log(mme_size[fit])], axis=1)
    coef, *_ = np.linalg.lstsq(X, np.log(bb[fit]), rcond=None)
    return np.exp(coef[0] + coef[1] * np.log(mme_size))


def np_marginal_nb_1d(SIZE, MU, m_observed, BETA):
    from scipy.stats import nbinom
    mu = np.asarray(BETA) * MU
    return float(np.sum(nbinom.logpmf(np.asarray(m_observed), SIZE, SIZE / (SIZE + mu))))


def load_fixture():
    d = np.load(_HERE.parent / "tests" / "golden" / "example_data.npz")
    return {k: d[k] for k in d.files}
