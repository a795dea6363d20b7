"""Host-side mirror of bayNorm's R interface over the C ABI.

The reference's host language is R (plus generated Rcpp glue); neither exists in this
image, so the layer above the C ABI is written in Python and keeps the reference's names,
argument meaning, defaults and error behaviour:

    bayNorm()            R/bayNorm.r:120-363        bayNorm_sup()      R/bayNorm.r:457-617
    Prior_fun()          R/PRIOR_FUNCTIONS.R:241    EstPrior()         R/PRIOR_FUNCTIONS.R:105
    BB_Fun()             R/PRIOR_FUNCTIONS.R:380    AdjustSIZE_fun()   R/sup_funs.R:27
    Check_input()        R/sup_funs.R:79            DownSampling()     src/bayNorm_main.cpp:892
    Main_NB_Bay / Main_mean_NB_Bay / Main_mode_NB_Bay / Main_mean_NB_spBay
                                                   src/bayNorm_main.cpp:1063-1299, :1519
    MarginalF_NB_1D/2D, GradientFun_NB_1D/2D       src/bayNorm_main.cpp:928-1013
    rowMeansFast, rowVarsFast, EstPrior_sprcpp, EstPrior_rcpp   :1303-1449

Everything numeric happens in libbaynorm_b200.so on the GPU; this file only marshals
pointers (numpy arrays, scipy.sparse matrices or CUDA torch tensors) and assembles the
Bay_out / PRIORS / input_params structures (a12 in SURVEY.md section 8).  R lists become
dicts with the same names; R matrices become Fortran-ordered numpy arrays (same memory
layout as R).  The R `.Call` shim that binds the same C ABI is in r_shim/ and INTEGRATION.md.
"""
from __future__ import annotations

import ctypes as C
import math
import warnings

import numpy as np

from . import _lib
from ._lib import BayNormError, FitOpts, PriorOut

_F64 = np.float64


# --------------------------------------------------------------------------------------
# pointers
# --------------------------------------------------------------------------------------
def _is_torch(a):
    return type(a).__module__.startswith("torch") and hasattr(a, "data_ptr")


def _ptr(a):
    """Address of a numpy array / torch tensor (host or CUDA), or None."""
    if a is None:
        return None
    if _is_torch(a):
        if not a.is_contiguous():
            raise ValueError("tensor arguments must be contiguous")
        return C.c_void_p(a.data_ptr())
    if not isinstance(a, np.ndarray):
        raise TypeError("expected numpy array or torch tensor, got %r" % type(a))
    if not (a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]):
        raise ValueError("array arguments must be contiguous")
    return C.c_void_p(a.ctypes.data)


def _vec(a, n=None, name="vector"):
    """float64 contiguous host vector (or a float64 CUDA tensor, passed through)."""
    if _is_torch(a):
        import torch
        if a.dtype != torch.float64:
            a = a.to(torch.float64)
        a = a.contiguous().reshape(-1)
    else:
        a = np.ascontiguousarray(np.asarray(a, dtype=_F64).reshape(-1))
    if n is not None and a.shape[0] != n:
        raise ValueError("%s has length %d, expected %d" % (name, a.shape[0], n))
    return a


class RMatrix(np.ndarray):
    """An R matrix with dimnames: a Fortran-ordered float64 array whose columns can also be taken by name
    (`BB_prior[:, 0]` and `BB_prior["BB_SIZE"]`), e.g. cbind(BB_SIZE, MME_MU) of R/PRIOR_FUNCTIONS.R:298."""

    def __new__(cls, columns, colnames, rownames=None):
        obj = np.asfortranarray(np.column_stack([np.asarray(c, dtype=_F64) for c in columns])).view(cls)
        obj.colnames = list(colnames)
        obj.rownames = None if rownames is None else list(rownames)
        return obj

    def __array_finalize__(self, obj):
        self.colnames = getattr(obj, "colnames", None)
        self.rownames = getattr(obj, "rownames", None)

    def __getitem__(self, key):
        if isinstance(key, str):
            return np.asarray(self)[:, self.colnames.index(key)]
        return np.asarray(self).__getitem__(key)


# --------------------------------------------------------------------------------------
# context / matrix handles
# --------------------------------------------------------------------------------------
class Context:
    """One GPU (bn_ctx).  Raises if the CUDA library or an sm_100 device is unavailable."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.bn_ctx_create(int(device), C.byref(h))
        if rc != 0:
            raise BayNormError(rc, self.lib.bn_last_error(None).decode())
        self.h = h
        self.device = int(device)
        self._hook = None          # keep the ctypes callback alive

    def check(self, rc):
        if rc != 0:
            raise BayNormError(rc, self.lib.bn_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.bn_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self.check(self.lib.bn_ctx_synchronize(self.h))

    @property
    def stream(self):
        return self.lib.bn_ctx_stream(self.h)

    @property
    def launch_count(self):
        return int(self.lib.bn_ctx_launch_count(self.h))

    def device_info(self):
        out = (C.c_int64 * 4)()
        self.check(self.lib.bn_ctx_device_info(self.h, out))
        return {"sms": out[0], "clock_khz": out[1], "cc": out[2], "mem_bytes": out[3]}

    def last_fit(self):
        """What the last likelihood fit launched: kernel variant, genes parked for the cluster kernel, ..."""
        out = (C.c_int64 * 4)()
        self.check(self.lib.bn_debug_last_fit(self.h, out))
        return {"variant": int(out[0]), "parked": int(out[1]), "park_after": int(out[2]), "analytic_gradient": bool(out[3])}

    def tune(self, knob, value):
        """Profiling A/B switch (bn_debug_tune); knob is one of TUNE_* below, value 0 restores the default."""
        self.check(self.lib.bn_debug_tune(self.h, int(knob), int(value)))

    def set_allreduce(self, fn, rank, world):
        """fn(buf_ptr:int, count:int, dtype:int, op:int, stream_ptr:int) -> int (0 = ok)."""
        if fn is None:
            self._hook = None
            self.check(self.lib.bn_ctx_set_allreduce(self.h, _lib.ALLREDUCE_FN(), None, 0, 1))
            return

        def _tramp(user, buf, count, dtype, op, stream):
            try:
                return int(fn(buf or 0, int(count), int(dtype), int(op), stream or 0) or 0)
            except Exception as e:  # never let an exception cross the C boundary
                warnings.warn("all-reduce hook failed: %r" % (e,))
                return 1

        self._hook = _lib.ALLREDUCE_FN(_tramp)
        self.check(self.lib.bn_ctx_set_allreduce(self.h, self._hook, None, int(rank), int(world)))


_default_ctx = {}


def default_context(device=0):
    c = _default_ctx.get(device)
    if c is None or c.h is None:
        c = _default_ctx[device] = Context(device)
    return c


class DeviceMatrix:
    """A gene x cell count matrix resident in HBM (bn_mat): the dgCMatrix of the reference."""

    def __init__(self, ctx, handle, rownames=None, colnames=None):
        self.ctx = ctx
        self.h = handle
        info = (C.c_int64 * 4)()
        ctx.check(ctx.lib.bn_mat_info(handle, info))
        self.G, self.C, self.nnz, self.is_count = int(info[0]), int(info[1]), int(info[2]), bool(info[3])
        self.rownames = rownames
        self.colnames = colnames

    @property
    def shape(self):
        return (self.G, self.C)

    def close(self):
        if getattr(self, "h", None):
            if getattr(self.ctx, "h", None):      # frees are ordered on the context's stream
                self.ctx.lib.bn_mat_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_origin(self, gene0=0, cell0=0):
        self.ctx.check(self.ctx.lib.bn_mat_set_origin(self.h, int(gene0), int(cell0)))

    def select_cols(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int64)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.bn_mat_select_cols(self.ctx.h, self.h, _ptr(idx), idx.size, C.byref(h)))
        cn = None if self.colnames is None else [self.colnames[k] for k in idx]
        return DeviceMatrix(self.ctx, h, self.rownames, cn)

    def export_csc(self):
        p = np.zeros(self.C + 1, np.int64)
        i = np.zeros(self.nnz, np.int32)
        x = np.zeros(self.nnz, _F64)
        self.ctx.check(self.ctx.lib.bn_mat_export_csc(self.ctx.h, self.h, _ptr(p), _ptr(i), _ptr(x)))
        return p, i, x

    @classmethod
    def from_csc(cls, ctx, G, Cn, p, i, x, **kw):
        if _is_torch(p):
            import torch
            p64 = 1 if p.dtype == torch.int64 else 0
        else:
            p = np.ascontiguousarray(p)
            if p.dtype not in (np.int32, np.int64):
                p = p.astype(np.int64)
            p64 = 1 if p.dtype == np.int64 else 0
            i = np.ascontiguousarray(i, dtype=np.int32)
            x = np.ascontiguousarray(x, dtype=_F64)
        h = C.c_void_p()
        ctx.check(ctx.lib.bn_mat_from_csc(ctx.h, int(G), int(Cn), _ptr(p), p64, _ptr(i), _ptr(x), C.byref(h)))
        return cls(ctx, h, **kw)

    @classmethod
    def from_dense(cls, ctx, A, **kw):
        if _is_torch(A):          # (C, G) row-major CUDA tensor == column-major G x C
            Cn, G = A.shape
            ptr = _ptr(A)
        else:
            A = np.asfortranarray(np.asarray(A, dtype=_F64))
            G, Cn = A.shape
            ptr = _ptr(A)
        h = C.c_void_p()
        ctx.check(ctx.lib.bn_mat_from_dense(ctx.h, int(G), int(Cn), ptr, C.byref(h)))
        return cls(ctx, h, **kw)

    @classmethod
    def synth_nb(cls, ctx, mu, size, beta, seed, gene0=0):
        mu, size, beta = _vec(mu), _vec(size), _vec(beta)
        h = C.c_void_p()
        ctx.check(ctx.lib.bn_mat_synth_nb(ctx.h, mu.shape[0], beta.shape[0], _ptr(mu), _ptr(size), _ptr(beta),
                                          int(seed) & (2 ** 64 - 1), int(gene0), C.byref(h)))
        return cls(ctx, h)


def Check_input(Data, ctx=None, rownames=None, colnames=None):
    """R/sup_funs.R:79-110: anything -> dgCMatrix; here: anything -> DeviceMatrix (CSC in HBM)."""
    if isinstance(Data, DeviceMatrix):
        return Data
    ctx = ctx or default_context()
    try:
        import scipy.sparse as sp
    except Exception:  # pragma: no cover
        sp = None
    if sp is not None and sp.issparse(Data):
        # never touch the caller's matrix: tocsc() hands back the SAME object when the input already is CSC, and the
        # canonical form the library wants (no duplicate entries, row indices increasing in a column) is made in place
        M = Data.tocsc(copy=True) if Data.format == "csc" else Data.tocsc()
        M.sum_duplicates()
        M.sort_indices()
        G, Cn = M.shape
        return DeviceMatrix.from_csc(ctx, G, Cn, M.indptr, M.indices, M.data, rownames=rownames, colnames=colnames)
    if isinstance(Data, (tuple, list)) and len(Data) == 5:     # (G, C, p, i, x)
        G, Cn, p, i, x = Data
        return DeviceMatrix.from_csc(ctx, G, Cn, p, i, x, rownames=rownames, colnames=colnames)
    return DeviceMatrix.from_dense(ctx, Data, rownames=rownames, colnames=colnames)


# --------------------------------------------------------------------------------------
# registered .Call routines (src/RcppExports.cpp:211-228), same names
# --------------------------------------------------------------------------------------
def _scalar_call(fn, ctx, *args):
    out = np.zeros(1, _F64)
    ctx.check(fn(ctx.h, *args, _ptr(out)))
    return float(out[0])


def MarginalF_NB_1D(SIZE, MU, m_observed, BETA, ctx=None):
    ctx = ctx or default_context()
    m, b = _vec(m_observed), _vec(BETA)
    return _scalar_call(ctx.lib.bn_marginal_nb_1d, ctx, float(SIZE), float(MU), _ptr(m), _ptr(b), m.shape[0])


def MarginalF_NB_2D(SIZE_MU, m_observed, BETA, ctx=None):
    ctx = ctx or default_context()
    sm, m, b = _vec(SIZE_MU, 2), _vec(m_observed), _vec(BETA)
    return _scalar_call(ctx.lib.bn_marginal_nb_2d, ctx, _ptr(sm), _ptr(m), _ptr(b), m.shape[0])


def GradientFun_NB_1D(SIZE, MU, m_observed, BETA, ctx=None):
    ctx = ctx or default_context()
    m, b = _vec(m_observed), _vec(BETA)
    return _scalar_call(ctx.lib.bn_gradient_nb_1d, ctx, float(SIZE), float(MU), _ptr(m), _ptr(b), m.shape[0])


def GradientFun_NB_2D(SIZE_MU, m_observed, BETA, ctx=None):
    ctx = ctx or default_context()
    sm, m, b = _vec(SIZE_MU, 2), _vec(m_observed), _vec(BETA)
    out = np.zeros(2, _F64)
    ctx.check(ctx.lib.bn_gradient_nb_2d(ctx.h, _ptr(sm), _ptr(m), _ptr(b), m.shape[0], _ptr(out)))
    return out


def rowMeansFast(x, ctx=None):
    M = Check_input(x, ctx)
    out = np.zeros(M.G, _F64)
    M.ctx.check(M.ctx.lib.bn_row_means(M.ctx.h, M.h, _ptr(out)))
    return out


def rowVarsFast(x, means, ctx=None):
    M = Check_input(x, ctx)
    mv = _vec(means, M.G, "means")
    out = np.zeros(M.G, _F64)
    M.ctx.check(M.ctx.lib.bn_row_vars(M.ctx.h, M.h, _ptr(mv), _ptr(out)))
    return out


def _mme(M, beta, nan_rule=True):
    MU, SIZE, V = (np.zeros(M.G, _F64) for _ in range(3))
    b = None if beta is None else _vec(beta, M.C, "BETA_vec")
    fn = M.ctx.lib.bn_mme if nan_rule else M.ctx.lib.bn_mme_raw     # raw: the quotient m^2 / (v - m) as EstPrior_sprcpp returns it
    M.ctx.check(fn(M.ctx.h, M.h, _ptr(b), _ptr(MU), _ptr(SIZE), _ptr(V)))
    return MU, SIZE, V


def EstPrior_sprcpp(Data, ctx=None):
    """src/bayNorm_main.cpp:1370-1394: list(MU, SIZE, v); SIZE is the raw quotient."""
    M = Check_input(Data, ctx)
    MU, SIZE, V = _mme(M, None, nan_rule=False)
    return {"MU": MU, "SIZE": SIZE, "v": V}


def EstPrior_rcpp(Data, ctx=None):
    """src/bayNorm_main.cpp:1423-1449 (dense twin)."""
    ctx = ctx or default_context()
    A = np.asfortranarray(np.asarray(Data, dtype=_F64))
    G, Cn = A.shape
    MU, SIZE, V = (np.zeros(G, _F64) for _ in range(3))
    ctx.check(ctx.lib.bn_mme_dense(ctx.h, G, Cn, _ptr(A), _ptr(MU), _ptr(SIZE), _ptr(V)))
    return {"MU": MU, "SIZE": SIZE, "v": V}


def EstPrior(Data, verbose=True, ctx=None):
    """R/PRIOR_FUNCTIONS.R:105-175: MME on data that is already normalised; SIZE NaN where v <= m."""
    M = Check_input(Data, ctx)
    MU, SIZE, _ = _mme(M, None)
    if verbose:
        print("Priors estimation based on MME method has completed.")
    return {"MU": MU, "SIZE": SIZE}


def _posterior(kind, Data, BETA_vec, size, mu, S=20, thres=None, seed=None, ctx=None, out=None, col0=0, ncol=None):
    M = Check_input(Data, ctx)
    ctx = M.ctx
    if mu is None:
        raise ValueError("mu must be supplied (the reference dereferences it unconditionally, "
                         "src/bayNorm_main.cpp:1118)")
    ncol = M.C - col0 if ncol is None else int(ncol)
    b, sz, m = _vec(BETA_vec, M.C, "BETA_vec"), _vec(size, M.G, "size"), _vec(mu, M.G, "mu")
    if kind == "sample":
        if out is None:
            out = np.empty((M.G, ncol, int(S)), dtype=_F64, order="F")
        if seed is None:        # the R shim draws the key from R's RNG; numpy's global RNG here
            seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
        ctx.check(ctx.lib.bn_post_sample(ctx.h, M.h, _ptr(b), _ptr(sz), _ptr(m), int(col0), ncol, int(S),
                                         int(seed) & (2 ** 64 - 1), _ptr(out)))
        return out
    if out is None:
        out = np.empty((M.G, ncol), dtype=_F64, order="F")
    fn = ctx.lib.bn_post_mean if kind == "mean" else ctx.lib.bn_post_mode
    ctx.check(fn(ctx.h, M.h, _ptr(b), _ptr(sz), _ptr(m), int(col0), ncol, _ptr(out)))
    return out


def Main_NB_Bay(Data, BETA_vec, size, mu, S=20, thres=None, **kw):
    """3-D array dim=c(G, C, S) of draws x + NB(size+x, tempmu)  (src/bayNorm_main.cpp:1063-1139)."""
    return _posterior("sample", Data, BETA_vec, size, mu, S, thres, **kw)


def Main_mean_NB_Bay(Data, BETA_vec, size, mu, S=20, thres=None, **kw):
    return _posterior("mean", Data, BETA_vec, size, mu, S, thres, **kw)


def Main_mode_NB_Bay(Data, BETA_vec, size, mu, S=20, thres=None, **kw):
    return _posterior("mode", Data, BETA_vec, size, mu, S, thres, **kw)


def Main_mean_NB_spBay(Data, BETA_vec, size, mu, S=20, thres=None, ctx=None, col0=0, ncol=None):
    """src/bayNorm_main.cpp:1519-1601: the posterior mean as a sparse matrix holding every non-zero value (:1598).  Built on the
    device (bn_post_mean_csc): no dense G x C array crosses the bus or sits in host memory."""
    import scipy.sparse as sp
    M = Check_input(Data, ctx)
    ctx = M.ctx
    if mu is None:
        raise ValueError("mu must be supplied (the reference dereferences it unconditionally, src/bayNorm_main.cpp:1118)")
    ncol = M.C - col0 if ncol is None else int(ncol)
    b, sz, m = _vec(BETA_vec, M.C, "BETA_vec"), _vec(size, M.G, "size"), _vec(mu, M.G, "mu")
    res = C.c_void_p()
    ctx.check(ctx.lib.bn_post_mean_csc(ctx.h, M.h, _ptr(b), _ptr(sz), _ptr(m), int(col0), ncol, C.byref(res)))
    try:
        info = (C.c_int64 * 3)()
        ctx.check(ctx.lib.bn_csc_result_info(res, info))
        nnz = int(info[2])
        p = np.zeros(ncol + 1, np.int64)
        i = np.zeros(nnz, np.int32)
        x = np.zeros(nnz, _F64)
        ctx.check(ctx.lib.bn_csc_result_fetch(ctx.h, res, _ptr(p), None, _ptr(i), _ptr(x)))
    finally:
        ctx.lib.bn_csc_result_free(res)
    return sp.csc_matrix((x, i, p), shape=(M.G, ncol))


def DownSampling(Data, BETA_vec, seed=None, ctx=None):
    """src/bayNorm_main.cpp:892-915: dense matrix in, dense matrix of Binomial(x, beta_j) out."""
    ctx = ctx or default_context()
    A = np.asfortranarray(np.asarray(Data, dtype=_F64))
    G, Cn = A.shape
    b = _vec(BETA_vec, Cn, "BETA_vec")
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
    out = np.empty((G, Cn), dtype=_F64, order="F")
    ctx.check(ctx.lib.bn_downsample(ctx.h, G, Cn, _ptr(A), _ptr(b), int(seed) & (2 ** 64 - 1), _ptr(out)))
    return out


def BetaFun(Data, MeanBETA, ctx=None, rownames=None, with_stats=False):
    """R/PRIOR_FUNCTIONS.R:30-71: capture efficiencies from the library sizes of the genes whose largest
    log(normalised count + 1) is below median + 3 MAD and that are zero in fewer than 35 % of the cells.
    Returns {"BETA", "Selected_genes"} like the reference's list (gene names if `rownames` is given or Data
    carries them, else 0-based row indices); with_stats adds the per-gene median / mad / max / zero fraction."""
    ctx = ctx or default_context()
    M = Data if isinstance(Data, DeviceMatrix) else Check_input(Data, ctx, rownames=rownames)
    beta = np.empty(M.C, dtype=_F64)
    sel = np.empty(M.G, dtype=np.int32)
    stats = np.empty(4 * M.G, dtype=_F64)
    ctx.check(ctx.lib.bn_beta_fun(ctx.h, M.h, float(MeanBETA), _ptr(beta), _ptr(sel), _ptr(stats)))
    idx = np.nonzero(sel)[0]
    names = rownames if rownames is not None else getattr(M, "rownames", None)
    out = {"BETA": beta, "Selected_genes": [names[i] for i in idx] if names is not None else idx}
    if with_stats:
        st = stats.reshape(4, M.G)
        out["stats"] = {"median": st[0], "mad": st[1], "max": st[2], "dropout": st[3]}
    return out


def SyntheticControl(Data, BETA_vec, seed=None, ctx=None):
    """R/synthetic_controls.r:57-70: Poisson control counts with the genes' normalised means, binomially
    down-sampled with the cells' capture efficiencies.  Returns {"N_c", "beta_c", "lambda"} like the
    reference's list; the G x C Poisson matrix is never materialised (one fused kernel)."""
    ctx = ctx or default_context()
    M = Data if isinstance(Data, DeviceMatrix) else Check_input(Data, ctx)
    b = _vec(BETA_vec, M.C, "BETA_vec")
    if seed is None:
        seed = int(np.random.randint(0, 2 ** 63 - 1, dtype=np.int64))
    lam = np.empty(M.G, dtype=_F64)
    N_c = np.empty((M.G, M.C), dtype=_F64, order="F")
    ctx.check(ctx.lib.bn_synthetic_control(ctx.h, M.h, _ptr(b), int(seed) & (2 ** 64 - 1), _ptr(lam), _ptr(N_c)))
    return {"N_c": N_c, "beta_c": b, "lambda": lam}


# --------------------------------------------------------------------------------------
# one process, several GPUs (bn_group): what `parallel = TRUE, NCores = n` is to the reference
# (R/PRIOR_FUNCTIONS.R:426-427), with gene blocks resident on the devices instead of a SOCK cluster
# --------------------------------------------------------------------------------------
class DeviceGroup:
    """ndev GPUs driven from this process; collectives over NCCL inside the library (bn_group_create)."""

    def __init__(self, ndev, devices=None):
        self.lib = _lib.load()
        h = C.c_void_p()
        dv = (C.c_int * ndev)(*devices) if devices is not None else None
        rc = self.lib.bn_group_create(int(ndev), dv, C.byref(h))
        if rc != 0:
            raise BayNormError(rc, self.lib.bn_group_last_error(None).decode())
        self.h, self.ndev, self._mats = h, int(ndev), []

    def check(self, rc):
        if rc != 0:
            raise BayNormError(rc, self.lib.bn_group_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            for gm in self._mats:
                gm.close()
            self.lib.bn_group_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def matrix(self, Data):
        """dgCMatrix-like (scipy sparse / dense array) -> gene blocks resident on the devices."""
        import scipy.sparse as sp
        A = Data.tocsc(copy=True) if sp.issparse(Data) else sp.csc_matrix(np.asarray(Data, dtype=_F64))
        A.sum_duplicates()
        A.sort_indices()
        G, Cn = A.shape
        p, i, x = np.ascontiguousarray(A.indptr, np.int64), np.ascontiguousarray(A.indices, np.int32), np.ascontiguousarray(A.data, _F64)
        h = C.c_void_p()
        self.check(self.lib.bn_group_mat_from_csc(self.h, G, Cn, _ptr(p), 1, _ptr(i), _ptr(x), C.byref(h)))
        gm = GroupMatrix(self, h, G, Cn)
        self._mats.append(gm)
        return gm

    def beta_default(self, gm, mean_beta=0.06):
        beta = np.zeros(gm.C, _F64)
        self.check(self.lib.bn_group_beta_default(self.h, gm.h, float(mean_beta), _ptr(beta)))
        return beta

    def Prior_fun(self, gm, BETA_vec, BB_SIZE=True, GR=False, method="spg"):
        """Prior_fun(FIX_MU = TRUE) over the group: the dict of api.Prior_fun(return_info=True)."""
        b = _vec(BETA_vec, gm.C, "BETA_vec")
        G = gm.G
        MU, SIZE = np.zeros(G, _F64), np.zeros(G, _F64)
        BB, ADJ = (np.zeros(G, _F64), np.zeros(G, _F64)) if BB_SIZE else (None, None)
        conv, nfe = np.zeros(G, np.int32), np.zeros(G, np.int32)
        po = PriorOut()
        po.MME_MU, po.MME_SIZE = _ptr(MU), _ptr(SIZE)
        po.BB_SIZE, po.MME_SIZE_adjust = _ptr(BB), _ptr(ADJ)
        po.conv, po.nfeval = _ptr(conv), _ptr(nfe)
        o = _fit_opts(GR, method)
        self.check(self.lib.bn_group_prior(self.h, gm.h, _ptr(b), 1 if BB_SIZE else 0, C.byref(o), C.byref(po)))
        out = {"MME_prior": {"MME_MU": MU, "MME_SIZE": SIZE}, "BETA_vec": b}
        if BB_SIZE:
            out["BB_prior"] = RMatrix([BB, MU], ["BB_SIZE", "MME_MU"])
            out["MME_SIZE_adjust"] = ADJ
        out["_info"] = {"conv": conv, "nfeval": nfe, "size_lower": po.size_lower, "size_upper": po.size_upper,
                        "coef": tuple(po.coef), "seconds": tuple(po.seconds)}
        return out

    def posterior(self, kind, gm, BETA_vec, size, mu, S=20, seed=0, col0=0, ncol=None):
        """kind: "mean" | "mode" | "sample" -> G x ncol (x S) array, Fortran order like the reference's mat / cube."""
        k = {"mean": 0, "mode": 1, "sample": 2}[kind]
        ncol = gm.C - col0 if ncol is None else ncol
        b, sz, mm = _vec(BETA_vec, gm.C, "BETA_vec"), _vec(size, gm.G, "size"), _vec(mu, gm.G, "mu")
        shape = (gm.G, ncol, S) if k == 2 else (gm.G, ncol)
        out = np.zeros(shape, _F64, order="F")
        self.check(self.lib.bn_group_post(self.h, gm.h, k, _ptr(b), _ptr(sz), _ptr(mm), int(col0), int(ncol), int(S),
                                          C.c_uint64(int(seed) & (2 ** 64 - 1)), _ptr(out)))
        return out


class GroupMatrix:
    def __init__(self, group, handle, G, Cn):
        self.group, self.h, self.G, self.C = group, handle, int(G), int(Cn)

    def bounds(self):
        b = np.zeros(self.group.ndev + 1, np.int64)
        self.group.check(self.group.lib.bn_gmat_bounds(self.h, _ptr(b)))
        return b

    def close(self):
        if getattr(self, "h", None):
            self.group.lib.bn_gmat_free(self.h)
            self.h = None


# --------------------------------------------------------------------------------------
# R-level functions
# --------------------------------------------------------------------------------------
TUNE_FIT_VARIANT, TUNE_SAMPLE_NCS, TUNE_POST_NCS, TUNE_SAMPLE_SERIAL, TUNE_FIT_SPLIT, TUNE_SAMPLE_CHUNKS = 0, 1, 2, 3, 4, 5


def _fit_opts(GR=False, method="spg", maxfeval=500, park_after=0):
    o = FitOpts()
    _lib.load().bn_fit_opts_default(C.byref(o))
    o.method = {"spg": 0, "brent": 1}[method]
    o.maxfeval = int(maxfeval)
    o.use_gradient = 1 if GR else 0
    o.park_after = int(park_after)
    return o


def BB_Fun(Data, BETA_vec, INITIAL_MU_vec, INITIAL_SIZE_vec, MU_lower=0.01, MU_upper=500, SIZE_lower=0.01,
           SIZE_upper=30, parallel=False, NCores=5, FIX_MU=True, GR=False, method="spg", ctx=None,
           return_info=False, park_after=0):
    """R/PRIOR_FUNCTIONS.R:380-549.  `NCores` is accepted and ignored: the fit of all genes is one kernel
    launch.  FIX_MU=False returns the G x 2 matrix (size, mu) of the reference (:543-546).  `parallel` keeps its one
    visible effect upstream: the parallel=TRUE branch boxes (size, mu) by c(SIZE_lower, MU_lower) .. c(SIZE_upper,
    MU_upper) (:414-419, :457-470), the parallel=FALSE branch hands spg the scalars SIZE_lower / SIZE_upper (:503-526),
    which BB::spg recycles over BOTH parameters."""
    M = Check_input(Data, ctx)
    ctx = M.ctx
    if not FIX_MU and not parallel:
        MU_lower, MU_upper = SIZE_lower, SIZE_upper
    b, mu0, s0 = _vec(BETA_vec, M.C, "BETA_vec"), _vec(INITIAL_MU_vec, M.G), _vec(INITIAL_SIZE_vec, M.G)
    out = np.zeros(M.G, _F64)
    conv = np.zeros(M.G, np.int32)
    nfe = np.zeros(M.G, np.int32)
    o = _fit_opts(GR, method, park_after=park_after)
    if not FIX_MU:
        mu_out = np.zeros(M.G, _F64)
        ctx.check(ctx.lib.bn_fit_size_mu(ctx.h, M.h, _ptr(b), _ptr(mu0), _ptr(s0), float(SIZE_lower), float(SIZE_upper),
                                         float(MU_lower), float(MU_upper), C.byref(o), _ptr(out), _ptr(mu_out),
                                         _ptr(conv), _ptr(nfe)))
        par = np.asfortranarray(np.stack([out, mu_out], axis=1))        # colnames c('size', 'mu')
        return (par, conv, nfe) if return_info else par
    ctx.check(ctx.lib.bn_fit_size(ctx.h, M.h, _ptr(b), _ptr(mu0), _ptr(s0), float(SIZE_lower), float(SIZE_upper),
                                  C.byref(o), _ptr(out), _ptr(conv), _ptr(nfe)))
    return (out, conv, nfe) if return_info else out


def AdjustSIZE_fun(BB_SIZE, MME_MU, MME_SIZE, ctx=None):
    """R/sup_funs.R:27-33 (MME_MU is unused upstream too)."""
    ctx = ctx or default_context()
    bb, ms = _vec(BB_SIZE), _vec(MME_SIZE)
    out = np.zeros(bb.shape[0], _F64)
    coef = np.zeros(3, _F64)
    ctx.check(ctx.lib.bn_adjust_size(ctx.h, bb.shape[0], _ptr(bb), _ptr(ms), _ptr(out), _ptr(coef)))
    return out


def Prior_fun(Data, BETA_vec, parallel=True, NCores=5, FIX_MU=True, GR=False, BB_SIZE=True, verbose=True,
              method="spg", ctx=None, return_info=False):
    """R/PRIOR_FUNCTIONS.R:241-311 in one native call (bn_prior).  FIX_MU=False runs the MME step natively,
    then the 2-D fit in the box the reference builds from the MME estimates (:275-279) and the size
    adjustment on BB_prior[, 1] (:291-296); BB_prior is then the G x 2 (size, mu) matrix."""
    M = Check_input(Data, ctx)
    ctx = M.ctx
    b = _vec(BETA_vec, M.C, "BETA_vec")
    if not FIX_MU and BB_SIZE:
        base = Prior_fun(M, b, FIX_MU=True, BB_SIZE=False, verbose=False, ctx=ctx, return_info=True)
        MU, SIZE = base["MME_prior"]["MME_MU"], base["MME_prior"]["MME_SIZE"]
        rng = np.zeros(2, _F64)
        ctx.check(ctx.lib.bn_vec_range(ctx.h, _ptr(MU), M.G, _ptr(rng)))      # min / max over all gene shards
        if verbose:
            print("Start optimization using spg from BB package.\nThis part may be time-consuming.")
        par, conv, nfe = BB_Fun(M, b, MU, SIZE, MU_lower=rng[0], MU_upper=rng[1], SIZE_lower=base["_info"]["size_lower"],
                                SIZE_upper=base["_info"]["size_upper"], parallel=parallel, FIX_MU=False, GR=GR, method=method,
                                ctx=ctx, return_info=True)
        ADJ = AdjustSIZE_fun(np.ascontiguousarray(par[:, 0]), MU, SIZE, ctx=ctx)
        if verbose:
            print("Prior estimation has completed!")
        out = {"MME_prior": base["MME_prior"], "BB_prior": RMatrix([par[:, 0], par[:, 1]], ["size", "mu"], M.rownames), "MME_SIZE_adjust": ADJ,
               "BETA_vec": base["BETA_vec"]}
        if return_info:
            out["_info"] = dict(base["_info"], conv=conv, nfeval=nfe, mu_lower=float(rng[0]), mu_upper=float(rng[1]))
        return out
    MU, SIZE = np.zeros(M.G, _F64), np.zeros(M.G, _F64)
    BB, ADJ = (np.zeros(M.G, _F64), np.zeros(M.G, _F64)) if BB_SIZE else (None, None)
    conv, nfe = np.zeros(M.G, np.int32), np.zeros(M.G, np.int32)
    po = PriorOut()
    po.MME_MU, po.MME_SIZE = _ptr(MU), _ptr(SIZE)
    po.BB_SIZE, po.MME_SIZE_adjust = _ptr(BB), _ptr(ADJ)
    po.conv, po.nfeval = _ptr(conv), _ptr(nfe)
    o = _fit_opts(GR, method)
    if verbose and BB_SIZE:
        print("Start optimization using spg from BB package.\nThis part may be time-consuming.")
    ctx.check(ctx.lib.bn_prior(ctx.h, M.h, _ptr(b), 1 if BB_SIZE else 0, C.byref(o), C.byref(po)))
    beta_host = b.cpu().numpy() if _is_torch(b) else b
    out = {"MME_prior": {"MME_MU": MU, "MME_SIZE": SIZE}, "BETA_vec": beta_host}
    if BB_SIZE:
        out["BB_prior"] = RMatrix([BB, MU], ["BB_SIZE", "MME_MU"], M.rownames)   # cbind(BB_SIZE, MME_MU), :298
        out["MME_SIZE_adjust"] = ADJ
        out = {"MME_prior": out["MME_prior"], "BB_prior": out["BB_prior"], "MME_SIZE_adjust": ADJ,
               "BETA_vec": beta_host}
        if verbose:
            print("Prior estimation has completed!")
    if return_info:
        out["_info"] = {"conv": conv, "nfeval": nfe, "size_lower": po.size_lower, "size_upper": po.size_upper,
                        "coef": tuple(po.coef), "seconds": tuple(po.seconds)}
    return out


def _myFunc(mode_version, mean_version, out_sparse):
    if mode_version and mean_version:
        raise ValueError("Only one of mode_version and mean_version should be specified to be TRUE, otherwise "
                         "both should be set to FALSE so that 3D array normalized data will be returned.")
    if not mode_version and not mean_version:
        return Main_NB_Bay
    if mode_version:
        return Main_mode_NB_Bay
    return Main_mean_NB_spBay if out_sparse else Main_mean_NB_Bay


def _default_beta(M):
    beta = np.zeros(M.C, _F64)
    M.ctx.check(M.ctx.lib.bn_beta_default(M.ctx.h, M.h, 0.06, _ptr(beta)))
    return beta


def _scale_round(Data, sffl):
    """round(Data / UMI_sffl) on the host (R/bayNorm.r:200, :281): input preparation, not hot path."""
    import scipy.sparse as sp
    if isinstance(Data, DeviceMatrix):
        p, i, x = Data.export_csc()
        return (Data.G, Data.C, p, i, np.round(x / sffl))
    if sp.issparse(Data):
        M = Data.tocsc(copy=True)
        M.data = np.round(M.data / sffl)
        return M
    return np.round(np.asarray(Data, dtype=_F64) / sffl)


def bayNorm(Data, BETA_vec=None, Conditions=None, UMI_sffl=None, Prior_type=None, mode_version=False,
            mean_version=False, S=20, parallel=True, NCores=5, FIX_MU=True, GR=False, BB_SIZE=True, verbose=True,
            out_sparse=False, seed=None, ctx=None, rownames=None, colnames=None):
    """R/bayNorm.r:120-363.  Returns dict(Bay_out, PRIORS, input_params) or, with Conditions,
    dict(Bay_out_list, PRIORS_LIST, input_params) keyed "Group <level>"."""
    myFunc = _myFunc(mode_version, mean_version, out_sparse)
    input_params = {"BETA_vec": BETA_vec, "Conditions": Conditions, "UMI_sffl": UMI_sffl, "Prior_type": Prior_type,
                    "FIX_MU": FIX_MU, "BB_SIZE": BB_SIZE, "GR": GR}
    M = Check_input(Data, ctx, rownames, colnames)
    ctx = M.ctx
    if BETA_vec is None:
        BETA_vec = _default_beta(M)                       # :164-170 (+ range check :185-187 in the library)
    BETA_vec = _vec(BETA_vec)
    bh = BETA_vec.cpu().numpy() if _is_torch(BETA_vec) else BETA_vec
    if M.rownames is not None and len(set(M.rownames)) != len(M.rownames):
        warnings.warn("There are duplicated row names in Data")
    if M.colnames is not None and len(set(M.colnames)) != len(M.colnames):
        warnings.warn("There are duplicated column names in Data")
    if bh.min() <= 0 or bh.max() >= 1:
        raise ValueError("The range of BETA must be within (0,1).")
    if M.C != bh.shape[0]:
        raise ValueError("The number of cells(columns) in Data is not consistent with the number of elements "
                         "in BETA_vec.")
    skw = {"seed": seed} if myFunc is Main_NB_Bay else {}
    pkw = dict(parallel=parallel, NCores=NCores, FIX_MU=FIX_MU, GR=GR, BB_SIZE=BB_SIZE, verbose=verbose)
    if Conditions is None:
        Data_sr = M if UMI_sffl is None else Check_input(_scale_round(M, UMI_sffl), ctx, M.rownames, M.colnames)
        PRIORS = Prior_fun(Data_sr, BETA_vec, **pkw)
        MU_input = PRIORS["MME_prior"]["MME_MU"]
        SIZE_input = PRIORS["MME_SIZE_adjust"] if BB_SIZE else PRIORS["MME_prior"]["MME_SIZE"]
        Bay_out = myFunc(Data_sr, BETA_vec, SIZE_input, MU_input, S=S, **skw)
        if verbose:
            print("bayNorm has completed!")
        return {"Bay_out": Bay_out, "PRIORS": PRIORS, "input_params": input_params}
    Conditions = np.asarray(Conditions)
    if M.C != Conditions.shape[0]:
        raise ValueError("Number of columns in expression matrix must match length of conditions vector!")
    if Prior_type is None:
        warnings.warn("Prior_type needs to be specified when Conditions are specified, now Prior_type is set "
                      "to be LL")
        Prior_type = "LL"
    _, first = np.unique(Conditions, return_index=True)
    Levels = Conditions[np.sort(first)]
    idx = [np.nonzero(Conditions == lv)[0] for lv in Levels]
    BETAList = [np.ascontiguousarray(bh[ix]) for ix in idx]
    DataList_sr = []
    for k, ix in enumerate(idx):
        d = M.select_cols(ix)
        if UMI_sffl is not None:
            sf = np.asarray(UMI_sffl, dtype=_F64).reshape(-1)
            d = Check_input(_scale_round(d, sf[k] if sf.size > 1 else sf[0]), ctx, d.rownames, d.colnames)
        DataList_sr.append(d)
    if Prior_type == "LL":
        PRIORS_LIST = [Prior_fun(d, b, **pkw) for d, b in zip(DataList_sr, BETAList)]
    elif Prior_type == "GG":
        allidx = np.concatenate(idx)
        if UMI_sffl is None:
            dall = M.select_cols(allidx)
        else:
            import scipy.sparse as sp
            parts = []
            for d in DataList_sr:
                p, i, x = d.export_csc()
                parts.append(sp.csc_matrix((x, i, p), shape=(d.G, d.C)))
            dall = Check_input(sp.hstack(parts, format="csc"), ctx)
        tmp = Prior_fun(dall, np.concatenate(BETAList), **pkw)
        PRIORS_LIST = [tmp for _ in Levels]
    else:
        raise ValueError("Prior_type must be 'LL' or 'GG'")
    names = ["Group %s" % lv for lv in Levels]
    Bay_out_list = {}
    for nm, d, b, P in zip(names, DataList_sr, BETAList, PRIORS_LIST):
        MU_input = P["MME_prior"]["MME_MU"]
        SIZE_input = P["MME_SIZE_adjust"] if BB_SIZE else P["MME_prior"]["MME_SIZE"]
        Bay_out_list[nm] = myFunc(d, b, SIZE_input, MU_input, S=S, **skw)
    if verbose:
        print("bayNorm has completed!")
    return {"Bay_out_list": Bay_out_list, "PRIORS_LIST": dict(zip(names, PRIORS_LIST)), "input_params": input_params}


def bayNorm_sup(Data, PRIORS=None, input_params=None, mode_version=False, mean_version=False, S=20, parallel=True,
                NCores=5, BB_SIZE=True, verbose=True, out_sparse=False, seed=None, ctx=None):
    """R/bayNorm.r:457-617: posterior only, re-using PRIORS / input_params of an earlier bayNorm()."""
    myFunc = _myFunc(mode_version, mean_version, out_sparse)
    Conditions = input_params.get("Conditions")
    UMI_sffl = input_params.get("UMI_sffl")
    if Conditions is None:
        BETA_vec = PRIORS["BETA_vec"]
    else:
        BETA_vec = np.concatenate([PRIORS[k]["BETA_vec"] for k in PRIORS])
    if (not input_params.get("BB_SIZE")) and BB_SIZE:
        raise ValueError("Previous priors does not contain Adjusted MME size. Try to run bayNorm with "
                         "BB_SIZE=TRUE.")
    M = Check_input(Data, ctx)
    ctx = M.ctx
    skw = {"seed": seed} if myFunc is Main_NB_Bay else {}
    if Conditions is None:
        Data_sr = M if UMI_sffl is None else Check_input(_scale_round(M, UMI_sffl), ctx)
        MU_input = PRIORS["MME_prior"]["MME_MU"]
        SIZE_input = PRIORS["MME_SIZE_adjust"] if BB_SIZE else PRIORS["MME_prior"]["MME_SIZE"]
        Bay_out = myFunc(Data_sr, BETA_vec, SIZE_input, MU_input, S=S, **skw)
        if verbose:
            print("bayNorm has completed!")
        return {"Bay_out": Bay_out, "PRIORS": PRIORS, "input_params": input_params}
    Conditions = np.asarray(Conditions)
    if M.C != Conditions.shape[0]:
        raise ValueError("Number of columns in expression matrix must match length of conditions vector!")
    _, first = np.unique(Conditions, return_index=True)
    Levels = Conditions[np.sort(first)]
    names = ["Group %s" % lv for lv in Levels]
    # BETA_vec is the concatenation of the per-group vectors in PRIORS order; upstream indexes
    # it with which(Conditions == level) (:559-561), so positions follow the ORIGINAL column order
    bh = np.asarray(BETA_vec, dtype=_F64)
    Bay_out_list = {}
    for k, (nm, lv) in enumerate(zip(names, Levels)):
        ix = np.nonzero(Conditions == lv)[0]
        d = M.select_cols(ix)
        if UMI_sffl is not None:
            sf = np.asarray(UMI_sffl, dtype=_F64).reshape(-1)
            d = Check_input(_scale_round(d, sf[k] if sf.size > 1 else sf[0]), ctx)
        P = PRIORS[nm]
        MU_input = P["MME_prior"]["MME_MU"]
        SIZE_input = P["MME_SIZE_adjust"] if BB_SIZE else P["MME_prior"]["MME_SIZE"]
        Bay_out_list[nm] = myFunc(d, bh[ix], SIZE_input, MU_input, S=S, **skw)
    if verbose:
        print("bayNorm has completed!")
    return {"Bay_out_list": Bay_out_list, "PRIORS_LIST": PRIORS, "input_params": input_params}
