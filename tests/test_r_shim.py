"""r_shim/baynorm_b200_shim.c EXECUTED against a stand-in for R's C API (tests/rmock/): there is no R in this image, but
the shim's own logic -- argument coercion, length checks, external pointers, dgCMatrix construction, the registration
table, error paths -- runs for real, on the GPU box against the real libbaynorm_b200.so.  The CPU tests check what needs no
device: the table is the reference's (src/RcppExports.cpp:211-228), host-only routines (t_sp, asMatrix), and that a missing
GPU surfaces as an R error, not a crash or a fallback."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SO = ROOT / "tests" / "rmock" / "librmock_shim.so"


@pytest.fixture(scope="module")
def R():
    lib_dir = ROOT / "baynorm_b200"
    cmd = ["gcc", "-O1", "-shared", "-fPIC", "-Wall", "-I", str(ROOT / "tests" / "rmock"), "-I", str(ROOT / "include"), "-o", str(SO),
           str(ROOT / "r_shim" / "baynorm_b200_shim.c"), str(ROOT / "tests" / "rmock" / "rmock.c"), "-L", str(lib_dir),
           "-lbaynorm_b200", "-Wl,-rpath," + str(lib_dir)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lib = C.CDLL(str(SO))
    V = C.c_void_p
    for name, res, args in [("rmock_real", V, [V, C.c_long]), ("rmock_int", V, [V, C.c_long]), ("rmock_lgl", V, [C.c_int]),
                            ("rmock_nil", V, []), ("rmock_dgC", V, [C.c_int, C.c_int, V, V, V]), ("rmock_type", C.c_int, [V]),
                            ("rmock_len", C.c_long, [V]), ("rmock_data", V, [V]), ("rmock_attr", V, [V, C.c_char_p]),
                            ("rmock_elt", V, [V, C.c_long]), ("rmock_text", C.c_char_p, [V]), ("rmock_last_error", C.c_char_p, []),
                            ("rmock_call", V, [C.c_char_p, C.c_int, C.POINTER(V)]), ("rmock_n_routines", C.c_int, []),
                            ("rmock_routine", C.c_char_p, [C.c_int, C.POINTER(C.c_int)]), ("rmock_run_finalizers", C.c_int, []),
                            ("rmock_set_dim2", None, [V, C.c_int, C.c_int]), ("rmock_set_seed", None, [C.c_ulonglong]),
                            ("rmock_unload", None, [])]:
        f = getattr(lib, name)
        f.restype, f.argtypes = res, args
    return Rmock(lib)


class RError(RuntimeError):
    pass


class Rmock:
    REALSXP, INTSXP, LGLSXP, VECSXP, S4SXP = 14, 13, 10, 19, 25

    def __init__(self, lib):
        self.lib = lib

    def real(self, a):
        a = np.ascontiguousarray(a, np.float64).reshape(-1, order="F")
        return self.lib.rmock_real(a.ctypes.data, a.size)

    def integer(self, a):
        a = np.ascontiguousarray(a, np.int32).reshape(-1, order="F")
        return self.lib.rmock_int(a.ctypes.data, a.size)

    def lgl(self, v):
        return self.lib.rmock_lgl(int(bool(v)))

    def matrix(self, A, integer=False):
        A = np.asarray(A)
        s = (self.integer if integer else self.real)(np.asfortranarray(A).reshape(-1, order="F"))
        self.lib.rmock_set_dim2(s, A.shape[0], A.shape[1])
        return s

    def dgC(self, X):
        import scipy.sparse as sp
        M = sp.csc_matrix(np.asarray(X, np.float64))
        M.sort_indices()
        return self.lib.rmock_dgC(M.shape[0], M.shape[1], self.integer(M.indptr), self.integer(M.indices), self.real(M.data))

    def call(self, name, *args):
        arr = (C.c_void_p * len(args))(*args)
        r = self.lib.rmock_call(name.encode(), len(args), arr)
        if not r:
            raise RError(self.lib.rmock_last_error().decode())
        return r

    def value(self, s):
        t, n = self.lib.rmock_type(s), self.lib.rmock_len(s)
        if t == self.REALSXP:
            v = np.ctypeslib.as_array(C.cast(self.lib.rmock_data(s), C.POINTER(C.c_double)), shape=(n,)).copy() if n else np.zeros(0)
        elif t in (self.INTSXP, self.LGLSXP):
            v = np.ctypeslib.as_array(C.cast(self.lib.rmock_data(s), C.POINTER(C.c_int)), shape=(n,)).copy() if n else np.zeros(0, np.int32)
        elif t == self.VECSXP:
            return [self.value(self.lib.rmock_elt(s, k)) for k in range(n)]
        else:
            raise TypeError("SEXP type %d" % t)
        dim = self.lib.rmock_attr(s, b"dim")
        if dim:
            v = v.reshape(tuple(self.value(dim)), order="F")
        return v

    def slots(self, obj):
        assert self.lib.rmock_type(obj) == self.S4SXP and self.lib.rmock_text(obj) == b"dgCMatrix"
        return {k: self.value(self.lib.rmock_attr(obj, k.encode())) for k in ("i", "p", "x", "Dim")}


def test_registration_table_is_the_references(R):
    """All 15 routines of src/RcppExports.cpp:211-228, same names and arities, plus the batched entries."""
    got = {}
    for k in range(R.lib.rmock_n_routines()):
        n = C.c_int()
        name = R.lib.rmock_routine(k, C.byref(n)).decode()
        got[name] = n.value
    want = {"_bayNorm_DownSampling": 2, "_bayNorm_MarginalF_NB_1D": 4, "_bayNorm_MarginalF_NB_2D": 3, "_bayNorm_GradientFun_NB_1D": 4,
            "_bayNorm_GradientFun_NB_2D": 3, "_bayNorm_Main_NB_Bay": 6, "_bayNorm_Main_mean_NB_Bay": 6, "_bayNorm_Main_mode_NB_Bay": 6,
            "_bayNorm_rowMeansFast": 1, "_bayNorm_rowVarsFast": 2, "_bayNorm_EstPrior_sprcpp": 1, "_bayNorm_EstPrior_rcpp": 1,
            "_bayNorm_t_sp": 1, "_bayNorm_asMatrix": 5, "_bayNorm_Main_mean_NB_spBay": 6}
    for n, a in want.items():
        assert got.get(n) == a, (n, got.get(n))
    with pytest.raises(RError, match="Incorrect number of arguments"):
        R.call("_bayNorm_DownSampling", R.real([1.0]))
    with pytest.raises(RError, match="not available"):
        R.call("_bayNorm_nope", R.real([1.0]))


def test_host_side_routines_t_sp_and_asMatrix(R):
    """t_sp (src/bayNorm_main.cpp:1472-1477) and asMatrix (:1497-1513) need no device."""
    rng = np.random.default_rng(0)
    X = (rng.random((13, 9)) < 0.3) * rng.integers(1, 9, (13, 9))
    t = R.slots(R.call("_bayNorm_t_sp", R.dgC(X)))
    import scipy.sparse as sp
    T = sp.csc_matrix((t["x"], t["i"], t["p"]), shape=tuple(t["Dim"]))
    assert tuple(t["Dim"]) == (9, 13) and np.array_equal(T.toarray(), X.T) and T.has_sorted_indices
    r, c = np.nonzero(X)
    m = R.value(R.call("_bayNorm_asMatrix", R.real(r), R.integer(c), R.real(X[r, c]), R.integer([13]), R.real([9.0])))
    assert m.dtype == np.int32 and np.array_equal(m, X)
    with pytest.raises(RError, match="out of bounds"):
        R.call("_bayNorm_asMatrix", R.real([13.0]), R.real([0.0]), R.real([1.0]), R.integer([13]), R.integer([9]))


def test_missing_gpu_is_an_r_error_not_a_fallback(R):
    """Without a usable sm_100 device the first device routine raises an R condition with the library's message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: covered by the gpu tests")
    with pytest.raises(RError, match="no CPU fallback|sm_100"):
        R.call("_bayNorm_rowMeansFast", R.dgC(np.eye(3)))


@pytest.mark.gpu
def test_shim_routines_against_the_python_binding(R, oracle, fixture_data):
    """Every reference routine through the shim equals the same library call through ctypes, with the argument types R users
    actually pass: integer count matrices, integer S, BETA as a plain double vector."""
    import baynorm_b200 as bn
    X, beta, size, mu = (fixture_data[k] for k in ("inputdata", "inputbeta", "size", "mu"))
    D = R.dgC(X)
    S, thres = R.integer([4]), R.integer([10000000])
    mean = R.value(R.call("_bayNorm_Main_mean_NB_Bay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
    assert mean.shape == X.shape and np.array_equal(mean, bn.Main_mean_NB_Bay(X, beta, size, mu))
    assert np.array_equal(mean, oracle.Main_mean_NB_Bay(oracle.CSC.from_dense(X), beta, size, mu))
    mode = R.value(R.call("_bayNorm_Main_mode_NB_Bay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
    assert np.array_equal(mode, oracle.Main_mode_NB_Bay(oracle.CSC.from_dense(X), beta, size, mu))
    # size given as an integer vector: Rcpp's NumericVector coerced it, so does the shim
    isz = np.maximum(np.round(size), 1).astype(np.int32)
    m2 = R.value(R.call("_bayNorm_Main_mean_NB_Bay", D, R.real(beta), R.integer(isz), R.real(mu), S, thres))
    assert np.array_equal(m2, bn.Main_mean_NB_Bay(X, beta, isz.astype(float), mu))
    # 3-D draws: set.seed() determinism (tests/testthat/test-bayNorm.r:9-17) and the reference's dim = c(G, C, S)
    R.lib.rmock_set_seed(7)
    d1 = R.value(R.call("_bayNorm_Main_NB_Bay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
    R.lib.rmock_set_seed(7)
    d2 = R.value(R.call("_bayNorm_Main_NB_Bay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
    d3 = R.value(R.call("_bayNorm_Main_NB_Bay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
    assert d1.shape == X.shape + (4,) and np.array_equal(d1, d2) and not np.array_equal(d1, d3) and (d1 >= X[:, :, None]).all()
    # sparse mean: every non-zero of the dense mean, as a dgCMatrix (src/bayNorm_main.cpp:1598)
    sp_ = R.slots(R.call("_bayNorm_Main_mean_NB_spBay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
    import scipy.sparse as sps
    got = sps.csc_matrix((sp_["x"], sp_["i"], sp_["p"]), shape=tuple(sp_["Dim"])).toarray()
    assert np.array_equal(got, mean) and sp_["p"][-1] == np.count_nonzero(mean)
    # MME: the sparse and the dense routine both return the RAW size m^2 / (v - m) (:1386, :1441)
    Xn = X / beta[None, :]
    es = R.value(R.call("_bayNorm_EstPrior_sprcpp", R.dgC(Xn)))
    ed = R.value(R.call("_bayNorm_EstPrior_rcpp", R.matrix(Xn)))
    raw = dict(oracle.EstPrior(oracle.CSC.from_dense(Xn)))
    with np.errstate(divide="ignore", invalid="ignore"):
        raw["SIZE"] = raw["MU"] ** 2 / (raw["v"] - raw["MU"])      # EstPrior_sprcpp returns the quotient as is (:1386)
    for k, nm in enumerate(("MU", "SIZE", "v")):
        np.testing.assert_allclose(es[k], raw[nm], rtol=1e-12)
        np.testing.assert_allclose(ed[k], raw[nm], rtol=1e-10)
    assert not np.isnan(es[1]).any()                               # raw quotient: negative where v < m, not NaN-ruled
    rm = R.value(R.call("_bayNorm_rowMeansFast", D))
    np.testing.assert_allclose(rm, X.mean(1), rtol=1e-13)
    rv = R.value(R.call("_bayNorm_rowVarsFast", D, R.real(rm)))
    np.testing.assert_allclose(rv, X.var(1, ddof=1), rtol=1e-12)
    # likelihood entry points
    g = 3
    ll = R.value(R.call("_bayNorm_MarginalF_NB_1D", R.real([size[g]]), R.real([mu[g]]), R.integer(X[g].astype(np.int32)), R.real(beta)))
    assert abs(ll[0] - oracle.MarginalF_NB_1D(size[g], mu[g], X[g], beta)) <= 1e-12 * abs(ll[0])
    g2 = R.value(R.call("_bayNorm_GradientFun_NB_2D", R.real([size[g], mu[g]]), R.real(X[g]), R.real(beta)))
    np.testing.assert_allclose(g2, oracle.GradientFun_NB_2D([size[g], mu[g]], X[g], beta), rtol=1e-9, atol=1e-9)
    # DownSampling on an INTEGER matrix (what as.matrix of counts often is in R)
    ds = R.value(R.call("_bayNorm_DownSampling", R.matrix(X.astype(np.int32), integer=True), R.real(beta)))
    assert ds.shape == X.shape and (ds <= X).all() and (ds >= 0).all() and np.array_equal(ds, np.floor(ds))
    # batched entries: Prior_fun in one call
    pr = R.value(R.call("bnb_prior", D, R.real(beta), R.lgl(1), R.lgl(0)))
    P = bn.Prior_fun(X, beta, verbose=False)
    assert np.array_equal(pr[0], P["MME_prior"]["MME_MU"]) and np.array_equal(pr[3], P["MME_SIZE_adjust"])
    bd = R.value(R.call("bnb_beta_default", D))
    np.testing.assert_allclose(bd.mean(), 0.06, rtol=1e-12)
    assert R.lib.rmock_run_finalizers() == 0                      # every device matrix was released eagerly


@pytest.mark.gpu
def test_shim_errors_do_not_leak_device_matrices(R, fixture_data):
    """An R error raised between upload and return leaves the device matrix to the external pointer's finalizer."""
    X, beta, size, mu = (fixture_data[k] for k in ("inputdata", "inputbeta", "size", "mu"))
    D = R.dgC(X)
    R.lib.rmock_run_finalizers()
    with pytest.raises(RError, match="BETA_vec has length 73, expected 74"):
        R.call("_bayNorm_Main_mean_NB_Bay", D, R.real(beta[:-1]), R.real(size), R.real(mu), R.integer([1]), R.integer([1]))
    with pytest.raises(RError, match="mu must be supplied"):
        R.call("_bayNorm_Main_mode_NB_Bay", D, R.real(beta), R.real(size), R.lib.rmock_nil(), R.integer([1]), R.integer([1]))
    assert R.lib.rmock_run_finalizers() == 2                      # two uploads were pending: the "GC" frees them
    assert R.lib.rmock_run_finalizers() == 0
    bad = beta.copy()
    bad[0] = 1.5
    with pytest.raises(RError, match="baynorm_b200"):
        R.call("bnb_fit_size", D, R.real(bad), R.real(mu), R.real(size), R.real([-1.0]), R.real([2.0]), R.lgl(0))
    with pytest.raises(RError, match="Data@p has length"):
        R.call("_bayNorm_rowMeansFast", R.lib.rmock_dgC(20, 75, R.integer(np.zeros(75)), R.integer([]), R.real([])))
    R.lib.rmock_run_finalizers()


@pytest.mark.gpu
def test_shim_multi_device_route(R, fixture_data):
    """bnb_set_devices(n): Prior_fun and the posterior routines over n GPUs of this session (the stand-in for the reference's
    SOCK cluster, R/PRIOR_FUNCTIONS.R:426-427) return what the one-device route returns.  With one GPU in the box the call
    must still be accepted for n = 1 and refuse n = 2 as an R error."""
    import ctypes as CT
    try:
        rt = CT.CDLL("libcudart.so.12")
    except OSError:
        rt = CT.CDLL("libcudart.so")
    n = CT.c_int(0)
    ndev = n.value if rt.cudaGetDeviceCount(CT.byref(n)) == 0 else 0
    X, beta, size, mu = (fixture_data[k] for k in ("inputdata", "inputbeta", "size", "mu"))
    D = R.dgC(X)
    S, thres = R.integer([3]), R.integer([10000000])
    assert R.value(R.call("bnb_set_devices", R.integer([1])))[0] == 1
    one = R.value(R.call("bnb_prior", D, R.real(beta), R.lgl(1), R.lgl(0)))
    mean1 = R.value(R.call("_bayNorm_Main_mean_NB_Bay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
    if ndev < 2:
        with pytest.raises(RError):
            R.call("bnb_set_devices", R.integer([2]))
        return
    try:
        assert R.value(R.call("bnb_set_devices", R.integer([2])))[0] == 2
        two = R.value(R.call("bnb_prior", D, R.real(beta), R.lgl(1), R.lgl(0)))
        for k in range(3):
            assert np.array_equal(one[k], two[k])
        np.testing.assert_allclose(two[3], one[3], rtol=1e-12)
        mean2 = R.value(R.call("_bayNorm_Main_mean_NB_Bay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
        assert np.array_equal(mean1, mean2)
        R.lib.rmock_set_seed(11)
        d2 = R.value(R.call("_bayNorm_Main_NB_Bay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
    finally:
        R.call("bnb_set_devices", R.integer([1]))
    R.lib.rmock_set_seed(11)
    d1 = R.value(R.call("_bayNorm_Main_NB_Bay", D, R.real(beta), R.real(size), R.real(mu), S, thres))
    assert np.array_equal(d1, d2)
    assert R.lib.rmock_run_finalizers() == 0
