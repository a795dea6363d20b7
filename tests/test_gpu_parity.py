"""GPU parity tests: the CUDA path, called through the C ABI (ctypes), against the CPU oracle
on the same inputs.  Tolerances are BASELINE.json's: posterior mode bit-exact; posterior mean,
MME mu/size and adjusted size <= 1e-6 relative (asserted much tighter where the arithmetic
allows); fitted BB_SIZE against the oracle running the same solver; draws in distribution."""
import numpy as np
import pytest

from util import PHILOX_KAT, rel_err, synth_counts

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bn():
    import baynorm_b200
    return baynorm_b200


@pytest.fixture(scope="module")
def small(oracle):
    X, mu, size, beta = synth_counts(300, 500, seed=11)
    return {"X": X, "mu": mu, "size": size, "beta": beta, "Mo": oracle.CSC.from_dense(X)}


@pytest.fixture(scope="module")
def dense50(oracle):
    X, mu, size, beta = synth_counts(120, 260, seed=5, m0=3.013)      # ~50 % non-zero, like config 2
    return {"X": X, "mu": mu, "size": size, "beta": beta, "Mo": oracle.CSC.from_dense(X)}


# ---------------------------------------------------------------------------- plumbing
def test_library_loaded_and_device(bn, ctx):
    info = ctx.device_info()
    assert info["cc"] // 10 == 10, info
    assert info["sms"] >= 100


def test_philox_known_answers(bn, ctx):
    import ctypes as C
    for ctr, key, want in PHILOX_KAT:
        c = np.array(ctr, np.uint32)
        k = np.array(key, np.uint32)
        o = np.zeros(4, np.uint32)
        ctx.check(ctx.lib.bn_debug_philox(ctx.h, C.c_void_p(c.ctypes.data), C.c_void_p(k.ctypes.data),
                                          C.c_void_p(o.ctypes.data)))
        assert tuple(int(v) for v in o) == want


def test_device_log(bn, ctx):
    """The table-driven FP64 logarithms of the fit kernel against 50-digit mpmath."""
    import ctypes as C
    import mpmath as mp
    mp.mp.dps = 50
    rng = np.random.default_rng(0)

    def run(which, x):
        x = np.ascontiguousarray(x, np.float64)
        o = np.zeros_like(x)
        ctx.check(ctx.lib.bn_debug_log(ctx.h, which, C.c_void_p(x.ctypes.data), x.size, C.c_void_p(o.ctypes.data)))
        return o

    x = np.concatenate([np.exp(rng.uniform(np.log(1e-300), np.log(1e300), 4000)), rng.uniform(0.5, 2.0, 4000),
                        1 + rng.uniform(-1, 1, 2000) * 1e-6, [1.0, 0.6875, 1.375, 2.0, 0.5, 1e-3, 64.0]])
    got = run(0, x)
    ex = np.array([float(mp.log(mp.mpf(float(v)))) for v in x])
    err = np.abs(got - ex)
    assert (err <= 2.5e-16 * np.abs(ex) + 2e-18).all(), float((err / (2.5e-16 * np.abs(ex) + 2e-18)).max())
    mod = (x > 1e-6) & (x < 1e6)          # the range the likelihood uses: within 1 ulp
    assert (err[mod] <= 2.3e-16 * np.abs(ex[mod]) + 2e-18).all()     # <= 1 ulp (an ulp is 1.1e-16 .. 2.2e-16 relative)
    y = np.concatenate([np.exp(rng.uniform(np.log(1e-30), np.log(1e6), 6000)), rng.uniform(0, 3, 3000), [0.0, 1.0, 1e-17]])
    got = run(1, y)
    ex = np.array([float(mp.log1p(mp.mpf(float(v)))) for v in y])
    err = np.abs(got - ex)
    assert (err <= 2.3e-16 * np.abs(ex) + 2e-18).all(), float((err / (2.3e-16 * np.abs(ex) + 2e-18)).max())
    ys = np.concatenate([np.exp(rng.uniform(np.log(1e-30), np.log(2.0 ** -6), 4000)), [0.0, 2.0 ** -6]])
    got = run(2, ys)
    ex = np.array([float(mp.log1p(mp.mpf(float(v)))) for v in ys])
    assert (np.abs(got - ex) <= 2.3e-16 * np.abs(ex)).all()
    z = np.concatenate([rng.uniform(64, 200, 2000), np.exp(rng.uniform(np.log(64), np.log(1e7), 2000))])
    got = run(3, z)
    ex = np.array([float(mp.loggamma(mp.mpf(float(v)))) for v in z])
    assert (np.abs(got - ex) <= 6e-16 * np.abs(ex)).all()


def test_upload_roundtrip_and_validation(bn, ctx, small, oracle):
    Mo = small["Mo"]
    for p in (Mo.p.astype(np.int32), Mo.p):
        M = bn.DeviceMatrix.from_csc(ctx, Mo.G, Mo.C, p, Mo.i, Mo.x)
        assert (M.G, M.C, M.nnz, M.is_count) == (Mo.G, Mo.C, Mo.nnz, True)
        p2, i2, x2 = M.export_csc()
        assert np.array_equal(p2, Mo.p) and np.array_equal(i2, Mo.i) and np.array_equal(x2, Mo.x)
    Md = bn.DeviceMatrix.from_dense(ctx, small["X"])
    p2, i2, x2 = Md.export_csc()
    assert np.array_equal(p2, Mo.p) and np.array_equal(i2, Mo.i) and np.array_equal(x2, Mo.x)
    bad = Mo.i.copy()
    bad[[0, 1]] = bad[[1, 0]]                       # unsorted rows inside column 0
    if Mo.p[1] >= 2:
        with pytest.raises(bn.BayNormError):
            bn.DeviceMatrix.from_csc(ctx, Mo.G, Mo.C, Mo.p, bad, Mo.x)
    frac = bn.DeviceMatrix.from_csc(ctx, Mo.G, Mo.C, Mo.p, Mo.i, Mo.x * 0.5)
    assert not frac.is_count


def test_select_cols(bn, ctx, small, oracle):
    Mo = small["Mo"]
    M = bn.Check_input(small["X"], ctx)
    idx = np.array([5, 3, 499, 0, 17, 17], np.int64)
    S = M.select_cols(idx)
    p, i, x = S.export_csc()
    So = Mo.select_cols(idx)
    assert np.array_equal(p, So.p) and np.array_equal(i, So.i) and np.array_equal(x, So.x)


# ------------------------------------------------------------------------ beta, MME, OLS
def test_default_beta(bn, ctx, small, oracle, fixture_data):
    for X in (small["X"], fixture_data["inputdata"]):
        M = bn.Check_input(X, ctx)
        got = bn.api._default_beta(M)
        want, bad = oracle.beta_default(oracle.CSC.from_dense(X))
        assert not bad
        assert rel_err(got, want).max() < 1e-13
    # a cell with no counts is clamped to the smallest positive beta (R/bayNorm.r:168)
    X = small["X"].copy()
    X[:, 7] = 0
    got = bn.api._default_beta(bn.Check_input(X, ctx))
    want, _ = oracle.beta_default(oracle.CSC.from_dense(X))
    assert rel_err(got, want).max() < 1e-13 and got[7] == got[got > 0].min()


def test_beta_range_error(bn, ctx):
    X = np.zeros((4, 3))
    X[0, 0] = 1000.0                                  # beta = (0.18, 0, 0) -> clamps fix it
    X[1, 1] = 1.0
    X[2, 2] = 1.0
    b = bn.api._default_beta(bn.Check_input(X, ctx))
    assert 0 < b.min() and b.max() < 1
    with pytest.raises(ValueError, match="range of BETA"):
        bn.bayNorm(X, BETA_vec=np.array([0.5, 1.0, 0.2]), verbose=False)


def test_colsums(bn, ctx, small, oracle):
    import ctypes as C
    M = bn.Check_input(small["X"], ctx)
    out = np.zeros(M.C)
    ctx.check(ctx.lib.bn_colsums(ctx.h, M.h, C.c_void_p(out.ctypes.data)))
    assert np.array_equal(out, small["X"].sum(axis=0))          # integer counts: exact


@pytest.mark.parametrize("which", ["small", "dense50", "fixture"])
def test_mme(bn, ctx, oracle, small, dense50, fixture_data, which):
    if which == "fixture":
        X, beta = fixture_data["inputdata"], fixture_data["inputbeta"]
    else:
        d = small if which == "small" else dense50
        X, beta = d["X"], d["beta"]
    Mo = oracle.CSC.from_dense(X)
    M = bn.Check_input(X, ctx)
    MU, SIZE, V = bn.api._mme(M, beta)
    want = oracle.EstPrior(Mo, beta)
    assert rel_err(MU, want["MU"]).max() < 1e-13
    assert rel_err(V, want["v"]).max() < 1e-12
    assert np.array_equal(np.isnan(SIZE), np.isnan(want["SIZE"]))
    ok = ~np.isnan(SIZE)
    # size = m^2/(v-m): the subtraction amplifies the ulp-level differences of v
    amp = np.abs(want["v"][ok]) / np.abs(want["v"][ok] - want["MU"][ok])
    assert (rel_err(SIZE[ok], want["SIZE"][ok]) < 1e-12 * np.maximum(amp, 1)).all()
    assert rel_err(SIZE[ok], want["SIZE"][ok]).max() < 1e-6
    # registered routines
    assert rel_err(bn.rowMeansFast(M), X.mean(axis=1)).max() < 1e-13
    rv = bn.rowVarsFast(M, X.mean(axis=1))
    assert rel_err(rv, X.var(axis=1, ddof=1)).max() < 1e-12
    raw = bn.EstPrior_sprcpp(M)
    w = oracle.EstPrior(Mo, None)
    assert rel_err(raw["MU"], w["MU"]).max() < 1e-13
    d = bn.EstPrior_rcpp(X)
    assert rel_err(d["MU"], w["MU"]).max() < 1e-13 and rel_err(d["v"], w["v"]).max() < 1e-12


def test_all_zero_gene_and_nan_fill(bn, ctx, oracle, small):
    X = small["X"].copy()
    X[3, :] = 0
    X[9, :] = 0
    Mo = oracle.CSC.from_dense(X)
    P = bn.Prior_fun(bn.Check_input(X, ctx), small["beta"], BB_SIZE=False, verbose=False)
    Po = oracle.Prior_fun(Mo, small["beta"], BB_SIZE=False)
    assert P["MME_prior"]["MME_MU"][3] == 0 and not np.isnan(P["MME_prior"]["MME_SIZE"]).any()
    assert rel_err(P["MME_prior"]["MME_SIZE"], Po["MME_prior"]["MME_SIZE"]).max() < 1e-6
    assert P["MME_prior"]["MME_SIZE"][3] == P["MME_prior"]["MME_SIZE"].min()


def test_adjust_size(bn, ctx, oracle):
    rng = np.random.default_rng(3)
    ms = rng.lognormal(0, 0.8, 5000)
    bb = np.clip(np.exp(0.4 + 0.7 * np.log(ms) + rng.normal(0, 0.2, 5000)), 0.05, 6.0)
    got = bn.AdjustSIZE_fun(bb, None, ms, ctx=ctx)
    assert rel_err(got, oracle.AdjustSIZE_fun(bb, None, ms)).max() < 1e-11
    assert rel_err(got, oracle.np_adjust_size(bb, ms)).max() < 1e-11


# ------------------------------------------------------------------------ likelihood, fit
def test_likelihood_and_gradient(bn, ctx, oracle, small, fixture_data):
    rng = np.random.default_rng(0)
    cases = [(fixture_data["inputdata"][g], fixture_data["inputbeta"], float(fixture_data["mu"][g])) for g in range(6)]
    for g in rng.choice(300, 12, replace=False):
        if small["X"][g].sum() > 0:
            cases.append((small["X"][g], small["beta"], float(small["X"][g].sum() / small["beta"].sum())))
    worst = 0.0
    for x, beta, mu in cases:
        for size in (0.05, 0.37, 1.0, 2.5, 11.0, 140.0):
            want = oracle.MarginalF_NB_1D(size, mu, x, beta)
            got = bn.MarginalF_NB_1D(size, mu, x, beta, ctx=ctx)
            worst = max(worst, abs(got - want) / abs(want))
            assert abs(got - want) <= 1e-12 * abs(want) + 1e-12, (size, mu, got, want)
            assert abs(bn.MarginalF_NB_2D([size, mu], x, beta, ctx=ctx) - want) <= 1e-12 * abs(want) + 1e-12
            gw = oracle.GradientFun_NB_1D(size, mu, x, beta)
            gg = bn.GradientFun_NB_1D(size, mu, x, beta, ctx=ctx)
            scale = np.abs(x).sum() / size + len(x)          # size of the terms that cancel
            assert abs(gg - gw) <= 1e-11 * scale, (size, mu, gg, gw)
            g2w = oracle.GradientFun_NB_2D([size, mu], x, beta)
            g2g = bn.GradientFun_NB_2D([size, mu], x, beta, ctx=ctx)
            assert abs(g2g[0] - g2w[0]) <= 1e-11 * scale and abs(g2g[1] - g2w[1]) <= 1e-11 * (np.abs(x).sum() / mu + 1)
    print("worst relative likelihood difference GPU vs oracle: %.3g" % worst)


def _fit_case(bn, ctx, oracle, X, beta):
    Mo = oracle.CSC.from_dense(X)
    Po = oracle.Prior_fun(Mo, beta, BB_SIZE=False)
    mu0, s0 = Po["MME_prior"]["MME_MU"], Po["MME_prior"]["MME_SIZE"]
    lo, hi = float(s0.min()), float(np.ceil(s0.max()))
    want, conv_o, cnt = oracle.BB_Fun(Mo, beta, mu0, s0, lo, hi, want_counts=True)
    got, conv, nfe = bn.BB_Fun(bn.Check_input(X, ctx), beta, mu0, s0, SIZE_lower=lo, SIZE_upper=hi, return_info=True)
    return Mo, mu0, s0, lo, hi, want, got, conv, conv_o, nfe, cnt


@pytest.mark.parametrize("which", ["fixture", "small", "dense50"])
def test_fit_spg_against_oracle(bn, ctx, oracle, small, dense50, fixture_data, which):
    if which == "fixture":
        X, beta = fixture_data["inputdata"], fixture_data["inputbeta"]
    else:
        d = small if which == "small" else dense50
        X, beta = d["X"], d["beta"]
    Mo, mu0, s0, lo, hi, want, got, conv, conv_o, nfe, cnt = _fit_case(bn, ctx, oracle, X, beta)
    re = rel_err(got, want)
    print("%s: BB_SIZE rel diff GPU vs oracle spg: median %.2e  p99 %.2e  max %.2e; <=1e-6: %.1f%%; feval GPU %d vs oracle %d"
          % (which, np.median(re), np.quantile(re, 0.99), re.max(), 100 * (re <= 1e-6).mean(), nfe.sum(),
             (cnt[:, 0] + cnt[:, 1]).sum()))
    assert np.array_equal(conv, conv_o)
    # same solver, same stopping rules: the two end on the same plateau of the likelihood ...
    dense = Mo.to_dense()
    for g in np.argsort(re)[-5:]:
        fa = oracle.MarginalF_NB_1D(got[g], mu0[g], dense[g], beta)
        fb = oracle.MarginalF_NB_1D(want[g], mu0[g], dense[g], beta)
        assert abs(fa - fb) <= 1e-9 * abs(fb) + 1e-9
    # ... and the stopping points agree to the north-star tolerance for nearly every gene; the
    # rest differ only through last-bit differences of the summation order feeding BB::spg's
    # forward-difference gradient (eps = 1e-7), which the reference itself is subject to.
    assert (re <= 1e-6).mean() >= 0.90
    assert re.max() <= 2e-4
    assert ((got >= lo) & (got <= hi)).all()
    # genes the oracle leaves on a bound are on the same bound here
    assert np.array_equal(want == hi, got == hi) and np.array_equal(want == lo, got == lo)


@pytest.fixture(scope="module")
def long_sparse(oracle):
    """Long genes, single-cell sparsity (C >= 8192, ~8 % non-zero): the shape class of BASELINE configs 3-5, which
    the launcher sends to the 4-genes-per-sweep kernel (k_fit_size_ms<4,4,4>) and, past park_after evaluations, to
    the cluster kernel."""
    X, mu, size, beta = synth_counts(96, 8400, seed=31)
    Mo = oracle.CSC.from_dense(X)
    Po = oracle.Prior_fun(Mo, beta, BB_SIZE=False)
    mu0, s0 = Po["MME_prior"]["MME_MU"], Po["MME_prior"]["MME_SIZE"]
    # BB_Fun's own default upper bound (R/PRIOR_FUNCTIONS.R:383): 85 % of the genes end strictly inside the box, up to
    # 129 evaluations per gene (Prior_fun's ceiling(max MME size) = 1 here would leave two thirds of them on the bound)
    lo, hi = float(s0.min()), 30.0
    want, conv_o, cnt = oracle.BB_Fun(Mo, beta, mu0, s0, lo, hi, nthreads=8, want_counts=True)
    assert ((want > lo) & (want < hi)).mean() > 0.7
    # the reference algorithm against ITSELF with the cells in another order (the same likelihood, summed in a different
    # order): how far two correct runs of BB::spg on this matrix land apart (tests/test_oracle.py::test_spg_jitter...)
    jit = []
    for k in range(2):
        perm = np.random.default_rng(100 + k).permutation(X.shape[1])
        w2 = oracle.BB_Fun(oracle.CSC.from_dense(X[:, perm]), beta[perm], mu0, s0, lo, hi, nthreads=8)
        jit.append(rel_err(w2, want))
    return dict(X=X, beta=beta, Mo=Mo, mu0=mu0, s0=s0, lo=lo, hi=hi, want=want, conv_o=conv_o, cnt=cnt,
                jitter_frac=min(float((j <= 1e-6).mean()) for j in jit), jitter_max=max(float(j.max()) for j in jit))


def _assert_fit_matches_oracle(oracle, d, got, conv, tag):
    """The assertions of test_fit_spg_against_oracle on a prepared case."""
    want, lo, hi, beta, mu0 = d["want"], d["lo"], d["hi"], d["beta"], d["mu0"]
    re = rel_err(got, want)
    print("%s: BB_SIZE rel diff GPU vs oracle spg: median %.2e  p99 %.2e  max %.2e; <=1e-6: %.1f%%"
          % (tag, np.median(re), np.quantile(re, 0.99), re.max(), 100 * (re <= 1e-6).mean()))
    assert np.array_equal(conv, d["conv_o"])
    for g in np.argsort(re)[-5:]:
        fa = oracle.MarginalF_NB_1D(got[g], mu0[g], d["X"][g], beta)
        fb = oracle.MarginalF_NB_1D(want[g], mu0[g], d["X"][g], beta)
        assert abs(fa - fb) <= 1e-9 * abs(fb) + 1e-9
    # North-star tolerance 1e-6 -- up to the jitter of the reference algorithm itself: BB::spg differentiates by forward
    # differences with eps = 1e-7, so the last bits of l(phi) (summation order) move its stopping point.  The oracle run
    # on the same matrix with permuted cells agrees with itself to 1e-6 for jitter_frac of the genes (~80 % at 8400
    # cells) with a maximum of jitter_max; the GPU must be no further from the oracle than the oracle is from itself.
    print("    oracle vs oracle (cells permuted): <=1e-6 for %.1f%%, max %.2e" % (100 * d["jitter_frac"], d["jitter_max"]))
    assert (re <= 1e-6).mean() >= min(0.90, d["jitter_frac"] - 0.05)
    assert re.max() <= max(3.0 * d["jitter_max"], 1e-6)
    assert ((got >= lo) & (got <= hi)).all()
    assert np.array_equal(want == hi, got == hi) and np.array_equal(want == lo, got == lo)


def test_fit_long_sparse_genes(bn, ctx, oracle, long_sparse):
    """The shape class of BASELINE configs 3, 4 and 5 (R/PRIOR_FUNCTIONS.R:446-456 per gene): version 2 of the fit (dense part
    from the per-call table, non-zeros staged by count: k_fit2) by default, and the first-generation kernels -- CTA per gene,
    2 and 4 genes per sweep (k_fit_size_ms<4,4,4>) -- behind the tuning switch.  All are the same function of the data."""
    d = long_sparse
    assert (d["X"] != 0).mean() < 0.2
    M = bn.Check_input(d["X"], ctx)
    got, conv, nfe = bn.BB_Fun(M, d["beta"], d["mu0"], d["s0"], SIZE_lower=d["lo"], SIZE_upper=d["hi"], return_info=True)
    info = ctx.last_fit()
    assert info["variant"] == 5, info
    _assert_fit_matches_oracle(oracle, d, got, conv, "version 2")
    tot = (d["cnt"][:, 0] + d["cnt"][:, 1]).sum()
    # evaluation counts: same order, not equal -- near the optimum the oracle's forward-difference gradient is dominated
    # by the rounding noise of its sequential 8400-term sum (~1e-3 >> gtol), so it stops on ftol a few steps later than
    # the GPU, whose sums are less noisy
    print("    evaluations: GPU %d, oracle %d" % (int(nfe.sum()), int(tot)))
    assert 0.7 * tot <= int(nfe.sum()) <= 1.1 * tot, (int(nfe.sum()), int(tot))
    res = {}
    for variant in (1, 2, 3):
        ctx.tune(bn.api.TUNE_FIT_VARIANT, variant)
        try:
            alt, conv_a, _ = bn.BB_Fun(M, d["beta"], d["mu0"], d["s0"], SIZE_lower=d["lo"], SIZE_upper=d["hi"],
                                       return_info=True)
            assert ctx.last_fit()["variant"] == variant
            if variant == 3:
                assert ctx.last_fit()["park_after"] == 128
        finally:
            ctx.tune(bn.api.TUNE_FIT_VARIANT, 0)
        res[variant] = alt
        assert np.array_equal(conv_a, conv)
        _assert_fit_matches_oracle(oracle, d, alt, conv_a, "first-generation kernel %d" % variant)
        ra = rel_err(alt, got)
        assert (ra <= 1e-6).mean() >= d["jitter_frac"] - 0.05 and ra.max() <= 3.0 * d["jitter_max"]
    assert np.array_equal(res[2], res[3])       # same per-thread cell map and summation order: bit-identical


def test_fit_dense_part_table(bn, ctx):
    """Version 2 of the fit reads the all-cells term F(r) = sum_j log(1 + r BETA_j) of the likelihood
    (the x = 0 branch of dnbinom_mu summed over the cells, src/bayNorm_main.cpp:928-943) from a per-call Chebyshev table:
    pinned against 80-bit sums over eleven decades of r, for a short and a long BETA vector."""
    import ctypes as CT
    from baynorm_b200.api import _ptr
    rng = np.random.default_rng(3)
    for C in (137, 60000):
        beta = np.clip(rng.lognormal(0.0, 0.4, C) * 0.06, 0.005, 0.5)
        r = np.exp(rng.uniform(np.log(2e-6), np.log(3e5), 4000))
        out = np.empty_like(r)
        ctx.check(ctx.lib.bn_debug_dense_table(ctx.h, _ptr(beta), C, 1e-6, 1e6, _ptr(r), r.size, _ptr(out)))
        bl = beta.astype(np.longdouble)
        want = np.array([float(np.log1p(np.longdouble(v) * bl).sum()) for v in r[:400]])
        err = np.abs(out[:400] - want) / want
        print("C = %d: table vs 80-bit sum, max rel %.2e, median %.2e" % (C, err.max(), np.median(err)))
        assert np.isfinite(out).all() and err.max() <= 4e-15
        # outside the built range: NaN (the kernels then sum the cells directly)
        r2 = np.array([1e-9, 1e9])
        o2 = np.empty(2)
        ctx.check(ctx.lib.bn_debug_dense_table(ctx.h, _ptr(beta), C, 1e-6, 1e6, _ptr(r2), 2, _ptr(o2)))
        assert np.isnan(o2).all()


@pytest.mark.parametrize("which", ["fixture", "dense50"])
def test_fit_versions_agree(bn, ctx, oracle, dense50, fixture_data, which):
    """Version 2 against the first-generation kernels on short genes (warp per gene), spg and Brent: the same likelihood
    evaluated two ways -- Brent's bracketed arg-max must agree to ~1e-7, spg up to its forward-difference jitter."""
    d = fixture_data if which == "fixture" else dense50
    X, beta = (d["inputdata"], d["inputbeta"]) if which == "fixture" else (d["X"], d["beta"])
    Mo = oracle.CSC.from_dense(X)
    Po = oracle.Prior_fun(Mo, beta, BB_SIZE=False)
    mu0, s0 = Po["MME_prior"]["MME_MU"], Po["MME_prior"]["MME_SIZE"]
    lo, hi = float(s0.min()), float(np.ceil(s0.max()))
    M = bn.Check_input(X, ctx)
    for method in ("spg", "brent"):
        kw = dict(SIZE_lower=lo, SIZE_upper=hi, return_info=True, method=method)
        ctx.tune(bn.api.TUNE_FIT_VARIANT, 5)
        try:
            a, conv_a, nfe_a = bn.BB_Fun(M, beta, mu0, s0, **kw)
            assert ctx.last_fit()["variant"] == 5
        finally:
            ctx.tune(bn.api.TUNE_FIT_VARIANT, 0)
        ctx.tune(bn.api.TUNE_FIT_VARIANT, 4)
        try:
            b, conv_b, nfe_b = bn.BB_Fun(M, beta, mu0, s0, **kw)
            assert ctx.last_fit()["variant"] in (0, 2, 3)
        finally:
            ctx.tune(bn.api.TUNE_FIT_VARIANT, 0)
        re = rel_err(a, b)
        print("%s %s: version 2 vs first generation: median %.2e max %.2e, <=1e-6 %.1f%%; evaluations %d vs %d"
              % (which, method, np.median(re), re.max(), 100 * (re <= 1e-6).mean(), nfe_a.sum(), nfe_b.sum()))
        assert np.array_equal(conv_a, conv_b)
        if method == "brent":
            assert re.max() <= 1e-5
        else:
            assert (re <= 1e-6).mean() >= 0.9 and re.max() <= 2e-4


@pytest.mark.parametrize("method", ["spg", "brent"])
def test_fit_cluster_kernel_takes_parked_genes(bn, ctx, oracle, long_sparse, method):
    """k_fit_cluster: genes still running after park_after evaluations are finished by 8-CTA clusters.  With
    park_after = 6 nearly every gene is parked mid-solve, so the hand-over of the solver state is what is tested."""
    d = long_sparse
    M = bn.Check_input(d["X"], ctx)
    kw = dict(SIZE_lower=d["lo"], SIZE_upper=d["hi"], return_info=True, method=method)
    ctx.tune(bn.api.TUNE_FIT_VARIANT, 3)        # the first-generation sweep kernel (version 2 parks nothing)
    try:
        ref, conv_r, nfe_r = bn.BB_Fun(M, d["beta"], d["mu0"], d["s0"], **kw)
        assert ctx.last_fit()["parked"] == 0
        got, conv, nfe = bn.BB_Fun(M, d["beta"], d["mu0"], d["s0"], park_after=6, **kw)
        info = ctx.last_fit()
    finally:
        ctx.tune(bn.api.TUNE_FIT_VARIANT, 0)
    assert info["variant"] == 3 and info["park_after"] == 6
    assert info["parked"] >= 0.5 * int((nfe_r > 6).sum()) > 0, info
    assert np.array_equal(conv, conv_r)
    if method == "spg":
        _assert_fit_matches_oracle(oracle, d, got, conv, "cluster kernel")
        assert abs(int(nfe.sum()) - int(nfe_r.sum())) <= 0.1 * nfe_r.sum()
    else:
        # Brent under the sweep kernel and under the cluster kernel: the bounded arg-max of the oracle
        for res in (ref, got):
            for g in range(0, 96, 5):
                if d["X"][g].sum() == 0:
                    continue
                xb, fb = oracle.brent_gene_1d(d["mu0"][g], d["X"][g], d["beta"], d["lo"], d["hi"], xatol=1e-10)
                fg = oracle.MarginalF_NB_1D(res[g], d["mu0"][g], d["X"][g], d["beta"])
                assert fg >= fb - 1e-9 * abs(fb)
                assert abs(res[g] - xb) <= 1e-5 * xb + 1e-7 or abs(fg - fb) <= 1e-10 * abs(fb)
        assert rel_err(got, ref).max() <= 1e-5      # Brent brackets the arg-max: no finite-difference jitter


def test_fit_gene_without_entries_but_positive_mean(bn, ctx, oracle, small):
    """BB_Fun takes a caller-supplied INITIAL_MU_vec: a gene with no stored entry and mu > 0 has the likelihood
    -phi sum log1p(mu beta_j / phi), which is not flat (it decreases in phi) -- BB::spg walks to the lower bound."""
    X = small["X"][:40].copy()
    X[5, :] = 0
    X[17, :] = 0
    beta = small["beta"]
    Mo = oracle.CSC.from_dense(X)
    mu0 = np.maximum(X.sum(1) / beta.sum(), 0.0)
    mu0[5] = 0.7                      # all-zero row, positive mean
    s0 = np.full(40, 1.3)
    lo, hi = 0.05, 9.0
    want, conv_o, _ = oracle.BB_Fun(Mo, beta, mu0, s0, lo, hi, want_counts=True)
    got, conv, nfe = bn.BB_Fun(bn.Check_input(X, ctx), beta, mu0, s0, SIZE_lower=lo, SIZE_upper=hi, return_info=True)
    assert got[17] == 1.3 and nfe[17] == 2                 # mean 0: flat, the start is returned
    # decreasing in phi: both end on the lower bound (up to the rounding of par + alpha * d in spg's last step)
    assert abs(want[5] - lo) <= 1e-15 and abs(got[5] - lo) <= 1e-15 and conv[5] == conv_o[5] and nfe[5] > 2
    assert np.median(rel_err(got, want)) < 1e-6


def test_fit_brent_is_argmax(bn, ctx, oracle, small):
    X, beta = small["X"], small["beta"]
    Mo, mu0, s0, lo, hi, want, got, *_ = _fit_case(bn, ctx, oracle, X, beta)
    gb = bn.BB_Fun(bn.Check_input(X, ctx), beta, mu0, s0, SIZE_lower=lo, SIZE_upper=hi, method="brent")
    dense = Mo.to_dense()
    for g in range(0, 300, 13):
        if dense[g].sum() == 0:
            continue
        xb, fb = oracle.brent_gene_1d(mu0[g], dense[g], beta, lo, hi, xatol=1e-10)
        fg = oracle.MarginalF_NB_1D(gb[g], mu0[g], dense[g], beta)
        assert fg >= fb - 1e-9 * abs(fb)          # at least as good a maximum as the CPU Brent
        assert abs(gb[g] - xb) <= 1e-5 * xb + 1e-7 or abs(fg - fb) <= 1e-10 * abs(fb)


def test_fit_with_analytic_gradient(bn, ctx, oracle, small):
    X, beta = small["X"], small["beta"]
    Mo, mu0, s0, lo, hi, want, got, *_ = _fit_case(bn, ctx, oracle, X, beta)
    gr = bn.BB_Fun(bn.Check_input(X, ctx), beta, mu0, s0, SIZE_lower=lo, SIZE_upper=hi, GR=True)
    inner = (want > lo) & (want < hi)
    assert np.median(rel_err(gr[inner], want[inner])) < 1e-4       # same optimum, different gradient noise
    assert np.array_equal(gr == hi, want == hi)


def test_fit_rejects_non_counts(bn, ctx, small):
    M = bn.Check_input(small["X"] * 0.5, ctx)
    with pytest.raises(bn.BayNormError, match="integer counts"):
        bn.BB_Fun(M, small["beta"], np.ones(300), np.ones(300), SIZE_lower=0.1, SIZE_upper=5)


@pytest.mark.parametrize("which", ["fixture", "small"])
def test_prior_fun_end_to_end(bn, ctx, oracle, small, fixture_data, which):
    if which == "fixture":
        X, beta = fixture_data["inputdata"], fixture_data["inputbeta"]
    else:
        X, beta = small["X"], small["beta"]
    P = bn.Prior_fun(bn.Check_input(X, ctx), beta, verbose=False, return_info=True)
    Po = oracle.Prior_fun(oracle.CSC.from_dense(X), beta)
    assert rel_err(P["MME_prior"]["MME_MU"], Po["MME_prior"]["MME_MU"]).max() < 1e-13
    assert rel_err(P["MME_prior"]["MME_SIZE"], Po["MME_prior"]["MME_SIZE"]).max() < 1e-6
    assert (rel_err(P["BB_prior"]["BB_SIZE"], Po["BB_prior"]["BB_SIZE"]) <= 1e-6).mean() >= 0.9
    assert rel_err(P["MME_SIZE_adjust"], Po["MME_SIZE_adjust"]).max() < 1e-6
    # BB_prior is the reference's G x 2 matrix cbind(BB_SIZE, MME_MU) (R/PRIOR_FUNCTIONS.R:298): columns by position and by name
    assert P["BB_prior"].shape == (P["MME_prior"]["MME_MU"].shape[0], 2) and P["BB_prior"].colnames == ["BB_SIZE", "MME_MU"]
    assert np.array_equal(P["BB_prior"]["MME_MU"], P["MME_prior"]["MME_MU"]) and np.array_equal(P["BB_prior"][:, 0], P["BB_prior"]["BB_SIZE"])
    assert P["_info"]["size_upper"] == np.ceil(Po["MME_prior"]["MME_SIZE"].max())


# ------------------------------------------------------------------------ posterior
@pytest.mark.parametrize("which", ["fixture", "small", "dense50", "tall"])
def test_posterior_mean_and_mode(bn, ctx, oracle, small, dense50, fixture_data, which):
    if which == "fixture":
        X, beta, size, mu = (fixture_data[k] for k in ("inputdata", "inputbeta", "size", "mu"))
    elif which == "tall":                   # several 2048-gene tiles, ragged last tile, empty columns
        X, mu, size, beta = synth_counts(4500, 37, seed=2)
        X[:, 5] = 0
        X[2047:2050, :] = 3
    else:
        d = small if which == "small" else dense50
        X, beta, size, mu = d["X"], d["beta"], d["size"], d["mu"]
    size = size.copy()
    size[::7] = 0.6                         # size + x <= 1 branch of the mode (:1275)
    size[1] = 1.0
    Mo = oracle.CSC.from_dense(X)
    M = bn.Check_input(X, ctx)
    mean = bn.Main_mean_NB_Bay(M, beta, size, mu)
    mode = bn.Main_mode_NB_Bay(M, beta, size, mu)
    wmean = oracle.Main_mean_NB_Bay(Mo, beta, size, mu)
    wmode = oracle.Main_mode_NB_Bay(Mo, beta, size, mu)
    assert mean.shape == X.shape and mean.flags["F_CONTIGUOUS"]
    assert np.array_equal(mode, wmode), "posterior mode must be bit-exact"
    assert np.array_equal(mean, wmean), "same IEEE operations in the same order: mean is bit-exact too"
    assert rel_err(mean, oracle.np_post_mean(X, beta, size, mu)).max() < 1e-14
    # column blocks give the same values as the whole matrix
    blk = bn.Main_mode_NB_Bay(M, beta, size, mu, col0=3, ncol=11)
    assert np.array_equal(blk, wmode[:, 3:14])
    sp = bn.Main_mean_NB_spBay(M, beta, size, mu)
    assert np.array_equal(sp.toarray(), wmean)


def test_posterior_device_pointers(bn, ctx, oracle, small):
    torch = pytest.importorskip("torch")
    X, beta, size, mu = small["X"], small["beta"], small["size"], small["mu"]
    Mo = small["Mo"]
    dev = "cuda:0"
    M = bn.DeviceMatrix.from_csc(ctx, Mo.G, Mo.C, torch.from_numpy(Mo.p).to(dev), torch.from_numpy(Mo.i).to(dev),
                                 torch.from_numpy(Mo.x).to(dev))
    out = torch.empty((Mo.C, Mo.G), dtype=torch.float64, device=dev)
    bn.Main_mean_NB_Bay(M, torch.from_numpy(beta).to(dev), torch.from_numpy(size).to(dev),
                        torch.from_numpy(mu).to(dev), out=out)
    assert np.array_equal(out.cpu().numpy().T, oracle.Main_mean_NB_Bay(Mo, beta, size, mu))


def _pooled_pit_chi2(draws, nb_size, nb_mu, x, rng):
    """Randomised probability-integral transform of every draw under its exact NB law; returns the
    chi-square statistic of the PIT values over 50 equal bins (dof 49)."""
    from scipy.stats import nbinom
    k = draws - x[..., None]
    n = nb_size[..., None]
    p = (nb_size / (nb_size + nb_mu))[..., None]
    lo = nbinom.cdf(k - 1, n, p)
    hi = nbinom.cdf(k, n, p)
    u = lo + (hi - lo) * rng.random(k.shape)
    h, _ = np.histogram(u.ravel(), bins=50, range=(0, 1))
    e = u.size / 50
    return float(((h - e) ** 2 / e).sum())


@pytest.mark.parametrize("which", ["fixture", "small"])
def test_posterior_samples_distribution(bn, ctx, oracle, small, fixture_data, which):
    from scipy.stats import chi2
    if which == "fixture":
        X, beta, size, mu = (fixture_data[k] for k in ("inputdata", "inputbeta", "size", "mu"))
        S = 3000
    else:
        X, beta, size, mu = small["X"][:60, :80], small["beta"][:80], small["size"][:60], small["mu"][:60] * 40
        S = 800                                                     # mu*40: exercises the gamma-Poisson branch
    Mo = oracle.CSC.from_dense(X)
    M = bn.Check_input(X, ctx)
    d = bn.Main_NB_Bay(M, beta, size, mu, S=S, seed=12345)
    assert d.shape == (X.shape[0], X.shape[1], S)
    assert np.array_equal(d, np.floor(d)) and (d >= X[:, :, None]).all()
    nb_size, nb_mu = oracle.post_sample_params(Mo, beta, size, mu)
    # per-element mean / variance against the exact NB moments
    mean_th = nb_mu + X
    var_th = nb_mu + nb_mu ** 2 / nb_size
    z = (d.mean(axis=2) - mean_th) / np.sqrt(var_th / S + 1e-300)
    assert np.abs(z).max() < 5.5 and abs(z.mean()) < 4 / np.sqrt(z.size)
    # Var(s^2) ~ (kappa4 + 2 sigma^4)/S with the NB fourth cumulant kappa4 = sigma^2 (1 + 6 sigma^2 / size)
    kap4 = var_th * (1 + 6 * var_th / nb_size)
    vz = (d.var(axis=2, ddof=1) - var_th) / np.sqrt((kap4 + 2 * var_th ** 2) / S + 1e-300)
    assert np.abs(vz).max() < 8
    # pooled randomised-PIT chi-square against the exact pmf, and the same statistic for the
    # oracle's own (gamma-Poisson) draws: "matched in distribution"
    rng = np.random.default_rng(1)
    stat = _pooled_pit_chi2(d, nb_size, nb_mu, X, rng)
    assert stat < chi2.ppf(1 - 1e-6, 49), stat
    So = min(S, 200)
    do = oracle.Main_NB_Bay(Mo, beta, size, mu, S=So, seed=99)
    stat_o = _pooled_pit_chi2(do, nb_size, nb_mu, X, rng)
    assert stat_o < chi2.ppf(1 - 1e-6, 49), stat_o
    # two-sample check GPU vs oracle draws on the pooled, standardised values
    from scipy.stats import ks_2samp
    a = ((d[:, :, :So] - mean_th[..., None]) / np.sqrt(var_th[..., None] + 1e-300)).ravel()
    b = ((do - mean_th[..., None]) / np.sqrt(var_th[..., None] + 1e-300)).ravel()
    assert ks_2samp(a, b).pvalue > 1e-4


def test_posterior_samples_are_a_pure_function(bn, ctx, oracle, small):
    X, beta, size, mu = small["X"], small["beta"], small["size"], small["mu"]
    M = bn.Check_input(X, ctx)
    a = bn.Main_NB_Bay(M, beta, size, mu, S=8, seed=42)
    b = bn.Main_NB_Bay(M, beta, size, mu, S=8, seed=42)
    assert np.array_equal(a, b)                                           # tests/testthat/test-bayNorm.r:9-17
    c = bn.Main_NB_Bay(M, beta, size, mu, S=8, seed=43)
    assert not np.array_equal(a, c)
    # draws do not depend on S, on the column block, or on how the matrix was sharded
    assert np.array_equal(bn.Main_NB_Bay(M, beta, size, mu, S=5, seed=42), a[:, :, :5])
    blk = bn.Main_NB_Bay(M, beta, size, mu, S=8, seed=42, col0=100, ncol=50)
    assert np.array_equal(blk, a[:, 100:150, :])
    Mo = small["Mo"].select_rows(100, 220)
    Ms = bn.DeviceMatrix.from_csc(ctx, Mo.G, Mo.C, Mo.p, Mo.i, Mo.x)
    Ms.set_origin(gene0=100)
    shard = bn.Main_NB_Bay(Ms, beta, size[100:220], mu[100:220], S=8, seed=42)
    assert np.array_equal(shard, a[100:220])
    np.random.seed(5)
    r1 = bn.Main_NB_Bay(M, beta, size, mu, S=2)
    np.random.seed(5)
    assert np.array_equal(r1, bn.Main_NB_Bay(M, beta, size, mu, S=2))      # set.seed() analogue


# ------------------------------------------------------------------------ DownSampling, synth
def test_downsampling(bn, ctx, small):
    from scipy.stats import binom, chi2
    X, beta = small["X"][:40, :60].copy(), small["beta"][:60].copy()
    X[0, :] = 400                            # BTRS branch (n*p >= 30 for beta > 0.075)
    beta[:5] = [0.3, 0.45, 0.7, 0.93, 0.08]
    R = 600
    reps = np.stack([bn.DownSampling(X, beta, seed=1000 + r, ctx=ctx) for r in range(R)], axis=2)
    assert np.array_equal(reps, np.floor(reps)) and (reps >= 0).all() and (reps <= X[:, :, None]).all()
    assert (reps[X == 0] == 0).all()
    assert np.array_equal(bn.DownSampling(X, beta, seed=1000, ctx=ctx), reps[:, :, 0])
    mean_th = X * beta[None, :]
    var_th = mean_th * (1 - beta[None, :])
    nz = var_th > 0
    z = (reps.mean(axis=2)[nz] - mean_th[nz]) / np.sqrt(var_th[nz] / R)
    assert np.abs(z).max() < 5.5
    rng = np.random.default_rng(0)
    k = reps[nz]
    n = X[nz][:, None]
    p = np.broadcast_to(beta[None, :], X.shape)[nz][:, None]
    lo, hi = binom.cdf(k - 1, n, p), binom.cdf(k, n, p)
    u = lo + (hi - lo) * rng.random(k.shape)
    h, _ = np.histogram(u.ravel(), bins=50, range=(0, 1))
    stat = ((h - u.size / 50) ** 2 / (u.size / 50)).sum()
    assert stat < chi2.ppf(1 - 1e-6, 49), stat


def test_synthetic_generator(bn, ctx):
    from util import synth_params
    mu, size, beta = synth_params(3000, 400, seed=8)
    M = bn.DeviceMatrix.synth_nb(ctx, mu, size, beta, seed=77)
    p, i, x = M.export_csc()
    assert M.is_count and p[0] == 0 and p[-1] == M.nnz == x.size
    for j in (0, 1, 199, 399):
        r = i[p[j]:p[j + 1]]
        assert (np.diff(r) > 0).all()
    assert (x > 0).all() and np.array_equal(x, np.floor(x))
    dens = M.nnz / (3000 * 400)
    assert 0.04 < dens < 0.14                 # "~8 % nnz" calibration of SURVEY.md 8(d)
    tot = np.zeros(3000)
    np.add.at(tot, i, x)
    exp = mu * beta.sum()
    big = exp > 50
    assert np.abs(tot[big] / exp[big] - 1).max() < 1.0 and abs(np.median(tot[big] / exp[big]) - 1) < 0.15
    # a gene shard generated on its own equals the same rows of the whole matrix
    Ms = bn.DeviceMatrix.synth_nb(ctx, mu[1000:1500], size[1000:1500], beta, seed=77, gene0=1000)
    ps, is_, xs = Ms.export_csc()
    keep = (i >= 1000) & (i < 1500)
    assert np.array_equal(xs, x[keep]) and np.array_equal(is_, i[keep] - 1000)


# ------------------------------------------------------------------------ bayNorm API
def test_baynorm_gg_mean(bn, ctx, oracle, fixture_data):
    """BASELINE.json configs[0]: fixture, GG, mean_version=TRUE, reference BETA_vec."""
    X, beta = fixture_data["inputdata"], fixture_data["inputbeta"]
    r = bn.bayNorm(X, BETA_vec=beta, mean_version=True, verbose=False, ctx=ctx)
    ro = oracle.bayNorm(oracle.CSC.from_dense(X), BETA_vec=beta, mean_version=True)
    assert set(r) == {"Bay_out", "PRIORS", "input_params"}
    assert set(r["PRIORS"]) == {"MME_prior", "BB_prior", "MME_SIZE_adjust", "BETA_vec"}
    assert rel_err(r["PRIORS"]["MME_SIZE_adjust"], ro["PRIORS"]["MME_SIZE_adjust"]).max() < 1e-6
    assert rel_err(r["Bay_out"], ro["Bay_out"]).max() < 1e-6
    # posterior-only entry re-uses the priors untouched (tests/testthat/test-bayNorm.r:20-29)
    r2 = bn.bayNorm_sup(X, PRIORS=r["PRIORS"], input_params=r["input_params"], mode_version=True, verbose=False, ctx=ctx)
    assert r2["PRIORS"] is r["PRIORS"]
    want = oracle.Main_mode_NB_Bay(oracle.CSC.from_dense(X), beta, r["PRIORS"]["MME_SIZE_adjust"],
                                   r["PRIORS"]["MME_prior"]["MME_MU"])
    assert np.array_equal(r2["Bay_out"], want)


def test_baynorm_default_beta_3d_and_determinism(bn, ctx, fixture_data):
    X = fixture_data["inputdata"][:10, :30]                     # the slice the reference's tests use
    np.random.seed(1258484)
    a = bn.bayNorm(X, BETA_vec=fixture_data["inputbeta"][:30], S=20, parallel=False, verbose=False, ctx=ctx)
    np.random.seed(1258484)
    b = bn.bayNorm(X, BETA_vec=fixture_data["inputbeta"][:30], S=20, parallel=False, verbose=False, ctx=ctx)
    assert a["Bay_out"].shape == (10, 30, 20)
    assert np.array_equal(a["Bay_out"], b["Bay_out"])
    assert np.array_equal(a["PRIORS"]["MME_SIZE_adjust"], b["PRIORS"]["MME_SIZE_adjust"])
    c = bn.bayNorm(X, verbose=False, mean_version=True, ctx=ctx)   # default beta path
    assert c["input_params"]["BETA_vec"] is None and c["PRIORS"]["BETA_vec"].shape == (30,)


def test_baynorm_conditions_ll_and_gg(bn, ctx, oracle, small):
    X, beta = small["X"][:120], small["beta"]
    cond = np.array((["b"] * 150 + ["a"] * 200 + ["b"] * 150))
    Mo = oracle.CSC.from_dense(X)
    for pt in ("LL", "GG"):
        r = bn.bayNorm(X, BETA_vec=beta, Conditions=cond, Prior_type=pt, mode_version=True, verbose=False, ctx=ctx)
        ro = oracle.bayNorm(Mo, BETA_vec=beta, Conditions=cond, Prior_type=pt, mode_version=True)
        assert list(r["Bay_out_list"]) == ["Group b", "Group a"] == list(ro["Bay_out_list"])
        for k in r["Bay_out_list"]:
            P, Pq = r["PRIORS_LIST"][k], ro["PRIORS_LIST"][k]
            # 120 genes x 200-300 cells: the regression of AdjustSIZE_fun sees only ~60 interior genes,
            # so single-gene spg stopping-point jitter (see test_fit_spg_against_oracle) is not averaged out
            assert rel_err(P["MME_SIZE_adjust"], Pq["MME_SIZE_adjust"]).max() < 5e-5
            assert r["Bay_out_list"][k].shape == ro["Bay_out_list"][k].shape
            # the mode is an integer step function of size: compare it on the GPU's own priors
            ix = np.nonzero(cond == k.split()[1])[0]
            want = oracle.Main_mode_NB_Bay(Mo.select_cols(ix), beta[ix], P["MME_SIZE_adjust"], P["MME_prior"]["MME_MU"])
            assert np.array_equal(r["Bay_out_list"][k], want)
    with pytest.raises(ValueError, match="Only one of mode_version"):
        bn.bayNorm(X, BETA_vec=beta, mode_version=True, mean_version=True, ctx=ctx)
    with pytest.raises(ValueError, match="conditions vector"):
        bn.bayNorm(X, BETA_vec=beta, Conditions=cond[:10], ctx=ctx)


# ------------------------------------------------------------------------ gene-sharded path
def _run_gene_sharded(bn, Mo, X, world, work):
    """Runs work(context, shard matrix, rank) on `world` gene shards, each with its own context and the
    all-reduce hook installed (host threads on one GPU; the hook is a barrier + reduce over the shards'
    device buffers).  Returns the list of results."""
    import threading
    torch = pytest.importorskip("torch")
    from baynorm_b200 import dist
    bounds = dist.balanced_gene_blocks(np.diff(Mo.p[:1]).sum() + (X > 0).sum(axis=1), world, n_cells=Mo.C)
    barrier = threading.Barrier(world)
    slots = [None] * world
    results = [None] * world
    errors = []

    def make_hook(rank):
        def hook(buf, count, dtype, op, stream):
            t = torch.as_tensor(dist._CudaBuf(buf, count), device="cuda:0")
            torch.cuda.synchronize()
            slots[rank] = t
            barrier.wait()
            st = torch.stack([s.clone() for s in slots])
            red = st.sum(0) if op == 0 else (st.min(0).values if op == 1 else st.max(0).values)
            torch.cuda.synchronize()
            barrier.wait()
            t.copy_(red)
            torch.cuda.synchronize()
            barrier.wait()
            return 0
        return hook

    def run(rank):
        try:
            c = bn.Context(0)
            c.set_allreduce(make_hook(rank), rank, world)
            g0, g1 = int(bounds[rank]), int(bounds[rank + 1])
            G, Cn, p, i, x = dist.split_csc_rows(Mo.G, Mo.C, Mo.p, Mo.i, Mo.x, g0, g1)
            M = bn.DeviceMatrix.from_csc(c, G, Cn, p, i, x)
            M.set_origin(gene0=g0)
            results[rank] = work(c, M, rank, g0)
            M.close()
            c.close()
        except Exception as e:      # pragma: no cover
            errors.append(e)
            barrier.abort()

    th = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert not errors, errors
    return results


def test_gene_sharded_prior_matches_unsharded(bn, ctx, oracle, small):
    """Two gene shards: BETA, MME, BB_SIZE are bit-identical to the unsharded run, the adjusted size agrees
    to the last few ulps."""
    X = small["X"]
    Mo = small["Mo"]
    Mfull = bn.Check_input(X, ctx)
    beta_full = bn.api._default_beta(Mfull)
    full = bn.Prior_fun(Mfull, beta_full, verbose=False)
    world = 2

    def work(c, M, rank, g0):
        beta = bn.api._default_beta(M)
        return beta, bn.Prior_fun(M, beta, verbose=False)

    results = _run_gene_sharded(bn, Mo, X, world, work)
    for r in range(world):
        assert np.array_equal(results[r][0], beta_full)
    cat = lambda f: np.concatenate([f(results[r][1]) for r in range(world)])
    assert np.array_equal(cat(lambda P: P["MME_prior"]["MME_MU"]), full["MME_prior"]["MME_MU"])
    assert np.array_equal(cat(lambda P: P["MME_prior"]["MME_SIZE"]), full["MME_prior"]["MME_SIZE"])
    assert np.array_equal(cat(lambda P: P["BB_prior"]["BB_SIZE"]), full["BB_prior"]["BB_SIZE"])
    assert rel_err(cat(lambda P: P["MME_SIZE_adjust"]), full["MME_SIZE_adjust"]).max() < 1e-12


def test_gene_sharded_failure_reaches_every_shard(bn, ctx, oracle, small):
    """One shard's gene block holds a non-integer value: Prior_fun must fail on BOTH shards (the status is agreed before the
    first data-dependent collective) instead of leaving the clean shard waiting in the all-reduce hook."""
    X = small["X"].copy()
    X[3, 10] = 2.5                                   # rank 0's block only
    Mo = oracle.CSC.from_dense(X)

    def work(c, M, rank, g0):
        try:
            bn.Prior_fun(M, small["beta"], verbose=False)
        except bn.BayNormError as e:
            return e.code, str(e)
        return 0, ""

    res = _run_gene_sharded(bn, Mo, X, 2, work)
    assert [r[0] for r in res] == [6, 6], res      # BN_ERR_NONCOUNT everywhere
    assert "integer counts" in res[0][1] and "another gene shard failed" in res[1][1]


def test_gene_sharded_betafun_and_controls(bn, ctx, oracle, dense50):
    """BetaFun and SyntheticControl on two gene shards: the column sums cross the shards through the
    all-reduce hook; selection, BETA, lambda and the control counts equal the unsharded run."""
    X = dense50["X"].copy()
    X[:, X.sum(axis=0) == 0] += 1.0
    Mo = oracle.CSC.from_dense(X)
    full = bn.BetaFun(X, 0.06, ctx=ctx)
    ctl = bn.SyntheticControl(X, dense50["beta"], seed=5, ctx=ctx)
    assert len(full["Selected_genes"]) > 0

    def work(c, M, rank, g0):
        bf = bn.BetaFun(M, 0.06, ctx=c)
        sc = bn.SyntheticControl(M, dense50["beta"], seed=5, ctx=c)
        return bf["BETA"], np.asarray(bf["Selected_genes"]) + g0, sc["lambda"], sc["N_c"]

    res = _run_gene_sharded(bn, Mo, X, 2, work)
    for r in range(2):
        np.testing.assert_allclose(res[r][0], full["BETA"], rtol=1e-13)
    assert np.array_equal(np.concatenate([res[r][1] for r in range(2)]), np.asarray(full["Selected_genes"]))
    np.testing.assert_allclose(np.concatenate([res[r][2] for r in range(2)]), ctl["lambda"], rtol=1e-13)
    # the draws are keyed by global gene / cell indices: identical whenever lambda is bit-identical
    same = np.concatenate([res[r][2] for r in range(2)]) == ctl["lambda"]
    assert np.array_equal(np.concatenate([res[r][3] for r in range(2)])[same], ctl["N_c"][same])


# ------------------------------------------------------------------------ edge cases
def test_large_counts_go_through_stirling(bn, ctx, oracle):
    """Counts above the tail table (64 for warp groups, 256 for CTA groups) use Stirling's series;
    the likelihood fit must still agree with the oracle.  C = 9000 selects the multi-slot CTA kernel."""
    rng = np.random.default_rng(4)
    for C in (300, 9000):
        G = 24
        beta = np.clip(rng.lognormal(0, 0.4, C) * 0.06, 0.005, 0.5)
        mu = np.array([2000.0, 9000.0, 300.0, 4000.0] * 6)[:G] * rng.uniform(0.8, 1.2, G)
        size = rng.lognormal(0.5, 0.5, G)
        lam = rng.gamma(size[:, None], (mu[:, None] * beta[None, :]) / size[:, None])
        X = rng.poisson(lam).astype(np.float64)
        assert X.max() > 300
        Mo = oracle.CSC.from_dense(X)
        Po = oracle.Prior_fun(Mo, beta, BB_SIZE=False)
        mu0, s0 = Po["MME_prior"]["MME_MU"], Po["MME_prior"]["MME_SIZE"]
        lo, hi = float(s0.min()), float(np.ceil(s0.max()))
        want = oracle.BB_Fun(Mo, beta, mu0, s0, lo, hi, nthreads=8)
        got = bn.BB_Fun(bn.Check_input(X, ctx), beta, mu0, s0, SIZE_lower=lo, SIZE_upper=hi)
        assert np.median(rel_err(got, want)) < 1e-6 and rel_err(got, want).max() < 2e-4
        g = 1
        f_gpu = bn.MarginalF_NB_1D(got[g], mu0[g], X[g], beta, ctx=ctx)
        f_cpu = oracle.MarginalF_NB_1D(got[g], mu0[g], X[g], beta)
        assert abs(f_gpu - f_cpu) <= 1e-11 * abs(f_cpu)


def test_ragged_shapes_and_empty_columns(bn, ctx, oracle):
    """One gene, one cell, tile boundaries +-1, all-zero matrix columns and rows."""
    rng = np.random.default_rng(6)
    for G, C in ((1, 5), (7, 1), (2049, 3), (1023, 65), (1025, 130)):
        X = rng.poisson(0.3, (G, C)).astype(np.float64)
        X[:, C // 2] = 0
        X[G // 2, :] = 0
        beta = rng.uniform(0.02, 0.4, C)
        size = rng.uniform(0.2, 4, G)
        mu = rng.uniform(0.0, 6, G)
        mu[0] = 0.0
        Mo = oracle.CSC.from_dense(X)
        M = bn.Check_input(X, ctx)
        assert np.array_equal(bn.Main_mode_NB_Bay(M, beta, size, mu), oracle.Main_mode_NB_Bay(Mo, beta, size, mu))
        assert np.array_equal(bn.Main_mean_NB_Bay(M, beta, size, mu), oracle.Main_mean_NB_Bay(Mo, beta, size, mu))
        d = bn.Main_NB_Bay(M, beta, size, mu, S=3, seed=1)
        assert d.shape == (G, C, 3) and (d >= X[:, :, None]).all()
        assert (d[0] == X[0][:, None]).all()                       # mu = 0: the draw is x itself
    Z = np.zeros((5, 4))
    Mz = bn.Check_input(Z, ctx)
    assert Mz.nnz == 0
    assert np.array_equal(bn.Main_mean_NB_Bay(Mz, np.full(4, 0.1), np.ones(5), np.ones(5)),
                          oracle.Main_mean_NB_Bay(oracle.CSC.from_dense(Z), np.full(4, 0.1), np.ones(5), np.ones(5)))
    with pytest.raises(bn.BayNormError):
        bn.api._default_beta(Mz)                                    # all column sums zero: no beta in (0,1)
    with pytest.raises(bn.BayNormError):
        bn.Prior_fun(bn.Check_input(np.ones((3, 1)), ctx), np.array([0.1]), verbose=False)   # MME needs >= 2 cells


def test_explicitly_stored_zeros(bn, ctx, oracle, small):
    """A dgCMatrix may carry explicitly stored zeros (e.g. after `x[x == 1] <- 0`).  They are zeros for every
    stage: BETA, MME, the fit (no lgamma term, no tail-table entry), the posterior and BetaFun's zero fraction."""
    import scipy.sparse as sp
    X = small["X"][:90, :260].copy()
    beta = small["beta"][:260]
    keep = X.copy()
    keep[keep == 1] = 0.0                                      # the values the explicit zeros stand for
    S = sp.csc_matrix(X)
    S.data[S.data == 1] = 0.0                                  # same pattern, zeros stored
    assert (S.data == 0).sum() > 100
    Mz = bn.Check_input(S, ctx)
    Mk = bn.Check_input(keep, ctx)
    assert Mz.nnz > Mk.nnz
    Pz = bn.Prior_fun(Mz, beta, verbose=False)
    Pk = bn.Prior_fun(Mk, beta, verbose=False)
    for a, b in ((Pz["MME_prior"]["MME_MU"], Pk["MME_prior"]["MME_MU"]), (Pz["MME_prior"]["MME_SIZE"], Pk["MME_prior"]["MME_SIZE"])):
        np.testing.assert_allclose(a, b, rtol=1e-12)
    assert (rel_err(Pz["BB_prior"]["BB_SIZE"], Pk["BB_prior"]["BB_SIZE"]) <= 1e-6).mean() >= 0.9
    np.testing.assert_allclose(Pz["MME_SIZE_adjust"], Pk["MME_SIZE_adjust"], rtol=1e-5)
    size, mu = Pk["MME_SIZE_adjust"], Pk["MME_prior"]["MME_MU"]
    assert np.array_equal(bn.Main_mode_NB_Bay(Mz, beta, size, mu), bn.Main_mode_NB_Bay(Mk, beta, size, mu))
    assert np.array_equal(bn.Main_mean_NB_Bay(Mz, beta, size, mu), bn.Main_mean_NB_Bay(Mk, beta, size, mu))
    assert np.array_equal(bn.Main_NB_Bay(Mz, beta, size, mu, S=4, seed=3), bn.Main_NB_Bay(Mk, beta, size, mu, S=4, seed=3))
    bz, bk = bn.BetaFun(Mz, 0.06, ctx=ctx, with_stats=True), bn.BetaFun(Mk, 0.06, ctx=ctx, with_stats=True)
    assert np.array_equal(bz["stats"]["dropout"], bk["stats"]["dropout"])
    f2z = bn.BB_Fun(Mz, beta, mu, Pk["MME_prior"]["MME_SIZE"], FIX_MU=False, SIZE_lower=0.01, SIZE_upper=10, MU_lower=0.0,
                    MU_upper=float(mu.max()), ctx=ctx)
    f2k = bn.BB_Fun(Mk, beta, mu, Pk["MME_prior"]["MME_SIZE"], FIX_MU=False, SIZE_lower=0.01, SIZE_upper=10, MU_lower=0.0,
                    MU_upper=float(mu.max()), ctx=ctx)
    lz = np.array([oracle.MarginalF_NB_2D(f2z[g], keep[g], beta) for g in range(90)])
    lk = np.array([oracle.MarginalF_NB_2D(f2k[g], keep[g], beta) for g in range(90)])
    assert (np.abs(lz - lk) <= 1e-4 * np.abs(lk) + 1e-6).all()


def test_posterior_on_non_integer_data(bn, ctx, oracle, small):
    """The posterior closed forms accept any non-negative doubles (e.g. UMI_sffl-scaled data)."""
    X = small["X"][:50, :40] * 0.37
    beta, size, mu = small["beta"][:40], small["size"][:50], small["mu"][:50]
    Mo = oracle.CSC.from_dense(X)
    M = bn.Check_input(X, ctx)
    assert not M.is_count
    assert np.array_equal(bn.Main_mean_NB_Bay(M, beta, size, mu), oracle.Main_mean_NB_Bay(Mo, beta, size, mu))
    assert np.array_equal(bn.Main_mode_NB_Bay(M, beta, size, mu), oracle.Main_mode_NB_Bay(Mo, beta, size, mu))
    pri = bn.EstPrior(M, verbose=False)
    w = oracle.EstPrior(Mo, None)
    assert rel_err(pri["MU"], w["MU"]).max() < 1e-13


def test_mode_exact_fallback_near_integers(bn, ctx, oracle):
    """Parameters chosen so that (mu - mu*beta)(size+x-1)/(size+mu*beta) is an exact or near integer:
    the fast reciprocal path must hand over to the exact IEEE sequence."""
    G, C = 64, 33
    size = np.full(G, 3.0)
    mu = np.arange(1, G + 1, dtype=np.float64)
    beta = np.full(C, 0.5)                       # v = (mu/2)(2 + x)/(3 + mu/2): many exact rationals
    X = np.tile(np.arange(C, dtype=np.float64) % 5, (G, 1))
    size[::3] = 2.0
    beta[::4] = 0.25
    Mo = oracle.CSC.from_dense(X)
    got = bn.Main_mode_NB_Bay(bn.Check_input(X, ctx), beta, size, mu)
    assert np.array_equal(got, oracle.Main_mode_NB_Bay(Mo, beta, size, mu))
    assert np.array_equal(got, oracle.np_post_mode(X, beta, size, mu))


def test_posterior_samples_every_path(bn, ctx, oracle):
    """The three draw paths -- 32-point table (tiny means), warp-built 1024-point table (means in the tens)
    and the gamma-Poisson fallback (means in the thousands) -- against the exact NB moments, on a matrix whose
    non-zero fraction under-predicts the number of big-mean elements (forces the list-overflow repeat)."""
    rng = np.random.default_rng(11)
    G, C, S = 300, 120, 600
    X = (rng.random((G, C)) < 0.03) * rng.integers(1, 4, (G, C)).astype(np.float64)
    beta = rng.uniform(0.03, 0.2, C)
    size = rng.uniform(0.5, 6.0, G)
    mu = np.where(np.arange(G) % 3 == 0, 0.5, np.where(np.arange(G) % 3 == 1, 60.0, 4000.0)) * rng.uniform(0.5, 2.0, G)
    Mo = oracle.CSC.from_dense(X)
    M = bn.Check_input(X, ctx)
    d = bn.Main_NB_Bay(M, beta, size, mu, S=S, seed=2024)
    assert np.array_equal(d, np.floor(d)) and (d >= X[:, :, None]).all()
    nb_size, nb_mu = oracle.post_sample_params(Mo, beta, size, mu)
    mean_th = nb_mu + X
    var_th = nb_mu + nb_mu ** 2 / nb_size
    z = (d.mean(axis=2) - mean_th) / np.sqrt(var_th / S)
    for cls in range(3):
        zc = z[np.arange(G) % 3 == cls]
        assert np.abs(zc).max() < 5.5 and abs(zc.mean()) < 4.5 / np.sqrt(zc.size), (cls, np.abs(zc).max(), zc.mean())
    kap4 = var_th * (1 + 6 * var_th / nb_size)
    vz = (d.var(axis=2, ddof=1) - var_th) / np.sqrt((kap4 + 2 * var_th ** 2) / S)
    assert np.abs(vz).max() < 8
    rng2 = np.random.default_rng(3)
    from scipy.stats import chi2
    for cls in range(3):
        sel = np.arange(G) % 3 == cls
        stat = _pooled_pit_chi2(d[sel], nb_size[sel], nb_mu[sel], X[sel], rng2)
        assert stat < chi2.ppf(1 - 1e-6, 49), (cls, stat)
    # purity across S and column blocks holds on every path
    assert np.array_equal(bn.Main_NB_Bay(M, beta, size, mu, S=7, seed=2024), d[:, :, :7])
    assert np.array_equal(bn.Main_NB_Bay(M, beta, size, mu, S=S, seed=2024, col0=17, ncol=40), d[:, 17:57, :])


# ------------------------------------------------------------------------ SyntheticControl
def test_synthetic_control(bn, ctx, oracle, small):
    """R/synthetic_controls.r:57-70: lambda against the numpy restatement (closed form), the counts against
    their exact law -- rbinom(rpois(lambda_i), beta_j) is Poisson(lambda_i beta_j)."""
    from scipy.stats import chi2, poisson
    X, beta = small["X"][:150, :200].copy(), small["beta"][:200].copy()
    X[:, X.sum(axis=0) == 0] += 1.0                                  # every cell needs a count
    out = bn.SyntheticControl(X, beta, seed=77, ctx=ctx)
    ref = oracle.SyntheticControl(X, beta, seed=1)
    np.testing.assert_allclose(out["lambda"], ref["lambda"], rtol=1e-12)
    assert np.array_equal(out["beta_c"], beta)
    N = out["N_c"]
    assert N.shape == X.shape and np.array_equal(N, np.floor(N)) and (N >= 0).all()
    assert np.array_equal(bn.SyntheticControl(X, beta, seed=77, ctx=ctx)["N_c"], N)          # pure function of the seed
    assert not np.array_equal(bn.SyntheticControl(X, beta, seed=78, ctx=ctx)["N_c"], N)
    lam = out["lambda"][:, None] * beta[None, :]
    z = (N.sum(axis=1) - lam.sum(axis=1)) / np.sqrt(lam.sum(axis=1) + 1e-300)               # per gene
    assert np.abs(z).max() < 5 and abs(z.mean()) < 4 / np.sqrt(z.size)
    zc = (N.sum(axis=0) - lam.sum(axis=0)) / np.sqrt(lam.sum(axis=0))                         # per cell
    assert np.abs(zc).max() < 5
    rng = np.random.default_rng(2)                                                           # pooled randomised PIT
    u = poisson.cdf(N - 1, lam) + (poisson.cdf(N, lam) - poisson.cdf(N - 1, lam)) * rng.random(N.shape)
    h, _ = np.histogram(u.ravel(), bins=50, range=(0, 1))
    stat = float(((h - u.size / 50) ** 2 / (u.size / 50)).sum())
    assert stat < chi2.ppf(1 - 1e-6, 49), stat
    # the oracle's own draws follow the same law
    No = ref["N_c"]
    zo = (No.sum(axis=1) - (ref["lambda"][:, None] * beta[None, :]).sum(axis=1)) / np.sqrt(lam.sum(axis=1) + 1e-300)
    assert np.abs(zo).max() < 5
    Xz = X.copy()
    Xz[:, 3] = 0.0
    with pytest.raises(bn.BayNormError):
        bn.SyntheticControl(Xz, beta, ctx=ctx)


# ------------------------------------------------------------------------ BetaFun
@pytest.mark.parametrize("which", ["fixture", "dense_even", "dense_odd", "clamp"])
def test_beta_fun(bn, ctx, oracle, fixture_data, which):
    """R/PRIOR_FUNCTIONS.R:30-71 against the dense numpy restatement: zero fractions, the order statistics of
    the genes that can be selected, the selected set and BETA."""
    mean_beta = 0.06
    if which == "fixture":
        X = fixture_data["inputdata"]                                # 20 x 74, the reference's own example call
    elif which == "clamp":                                           # BETA >= 1 and BETA <= 0 both occur before clamping
        rng = np.random.default_rng(8)
        depth = np.ones(30)
        depth[0], depth[1] = 40.0, 0.01
        X = rng.poisson(6.0 * depth[None, :] * rng.uniform(0.5, 2.0, (40, 1))).astype(np.float64)
        X[39, :] = 0.0
        X[39, 1] = 3.0
        X[:39, 1] = 0.0
        mean_beta = 0.5
    else:
        rng = np.random.default_rng(21)
        G, Cn = 260, (150 if which == "dense_even" else 151)
        mu = rng.lognormal(1.5, 1.2, G)
        X = rng.poisson(rng.gamma(2.0, mu[:, None] / 2.0, (G, Cn)) * rng.uniform(0.5, 1.5, Cn)[None, :]).astype(np.float64)
        X[:, X.sum(axis=0) == 0] += 1.0
    got = bn.BetaFun(X, mean_beta, ctx=ctx, with_stats=True)
    ref = oracle.BetaFun(X, mean_beta)
    np.testing.assert_allclose(got["stats"]["dropout"], ref["stats"]["dropout"], rtol=0, atol=0)
    cand = ref["stats"]["dropout"] < 0.35
    assert cand.sum() > 3, "test matrix has no candidate genes"
    for k in ("median", "mad", "max"):
        np.testing.assert_allclose(got["stats"][k][cand], ref["stats"][k][cand], rtol=1e-13, atol=1e-15, err_msg=k)
        assert np.isnan(got["stats"][k][~cand]).all()
    assert np.array_equal(np.asarray(got["Selected_genes"]), ref["Selected_genes"])
    assert 0 < len(ref["Selected_genes"]) < X.shape[0]
    np.testing.assert_allclose(got["BETA"], ref["BETA"], rtol=1e-12)
    assert (got["BETA"] > 0).all() and (got["BETA"] < 1).all()
    if which != "clamp":
        assert abs(got["BETA"].mean() - mean_beta) < 1e-12


# ------------------------------------------------------------------------ FIX_MU = FALSE
@pytest.mark.parametrize("which", ["fixture", "small"])
def test_fit_2d_against_oracle(bn, ctx, oracle, small, fixture_data, which):
    """BB_Fun(FIX_MU=FALSE): spg over (size, mu) in the reference's box, against the restated spg running the
    same algorithm on the CPU.  Like the 1-D fit, the stopping point of a forward-difference spg moves with the
    last bits of the likelihood, so the assertions are: most genes within 1e-6, all close, and the two stopping
    points equally good (same likelihood to 1e-9 relative)."""
    if which == "fixture":
        X, beta = fixture_data["inputdata"], fixture_data["inputbeta"]
    else:
        X, beta = small["X"][:80, :300], small["beta"][:300]
    Mo = oracle.CSC.from_dense(X)
    pri = oracle.EstPrior(Mo, beta)
    size0 = pri["SIZE"].copy()
    size0[np.isnan(size0)] = np.nanmin(size0)
    mu0 = pri["MU"]
    box = dict(SIZE_lower=float(size0.min()), SIZE_upper=float(np.ceil(size0.max())), MU_lower=float(mu0.min()),
               MU_upper=float(mu0.max()))
    got, conv, nfe = bn.BB_Fun(X, beta, mu0, size0, FIX_MU=False, parallel=True, ctx=ctx, return_info=True, **box)
    want, wconv, wcnt = oracle.BB_Fun_2d(X, beta, mu0, size0, **box)
    assert got.shape == (X.shape[0], 2)
    assert (got[:, 0] >= box["SIZE_lower"]).all() and (got[:, 0] <= box["SIZE_upper"]).all()
    assert (got[:, 1] >= box["MU_lower"]).all() and (got[:, 1] <= box["MU_upper"]).all()
    rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
    live = (X > 0).sum(axis=1) > 0
    assert np.array_equal(got[~live], want[~live])                       # all-zero genes: the projected start
    assert (rel[live].max(axis=1) <= 1e-6).mean() >= 0.8, rel[live].max(axis=1)
    # along the flat size-mu ridge of a sparse gene two equally good stopping points can sit percents apart:
    # what has to agree everywhere is the likelihood they reach
    lg = np.array([oracle.MarginalF_NB_2D(got[g], X[g], beta) for g in range(X.shape[0])])
    lw = np.array([oracle.MarginalF_NB_2D(want[g], X[g], beta) for g in range(X.shape[0])])
    done = (conv == 0) & (wconv == 0)                # genes whose spg converged (the others ran into maxfeval = 500)
    assert done.mean() > 0.8
    assert (np.abs(lg - lw)[done] <= 1e-8 * np.abs(lw[done]) + 1e-8).all(), np.abs(lg - lw)[done].max()
    # a run cut off at 500 evaluations on a near-flat ridge (genes with one or two counts) stops wherever it is:
    # the two walks still end at the same height to 1e-4
    assert (np.abs(lg - lw) <= 1e-4 * np.abs(lw) + 1e-8).all(), np.abs(lg - lw).max()
    assert rel.max() < 0.5, rel.max()
    assert (conv == wconv).mean() >= 0.9
    assert abs(nfe.sum() / (wcnt[:, 0] + wcnt[:, 1]).sum() - 1) < 0.05     # same amount of work
    # the fitted point is at least as good as the start (and as the 1-D fit with mu fixed)
    l0 = np.array([oracle.MarginalF_NB_2D([min(max(size0[g], box["SIZE_lower"]), box["SIZE_upper"]), mu0[g]], X[g], beta)
                   for g in range(X.shape[0])])
    assert (lg >= l0 - 1e-9 * np.abs(l0)).all()


def test_fit_2d_analytic_gradient_and_prior_fun(bn, ctx, oracle, small):
    """GR=TRUE (GradientFun_NB_2D) reaches the same optimum as the forward-difference run; Prior_fun(FIX_MU=FALSE)
    returns the reference's structure: MME_prior, BB_prior (size, mu), MME_SIZE_adjust from BB_prior[, 1]."""
    X, beta = small["X"][:60, :250], small["beta"][:250]
    P = bn.Prior_fun(X, beta, FIX_MU=False, verbose=False, ctx=ctx, return_info=True)
    Pg = bn.Prior_fun(X, beta, FIX_MU=False, GR=True, verbose=False, ctx=ctx)
    assert set(P) >= {"MME_prior", "BB_prior", "MME_SIZE_adjust", "BETA_vec"} and P["BB_prior"].colnames == ["size", "mu"] and P["BB_prior"].shape[1] == 2
    Po = oracle.Prior_fun(oracle.CSC.from_dense(X), beta, BB_SIZE=False)
    np.testing.assert_allclose(P["MME_prior"]["MME_MU"], Po["MME_prior"]["MME_MU"], rtol=1e-13)
    adj = oracle.AdjustSIZE_fun(P["BB_prior"]["size"], P["MME_prior"]["MME_MU"], P["MME_prior"]["MME_SIZE"])
    np.testing.assert_allclose(P["MME_SIZE_adjust"], adj, rtol=1e-10)
    info = P["_info"]
    assert info["mu_lower"] == P["MME_prior"]["MME_MU"].min() and info["mu_upper"] == P["MME_prior"]["MME_MU"].max()
    lf = np.array([oracle.MarginalF_NB_2D([P["BB_prior"]["size"][g], P["BB_prior"]["mu"][g]], X[g], beta) for g in range(60)])
    lgr = np.array([oracle.MarginalF_NB_2D([Pg["BB_prior"]["size"][g], Pg["BB_prior"]["mu"][g]], X[g], beta) for g in range(60)])
    conv_ok = (P["_info"]["conv"] == 0)
    assert (np.abs(lf - lgr)[conv_ok] <= 1e-6 * np.abs(lf[conv_ok]) + 1e-6).all(), np.abs(lf - lgr)[conv_ok].max()
    assert (np.abs(lf - lgr) <= 1e-4 * np.abs(lf) + 1e-3).all(), np.abs(lf - lgr).max()      # runs cut off at maxfeval
    # device gradient against the oracle's GradientFun_NB_2D on one gene
    g = int(np.argmax((X > 0).sum(axis=1)))
    par = np.array([P["BB_prior"]["size"][g] * 1.3, P["BB_prior"]["mu"][g] * 0.8])
    np.testing.assert_allclose(bn.GradientFun_NB_2D(par, X[g], beta, ctx=ctx), oracle.GradientFun_NB_2D(par, X[g], beta), rtol=1e-9)


# ------------------------------------------------------------------------ randomized shapes
def test_random_shapes_and_column_blocks(bn, ctx, oracle):
    """Posterior mean / mode (bit-exact), draws (support, purity across column blocks) and DownSampling (support) on
    random shapes around the kernels' block sizes (128-gene row blocks, 1024-gene CTAs, 32 x 32 tiles, 64-column
    chunks), densities from 1 % to 70 %, with empty columns/rows and arbitrary column sub-blocks."""
    rng = np.random.default_rng(2026)
    shapes = [(127, 33), (128, 64), (129, 65), (255, 31), (1024, 40), (1151, 70), (1153, 129), (2050, 17), (300, 200), (96, 257)]
    for n, (G, C) in enumerate(shapes):
        dens = [0.01, 0.08, 0.3, 0.7][n % 4]
        X = (rng.random((G, C)) < dens) * rng.integers(1, 30, (G, C)).astype(np.float64)
        X[:, rng.integers(0, C)] = 0
        X[rng.integers(0, G), :] = 0
        beta = rng.uniform(0.02, 0.6, C)
        size = rng.uniform(0.1, 5, G)
        mu = rng.uniform(0.0, 20, G)
        Mo = oracle.CSC.from_dense(X)
        M = bn.Check_input(X, ctx)
        wmean, wmode = oracle.Main_mean_NB_Bay(Mo, beta, size, mu), oracle.Main_mode_NB_Bay(Mo, beta, size, mu)
        assert np.array_equal(bn.Main_mean_NB_Bay(M, beta, size, mu), wmean), (G, C)
        assert np.array_equal(bn.Main_mode_NB_Bay(M, beta, size, mu), wmode), (G, C)
        c0 = int(rng.integers(0, C - 1))
        nc = int(rng.integers(1, C - c0 + 1))
        assert np.array_equal(bn.Main_mode_NB_Bay(M, beta, size, mu, col0=c0, ncol=nc), wmode[:, c0:c0 + nc]), (G, C, c0, nc)
        assert np.array_equal(bn.Main_mean_NB_Bay(M, beta, size, mu, col0=c0, ncol=nc), wmean[:, c0:c0 + nc]), (G, C, c0, nc)
        d = bn.Main_NB_Bay(M, beta, size, mu, S=5, seed=9)
        assert d.shape == (G, C, 5) and np.array_equal(d, np.floor(d)) and (d >= X[:, :, None]).all()
        assert (d[mu == 0] == X[mu == 0][:, :, None]).all()
        assert np.array_equal(bn.Main_NB_Bay(M, beta, size, mu, S=5, seed=9, col0=c0, ncol=nc), d[:, c0:c0 + nc, :]), (G, C, c0, nc)
        ds = bn.DownSampling(X, beta, seed=4, ctx=ctx)
        assert ds.shape == X.shape and np.array_equal(ds, np.floor(ds)) and (ds >= 0).all() and (ds <= X).all()
        assert (ds[X == 0] == 0).all()
        if n % 3 == 0:
            Xc = X.copy()
            Xc[0, Xc.sum(axis=0) == 0] = 1.0
            sc = bn.SyntheticControl(Xc, beta, seed=3, ctx=ctx)
            np.testing.assert_allclose(sc["lambda"], oracle.SyntheticControl(Xc, beta)["lambda"], rtol=1e-12)
            assert sc["N_c"].shape == X.shape and (sc["N_c"] >= 0).all() and (sc["N_c"][sc["lambda"] == 0] == 0).all()


# ------------------------------------------------------------------------ boundary details
def test_check_input_leaves_the_callers_matrix_alone(bn, ctx, small):
    """Check_input on a scipy CSC matrix with unsorted indices and duplicate entries: the caller's object is untouched and the
    device matrix holds the canonical form (duplicates summed, rows increasing), like as(Data, "dgCMatrix") upstream."""
    import scipy.sparse as sp
    rows = np.array([3, 1, 1, 0, 2, 2], np.int32)
    cols_p = np.array([0, 3, 3, 6], np.int32)
    vals = np.array([1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    A = sp.csc_matrix((vals.copy(), rows.copy(), cols_p.copy()), shape=(4, 3))
    M = bn.Check_input(A, ctx)
    assert np.array_equal(A.indices, rows) and np.array_equal(A.data, vals) and not A.has_canonical_format
    p, i, x = M.export_csc()
    want = sp.csc_matrix(A.toarray())
    assert np.array_equal(p, want.indptr) and np.array_equal(i, want.indices) and np.array_equal(x, want.data)


def test_fit_2d_serial_branch_box(bn, ctx, oracle, small):
    """BB_Fun(FIX_MU=FALSE, parallel=FALSE) hands spg the scalar size bounds for BOTH parameters (R/PRIOR_FUNCTIONS.R:503-526);
    parallel=TRUE boxes mu by (MU_lower, MU_upper) (:414-419)."""
    X, beta = small["X"][:40], small["beta"]
    mu0 = np.maximum(X.sum(1) / beta.sum(), 0.05)
    s0 = np.full(40, 1.2)
    kw = dict(MU_lower=0.01, MU_upper=500.0, SIZE_lower=0.05, SIZE_upper=4.0, FIX_MU=False)
    ser = bn.BB_Fun(bn.Check_input(X, ctx), beta, mu0, s0, parallel=False, **kw)
    par = bn.BB_Fun(bn.Check_input(X, ctx), beta, mu0, s0, parallel=True, **kw)
    assert ser.shape == (40, 2) and (ser[:, 1] <= 4.0).all() and (ser[:, 1] >= 0.05).all()
    assert (par[:, 1] > 4.0).any()                   # expressed genes: the mean leaves the size box when it may
    want, conv, _ = oracle.BB_Fun_2d(X, beta, mu0, s0, 0.05, 4.0, 0.05, 4.0)
    ok = conv == 0
    assert (rel_err(ser[ok], want[ok]) <= 1e-5).mean() >= 0.8
    assert np.array_equal(ser[:, 1] == 4.0, want[:, 1] == 4.0)


def test_sparse_mean_result_and_tile_stream(bn, ctx, oracle, small):
    """bn_post_mean_csc (Main_mean_NB_spBay, src/bayNorm_main.cpp:1598: every non-zero of the dense mean, device-resident) and
    bn_post_stream (column tiles through pinned host memory) against the dense result."""
    import ctypes as C
    from baynorm_b200._lib import TILE_SINK
    from baynorm_b200.api import _ptr
    X, beta, size, mu = small["X"], small["beta"], small["size"].copy(), small["mu"].copy()
    mu[::9] = 0.0                                    # mean 0: the posterior mean is x, so all-zero rows give true zeros
    M = bn.Check_input(X, ctx)
    dense = bn.Main_mean_NB_Bay(M, beta, size, mu)
    sp_ = bn.Main_mean_NB_spBay(M, beta, size, mu)
    assert sp_.format == "csc" and sp_.has_sorted_indices and sp_.nnz == np.count_nonzero(dense) < dense.size
    assert np.array_equal(sp_.toarray(), dense)
    # streamed tiles: mean, mode and draws reassembled from 37-column tiles equal the one-shot results
    G, Cn, S = M.G, M.C, 6
    for kind, ref in ((0, dense), (1, bn.Main_mode_NB_Bay(M, beta, size, mu)), (2, bn.Main_NB_Bay(M, beta, size, mu, S=S, seed=99))):
        got = np.full(ref.shape, np.nan, order="F")
        seen = []

        def sink(user, col0, ncol, tile, got=got, kind=kind):
            planes = S if kind == 2 else 1
            t = np.ctypeslib.as_array(C.cast(tile, C.POINTER(C.c_double)), shape=(G * ncol * planes,))
            if kind == 2:
                got[:, col0:col0 + ncol, :] = t.reshape((G, ncol, S), order="F")
            else:
                got[:, col0:col0 + ncol] = t.reshape((G, ncol), order="F")
            seen.append((col0, ncol))
            return 0

        cb = TILE_SINK(sink)
        b, sz, m = (np.ascontiguousarray(v, np.float64) for v in (beta, size, mu))
        ctx.check(ctx.lib.bn_post_stream(ctx.h, M.h, kind, _ptr(b), _ptr(sz), _ptr(m), S, 99, 37, cb, None))
        assert [c for c, _ in seen] == list(range(0, Cn, 37)) and sum(n for _, n in seen) == Cn
        assert np.array_equal(got, ref)
    # a sink that gives up stops the iteration with an error
    stop = TILE_SINK(lambda user, col0, ncol, tile: 1)
    with pytest.raises(bn.BayNormError, match="sink"):
        ctx.check(ctx.lib.bn_post_stream(ctx.h, M.h, 0, _ptr(b), _ptr(sz), _ptr(m), 1, 0, 37, stop, None))


def test_cuda_against_reference_golden(bn, ctx):
    """The CUDA path against REAL R / bayNorm / BB output on the fixture (tests/golden/make_reference_golden.R).  Skips loudly
    while the file does not exist: R is not available in the build image."""
    from test_oracle import REFERENCE_GOLDEN, load_reference_golden
    if not REFERENCE_GOLDEN.exists():
        pytest.skip("PARITY UNPINNED: %s is missing -- produce it with Rscript tests/golden/make_reference_golden.R" % REFERENCE_GOLDEN)
    R = load_reference_golden()
    X, beta, size, mu = R["inputdata"], R["inputbeta"], R["size"], R["mu"]
    M = bn.Check_input(X, ctx)
    assert np.array_equal(bn.Main_mode_NB_Bay(M, beta, size, mu), R["Main_mode_NB_Bay"])             # bit-exact
    assert rel_err(bn.Main_mean_NB_Bay(M, beta, size, mu), R["Main_mean_NB_Bay"]).max() <= 1e-12
    P = bn.Prior_fun(M, beta, verbose=False)
    assert rel_err(P["MME_prior"]["MME_MU"], R["Prior_MME_MU"]).max() <= 1e-6
    assert rel_err(P["MME_prior"]["MME_SIZE"], R["Prior_MME_SIZE"]).max() <= 1e-6
    re = rel_err(P["BB_prior"]["BB_SIZE"], R["Prior_BB_SIZE"])
    assert (re <= 1e-6).mean() >= 0.8 and re.max() <= 2e-4
    assert rel_err(P["MME_SIZE_adjust"], R["Prior_MME_SIZE_adjust"]).max() <= 1e-4
    out = bn.bayNorm(X, BETA_vec=beta, mean_version=True, verbose=False, ctx=ctx)
    assert rel_err(out["Bay_out"], R["bayNorm_mean_Bay_out"]).max() <= 1e-4
    d = bn.Main_NB_Bay(M, beta, size, mu, S=2000, seed=1)
    z = (d.mean(2) - R["Main_NB_Bay_mean_over_2000"]) / np.sqrt((d.var(2, ddof=1) + R["Main_NB_Bay_var_over_2000"]) / 2000 + 1e-12)
    assert np.abs(z).max() < 6.0                                                                         # two-sample z per element


# ---------------------------------------------------------------------------------------------------------------------
# one process, several devices (bn_group): the library's own threads + NCCL instead of one process per GPU
# ---------------------------------------------------------------------------------------------------------------------
def _device_count():
    import ctypes as CT
    try:
        rt = CT.CDLL("libcudart.so.12")
    except OSError:
        rt = CT.CDLL("libcudart.so")
    n = CT.c_int(0)
    return n.value if rt.cudaGetDeviceCount(CT.byref(n)) == 0 else 0


@pytest.mark.parametrize("ndev", [1, 2])
def test_device_group_matches_one_device(bn, ctx, small, ndev):
    """bn_group_*: gene blocks on ndev devices driven by one host thread each, collectives over NCCL loaded inside the
    library (R/PRIOR_FUNCTIONS.R:426-427 is the reference's parallel branch).  BETA, MME and BB_SIZE equal the one-device
    run bit for bit, the adjusted size to a few ulps (its five OLS sums are added in another order); posterior mean,
    mode and draws are bit-identical (draws are pure functions of seed, gene, cell and plane)."""
    if _device_count() < ndev:
        pytest.skip("needs %d GPUs" % ndev)
    from baynorm_b200.api import DeviceGroup
    X = small["X"]
    M1 = bn.Check_input(X, ctx)
    beta1 = bn.api._default_beta(M1)
    full = bn.Prior_fun(M1, beta1, verbose=False, return_info=True)
    grp = DeviceGroup(ndev)
    try:
        gm = grp.matrix(X)
        b = gm.bounds()
        assert b[0] == 0 and b[-1] == X.shape[0] and np.all(np.diff(b) > 0)
        beta = grp.beta_default(gm)
        assert np.array_equal(beta, beta1)
        P = grp.Prior_fun(gm, beta)
        assert np.array_equal(P["MME_prior"]["MME_MU"], full["MME_prior"]["MME_MU"])
        assert np.array_equal(P["MME_prior"]["MME_SIZE"], full["MME_prior"]["MME_SIZE"])
        assert np.array_equal(P["BB_prior"]["BB_SIZE"], full["BB_prior"]["BB_SIZE"])
        assert np.array_equal(P["_info"]["conv"], full["_info"]["conv"])
        assert rel_err(P["MME_SIZE_adjust"], full["MME_SIZE_adjust"]).max() < 1e-12
        assert P["_info"]["size_lower"] == full["_info"]["size_lower"] and P["_info"]["size_upper"] == full["_info"]["size_upper"]
        size, mu = full["MME_SIZE_adjust"], full["MME_prior"]["MME_MU"]
        c0, nc = 7, 300
        mean = grp.posterior("mean", gm, beta, size, mu, col0=c0, ncol=nc)
        mode = grp.posterior("mode", gm, beta, size, mu, col0=c0, ncol=nc)
        draws = grp.posterior("sample", gm, beta, size, mu, S=5, seed=77, col0=c0, ncol=nc)
        assert np.array_equal(mean, bn.Main_mean_NB_Bay(M1, beta, size, mu, ctx=ctx, col0=c0, ncol=nc))
        assert np.array_equal(mode, bn.Main_mode_NB_Bay(M1, beta, size, mu, ctx=ctx, col0=c0, ncol=nc))
        assert np.array_equal(draws, bn.Main_NB_Bay(M1, beta, size, mu, S=5, seed=77, ctx=ctx, col0=c0, ncol=nc))
    finally:
        grp.close()
