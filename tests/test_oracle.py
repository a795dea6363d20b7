"""CPU tests (no GPU): pin the oracle before anything is compared with it.

The reference's own tests hold no numeric golden vectors (tests/testthat/test-bayNorm.r:9-29 only
checks run-to-run identity), and the reference cannot run in this image, so the oracle is pinned
against (a) independent implementations of the same published mathematics (scipy / mpmath /
numpy), (b) the decoded bundled fixture, (c) committed golden outputs of the oracle itself
(tests/golden/oracle_fixture_outputs.npz, made by tests/golden/make_oracle_golden.py) so that a
later edit of the oracle cannot silently move the target."""
import numpy as np
import pytest

from util import rel_err, synth_counts


def test_fixture_matches_survey_probe(fixture_data):
    A = fixture_data["inputdata"]
    assert A.shape == (20, 74) and A.max() == 20 and abs(A.mean() - 2.2264) < 1e-3
    assert abs((A == 0).mean() - 0.330) < 1e-3
    b = fixture_data["inputbeta"]
    assert b.shape == (74,) and abs(b.min() - 0.0765) < 1e-4 and abs(b.max() - 0.1975) < 1e-4
    assert fixture_data["inputdata_rownames"][0] == "0610009B22Rik"
    assert fixture_data["inputdata_colnames"][0] == "SC_2i_1"
    assert fixture_data["mu"].shape == (20,) and fixture_data["size"].shape == (20,)


def test_dnbinom_mu_against_scipy_and_mpmath(oracle):
    from scipy.stats import nbinom
    import mpmath as mp
    mp.mp.dps = 50
    rng = np.random.default_rng(0)
    x = np.concatenate([np.zeros(50), rng.integers(1, 8, 150), rng.integers(8, 400, 100)]).astype(float)
    size = np.exp(rng.uniform(np.log(0.02), np.log(300), x.size))
    mu = np.exp(rng.uniform(np.log(1e-4), np.log(500), x.size))
    got = oracle.dnbinom_mu_log(x, size, mu)
    direct = oracle.dnbinom_mu_log_direct(x, size, mu)
    sp = nbinom.logpmf(x, size, size / (size + mu))

    def exact(xv, s, m):
        xv, s, m = mp.mpf(xv), mp.mpf(s), mp.mpf(m)
        return float(mp.loggamma(xv + s) - mp.loggamma(s) - mp.loggamma(xv + 1) + s * mp.log(s / (s + m))
                     + xv * mp.log(m / (s + m)))

    ex = np.array([exact(a, b, c) for a, b, c in zip(x, size, mu)])
    assert rel_err(got, ex).max() < 2e-13          # restated saddle-point form
    assert rel_err(direct, ex).max() < 5e-13       # lgamma form (cancels more for large size)
    assert rel_err(sp, ex).max() < 1e-10           # scipy is the loosest of the three
    # x == 0 shortcut in both regimes
    assert rel_err(oracle.dnbinom_mu_log(0.0, 5.0, 1e-9), -1e-9).max() < 1e-8
    # non-integer x: density 0 (R warns and returns -Inf)
    assert oracle.dnbinom_mu_log(1.5, 2.0, 1.0) == -np.inf


def test_digamma(oracle):
    from scipy.special import digamma
    x = np.exp(np.linspace(np.log(0.01), np.log(500), 400))
    d = oracle.digamma(x)
    assert (np.abs(d - digamma(x)) <= 2e-14 * np.maximum(1, np.abs(digamma(x))) + 1e-13 / x).all()


def test_marginal_and_gradient(oracle, fixture_data):
    A, beta = fixture_data["inputdata"], fixture_data["inputbeta"]
    for g in (0, 4, 13):
        mu = A[g].sum() / beta.sum()
        for size in (0.3, 1.7, 9.0):
            f = oracle.MarginalF_NB_1D(size, mu, A[g], beta)
            assert abs(f - oracle.np_marginal_nb_1d(size, mu, A[g], beta)) < 1e-10 * abs(f)
            assert f == oracle.MarginalF_NB_2D([size, mu], A[g], beta)
            h = 1e-6
            num = (oracle.MarginalF_NB_1D(size + h, mu, A[g], beta) - oracle.MarginalF_NB_1D(size - h, mu, A[g], beta)) / (2 * h)
            assert abs(oracle.GradientFun_NB_1D(size, mu, A[g], beta) - num) < 1e-5 * (1 + abs(num))
            g2 = oracle.GradientFun_NB_2D([size, mu], A[g], beta)
            numm = (oracle.MarginalF_NB_1D(size, mu + h, A[g], beta) - oracle.MarginalF_NB_1D(size, mu - h, A[g], beta)) / (2 * h)
            assert abs(g2[0] - oracle.GradientFun_NB_1D(size, mu, A[g], beta)) < 1e-12 * (1 + abs(g2[0]))
            assert abs(g2[1] - numm) < 1e-5 * (1 + abs(numm))


def test_closed_forms_against_numpy(oracle, fixture_data):
    X, mu, size, beta = synth_counts(150, 90, seed=3)
    X[5] = 0
    M = oracle.CSC.from_dense(X)
    assert np.array_equal(M.to_dense(), X)
    b, bad = oracle.beta_default(M)
    assert not bad and rel_err(b, oracle.np_beta_default(X)).max() < 1e-14
    pri = oracle.EstPrior(M, beta)
    m, s, v = oracle.np_mme(X, beta)
    assert rel_err(pri["MU"], m).max() < 1e-13 and rel_err(pri["v"], v).max() < 1e-12
    assert np.array_equal(np.isnan(pri["SIZE"]), np.isnan(s)) and np.isnan(pri["SIZE"][5])
    mean = oracle.Main_mean_NB_Bay(M, beta, size, mu)
    mode = oracle.Main_mode_NB_Bay(M, beta, size, mu)
    assert np.array_equal(mean, oracle.np_post_mean(X, beta, size, mu))       # bit-for-bit
    assert np.array_equal(mode, oracle.np_post_mode(X, beta, size, mu))
    rng = np.random.default_rng(1)
    ms = rng.lognormal(0, 1, 400)
    bb = np.clip(np.exp(0.2 + 0.8 * np.log(ms) + rng.normal(0, 0.3, 400)), 0.1, 8)
    assert rel_err(oracle.AdjustSIZE_fun(bb, None, ms), oracle.np_adjust_size(bb, ms)).max() < 1e-12


def test_spg_lands_on_the_argmax(oracle, fixture_data):
    """SURVEY.md appendix B probe 4: restated spg vs bounded Brent on the fixture."""
    A, beta = fixture_data["inputdata"], fixture_data["inputbeta"]
    M = oracle.CSC.from_dense(A)
    P = oracle.Prior_fun(M, beta)
    s0 = P["MME_prior"]["MME_SIZE"]
    lo, hi = s0.min(), np.ceil(s0.max())
    assert abs(lo - 0.3005) < 1e-3 and hi == 5.0
    bb = P["BB_prior"]["BB_SIZE"]
    assert (bb == hi).sum() == 7
    rel = []
    for g in range(20):
        xb, fb = oracle.brent_gene_1d(P["MME_prior"]["MME_MU"][g], A[g], beta, lo, hi)
        fs = oracle.MarginalF_NB_1D(bb[g], P["MME_prior"]["MME_MU"][g], A[g], beta)
        assert fs >= fb - 1e-7 * abs(fb)
        rel.append(abs(bb[g] - xb) / xb)
    rel = np.array(rel)
    assert np.median(rel) < 1e-6 and rel.max() < 1e-4
    # serial and threaded fits are the same numbers (genes are independent)
    b2 = oracle.BB_Fun(M, beta, P["MME_prior"]["MME_MU"], s0, lo, hi, nthreads=4)
    assert np.array_equal(b2, bb)


def test_oracle_golden_outputs(oracle, fixture_data):
    from pathlib import Path
    g = np.load(Path(__file__).parent / "golden" / "oracle_fixture_outputs.npz")
    A, beta = fixture_data["inputdata"], fixture_data["inputbeta"]
    M = oracle.CSC.from_dense(A)
    r = oracle.bayNorm(M, BETA_vec=beta, mean_version=True)
    assert rel_err(r["PRIORS"]["MME_prior"]["MME_MU"], g["MME_MU"]).max() < 1e-14
    assert rel_err(r["PRIORS"]["MME_prior"]["MME_SIZE"], g["MME_SIZE"]).max() < 1e-12
    assert rel_err(r["PRIORS"]["BB_prior"]["BB_SIZE"], g["BB_SIZE"]).max() < 1e-9
    assert rel_err(r["PRIORS"]["MME_SIZE_adjust"], g["MME_SIZE_adjust"]).max() < 1e-9
    assert rel_err(r["Bay_out"], g["Bay_out_mean"]).max() < 1e-9
    # config 1 of BASELINE.json: Main_mean_NB_Bay(inputdata, inputbeta, size, mu) with the bundled priors
    assert np.array_equal(oracle.Main_mean_NB_Bay(M, beta, fixture_data["size"], fixture_data["mu"]), g["mean_bundled_priors"])
    assert np.array_equal(oracle.Main_mode_NB_Bay(M, beta, fixture_data["size"], fixture_data["mu"]), g["mode_bundled_priors"])


def test_oracle_samplers_have_the_right_law(oracle, fixture_data):
    from scipy.stats import chi2, nbinom
    A, beta, size, mu = (fixture_data[k] for k in ("inputdata", "inputbeta", "size", "mu"))
    M = oracle.CSC.from_dense(A)
    S = 400
    d = oracle.Main_NB_Bay(M, beta, size, mu, S=S, seed=5)
    nb_size, nb_mu = oracle.post_sample_params(M, beta, size, mu)
    z = (d.mean(axis=2) - (nb_mu + A)) / np.sqrt((nb_mu + nb_mu ** 2 / nb_size) / S)
    assert np.abs(z).max() < 5.5
    rng = np.random.default_rng(0)
    k = d - A[..., None]
    p = (nb_size / (nb_size + nb_mu))[..., None]
    lo, hi = nbinom.cdf(k - 1, nb_size[..., None], p), nbinom.cdf(k, nb_size[..., None], p)
    u = lo + (hi - lo) * rng.random(k.shape)
    h, _ = np.histogram(u.ravel(), bins=50, range=(0, 1))
    assert ((h - u.size / 50) ** 2 / (u.size / 50)).sum() < chi2.ppf(1 - 1e-6, 49)
    ds = np.stack([oracle.DownSampling(A, beta, seed=s) for s in range(300)], axis=2)
    zz = (ds.mean(axis=2) - A * beta) / np.sqrt(A * beta * (1 - beta) / 300 + 1e-300)
    assert np.abs(zz).max() < 5.5 and (ds <= A[..., None]).all()


def test_conditions_orchestration(oracle):
    X, mu, size, beta = synth_counts(40, 60, seed=9, m0=1.0)
    M = oracle.CSC.from_dense(X)
    cond = np.array([2] * 20 + [1] * 25 + [2] * 15)
    r = oracle.bayNorm(M, BETA_vec=beta, Conditions=cond, Prior_type="LL", mean_version=True)
    assert list(r["Bay_out_list"]) == ["Group 2", "Group 1"]
    ix = np.nonzero(cond == 1)[0]
    solo = oracle.bayNorm(M.select_cols(ix), BETA_vec=beta[ix], mean_version=True)
    assert np.array_equal(solo["Bay_out"], r["Bay_out_list"]["Group 1"])
    g = oracle.bayNorm(M, BETA_vec=beta, Conditions=cond, Prior_type="GG", mean_version=True)
    assert g["PRIORS_LIST"]["Group 1"] is g["PRIORS_LIST"]["Group 2"]
