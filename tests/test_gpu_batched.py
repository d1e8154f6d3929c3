"""The batched-beta path (bootstrap replicates, parallel MCMC chains): X.[b_1..b_m] on the
FP64 tensor-core path.  Every replicate must equal the single-beta fused pass and the
oracle; with resample counts it must equal the oracle on the physically resampled rows
(what apop_bootstrap_cov builds, apop_bootstrap.m4.c:143-149)."""
import numpy as np
import pytest

from conftest import gen_probit_logit_sample

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fam,n,k,m", [("probit", 20011, 20, 37), ("probit", 4000, 32, 130), ("logit", 9000, 24, 256),
                                       ("probit", 333, 5, 3), ("logit", 5000, 9, 17), ("poisson_reg", 5000, 16, 20)])
def test_batched_equals_single_and_oracle(b200, oracle, fam, n, k, m):
    ab = b200
    rng = np.random.default_rng(n + k + m)
    if fam == "poisson_reg":
        X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
        beta = rng.uniform(-1, 1, k) / np.sqrt(k)
        y = rng.poisson(np.exp(X @ beta)).astype(float)
        F, oll, osc = ab.POISSON_REG, oracle.poisson_reg_ll, oracle.poisson_reg_score
    else:
        X, y, beta = gen_probit_logit_sample(n, k, seed=n + k, method=fam)
        F = ab.PROBIT if fam == "probit" else ab.LOGIT
        oll = oracle.probit_ll if fam == "probit" else (lambda X, y, b, mode=1: oracle.logit_ll(X, y, b, mode=mode))
        osc = oracle.probit_score if fam == "probit" else (lambda X, y, b, mode=1: oracle.logit_score(X, y, b, mode=mode))
    d = ab.pin(ab.Data(matrix=X, vector=y))
    B = beta[None, :] + 0.05 * rng.standard_normal((m, k))
    ll, g = ab.ll_score_batched(F, d, B)
    for r in list(range(min(m, 6))) + [m - 1]:
        want = oll(X, y, B[r], oracle.EXACT)
        assert abs(ll[r] - float(want)) <= 1e-12 * abs(float(want)), (r, ll[r], want)
        ge, scale = osc(X, y, B[r], oracle.EXACT)
        assert np.all(np.abs(g[r] - ge.astype(float)) <= 1e-12 * scale.astype(float)), r
        l1, g1 = ab.ll_score(F, d, B[r])
        assert abs(l1 - ll[r]) <= 1e-13 * abs(l1)
    # LL only (what a Metropolis step wants): the same values up to the order of summation (the LL-only
    # kernel adds each lane's eight terms per tile before the compensated add)
    ll2, none = ab.ll_score_batched(F, d, B, want_score=False)
    assert none is None and np.allclose(ll2, ll, rtol=1e-13, atol=0)
    ab.unpin(d)


def test_bootstrap_counts_and_weighted_pass(b200, oracle):
    ab = b200
    n, k, m = 6007, 20, 24
    X, y, beta = gen_probit_logit_sample(n, k, seed=4)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    ab.bootstrap_counts(d, m, seed=11)
    c = ab.counts_download(d, m, n)
    assert np.all(c.sum(0) == n)                  # every replicate draws exactly N rows with replacement
    assert abs(c.mean() - 1.0) < 1e-12 and 0.9 < c.var() < 1.1 and c.max() < 12
    assert not np.array_equal(c[:, 0], c[:, 1])
    rng = np.random.default_rng(0)
    B = beta[None, :] + 0.05 * rng.standard_normal((m, k))
    ll, g = ab.ll_score_batched(ab.PROBIT, d, B, use_counts=True)
    for r in (0, 5, m - 1):
        idx = np.repeat(np.arange(n), c[:, r])    # the physically resampled data set
        want = oracle.probit_ll(X[idx], y[idx], B[r], oracle.EXACT)
        assert abs(ll[r] - float(want)) <= 1e-12 * abs(float(want))
        ge, scale = oracle.probit_score(X[idx], y[idx], B[r], oracle.EXACT)
        assert np.all(np.abs(g[r] - ge.astype(float)) <= 1e-12 * scale.astype(float))
    # same seed -> same resamples; different seed -> different
    ab.bootstrap_counts(d, m, seed=11)
    assert np.array_equal(ab.counts_download(d, m, n), c)
    ab.bootstrap_counts(d, m, seed=12)
    assert not np.array_equal(ab.counts_download(d, m, n), c)
    ab.unpin(d)
