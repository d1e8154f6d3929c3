"""The callers that loop over the path on the device entries:
closed-form estimates, covariance from one information pass, apop_bootstrap_cov with
resident resample counts and lock-step replicate fits, apop_model_metropolis and the
batched chains.  Statistical checks in the style of the reference's own tests
(tests/test_apop.c:709-749 jackknife/bootstrap, tests/update_via_rng.c)."""
import ctypes as C

import numpy as np
import pytest

from conftest import gen_probit_logit_sample

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    from apophenia_b200 import host as h
    h.lib()
    return h


def test_closed_form_estimates_and_covariance(b200, host, oracle):
    ab, h = b200, host.lib()
    from apophenia_b200 import abi
    lib = abi.load_library()
    rng = np.random.default_rng(1)
    # Poisson: lambda-hat = mean of every cell (model/apop_poisson.c:42-47)
    M = rng.poisson(3.1, (4000, 5)).astype(float)
    d = host.data_from_numpy(M)
    m = h.apop_model_copy(host.model_object(ab.POISSON)); lib.apop_b200_register(m, ab.POISSON)
    est = h.apop_estimate(d, m)
    assert abs(host.parameters_of(est)[0] - M.mean()) < 1e-13
    # Normal: (mean, sqrt var) (model/apop_normal.c:78-97)
    N = rng.normal(2.0, 3.0, (3000, 4))
    d = host.data_from_numpy(N)
    m = h.apop_model_copy(host.model_object(ab.NORMAL)); lib.apop_b200_register(m, ab.NORMAL)
    est = h.apop_estimate(d, m)
    assert np.allclose(host.parameters_of(est), [N.mean(), N.std()], rtol=1e-12)
    # OLS: (X'X)^-1 X'y with the outcome in column 0 (ols_shuffle), test_OLS style exact line
    X = rng.uniform(-1, 1, (5000, 4)); beta = np.array([0.5, -1.0, 2.0, 0.25])
    Xd = X.copy(); Xd[:, 0] = 1
    y = Xd @ beta
    raw = X.copy(); raw[:, 0] = y
    d = host.data_from_numpy(raw)
    m = h.apop_model_copy(host.model_object(ab.OLS)); lib.apop_b200_register(m, ab.OLS)
    est = h.apop_estimate(d, m)
    assert np.allclose(host.parameters_of(est), beta, atol=1e-12)
    # covariance of a probit fit = inverse observed information, one device pass
    Xp, yp, bp = gen_probit_logit_sample(20000, 4, seed=5)
    raw = Xp.copy(); raw[:, 0] = yp
    dp = host.data_from_numpy(raw)
    est, params, rep = host.estimate(ab.PROBIT, dp, method="Newton", tolerance=1e-9, relative_tolerance=True)
    assert rep["status"] == 0 and rep["iterations"] <= 8
    cov = h.apop_model_numerical_covariance(dp, est)
    want = np.linalg.inv(oracle.information(0, Xp, yp, params).astype(float))
    assert np.allclose(host.matrix_of(cov), want, rtol=1e-9)
    assert h.apop_data_get_page(est.contents.parameters, b"<Covariance>")


def test_bootstrap_cov_batched_probit(b200, host, oracle):
    ab, h = b200, host.lib()
    n, k, reps = 20000, 4, 400
    X, y, beta = gen_probit_logit_sample(n, k, seed=11)
    raw = X.copy(); raw[:, 0] = y
    d = host.data_from_numpy(raw)
    est, params, rep = host.estimate(ab.PROBIT, d, method="Newton", tolerance=1e-10, relative_tolerance=True)
    launches = ab.launch_count()
    cov, boots = host.bootstrap_cov(d, est, iterations=reps, seed=3)
    passes = (ab.launch_count() - launches)
    assert boots.shape == (reps, k) and passes < 80          # lock-step fits: a handful of batched passes
    # replicate estimates are the MLEs of the resampled data sets
    c = ab.counts_download(_DataPtr(d), reps, n)
    from scipy import optimize
    for r in (0, reps - 1):
        idx = np.repeat(np.arange(n), c[:, r])
        f = lambda b: -float(oracle.probit_ll(X[idx], y[idx], b))
        g = lambda b: -oracle.probit_score(X[idx], y[idx], b)[0].astype(float)
        ref = optimize.minimize(f, params, jac=g, method="BFGS", options={"gtol": 1e-8})
        assert np.max(np.abs(boots[r] - ref.x)) < 1e-6
    # and their covariance is the sandwich ~ inverse information for a correctly specified model
    want = np.linalg.inv(oracle.information(0, X, y, params).astype(float))
    assert np.allclose(np.diag(cov), np.diag(want), rtol=0.25)
    assert np.all(np.linalg.eigvalsh(cov) > 0)


class _DataPtr:
    """adapter: a raw apop_data* where the api expects an object with .ptr"""
    def __init__(self, p):
        self.ptr = p


def test_bootstrap_cov_host_loop_for_other_models(b200, host):
    """models outside the batched path take the reference's own loop: resample into the same
    scratch set, invalidate the resident copy, re-estimate (apop_bootstrap.m4.c:143-149)"""
    ab, h = b200, host.lib()
    from apophenia_b200 import abi
    lib = abi.load_library()
    rng = np.random.default_rng(2)
    n, k = 3000, 3
    X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
    beta = np.array([1.0, -2.0, 0.5]); y = X @ beta + rng.standard_normal(n)
    raw = X.copy(); raw[:, 0] = y
    d = host.data_from_numpy(raw)
    m = h.apop_model_copy(host.model_object(ab.OLS)); lib.apop_b200_register(m, ab.OLS)
    est = h.apop_estimate(d, m)
    cov, boots = host.bootstrap_cov(d, est, iterations=60, seed=1)
    want = np.linalg.inv(X.T @ X)          # sigma^2 = 1
    assert boots.shape == (60, 3)
    assert np.allclose(np.diag(cov), np.diag(want), rtol=0.6)
    assert np.allclose(boots.mean(0), beta, atol=0.1)


def test_metropolis_single_and_batched_chains(b200, host, oracle):
    ab, h = b200, host.lib()
    n, k = 20000, 3
    X, y, beta = gen_probit_logit_sample(n, k, seed=21, method="logit")
    raw = X.copy(); raw[:, 0] = y
    d = host.data_from_numpy(raw)
    est, mle, rep = host.estimate(ab.LOGIT, d, method="Newton", tolerance=1e-10, relative_tolerance=True)
    sd = np.sqrt(np.diag(np.linalg.inv(oracle.information(1, X, y, mle).astype(float))))
    # one chain: the unmodified caller, one LL pass per step
    launches = ab.launch_count()
    draws = host.metropolis(d, est, periods=3000, burnin=0.2, initial_scale=float(sd.mean()), seed=5)
    assert ab.launch_count() - launches >= 3000
    assert draws.shape == (1, 2400, k)
    assert np.all(np.abs(draws[0].mean(0) - mle) < 0.5 * sd + 1e-3)
    assert np.all((draws[0].std(0) > 0.4 * sd) & (draws[0].std(0) < 2.5 * sd))
    # 64 chains in lock-step: one batched pass per step
    launches = ab.launch_count()
    chains = host.metropolis(d, est, periods=1500, burnin=0.2, initial_scale=float(sd.mean()), seed=6, n_chains=64)
    assert ab.launch_count() - launches <= 2 * 1502
    assert chains.shape == (64, 1200, k)
    pooled = chains.reshape(-1, k)
    assert np.all(np.abs(pooled.mean(0) - mle) < 0.2 * sd)
    assert np.all((pooled.std(0) > 0.7 * sd) & (pooled.std(0) < 1.4 * sd))
    assert not np.array_equal(chains[0], chains[1])
