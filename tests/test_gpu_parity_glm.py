"""Parity of the fused CUDA kernels with the oracle, through the C-ABI.

Tolerances (BASELINE.json north_star): LL within 1e-12 relative of
the oracle in exact (x87 long double) mode; score within 1e-12 * sum_i |X_ij d_i|
per component.  The data cross the boundary as host `apop_data` structs.
"""
import numpy as np
import pytest

from conftest import gen_probit_logit_sample

pytestmark = pytest.mark.gpu

LL_RTOL = 1e-12
SCORE_RTOL = 1e-12


def _check_ll(got, want):
    want = np.longdouble(want)
    assert abs(np.longdouble(got) - want) <= LL_RTOL * abs(want) + 1e-300, (got, want)


def _check_score(got, want, scale):
    err = np.abs(np.asarray(got, dtype=np.longdouble) - want)
    assert np.all(err <= SCORE_RTOL * scale + 1e-300), (err / scale).max()


@pytest.mark.parametrize("n,k", [(80000, 8), (100003, 10), (31, 3), (32, 1), (33, 32), (50000, 32),
                                 (20001, 16), (20000, 20), (20000, 24), (1, 5), (4097, 7)])
def test_probit_ll_score(b200, oracle, n, k):
    ab = b200
    X, y, beta_true = gen_probit_logit_sample(n, k, seed=8 + n + k)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    for beta in (beta_true, oracle.probit_start(X), np.zeros(k)):
        m = ab.probit_model(beta, d)
        launches = ab.launch_count()
        ll = ab.log_likelihood(ab.PROBIT, d, m)
        g = ab.score(ab.PROBIT, d, m)
        assert ab.launch_count() == launches + 1  # f then df at the same beta: one fused pass
        _check_ll(ll, oracle.probit_ll(X, y, beta, oracle.EXACT))
        ge, scale = oracle.probit_score(X, y, beta, oracle.EXACT)
        _check_score(g, ge, scale)
        # and the gap to the reference's own operation order is the reference's rounding
        gf, _ = oracle.probit_score(X, y, beta, oracle.FAITHFUL)
        assert np.all(np.abs(g - gf) <= 1e-10 * scale + 1e-300)
        assert abs(ll - float(oracle.probit_ll(X, y, beta, oracle.FAITHFUL))) <= 1e-11 * abs(ll) + 1e-300
    ab.unpin(d)


def test_probit_clamps_and_extremes(b200, oracle):
    """rows that hit the 1e-10 clamps (model/apop_probit.c:85-86) and the cdf saturation.

    Here P(observed) = 1 - Phi(.) is formed from a double-rounded cdf, exactly as the
    reference does, so a one-ulp difference between CUDA's and libm's erfc moves the
    row's log by ~ulp/P(observed): the tolerance is built from that, per row, against
    the oracle in the reference's own (faithful) arithmetic.  (|xb| in 7.5..8.5 with the
    unlikely outcome is the knife edge where 1 - cdf has one or two significant bits in
    the reference itself; it is left out.)"""
    ab = b200
    xb = np.array([-40.0, -38.0, -37.0, -20.0, -9.0, -8.6, -6.0, -3.0, 0.0, 3.0, 6.0, 8.6, 9.0, 20.0, 37.0, 38.0, 40.0])
    X = np.stack([np.ones(xb.size * 2), np.tile(xb, 2)], axis=1)
    y = np.concatenate([np.zeros(xb.size), np.ones(xb.size)])
    beta = np.array([0.0, 1.0])
    d = ab.pin(ab.Data(matrix=X, vector=y))
    m = ab.probit_model(beta, d)
    ll = ab.log_likelihood(ab.PROBIT, d, m)
    g = ab.score(ab.PROBIT, d, m)
    assert np.isfinite(ll) and np.all(np.isfinite(g))
    # per-row P(observed) as the reference forms it
    n = np.array([oracle.gauss_cdf(-v) for v in X[:, 1]])
    n = np.where(n == 0, 1e-10, n); n = np.where(n < 1, n, 1 - 1e-10)
    arg = np.where(y != 0, 1 - n, n)
    # log(n): relative error of n; log(1-n): absolute error of n over 1-n
    tol = np.sum(8 * np.finfo(float).eps * np.where(y != 0, n / arg, 1.0)) + 1e-12 * 1300
    want = float(oracle.probit_ll(X, y, beta, oracle.FAITHFUL))
    assert abs(ll - want) <= tol, (ll, want, tol)
    assert tol < 1e-3           # the clamps differ by O(10) per row if they are wrong
    gf, scale = oracle.probit_score(X, y, beta, oracle.FAITHFUL)
    assert np.all(np.abs(g - gf.astype(float)) <= 1e-5 * scale.astype(float))
    ab.unpin(d)


def test_probit_views_and_unpinned(b200, oracle):
    """tda > size2 (matrix views), strided vector, unpinned upload-per-call, invalidate"""
    ab = b200
    X, y, beta = gen_probit_logit_sample(5000, 6, seed=3)
    big = np.zeros((5000, 9)); big[:, 2:8] = X
    ys = np.zeros(10000); ys[::2] = y
    d = ab.Data(matrix=big[:, 2:8], vector=ys[::2])
    assert d._m.tda == 9 and d._v.stride == 2
    m = ab.probit_model(beta, d)
    want = float(oracle.probit_ll(X, y, beta, oracle.EXACT))
    assert not ab.is_pinned(d)
    ll = ab.log_likelihood(ab.PROBIT, d, m)          # transient upload
    assert abs(ll - want) <= 1e-12 * abs(want)
    assert not ab.is_pinned(d)
    ab.pin(d)
    assert ab.is_pinned(d)
    # the bootstrap hazard: same pointer, new contents
    big[:, 2:8] = X[::-1]; ys[::2] = y[::-1]
    ab.invalidate(d)
    ll2 = ab.log_likelihood(ab.PROBIT, d, m)
    want2 = float(oracle.probit_ll(X[::-1], y[::-1], beta, oracle.EXACT))
    assert abs(ll2 - want2) <= 1e-12 * abs(want2)
    ab.unpin(d)


def test_error_conventions(b200):
    """NaN from LL on NULL inputs, NaN-filled score, error char on the model"""
    ab = b200
    X, y, beta = gen_probit_logit_sample(100, 4, seed=5)
    d = ab.Data(matrix=X, vector=y)
    assert np.isnan(ab.log_likelihood(ab.PROBIT, None, ab.probit_model(beta)))
    assert np.isnan(ab.log_likelihood(ab.PROBIT, d, None))
    m_bad = ab.probit_model(np.ones(7), d)  # wrong parameter count
    assert np.isnan(ab.log_likelihood(ab.PROBIT, d, m_bad))
    assert m_bad.error == b"d"
    assert np.all(np.isnan(ab.score(ab.PROBIT, d, m_bad)))
    assert ab.last_error()


@pytest.mark.parametrize("n,k", [(60000, 9), (1000, 2), (50001, 24)])
def test_binary_logit(b200, oracle, n, k):
    ab = b200
    X, y, beta = gen_probit_logit_sample(n, k, seed=n + k, method="logit")
    d = ab.pin(ab.Data(matrix=X, vector=y))
    for b in (beta, 3 * beta, np.zeros(k)):
        m = ab.logit_model(b, d)
        _check_ll(ab.log_likelihood(ab.LOGIT, d, m), oracle.logit_ll(X, y, b, mode=oracle.EXACT))
        ge, scale = oracle.logit_score(X, y, b, mode=oracle.EXACT)
        _check_score(ab.score(ab.LOGIT, d, m), ge, scale)
    ab.unpin(d)


@pytest.mark.parametrize("n,k", [(40000, 16), (999, 3)])
def test_poisson_regression(b200, oracle, n, k):
    ab = b200
    rng = np.random.default_rng(n)
    X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
    beta = rng.uniform(-1, 1, k) / np.sqrt(k)
    y = rng.poisson(np.exp(X @ beta)).astype(float)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    for b in (beta, np.zeros(k)):
        m = ab.vector_model(b, d)
        _check_ll(ab.log_likelihood(ab.POISSON_REG, d, m), oracle.poisson_reg_ll(X, y, b, oracle.EXACT))
        ge, scale = oracle.poisson_reg_score(X, y, b, oracle.EXACT)
        _check_score(ab.score(ab.POISSON_REG, d, m), ge, scale)
    ab.unpin(d)


@pytest.mark.parametrize("weights", [False, True])
def test_ols(b200, oracle, weights):
    ab = b200
    rng = np.random.default_rng(11)
    n, k = 30007, 12
    X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
    beta = rng.uniform(-1, 1, k)
    y = X @ beta + rng.standard_normal(n)
    w = rng.uniform(0.5, 2.0, n) if weights else None
    d = ab.pin(ab.Data(matrix=X, vector=y, weights=w))
    for b in (beta, beta + 0.1):
        m = ab.vector_model(b, d)
        want = oracle.ols_ll(X, y, b, w, oracle.EXACT)
        _check_ll(ab.log_likelihood(ab.OLS, d, m), want)
        ge, scale = oracle.ols_score(X, y, b, w, oracle.EXACT)
        _check_score(ab.score(ab.OLS, d, m), ge, scale)
        # the reference's own form (exp then log) agrees to its rounding
        wf = float(oracle.ols_ll(X, y, b, w, oracle.FAITHFUL))
        assert abs(float(want) - wf) <= 1e-10 * abs(wf)
    ab.unpin(d)


def test_iid_normal_and_poisson(b200, oracle):
    ab = b200
    rng = np.random.default_rng(5)
    M = rng.normal(3.0, 2.0, (20011, 7)); v = rng.normal(3.0, 2.0, 20011)
    for mat, vec in ((M, None), (None, v), (M, v)):
        d = ab.pin(ab.Data(matrix=mat, vector=vec))
        m = ab.vector_model([2.9, 2.1], d)
        _check_ll(ab.log_likelihood(ab.NORMAL, d, m), oracle.normal_ll(vec, mat, 2.9, 2.1, oracle.EXACT))
        g = ab.score(ab.NORMAL, d, m)
        ge = oracle.normal_score(vec, mat, 2.9, 2.1, oracle.EXACT)
        assert np.all(np.abs(g - ge) <= 1e-12 * np.maximum(np.abs(ge), 1e3))
        ab.unpin(d)
    P = rng.poisson(3.1, (30001, 5)).astype(float); pv = rng.poisson(3.1, 30001).astype(float)
    for mat, vec in ((P, None), (P, pv)):
        d = ab.pin(ab.Data(matrix=mat, vector=vec))
        m = ab.vector_model([3.05], d)
        _check_ll(ab.log_likelihood(ab.POISSON, d, m), oracle.poisson_ll(vec, mat, 3.05, oracle.EXACT))
        _check_ll(ab.log_likelihood(ab.POISSON, d, m), oracle.poisson_ll(vec, mat, 3.05, oracle.FAITHFUL) * (1 + 0))
        g = ab.score(ab.POISSON, d, m)
        ge = oracle.poisson_score(vec, mat, 3.05, oracle.EXACT)
        assert abs(g[0] - float(ge[0])) <= 1e-12 * abs(float(ge[0])) + 1e-6
        ab.unpin(d)
    # invalid cells give -inf ("look elsewhere", model/apop_poisson.c:15-16)
    bad = P.copy(); bad[17, 2] = 2.5
    d = ab.pin(ab.Data(matrix=bad))
    assert ab.log_likelihood(ab.POISSON, d, ab.vector_model([3.0], d)) == -np.inf
    ab.unpin(d)


def test_map_sum(b200, oracle):
    """apop_map_sum small exact cases (tests/test_apop.c:1037-1044 style) and big ones"""
    ab = b200
    M = np.arange(1.0, 13.0).reshape(4, 3); v = np.array([1.0, 2.0, 3.0, 4.0])
    d = ab.pin(ab.Data(matrix=M, vector=v))
    assert ab.map_sum(d, ab.MAP_IDENTITY) == 78 + 10
    assert ab.map_sum(d, ab.MAP_IDENTITY, part="m") == 78
    assert ab.map_sum(d, ab.MAP_IDENTITY, part="v") == 10
    assert ab.map_sum(d, ab.MAP_DEV2, 1.0, part="v") == 0 + 1 + 4 + 9
    assert abs(ab.map_sum(d, ab.MAP_LOG) - float(oracle.map_sum(v, M, 4))) < 1e-12
    assert abs(ab.map_sum(d, ab.MAP_EXP) - float(oracle.map_sum(v, M, 5))) < 1e-12 * float(oracle.map_sum(v, M, 5))
    ab.unpin(d)


def test_synth_download_roundtrip_and_parity(b200, oracle):
    """device-generated data (bench input) read back equals what the kernels saw"""
    ab = b200
    k = 10
    beta = np.linspace(-0.9, 0.9, k)
    s = ab.synth(ab.PROBIT, 100003, k, beta, seed=8)
    X, y = s.download()
    assert np.all(X[:, 0] == 1.0) and np.all(np.abs(X[:, 1:]) < 1) and set(np.unique(y)) <= {0.0, 1.0}
    assert abs(X[:, 1:].mean()) < 0.01 and 0.2 < y.mean() < 0.8
    ll, g = ab.ll_score(ab.PROBIT, s, beta)
    _check_ll(ll, oracle.probit_ll(X, y, beta, oracle.EXACT))
    ge, scale = oracle.probit_score(X, y, beta, oracle.EXACT)
    _check_score(g, ge, scale)
    # sharded generation reproduces the same global data set
    s2 = ab.synth(ab.PROBIT, 50003, k, beta, seed=8, row_offset=50000)
    X2, y2 = s2.download()
    assert np.array_equal(X2, X[50000:]) and np.array_equal(y2, y[50000:])
    s.free(); s2.free()


@pytest.mark.parametrize("n,k,ncat", [(40000, 32, 8), (30001, 5, 3), (9999, 24, 4), (20000, 8, 8), (33, 16, 5),
                                      (10000, 9, 6), (10000, 17, 7)])
def test_multinomial_logit(b200, oracle, n, k, ncat):
    """cfg2 shape (k=32, C=8) and friends: LL per one_logit_row, analytic score in pack order"""
    ab = b200
    X, y, B = gen_probit_logit_sample(n, k, seed=n + k + ncat, method="logit", ncat=ncat)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    # the first multinomial pass over a data set converts the resident tiles to the swizzled layout (one extra launch)
    ab.log_likelihood(ab.LOGIT, d, ab.logit_model(0.5 * B, d))
    for Bm in (B, np.zeros_like(B), 4 * B):
        m = ab.logit_model(Bm, d)
        launches = ab.launch_count()
        ll = ab.log_likelihood(ab.LOGIT, d, m)
        g = ab.score(ab.LOGIT, d, m)
        assert ab.launch_count() == launches + 1
        _check_ll(ll, oracle.logit_ll(X, y, Bm, mode=oracle.EXACT))
        ge, scale = oracle.logit_score(X, y, Bm, mode=oracle.EXACT)
        _check_score(g, ge, scale)
        assert abs(ll - float(oracle.logit_ll(X, y, Bm, mode=oracle.FAITHFUL))) <= 1e-11 * abs(ll)
    ab.unpin(d)


def test_logit_category_page_and_overflow(b200, oracle):
    """category values from the factor page (find_index), and rows whose exp(-max) overflows"""
    ab = b200
    X, y, B = gen_probit_logit_sample(5000, 4, seed=1, method="logit", ncat=3)
    cats = np.array([2.0, 5.0, 9.0])
    d = ab.Data(matrix=X, vector=cats[y.astype(int)])
    d.add_page(ab.Data(vector=cats, title="<categories for vector>"))
    ab.pin(d)
    m = ab.logit_model(B, d)
    _check_ll(ab.log_likelihood(ab.LOGIT, d, m), oracle.logit_ll(X, y, B, mode=oracle.EXACT))
    ab.unpin(d)
    # huge negative x.beta: expl(-max) is finite in x87 up to 11356, in double up to 709
    Xo = np.array([[1.0, -800.0], [1.0, 800.0], [1.0, -5000.0], [1.0, 0.0]])
    yo = np.array([1.0, 0.0, 0.0, 1.0])
    for Bo in (np.array([[0.0], [1.0]]), np.array([[0.0, 0.5], [1.0, 2.0]])):
        d = ab.pin(ab.Data(matrix=Xo, vector=yo))
        m = ab.logit_model(Bo, d)
        want = float(oracle.logit_ll(Xo, yo, Bo, mode=oracle.EXACT))
        got = ab.log_likelihood(ab.LOGIT, d, m)
        assert np.isfinite(got) and abs(got - want) <= 1e-12 * abs(want), (got, want)
        ge, scale = oracle.logit_score(Xo, yo, Bo, mode=oracle.EXACT)
        _check_score(ab.score(ab.LOGIT, d, m), ge, scale)
        ab.unpin(d)


@pytest.mark.parametrize("n,k", [(30011, 32), (5000, 5), (20000, 20)])
def test_information_and_gram(b200, oracle, n, k):
    """observed information by one device pass (replaces the numeric Hessian,
    apop_mle.m4.c:187-223) and the OLS normal equations (model/apop_ols.c:333-334)"""
    ab = b200
    X, y, beta = gen_probit_logit_sample(n, k, seed=n + k)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    for fam, ofam in ((ab.PROBIT, 0), (ab.LOGIT, 1), (ab.POISSON_REG, 5)):
        b = beta / (np.sqrt(k) if fam == ab.POISSON_REG else 1)
        H = ab.information(fam, d, b)
        want = oracle.information(ofam, X, y, b).astype(float)
        scale = np.abs(want).max()
        assert np.max(np.abs(H - want)) <= 1e-12 * scale * k, fam
        assert np.allclose(H, H.T, rtol=0, atol=1e-13 * scale)
    A, bvec = ab.ols_gram(d)
    wa, wb = oracle.ols_gram(X, y, oracle.EXACT)
    assert np.max(np.abs(A - wa.astype(float))) <= 1e-13 * np.abs(wa).max().astype(float) * 10
    assert np.max(np.abs(bvec - wb.astype(float))) <= 1e-12 * (np.abs(X).T @ np.abs(y)).max()
    ab.unpin(d)


def test_nist_pontius_through_device_gram(b200):
    """tests/nist_tests.c:20-35: certified OLS coefficients from the device X'X, X'y"""
    import json, os
    ab = b200
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "nist_strd.json")))["pontius"]
    yx = np.array(g["y_x"]); y, x = yx[:, 0], yx[:, 1]
    X = np.stack([np.ones_like(x), x, x * x], axis=1)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    A, b = ab.ols_gram(d)
    beta = np.linalg.solve(A, b)
    for got, want, tol in zip(beta, g["certified_beta"], g["tolerances"]):
        assert abs(got - want) < tol
    ab.unpin(d)


def test_prep_reductions_on_device(b200, oracle):
    """probit_prep's start point and the column sums, reduced on the resident data"""
    ab = b200
    X, y, beta = gen_probit_logit_sample(30011, 12, seed=9)
    X[5, 3] = 0.0; X[77, 11] = 0.0                    # zeros contribute 0 to the mean log
    d = ab.pin(ab.Data(matrix=X, vector=y))
    assert np.allclose(ab.probit_start(d, 12), oracle.probit_start(X), rtol=1e-13)
    assert np.allclose(ab.col_sums(d, 1, 12), X.sum(0), rtol=1e-13, atol=1e-9)
    assert np.allclose(ab.col_sums(d, 2, 12), (X * X).sum(0), rtol=1e-13)
    ab.unpin(d)


def test_dot_matches_reference_dgemm(b200):
    """apop_dot (tests/test_apop.c:527-550 checks exact equality of small products; here the
    index X.B against numpy, and the 'd' error char on a dimension mismatch)"""
    ab = b200
    rng = np.random.default_rng(3)
    X = rng.uniform(-1, 1, (10007, 9)); y = np.zeros(10007)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    for c in (1, 3):
        B = rng.uniform(-1, 1, (9, c))
        got = ab.dot(d, B, 10007)
        assert np.max(np.abs(got - X @ B)) < 1e-14
    Xi = np.arange(12.0).reshape(4, 3)
    di = ab.pin(ab.Data(matrix=Xi, vector=np.zeros(4)))
    assert np.array_equal(ab.dot(di, np.array([1.0, 2.0, 3.0]), 4).ravel(), Xi @ [1.0, 2.0, 3.0])  # exact on small integers
    with pytest.raises(ab.B200Error):
        ab.dot(d, np.ones((5, 1)), 10007)
    assert d.struct.error == b"d"
    ab.unpin(d); ab.unpin(di)


@pytest.mark.parametrize("fam", ["probit", "logit", "poisson_reg", "ols"])
def test_every_column_count_1_to_32(b200, oracle, fam):
    """each K = 1..32 is its own kernel instantiation: all of them against the oracle"""
    ab = b200
    rng = np.random.default_rng(17)
    n = 2500
    for k in range(1, 33):
        X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
        beta = rng.uniform(-1, 1, k) / np.sqrt(k)
        xb = X @ beta
        if fam == "probit":
            y = (rng.standard_normal(n) > -xb).astype(float)
            F, ll_o, sc_o = ab.PROBIT, oracle.probit_ll(X, y, beta), oracle.probit_score(X, y, beta)
            m = lambda d: ab.probit_model(beta, d)
        elif fam == "logit":
            y = (rng.uniform(size=n) < 1 / (1 + np.exp(-xb))).astype(float)
            F, ll_o, sc_o = ab.LOGIT, oracle.logit_ll(X, y, beta), oracle.logit_score(X, y, beta)
            m = lambda d: ab.logit_model(beta, d)
        elif fam == "poisson_reg":
            y = rng.poisson(np.exp(xb)).astype(float)
            F, ll_o, sc_o = ab.POISSON_REG, oracle.poisson_reg_ll(X, y, beta), oracle.poisson_reg_score(X, y, beta)
            m = lambda d: ab.vector_model(beta, d)
        else:
            y = xb + rng.standard_normal(n)
            F, ll_o, sc_o = ab.OLS, oracle.ols_ll(X, y, beta * 1.05), oracle.ols_score(X, y, beta * 1.05)
            m = lambda d: ab.vector_model(beta * 1.05, d)
        d = ab.pin(ab.Data(matrix=X, vector=y))
        mod = m(d)
        _check_ll(ab.log_likelihood(F, d, mod), ll_o)
        _check_score(ab.score(F, d, mod), sc_o[0], sc_o[1])
        ab.unpin(d)


def test_empty_and_too_wide_inputs(b200):
    """ragged edges: zero rows give a zero sum (apop_map_sum of nothing), k > 32 is refused loudly"""
    ab = b200
    d0 = ab.pin(ab.Data(matrix=np.zeros((0, 3)), vector=np.zeros(0)))
    assert ab.log_likelihood(ab.PROBIT, d0, ab.probit_model([0.1, 0.2, 0.3], d0)) == 0.0
    assert np.all(ab.score(ab.PROBIT, d0, ab.probit_model([0.1, 0.2, 0.3], d0)) == 0.0)
    assert ab.map_sum(d0, ab.MAP_IDENTITY) == 0.0
    ab.unpin(d0)
    X = np.ones((64, 1100)); y = np.zeros(64)
    dw = ab.pin(ab.Data(matrix=X, vector=y))
    mw = ab.probit_model(np.zeros(1100), dw)
    assert np.isnan(ab.log_likelihood(ab.PROBIT, dw, mw)) and mw.error == b"d"
    assert "1..1024" in ab.last_error()
    ab.unpin(dw)


@pytest.mark.parametrize("fam,n,k", [("probit", 20011, 40), ("probit", 5000, 33), ("logit", 9000, 64), ("poisson_reg", 6000, 100),
                                     ("ols", 7000, 48), ("probit", 3000, 257), ("probit", 4001, 129), ("logit", 3000, 160),
                                     ("poisson_reg", 2500, 200), ("ols", 3001, 225), ("probit", 2000, 256), ("logit", 2000, 700),
                                     ("logit", 4001, 49), ("probit", 1500, 513), ("poisson_reg", 1200, 1024), ("ols", 2001, 1000),
                                     ("probit", 31, 300), ("ols", 45, 520), ("probit", 1999, 768)])
def test_wide_data_beyond_32_columns(b200, oracle, fam, n, k):
    """k > 32: the register-resident kernel to k = 48, then glm_wide.cu (the lanes of a warp tile a rows x columns
    patch: whole tiles to k = 256, half tiles in CTAs of up to 8 warps to 512 and up to 16 warps to 1024); same parity bar"""
    ab = b200
    rng = np.random.default_rng(n + k)
    X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
    beta = rng.uniform(-1, 1, k) / np.sqrt(k)
    xb = X @ beta
    w = None
    if fam == "probit":
        y = (rng.standard_normal(n) > -xb).astype(float)
        F, want, sc, mk = ab.PROBIT, oracle.probit_ll(X, y, beta), oracle.probit_score(X, y, beta), ab.probit_model
    elif fam == "logit":
        y = (rng.uniform(size=n) < 1 / (1 + np.exp(-xb))).astype(float)
        F, want, sc, mk = ab.LOGIT, oracle.logit_ll(X, y, beta), oracle.logit_score(X, y, beta), ab.logit_model
    elif fam == "poisson_reg":
        y = rng.poisson(np.exp(xb)).astype(float)
        F, want, sc, mk = ab.POISSON_REG, oracle.poisson_reg_ll(X, y, beta), oracle.poisson_reg_score(X, y, beta), ab.vector_model
    else:
        y = xb + rng.standard_normal(n)
        w = rng.uniform(0.5, 2, n)
        F, want, sc, mk = ab.OLS, oracle.ols_ll(X, y, beta, w), oracle.ols_score(X, y, beta, w), ab.vector_model
    d = ab.pin(ab.Data(matrix=X, vector=y, weights=w))
    m = mk(beta, d)
    launches = ab.launch_count()
    _check_ll(ab.log_likelihood(F, d, m), want)
    _check_score(ab.score(F, d, m), sc[0], sc[1])
    # one fused pass for LL and score (+ the one-off lgamma constant of the Poisson regression)
    assert ab.launch_count() == launches + (2 if fam == "poisson_reg" else 1)
    ab.unpin(d)


def test_exponential_gamma_lognormal(b200, oracle):
    """the next iid models on the map_sum shape: LL and score at fixed
    parameters against the oracle; parameter-independent sums are reduced once per residency"""
    ab = b200
    rng = np.random.default_rng(6)
    M = rng.gamma(2.5, 1.7, (20011, 5)); v = rng.gamma(2.5, 1.7, 20011)
    for mat, vec in ((M, None), (M, v)):
        d = ab.pin(ab.Data(matrix=mat, vector=vec))
        launches = ab.launch_count()
        for fam, params, ofn in ((ab.EXPONENTIAL, [3.9], oracle.exponential), (ab.GAMMA, [2.4, 1.8], oracle.gamma),
                                 (ab.GAMMA, [2.6, 1.5], oracle.gamma)):
            m = ab.vector_model(params, d)
            want, gw = ofn(vec, mat, *params)
            _check_ll(ab.log_likelihood(fam, d, m), want)
            g = ab.score(fam, d, m)
            assert np.all(np.abs(g - gw.astype(float)) <= 1e-11 * np.maximum(np.abs(gw.astype(float)), mat.size))
        assert ab.launch_count() == launches + 1           # one pass serves every parameter value
        m = ab.vector_model([0.9, 0.7], d)
        want, gw = oracle.lognormal(vec, mat, 0.9, 0.7)
        _check_ll(ab.log_likelihood(ab.LOGNORMAL, d, m), want)
        g = ab.score(ab.LOGNORMAL, d, m)
        assert np.all(np.abs(g - gw.astype(float)) <= 1e-11 * np.maximum(np.abs(gw.astype(float)), mat.size))
        ab.unpin(d)
    Mz = M.copy(); Mz[3, 2] = 0.0                           # a zero cell: finite gamma LL, dLL/da = -inf
    d = ab.pin(ab.Data(matrix=Mz))
    m = ab.vector_model([2.4, 1.8], d)
    want, gw = oracle.gamma(None, Mz, 2.4, 1.8)
    _check_ll(ab.log_likelihood(ab.GAMMA, d, m), want)
    assert ab.score(ab.GAMMA, d, m)[0] == -np.inf
    ab.unpin(d)


@pytest.mark.gpu
def test_bernoulli_beta_zipf_dirichlet_t_yule(b200, oracle):
    """the remaining iid models of SURVEY 8(f) rank 3 through the reference-signature entries: LL (and
    score where the reference has one, plus the new Bernoulli score) at fixed parameters against the oracle"""
    ab = b200
    rng = np.random.default_rng(12)

    def check(fam, mat, vec, params, ofn, has_score=True, one_pass=True, launches_per_call=0):
        d = ab.pin(ab.Data(matrix=mat, vector=vec))
        launches = ab.launch_count()
        for prm in params:
            m = ab.vector_model(prm, d)
            want, gw = ofn(vec, mat, *prm)
            _check_ll(ab.log_likelihood(fam, d, m), want)
            if has_score:
                g = ab.score(fam, d, m)
                gwf = gw.astype(float)
                assert np.all(np.abs(g - gwf) <= 1e-11 * np.maximum(np.abs(gwf), mat.size)), (g, gwf)
        if one_pass:
            assert ab.launch_count() == launches + 1           # parameter-independent sums: one pass per residency
        else:
            assert ab.launch_count() == launches + len(params)  # LL and score at one parameter share a pass
        ab.unpin(d)

    M = rng.beta(2.0, 3.0, (20011, 5)); v = rng.beta(2.0, 3.0, 20011)
    check(ab.BETA, M, None, [[2.2, 3.1], [1.7, 2.9]], oracle.beta)
    check(ab.BETA, M, v, [[2.2, 3.1]], oracle.beta)
    b = (rng.uniform(size=(30001, 3)) < 0.3) * 2.0
    check(ab.BERNOULLI, b, None, [[0.31], [0.5]], oracle.bernoulli)
    Z = rng.zipf(2.3, (20000, 4)).astype(float)
    check(ab.ZIPF, Z, None, [[2.1], [1.5]], oracle.zipf, has_score=False)
    T = rng.standard_t(5, (20011, 4)) * 2 + 1
    check(ab.TDIST, T, None, [[0.9, 2.1, 5.5], [1.0, 2.0, 4.0]], oracle.tdist, has_score=False, one_pass=False)
    Y = rng.zipf(2.5, (20000, 3)).astype(float)
    check(ab.YULE, Y, Y[:, 0].copy(), [[2.4], [3.3]], oracle.yule, one_pass=False)
    # Dirichlet: rows of the matrix
    D = rng.dirichlet([1.5, 2.5, 3.0, 0.8], 30011)
    d = ab.pin(ab.Data(matrix=D))
    launches = ab.launch_count()
    for a in ([1.4, 2.6, 3.1, 0.9], [1.0, 1.0, 1.0, 1.0]):
        m = ab.vector_model(a, d)
        want, gw = oracle.dirichlet(D, a)
        _check_ll(ab.log_likelihood(ab.DIRICHLET, d, m), want)
        g = ab.score(ab.DIRICHLET, d, m)
        assert np.all(np.abs(g - gw.astype(float)) <= 1e-11 * np.maximum(np.abs(gw.astype(float)), D.shape[0]))
    assert ab.launch_count() == launches + 1
    ab.unpin(d)
    # beta: a cell outside [0,1] drops out of the LL but poisons the unguarded score sums, as in the reference
    Mo = M.copy(); Mo[5, 1] = 1.5
    d = ab.pin(ab.Data(matrix=Mo))
    m = ab.vector_model([2.2, 3.1], d)
    want, gw = oracle.beta(None, Mo, 2.2, 3.1)
    _check_ll(ab.log_likelihood(ab.BETA, d, m), want)
    assert np.isnan(ab.score(ab.BETA, d, m)[1])
    ab.unpin(d)
