"""Round-2 additions, through the C-ABI, against the oracle / numpy:

* a weights column is ignored by probit, logit and the Poisson regression (the reference's never read it);
* every branch of ols_log_likelihood / ols_score (model/apop_ols.c:129-205): un-prepped data, the
  <Predicted> and <Error variance> pages, the input_distribution of apop_lm_settings;
* apop_b200_ll_score for all 15 families;
* apop_map with a materialised result (cell and row functions);
* the multinomial-logit observed information (P x P) against the oracle, the numeric-Hessian fallback
  (apop_model_hessian) for families without an analytic form, Newton and covariance on top of both;
* the knife-edge band of the probit link (|x.beta| 7.5 .. 8.5 with the unlikely outcome);
* residency hygiene: a header whose buffers changed is re-uploaded, counts die with a re-upload.
"""
import ctypes as C

import numpy as np
import pytest

from conftest import gen_probit_logit_sample

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-300)


# ------------------------------------------------------------------------------------------------
def test_weights_column_is_ignored_by_probit_logit_poisson(b200, oracle):
    ab = b200
    n, k = 30011, 9
    X, y, beta = gen_probit_logit_sample(n, k, seed=21)
    w = np.random.default_rng(2).uniform(0.5, 2.0, n)
    d = ab.pin(ab.Data(matrix=X, vector=y, weights=w))
    m = ab.probit_model(beta, d)
    ll = ab.log_likelihood(ab.PROBIT, d, m)
    assert _rel(ll, oracle.probit_ll(X, y, beta, oracle.EXACT)) <= 1e-12
    ge, sc = oracle.probit_score(X, y, beta, oracle.EXACT)
    assert np.all(np.abs(ab.score(ab.PROBIT, d, m) - ge.astype(float)) <= 1e-12 * sc.astype(float))
    # binary logit and the batched path on the same weighted data set
    ml = ab.logit_model(beta, d)
    assert _rel(ab.log_likelihood(ab.LOGIT, d, ml), oracle.logit_ll(X, y, beta.reshape(-1, 1), mode=oracle.EXACT)) <= 1e-12
    Bs = beta[None, :] * np.array([1.0, 0.9, 1.1])[:, None]
    llb, gb = ab.ll_score_batched(ab.PROBIT, d, Bs)
    for r in range(3):
        assert _rel(llb[r], oracle.probit_ll(X, y, Bs[r], oracle.EXACT)) <= 1e-12
    H = ab.information(ab.PROBIT, d, beta)
    assert np.max(np.abs(H - oracle.information(0, X, y, beta).astype(float))) <= 1e-11 * np.abs(H).max()
    ab.unpin(d)
    # Poisson regression
    yp = np.floor(np.abs(3 * X[:, 1] + 2)) % 6
    dp = ab.pin(ab.Data(matrix=X, vector=yp, weights=w))
    bp = beta / 4
    assert _rel(ab.log_likelihood(ab.POISSON_REG, dp, ab.vector_model(bp, dp)), oracle.poisson_reg_ll(X, yp, bp, oracle.EXACT)) <= 1e-12
    ab.unpin(dp)
    # multinomial logit: the staged tile skips the weights column
    Xl, yl, B = gen_probit_logit_sample(20003, 12, seed=5, method="logit", ncat=5)
    wl = np.random.default_rng(3).uniform(0.5, 2.0, 20003)
    dl = ab.pin(ab.Data(matrix=Xl, vector=yl, weights=wl))
    mlg = ab.logit_model(B, dl)
    assert _rel(ab.log_likelihood(ab.LOGIT, dl, mlg), oracle.logit_ll(Xl, yl, B, mode=oracle.EXACT)) <= 1e-12
    gl, sl = oracle.logit_score(Xl, yl, B, mode=oracle.EXACT)
    assert np.all(np.abs(ab.score(ab.LOGIT, dl, mlg) - gl.astype(float).ravel()) <= 1e-12 * sl.astype(float).ravel())
    ab.unpin(dl)


# ------------------------------------------------------------------------------------------------
def _ols_reference(raw_or_X, y, beta, w=None, unprepped=False):
    """ols_log_likelihood / ols_score in long double, following model/apop_ols.c:129-205 line by line"""
    L = np.longdouble
    A = np.asarray(raw_or_X, dtype=L)
    b = np.asarray(beta, dtype=L)
    n = A.shape[0]
    if unprepped:
        actual = A[:, 0]
        expected = A @ b + b[0] * (1 - actual)
    else:
        actual = np.asarray(y, dtype=L)
        expected = A @ b
    e = expected - actual
    mean = e.sum() / n
    var = ((e - mean) ** 2).sum() / (n - 1)
    ww = np.ones(n, dtype=L) if w is None else np.asarray(w, dtype=L)
    sigma = np.sqrt(var)
    ll = (np.log(np.exp(-e * e / (2 * var)) / (sigma * np.sqrt(2 * L(np.pi))) * ww)).sum()
    g = ((ww * e / var)[:, None] * A).sum(0)       # apop_data_get(d, i, j): the RAW column 0 when un-prepped
    return ll, g, e, var


def test_ols_unprepped_branch(b200):
    """model/apop_ols.c:149-153 and :186-190: outcome still in column 0, no vector"""
    ab = b200
    rng = np.random.default_rng(4)
    n, k = 20017, 7
    X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
    beta = rng.uniform(-1, 1, k)
    yv = X @ beta + rng.standard_normal(n)
    raw = X.copy(); raw[:, 0] = yv
    b_eval = beta * 1.05
    for w in (None, rng.uniform(0.5, 2.0, n)):
        d = ab.Data(matrix=raw.copy(), weights=w)             # un-prepped: no vector
        m = ab.vector_model(b_eval, d)
        ll = ab.log_likelihood(ab.OLS, d, m)
        g = ab.score(ab.OLS, d, m)
        wl, wg, _, _ = _ols_reference(raw, None, b_eval, w, unprepped=True)
        assert _rel(ll, wl) <= 1e-12
        scale = np.abs(raw).sum(0) * 5
        assert np.all(np.abs(g - wg.astype(float)) <= 1e-12 * scale), (g, wg)
        # and the prepped evaluation agrees in LL (the shuffle is only a change of representation)
        dpz = ab.Data(matrix=X, vector=yv, weights=w)
        assert _rel(ab.log_likelihood(ab.OLS, dpz, ab.vector_model(b_eval, dpz)), ll) <= 1e-13


def test_ols_predicted_error_variance_and_input_distribution(b200, oracle):
    from apophenia_b200 import abi, host
    ab = b200
    h = host.lib()
    rng = np.random.default_rng(6)
    n, k = 15013, 5
    X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
    beta = rng.uniform(-1, 1, k)
    yv = X @ beta + 0.7 * rng.standard_normal(n)
    d = ab.pin(ab.Data(matrix=X, vector=yv))
    m = ab.vector_model(beta * 0.9, d)
    base, _, e, var = _ols_reference(X, yv, beta * 0.9)
    assert _rel(ab.log_likelihood(ab.OLS, d, m), base) <= 1e-12
    # ---- <Error variance> page on the parameters (:157)
    L = np.longdouble
    ev = ab.Data(matrix=np.array([[0.5]]), title="<Error variance>")
    m.parameters.add_page(ev)
    want = (-(e * e).sum() / (2 * L(0.5)) - n * np.log(np.sqrt(L(0.5)) * np.sqrt(2 * L(np.pi))))
    assert _rel(ab.log_likelihood(ab.OLS, d, m), want) <= 1e-12
    # ---- <Predicted> page on the info of a model whose data is d (:139-145): errors = column 2, zero padded
    m2 = ab.vector_model(beta * 0.9, d)
    stored = rng.standard_normal(n - 100) * 0.3                  # shorter than the data: the rest are zero errors
    page = np.zeros((n - 100, 3)); page[:, 2] = stored
    info = ab.Data(matrix=page, title="<Predicted>")
    m2.struct.info = info.ptr
    ee = np.concatenate([stored, np.zeros(100)]).astype(L)
    v2 = ((ee - ee.sum() / n) ** 2).sum() / (n - 1)
    want2 = -(ee * ee).sum() / (2 * v2) - n * np.log(np.sqrt(v2) * np.sqrt(2 * L(np.pi)))
    assert _rel(ab.log_likelihood(ab.OLS, d, m2), want2) <= 1e-12
    # ... but only for the model's own data set (d == p->data)
    d_other = ab.pin(ab.Data(matrix=X.copy(), vector=yv.copy()))
    assert _rel(ab.log_likelihood(ab.OLS, d_other, m2), base) <= 1e-12
    ab.unpin(d_other)
    # ---- input_distribution (:163-166): a device Normal over the regressors multiplies every row's density
    m3 = ab.vector_model(beta * 0.9, d)
    idist = h.apop_model_copy(host.model_object(abi.NORMAL))
    abi.check(abi.load_library().apop_b200_register(idist, abi.NORMAL), "register")
    idist.contents.parameters = h.apop_data_alloc(2, 0, 0)
    idist.contents.parameters.contents.vector.contents.data[0] = 0.1
    idist.contents.parameters.contents.vector.contents.data[1] = 0.8

    class LmSettings(C.Structure):
        _fields_ = [("destroy_data", C.c_int), ("instruments", C.c_void_p), ("want_cov", C.c_char),
                    ("want_expected_value", C.c_char), ("input_distribution", C.POINTER(abi.apop_model))]
    grp = LmSettings(0, None, b"n", b"n", idist)
    h.apop_settings_group_alloc(m3.ptr, b"apop_lm", C.byref(grp), C.sizeof(grp))
    xterm = oracle.normal_ll(None, X, 0.1, 0.8, oracle.EXACT)     # sum over every cell of the matrix part
    assert _rel(ab.log_likelihood(ab.OLS, d, m3), base + xterm) <= 1e-12
    ab.unpin(d)


# ------------------------------------------------------------------------------------------------
def test_ll_score_covers_every_family(b200, oracle):
    """apop_b200_ll_score's switch used to stop at the lognormal"""
    ab = b200
    rng = np.random.default_rng(9)
    M = rng.beta(2.0, 3.0, (5003, 3))
    d = ab.pin(ab.Data(matrix=M))
    wl, wg = oracle.beta(None, M, 2.2, 3.1)
    ll, g = ab.ll_score(ab.BETA, d, [2.2, 3.1])
    assert _rel(ll, wl) <= 1e-12 and np.all(np.abs(g - wg.astype(float)) <= 1e-11 * M.size)
    ab.unpin(d)
    Md = M / M.sum(1, keepdims=True)
    wl, wg = oracle.dirichlet(Md, np.array([1.5, 2.0, 2.5]))
    dd = ab.pin(ab.Data(matrix=Md))
    ll, g = ab.ll_score(ab.DIRICHLET, dd, [1.5, 2.0, 2.5])
    assert _rel(ll, wl) <= 1e-12 and np.all(np.abs(g - wg.astype(float)) <= 1e-11 * M.size)
    ab.unpin(dd)
    Z = np.floor(rng.pareto(1.5, (4001, 2))) + 1
    dz = ab.pin(ab.Data(matrix=Z))
    wl, _ = oracle.zipf(None, Z, 1.7)
    ll, g = ab.ll_score(ab.ZIPF, dz, [1.7])      # no analytic score: the numeric gradient over the device LL
    h = 1e-5
    fd = (float(oracle.zipf(None, Z, 1.7 + h)[0]) - float(oracle.zipf(None, Z, 1.7 - h)[0])) / (2 * h)
    assert _rel(ll, wl) <= 1e-12 and abs(g[0] - fd) <= 1e-5 * abs(fd)
    wl, wg = oracle.yule(None, Z, 2.3)
    ll, g = ab.ll_score(ab.YULE, dz, [2.3])
    assert _rel(ll, wl) <= 1e-12 and abs(g[0] - float(wg[0])) <= 1e-11 * Z.size
    ab.unpin(dz)
    T = rng.standard_t(5, (3001, 2)) * 1.3 + 0.4
    dt = ab.pin(ab.Data(matrix=T))
    wl, _ = oracle.tdist(None, T, 0.4, 1.3, 5.0)
    ll, g = ab.ll_score(ab.TDIST, dt, [0.4, 1.3, 5.0])
    assert _rel(ll, wl) <= 1e-12 and np.all(np.isfinite(g)) and g.size == 3
    ab.unpin(dt)
    Bn = (rng.uniform(size=(2001, 3)) < 0.3).astype(float)
    db = ab.pin(ab.Data(matrix=Bn))
    wl, wg = oracle.bernoulli(None, Bn, 0.35)
    ll, g = ab.ll_score(ab.BERNOULLI, db, [0.35])
    assert _rel(ll, wl) <= 1e-12 and abs(g[0] - float(wg[0])) <= 1e-11 * Bn.size
    ab.unpin(db)


# ------------------------------------------------------------------------------------------------
def test_map_materialised(b200, oracle):
    """apop_map's output shapes (apop_mapply.m4.c:300-313): same shape for cell functions, one number per row"""
    ab = b200
    rng = np.random.default_rng(12)
    n, k = 10007, 6
    X = rng.uniform(0.1, 2.0, (n, k))
    y = rng.uniform(0.1, 2.0, n)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    v, M = ab.map(d, ab.MAP_LOG, part="a")
    assert np.max(np.abs(M - np.log(X))) <= 4e-16 * np.max(np.abs(np.log(X))) + 1e-16 and np.max(np.abs(v - np.log(y))) <= 1e-15
    v, M = ab.map(d, ab.MAP_DEV2, 0.7, part="m")          # the vector is copied through
    assert np.allclose(M, (X - 0.7) ** 2, rtol=1e-15, atol=0) and np.array_equal(v, y)
    v, M = ab.map(d, ab.MAP_SCALE, 3.0, part="v")
    assert np.array_equal(M, X) and np.array_equal(v, 3.0 * y)
    assert np.allclose(ab.map(d, ab.MAP_EXP)[1], np.exp(X), rtol=4e-16)
    # the map then the sum is apop_map_sum
    assert _rel(ab.map_sum(d, ab.MAP_LOG, part="m"), np.log(X).astype(np.longdouble).sum()) <= 1e-13
    assert np.allclose(ab.map_rows(d, ab.ROW_SUM, n), X.sum(1), rtol=1e-15)
    assert np.array_equal(ab.map_rows(d, ab.ROW_MAX, n), X.max(1))
    assert np.allclose(ab.map_rows(d, ab.ROW_MEAN, n), X.mean(1), rtol=1e-15)
    p = rng.uniform(-1, 1, k)
    assert np.allclose(ab.map_rows(d, ab.ROW_DOT, n, p), X @ p, rtol=0, atol=1e-14)
    assert np.allclose(ab.map_rows(d, ab.ROW_OLS_ERROR, n, p), X @ p - y, rtol=0, atol=1e-14)
    ab.unpin(d)
    # the per-row likelihood terms (biprobit_ll_row) sum to the fused LL and match the oracle row by row
    Xp, yp, beta = gen_probit_logit_sample(3001, 5, seed=2)
    dp = ab.pin(ab.Data(matrix=Xp, vector=yp))
    rows = ab.map_rows(dp, ab.ROW_PROBIT_LL, 3001, beta)
    for i in (0, 17, 3000):
        assert _rel(rows[i], oracle.probit_ll(Xp[i:i + 1], yp[i:i + 1], beta, oracle.EXACT)) <= 1e-13
    assert _rel(rows.astype(np.longdouble).sum(), ab.ll_score(ab.PROBIT, dp, beta)[0]) <= 1e-13
    rl = ab.map_rows(dp, ab.ROW_LOGIT2_LL, 3001, beta)
    assert _rel(rl.astype(np.longdouble).sum(), oracle.logit_ll(Xp, yp, beta.reshape(-1, 1), mode=oracle.EXACT)) <= 1e-12
    # dimension error convention: error char 'd' on the data
    with pytest.raises(ab.B200Error):
        ab.map_rows(dp, ab.ROW_DOT, 3001, np.ones(4))
    assert dp.struct.error == b"d"
    ab.unpin(dp)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k,C", [(20011, 5, 3), (15000, 12, 5), (30001, 32, 8), (4001, 17, 4), (31, 8, 8)])
def test_multinomial_information_matches_oracle(b200, oracle, n, k, C):
    ab = b200
    X, y, B = gen_probit_logit_sample(n, k, seed=n + k, method="logit", ncat=C)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    for Bv in (B, np.zeros_like(B)):
        H = ab.information(ab.LOGIT, d, Bv.ravel())
        want = oracle.logit_information(X, Bv, C)
        assert np.max(np.abs(H - want.astype(float))) <= 1e-12 * float(np.abs(want).max()) * 8
        assert np.array_equal(H, H.T)
        assert np.all(np.linalg.eigvalsh(H) > -1e-9 * np.abs(H).max())
    ab.unpin(d)


def test_numeric_hessian_fallback_and_covariance(b200, oracle):
    """apop_model_hessian (apop_mle.m4.c:187-223) over the device score, for a family without an analytic
    information (gamma), and apop_model_numerical_covariance on top of it; for probit the device information
    and the numeric Hessian must agree"""
    from apophenia_b200 import abi, host
    ab = b200
    h, lib = host.lib(), abi.load_library()
    rng = np.random.default_rng(3)
    G = rng.gamma(2.5, 1.7, (20000, 2))
    d = ab.pin(ab.Data(matrix=G))
    m = h.apop_model_copy(host.model_object(abi.GAMMA))
    abi.check(lib.apop_b200_register(m, abi.GAMMA), "register")
    m.contents.parameters = h.apop_data_alloc(2, 0, 0)
    pv = m.contents.parameters.contents.vector.contents.data
    pv[0], pv[1] = 2.4, 1.8
    Hn = host.matrix_of(h.apop_model_hessian(d.ptr, m, 0.0))
    # d2LL/da2 = -N psi'(a), d2/dadb = -N/b, d2/db2 = -2 sum x / b^3 + N a / b^2
    from scipy.special import polygamma
    N = G.size
    want = np.array([[-N * polygamma(1, 2.4), -N / 1.8], [-N / 1.8, -2 * G.sum() / 1.8 ** 3 + N * 2.4 / 1.8 ** 2]])
    assert np.max(np.abs(Hn - want)) <= 1e-6 * np.abs(want).max()
    cov = host.matrix_of(h.apop_model_numerical_covariance(d.ptr, m))
    assert np.max(np.abs(cov - np.linalg.inv(-want))) <= 1e-5 * np.abs(np.linalg.inv(-want)).max()
    assert h.apop_data_get_page(m.contents.parameters, b"<Covariance>")
    ab.unpin(d)
    # probit: one device information pass == -(numeric Hessian of the device score)
    X, y, beta = gen_probit_logit_sample(40000, 6, seed=1)
    dp = ab.pin(ab.Data(matrix=X, vector=y))
    mp = h.apop_model_copy(host.model_object(abi.PROBIT))
    abi.check(lib.apop_b200_register(mp, abi.PROBIT), "register")
    mp.contents.parameters = h.apop_data_alloc(6, 1, 0)
    for j in range(6):
        mp.contents.parameters.contents.matrix.contents.data[j] = beta[j]
    Hd = ab.information(ab.PROBIT, dp, beta)
    Hnum = -host.matrix_of(h.apop_model_hessian(dp.ptr, mp, 0.0))
    assert np.max(np.abs(Hd - Hnum)) <= 1e-6 * np.abs(Hd).max()
    ab.unpin(dp)


def test_newton_and_covariance_for_multinomial_logit(b200, oracle):
    """cfg2 in small: Newton on the P x P device information reaches the BFGS estimates, and
    apop_model_numerical_covariance is the inverse of one information pass"""
    from apophenia_b200 import abi, host
    ab = b200
    n, k, Cn = 60000, 6, 4
    X, y, B = gen_probit_logit_sample(n, k, seed=77, method="logit", ncat=Cn)
    raw = X.copy(); raw[:, 0] = y
    ests = {}
    for method in ("BFGS cg", "Newton"):
        dd = host.data_from_numpy(raw)
        est, params, rep = host.estimate(abi.LOGIT, dd, method=method, tolerance=1e-11, relative_tolerance=True, max_iterations=500)
        assert rep["status"] == 0, (method, rep)
        ests[method] = (est, params, dd, rep)
    assert np.max(np.abs(ests["Newton"][1] - ests["BFGS cg"][1])) <= 1e-8 * np.max(np.abs(ests["BFGS cg"][1]))
    assert ests["Newton"][3]["iterations"] <= 25
    est, params, dd, _ = ests["Newton"]
    h = host.lib()
    cov = host.matrix_of(h.apop_model_numerical_covariance(dd, est))
    Hw = oracle.logit_information(X, params.reshape(k, Cn - 1), Cn).astype(float)
    assert np.max(np.abs(cov - np.linalg.inv(Hw))) <= 1e-9 * np.abs(np.linalg.inv(Hw)).max()
    se = np.sqrt(np.diag(cov))
    assert np.all(np.abs(params - B.ravel()) < 6 * se)


# ------------------------------------------------------------------------------------------------
def test_probit_knife_edge_band(b200):
    """|x.beta| in 7.5 .. 8.5 with the UNLIKELY outcome: n = Phi(-xb) lies within a few ulp of 1 and the
    reference takes log(1 - n) of the rounded double (model/apop_probit.c:84-87), so the row term depends
    on how n rounds.  GSL forms n = 1 - tail (cdf/gauss.c), i.e. the exact value rounded to a double; any
    implementation of the same formula may land on a neighbouring double.  Per row the device's term must
    be log(1 - n_c) for n_c among the three doubles nearest the exact 1 - Phi(xb) -- the tolerance is that
    set, not a number -- with the 1e-10 clamp where n_c == 1; the fused LL and score are then the sums of
    exactly those row terms."""
    mp = pytest.importorskip("mpmath")
    ab = b200
    mp.mp.dps = 50
    xb = -np.concatenate([np.linspace(7.5, 8.5, 401), [8.2923, 8.2924, 8.2925, 8.21, 8.3]])
    n = xb.size
    X = np.stack([np.ones(n), xb], axis=1)
    y = np.ones(n)                                   # P(y = 1) = Phi(xb) ~ 1e-14 .. 1e-17: the unlikely outcome
    beta = np.array([0.0, 1.0])
    d = ab.pin(ab.Data(matrix=X, vector=y))
    rows = ab.map_rows(d, ab.ROW_PROBIT_LL, n, beta)
    chosen = np.zeros(n)
    for i in range(n):
        tail = mp.ncdf(mp.mpf(float(xb[i])))         # exact 1 - n
        n0 = float(1 - tail)                         # correctly rounded n
        cands = {n0, np.nextafter(n0, 0.0), min(np.nextafter(n0, 2.0), 1.0)}
        terms = []
        for nc in cands:
            ncl = nc if nc < 1.0 else 1.0 - 1e-10    # the reference's clamp on the rounded double
            terms.append((float(mp.log(mp.mpf(1.0 - ncl))), 1.0 - ncl))
        errs = [abs(rows[i] - t) / abs(t) for t, _ in terms]
        j = int(np.argmin(errs))
        assert errs[j] <= 1e-14, (float(xb[i]), rows[i], terms)
        chosen[i] = terms[j][1]
    ll, g = ab.ll_score(ab.PROBIT, d, beta)
    assert abs(ll - rows.astype(np.longdouble).sum()) <= 1e-13 * abs(ll)
    # score: d_i = phi(xb) / (1 - n_c) with the same n_c the LL term used
    dvals = np.array([float(mp.npdf(mp.mpf(float(v)))) for v in xb]) / chosen
    want = np.array([dvals.sum(), (dvals * xb).sum()])
    assert np.all(np.abs(g - want) <= 1e-12 * np.abs(want)), (g, want)
    ab.unpin(d)


# ------------------------------------------------------------------------------------------------
def test_residency_identity_and_counts(b200, oracle):
    """a registry entry must not outlive the buffers it was uploaded from"""
    ab = b200
    X, y, beta = gen_probit_logit_sample(6000, 4, seed=31)
    X2, y2, _ = gen_probit_logit_sample(5000, 4, seed=32)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    assert _rel(ab.ll_score(ab.PROBIT, d, beta)[0], oracle.probit_ll(X, y, beta, oracle.EXACT)) <= 1e-12
    ab.bootstrap_counts(d, 8, seed=1)
    # the same apop_data address now describes other buffers (what a free + malloc at the same address does)
    from apophenia_b200 import abi
    m2, a2 = abi.make_matrix(X2)
    v2, b2 = abi.make_vector(y2)
    d._m, d._v = m2, v2
    d._keep += [a2, b2]
    d.struct.matrix = C.pointer(m2)
    d.struct.vector = C.pointer(v2)
    got = ab.ll_score(ab.PROBIT, d, beta)[0]       # no invalidate: the signature check must catch it
    assert _rel(got, oracle.probit_ll(X2, y2, beta, oracle.EXACT)) <= 1e-12
    # the resample counts belonged to the old rows
    with pytest.raises(ab.B200Error):
        ab.ll_score_batched(ab.PROBIT, d, beta[None, :], use_counts=True)
    ab.unpin(d)
    # a Data that dies releases its resident copy
    lib = abi.load_library()
    tmp = ab.pin(ab.Data(matrix=X, vector=y))
    p = tmp.ptr
    assert lib.apop_b200_is_pinned(p)
    del tmp
    import gc
    gc.collect()
    assert not lib.apop_b200_is_pinned(p)


def test_device_recode_survives_reupload(b200, oracle):
    """prep_outcome(sync_host=0) recodes on the device only; a dirty re-upload from the raw host copy must
    shuffle and recode again"""
    ab = b200
    Xl, yl, B = gen_probit_logit_sample(9001, 6, seed=8, method="logit", ncat=4)
    vals = np.array([-2.0, 0.5, 3.0, 10.0])
    raw = Xl.copy(); raw[:, 0] = vals[yl.astype(int)]
    d = ab.Data(matrix=raw)
    cats, start = ab.prep_outcome(d)
    assert np.array_equal(cats, vals)
    m = ab.logit_model(B, d)
    want = float(oracle.logit_ll(Xl, yl, B, mode=oracle.EXACT))
    assert _rel(ab.log_likelihood(ab.LOGIT, d, m), want) <= 1e-12
    ab.invalidate(d)
    assert _rel(ab.log_likelihood(ab.LOGIT, d, m), want) <= 1e-12
    ab.unpin(d)


# ------------------------------------------------------------------------------------------------
def test_registered_window_upload_roundtrip(b200, monkeypatch):
    """the zero-copy upload (cudaHostRegister'ed windows of the caller's matrix, DMA straight out of it):
    2 MB windows over a 30 MB matrix whose 296-byte rows straddle every window boundary and whose first byte is
    not page aligned; outcome and weights through the small pinned staging; bit-exact round trip, and the same
    through the device prep with the outcome in column 0"""
    ab = b200
    rng = np.random.default_rng(41)
    n, k = 100_003, 37
    big = rng.standard_normal((n + 3, k))
    X = big[3:]                                   # contiguous rows, base pointer 888 bytes past the allocation
    y = rng.standard_normal(n)
    w = rng.uniform(0.5, 2.0, n)
    monkeypatch.setenv("APOP_B200_PIN_MODE", "register")
    monkeypatch.setenv("APOP_B200_PIN_WINDOW_MB", "2")
    d = ab.pin(ab.Data(matrix=X, vector=y, weights=w))
    assert ab.last_pin_mode()[0] == "registered"
    Xd = np.zeros((n, k)); yd = np.zeros(n); wd = np.zeros(n)
    from apophenia_b200 import abi
    dp = C.POINTER(C.c_double)
    abi.check(abi.load_library().apop_b200_download(d.ptr, Xd.ctypes.data_as(dp), yd.ctypes.data_as(dp), wd.ctypes.data_as(dp)), "download")
    assert np.array_equal(Xd, X) and np.array_equal(yd, y) and np.array_equal(wd, w)
    ab.unpin(d)
    # staged and registered uploads give bitwise the same resident data, hence the same results
    Xp, yp, beta = gen_probit_logit_sample(90_001, 33, seed=4)
    res = {}
    for mode in ("staged", "register"):
        monkeypatch.setenv("APOP_B200_PIN_MODE", mode)
        dd = ab.pin(ab.Data(matrix=Xp, vector=yp))
        res[mode] = ab.ll_score(ab.PROBIT, dd, beta)
        ab.unpin(dd)
    assert res["staged"][0] == res["register"][0] and np.array_equal(res["staged"][1], res["register"][1])
    # outcome in column 0: the tiling kernel shuffles it out on the registered path as well
    raw = Xp[:, :20].copy(); raw[:, 0] = yp
    monkeypatch.setenv("APOP_B200_PIN_MODE", "register")
    dr = ab.Data(matrix=raw)
    cats, start = ab.prep_outcome(dr)
    assert ab.last_pin_mode()[0] == "registered" and np.array_equal(cats, [0.0, 1.0])
    ll_r = ab.log_likelihood(ab.PROBIT, dr, ab.probit_model(beta[:20], dr))
    X20 = Xp[:, :20].copy(); X20[:, 0] = 1.0
    ds = ab.pin(ab.Data(matrix=X20, vector=yp))
    assert ll_r == ab.log_likelihood(ab.PROBIT, ds, ab.probit_model(beta[:20], ds))
    ab.unpin(dr); ab.unpin(ds)


def test_page_locked_source_goes_zero_copy(b200, oracle):
    """a matrix allocated with apop_b200_host_alloc is read in place by the DMA engine (mode 2), chosen without any
    switch; results are those of the staged upload of the same numbers"""
    ab = b200
    X, y, beta = gen_probit_logit_sample(70_001, 30, seed=14)     # 16.8 MB: above the 8 MB threshold
    Xp = ab.host_empty(X.shape, pinned=True)
    Xp[:] = X
    d = ab.pin(ab.Data(matrix=Xp, vector=y))
    assert ab.last_pin_mode()[0].startswith("zero-copy")
    ll, g = ab.ll_score(ab.PROBIT, d, beta)
    ab.unpin(d)
    d2 = ab.pin(ab.Data(matrix=X, vector=y))
    assert ab.last_pin_mode()[0] == "staged"
    ll2, g2 = ab.ll_score(ab.PROBIT, d2, beta)
    ab.unpin(d2)
    assert ll == ll2 and np.array_equal(g, g2)
    assert _rel(ll, oracle.probit_ll(X, y, beta, oracle.EXACT)) <= 1e-12


# ------------------------------------------------------------------------------------------------
def test_batched_shared_centre_expansion_matches_exact(b200, oracle, monkeypatch):
    """the batched kernel's Taylor form of the link around x . mean(B): close parameter vectors take the Horner path
    (exact share 0), far ones the exact evaluation -- both must match the oracle replicate by replicate at 1e-12"""
    ab = b200
    n, k, m = 50_003, 13, 24
    rng = np.random.default_rng(17)
    for fam, method in ((ab.PROBIT, "probit"), (ab.LOGIT, "logit")):
        X, y, beta = gen_probit_logit_sample(n, k, seed=29, method=method)
        d = ab.pin(ab.Data(matrix=X, vector=y))
        ab.bootstrap_counts(d, m, seed=2)
        c = ab.counts_download(d, m, n)
        for spread, want_exact in ((1e-3, 0.0), (0.3, 1.0)):
            Bs = beta[None, :] + spread * rng.standard_normal((m, k))
            for use_counts in (False, True):
                llb, gb = ab.ll_score_batched(fam, d, Bs, use_counts=use_counts)
                frac = ab.batched_exact_fraction(d)
                assert (frac == 0.0) if want_exact == 0.0 else (frac > 0.5 or frac == 0.0), (spread, frac)   # 0.0 again once the exact kernel took over
                for r in (0, 11, m - 1):
                    if use_counts:
                        idx = np.repeat(np.arange(n), c[:, r]); Xr, yr = X[idx], y[idx]
                    else:
                        Xr, yr = X, y
                    if fam == ab.PROBIT:
                        wl = oracle.probit_ll(Xr, yr, Bs[r], oracle.EXACT); wg, ws = oracle.probit_score(Xr, yr, Bs[r], oracle.EXACT)
                    else:
                        wl = oracle.logit_ll(Xr, yr, Bs[r].reshape(-1, 1), mode=oracle.EXACT)
                        wg, ws = oracle.logit_score(Xr, yr, Bs[r].reshape(-1, 1), mode=oracle.EXACT); wg, ws = wg.ravel(), ws.ravel()
                    assert _rel(llb[r], wl) <= 1e-12
                    assert np.all(np.abs(gb[r] - wg.astype(float)) <= 1e-12 * ws.astype(float))
        ab.unpin(d)


def test_batched_poisson_regression_with_resample_counts(b200, oracle):
    """resample counts weight the lgamma(y+1) constant per replicate: carried per row by the batched kernels"""
    ab = b200
    n, k, m = 30_011, 9, 16
    rng = np.random.default_rng(23)
    X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
    beta = rng.uniform(-1, 1, k) / 3
    y = rng.poisson(np.exp(X @ beta)).astype(float)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    ab.bootstrap_counts(d, m, seed=4)
    c = ab.counts_download(d, m, n)
    for spread in (1e-3, 0.2):
        Bs = beta[None, :] + spread * rng.standard_normal((m, k))
        llb, gb = ab.ll_score_batched(ab.POISSON_REG, d, Bs, use_counts=True)
        for r in (0, 7, m - 1):
            idx = np.repeat(np.arange(n), c[:, r])
            wl = oracle.poisson_reg_ll(X[idx], y[idx], Bs[r], oracle.EXACT)
            wg, ws = oracle.poisson_reg_score(X[idx], y[idx], Bs[r], oracle.EXACT)
            assert _rel(llb[r], wl) <= 1e-12
            assert np.all(np.abs(gb[r] - wg.astype(float)) <= 1e-12 * ws.astype(float))
        ll1, g1 = ab.ll_score_batched(ab.POISSON_REG, d, Bs[:3], use_counts=False)
        assert _rel(ll1[1], oracle.poisson_reg_ll(X, y, Bs[1], oracle.EXACT)) <= 1e-12
    ab.unpin(d)


def test_multinomial_logit_replicate_weighted_pass_and_bootstrap(b200, oracle):
    """apop_b200_set_replicate: LL and score of the multinomial logit over the resident data weighted by one bootstrap
    replicate's counts == the oracle on the physically resampled rows; apop_bootstrap_cov of a multinomial model loops
    the replicates over that (no row copies) and lands near the inverse observed information"""
    from apophenia_b200 import abi, host
    ab = b200
    n, k, Cn, m = 40_003, 7, 4, 40
    X, y, B = gen_probit_logit_sample(n, k, seed=61, method="logit", ncat=Cn)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    ab.bootstrap_counts(d, m, seed=9)
    c = ab.counts_download(d, m, n)
    base = ab.ll_score(ab.LOGIT, d, B.ravel())
    for r in (0, 17, m - 1):
        ab.set_replicate(d, r)
        ll, g = ab.ll_score(ab.LOGIT, d, B.ravel())
        idx = np.repeat(np.arange(n), c[:, r])
        wl = oracle.logit_ll(X[idx], y[idx], B, mode=oracle.EXACT)
        wg, ws = oracle.logit_score(X[idx], y[idx], B, mode=oracle.EXACT)
        assert _rel(ll, wl) <= 1e-12
        assert np.all(np.abs(g - wg.astype(float).ravel()) <= 1e-12 * ws.astype(float).ravel())
    ab.set_replicate(d, -1)
    again = ab.ll_score(ab.LOGIT, d, B.ravel())
    assert again[0] == base[0] and np.array_equal(again[1], base[1])
    with pytest.raises(ab.B200Error):
        ab.set_replicate(d, m + 8)
    ab.unpin(d)
    # the driver: fit, then bootstrap covariance through the replicate loop
    raw = X.copy(); raw[:, 0] = y
    dd = host.data_from_numpy(raw)
    est, params, rep = host.estimate(abi.LOGIT, dd, method="BFGS cg", tolerance=1e-10, relative_tolerance=True, max_iterations=500)
    assert rep["status"] == 0
    cov, boots = host.bootstrap_cov(dd, est, iterations=60, seed=3)
    Hw = oracle.logit_information(X, params.reshape(k, Cn - 1), Cn).astype(float)
    want = np.diag(np.linalg.inv(Hw))
    assert boots.shape == (60, k * (Cn - 1))
    assert np.all(np.abs(boots.mean(0) - params) < 5 * np.sqrt(want / 60) + 0.2 * np.sqrt(want))
    ratio = np.diag(cov) / want
    assert 0.4 < ratio.min() and ratio.max() < 2.2, ratio        # 60 replicates: chi-square spread around 1


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k,C", [(4097, 18, 3), (30011, 21, 5), (33, 24, 8), (2049, 17, 2), (5000, 32, 4), (4097, 9, 6)])
def test_multinomial_logit_single_tile_per_warp(b200, oracle, n, k, C):
    """Few tiles on a full grid: every warp sees at most one tile, so the second ring stage of logit_hybrid_kernel is
    never filled.  A register block that overran KB = 24 read that stage (times B's zero padding) and returned NaN
    whenever it held a NaN bit pattern; other kernels run first here so that shared memory is not clean."""
    ab = b200
    X, y, B = gen_probit_logit_sample(n, k, seed=n + k, method="logit", ncat=C)
    nj = 32 * 148 * 32 * 2 + 7                                                  # two tiles per warp: both stages
    junk = ab.pin(ab.Data(matrix=np.full((nj, 32), np.nan), vector=np.arange(nj) % 3.0))
    try:
        ab.ll_score(ab.LOGIT, junk, np.ones(32 * 2))                               # NaN rows through the ring
    except Exception:
        pass
    ab.unpin(junk)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    ll, g = ab.ll_score(ab.LOGIT, d, B.ravel())
    assert _rel(ll, oracle.logit_ll(X, y, B, mode=oracle.EXACT)) <= 1e-12
    wg, ws = oracle.logit_score(X, y, B, mode=oracle.EXACT)
    assert np.all(np.abs(g - wg.astype(float).ravel()) <= 1e-12 * ws.astype(float).ravel())
    ab.unpin(d)
