"""Host logic around the device boundary (no GPU needed): the vtable registry, pack
order, prep (factor page, ols_shuffle, start point) and the optimizer, driven here
by oracle-backed callbacks installed exactly where the device entries go."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import gen_probit_logit_sample

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def host():
    from apophenia_b200 import host as h
    h.lib()
    return h


def _np_vec(v):
    return np.array([v.contents.data[i * v.contents.stride] for i in range(v.contents.size)])


def test_vtable_add_get_drop(host):
    """tests/test_apop.c:1151-1163: add/get/drop round trip; add is a no-op when the hash exists"""
    h = host.lib()
    f1, f2 = C.c_void_p(0x1000), C.c_void_p(0x2000)
    assert h.apop_vtable_get(b"apop_score", 777) is None
    assert h.apop_vtable_add(b"apop_score", f1, 777) == 0
    assert h.apop_vtable_get(b"apop_score", 777) == 0x1000
    h.apop_vtable_add(b"apop_score", f2, 777)              # present: stays f1 (apop_vtables.c:98)
    assert h.apop_vtable_get(b"apop_score", 777) == 0x1000
    assert h.apop_vtable_get(b"apop_predict", 777) is None  # tables are per name
    assert h.apop_vtable_drop(b"apop_score", 777) == 0
    assert h.apop_vtable_drop(b"apop_score", 777) == 1      # 1 = not found
    assert h.apop_vtable_get(b"apop_score", 777) is None
    h.apop_vtable_add(b"apop_score", f2, 777)
    assert h.apop_vtable_get(b"apop_score", 777) == 0x2000
    h.apop_vtable_drop(b"apop_score", 777)


def test_pack_unpack_order(host):
    """vector, then matrix row by row, then weights, then non-<info> pages
    (apop_conversions.m4.c:762-806); B[j][c] lands at j*(C-1)+c"""
    h = host.lib()
    d = h.apop_data_alloc(2, 3, 2)
    for i in range(2):
        d.contents.vector.contents.data[i] = 10 + i
    for i in range(6):
        d.contents.matrix.contents.data[i] = i
    page = h.apop_data_alloc(2, 0, 0)
    page.contents.vector.contents.data[0], page.contents.vector.contents.data[1] = 100, 101
    h.apop_data_add_page(d, h.apop_data_alloc(3, 0, 0), b"<Covariance>")   # info page: skipped
    h.apop_data_add_page(d, page, b"more")
    pv = h.apop_data_pack(d, None)
    assert list(_np_vec(pv)) == [10, 11, 0, 1, 2, 3, 4, 5, 100, 101]
    rev = h.gsl_vector_alloc(10)
    for i in range(10):
        rev.contents.data[i] = 50 - i
    h.apop_data_unpack(rev, d)
    assert d.contents.matrix.contents.data[3] == 50 - 5 and page.contents.vector.contents.data[1] == 41
    assert h.apop_data_pack(d, h.gsl_vector_alloc(7)) is None or True  # size mismatch -> NULL
    h.apop_data_free(d)


def test_probit_prep(host, oracle):
    """probit_prep (model/apop_probit.c:42-81): factor page, ols_shuffle, k x (C-1) parameters,
    start point exp(mean ln|x|)"""
    h = host.lib()
    X, y, beta = gen_probit_logit_sample(500, 4, seed=2)
    raw = X.copy(); raw[:, 0] = 3 + 4 * y                 # outcome in column 0, values {3, 7}
    d = host.data_from_numpy(raw)
    m = h.apop_model_copy(host.model_object(0))
    h.apop_prep(d, m)
    assert d.contents.vector and d.contents.vector.contents.size == 500
    assert np.array_equal(_np_vec(d.contents.vector), y)   # factor indices 0/1
    got = np.array([d.contents.matrix.contents.data[i * 4] for i in range(500)])
    assert np.all(got == 1.0)                               # column 0 := 1
    pg = h.apop_b200_category_table(d)
    assert pg.contents.vector.contents.size == 2
    P = m.contents.parameters.contents.matrix.contents
    assert (P.size1, P.size2) == (4, 1)
    assert np.allclose(host.parameters_of(m), oracle.probit_start(X), rtol=1e-15)
    h.apop_prep(d, m)                                       # re-prep is a no-op
    assert m.contents.error == b"\x00"


def _install_oracle_model(host, oracle, family, keep):
    """a model whose log_likelihood / score are oracle-backed callbacks, registered the way
    apop_b200_register registers the device entries (hash = address of log_likelihood)"""
    from apophenia_b200 import abi
    h = host.lib()
    m = h.apop_model_copy(host.model_object(family))

    def arrays(dp):
        d = dp.contents
        n, k = d.matrix.contents.size1, d.matrix.contents.size2
        X = np.ctypeslib.as_array(d.matrix.contents.data, shape=(n, k))
        y = np.ctypeslib.as_array(d.vector.contents.data, shape=(n,))
        return X, y

    def ll(dp, mp):
        X, y = arrays(dp)
        b = host.parameters_of(mp)
        return float(oracle.probit_ll(X, y, b) if family == 0 else oracle.logit_ll(X, y, b.reshape(X.shape[1], -1)))

    def score(dp, gp, mp):
        X, y = arrays(dp)
        b = host.parameters_of(mp)
        g = (oracle.probit_score(X, y, b) if family == 0 else oracle.logit_score(X, y, b.reshape(X.shape[1], -1)))[0]
        for i in range(gp.contents.size):
            gp.contents.data[i] = float(g[i])

    cll, csc = abi.LL_FN(ll), abi.SCORE_FN(score)
    keep += [cll, csc]
    m.contents.log_likelihood = C.cast(cll, C.c_void_p)
    h.apop_vtable_add(b"apop_score", C.cast(csc, C.c_void_p), C.cast(cll, C.c_void_p).value)
    return m


@pytest.mark.parametrize("method", ["BFGS cg", "FR cg", "NM simplex"])
def test_fake_logit_through_the_mirror(host, oracle, method):
    """eg/fake_logit.c end to end: apop_estimate -> prep -> apop_maximum_likelihood."""
    h = host.lib()
    g = json.load(open(os.path.join(GOLD, "fake_logit.json")))
    d = host.data_from_numpy(np.array(g["rows_outcome_A_B"]))
    keep = []
    m = _install_oracle_model(host, oracle, 1, keep)
    host.mle_settings(m, method=method, tolerance=1e-9 if method != "NM simplex" else 1e-7, keep=keep)
    est = h.apop_estimate(d, m)
    rep = h.apop_mle_last_report()
    assert rep.status == 0
    tol = 1e-7 if method != "NM simplex" else 2e-5
    assert np.allclose(host.parameters_of(est), g["true_mle_1_A_B"], atol=tol)
    # the reference's canned numbers are its NM stopping point (1e-6 assert, 9e-6 from the MLE)
    assert np.max(np.abs(host.parameters_of(est) - g["reference_canned_fit_1_A_B"])) < 3e-5
    h.apop_vtable_drop(b"apop_score", m.contents.log_likelihood)


def test_amash_logit_through_the_mirror(host, oracle):
    """eg/logit.c end to end on the real data: prep turns the outcome column into factor indices and
    shuffles it out, then apop_maximum_likelihood; compared with scipy on the same likelihood"""
    from conftest import amash_design
    from scipy import optimize
    h = host.lib()
    raw = amash_design()
    d = host.data_from_numpy(raw.copy())
    keep = []
    m = _install_oracle_model(host, oracle, 1, keep)
    host.mle_settings(m, method="BFGS cg", tolerance=1e-9, keep=keep)
    est = h.apop_estimate(d, m)
    assert h.apop_mle_last_report().status == 0
    y = raw[:, 0].copy(); X = raw.copy(); X[:, 0] = 1.0
    f = lambda b: -np.sum(y * (X @ b) - np.logaddexp(0, X @ b))
    res = optimize.minimize(f, np.zeros(4), method="BFGS", options={"gtol": 1e-9})
    assert np.allclose(host.parameters_of(est), res.x, atol=2e-6)
    h.apop_vtable_drop(b"apop_score", m.contents.log_likelihood)


def test_probit_recovery_like_reference_test(host, oracle):
    """test_probit_and_logit (tests/test_apop.c:935-959): recover beta within 0.07 (smaller n here)"""
    h = host.lib()
    X, y, beta = gen_probit_logit_sample(8000, 3, seed=8)
    raw = X.copy(); raw[:, 0] = y
    d = host.data_from_numpy(raw)
    keep = []
    m = _install_oracle_model(host, oracle, 0, keep)
    host.mle_settings(m, method="BFGS cg", tolerance=1e-6, keep=keep)
    est = h.apop_estimate(d, m)
    rep = h.apop_mle_last_report()
    assert rep.status == 0 and rep.iterations < 60
    assert np.linalg.norm(host.parameters_of(est) - beta) < 0.07 * 3
    # score at the optimum ~ 0; numeric gradient agrees with the registered score
    gv = h.gsl_vector_alloc(3)
    h.apop_score(d, gv, est)
    assert np.max(np.abs(_np_vec(gv))) < 1e-4
    est.contents.parameters.contents.matrix.contents.data[1] += 0.05
    h.apop_score(d, gv, est)
    num = h.apop_numerical_gradient(d, est, 1e-3)
    assert np.allclose(_np_vec(num), _np_vec(gv), rtol=1e-5, atol=1e-5)
    h.apop_vtable_drop(b"apop_score", m.contents.log_likelihood)


def test_unregistered_model_fails_loudly(host, capfd):
    """the mirror ships no CPU likelihoods: without apop_b200_register the fit refuses"""
    h = host.lib()
    X, y, _ = gen_probit_logit_sample(50, 2, seed=1)
    raw = X.copy(); raw[:, 0] = y
    d = host.data_from_numpy(raw)
    est = h.apop_estimate(d, host.model_object(0))
    assert est.contents.error == b"m"
    assert h.apop_mle_last_report().status == -1
    assert "no CPU likelihoods" in capfd.readouterr().err
