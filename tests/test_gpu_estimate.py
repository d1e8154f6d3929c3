"""apop_estimate through the reference's own dispatch with the device entries registered:
fitted parameters agree to 1e-8 relative with the same optimizer driven by the oracle
(BASELINE.json north_star; SURVEY F9: compare two tight runs of one optimizer)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import gen_probit_logit_sample
from test_host_mirror import _install_oracle_model, _np_vec

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def host():
    from apophenia_b200 import host as h
    h.lib()
    return h


def _fit_oracle(host, oracle, family, raw, **settings):
    h = host.lib()
    d = host.data_from_numpy(raw)
    keep = []
    m = _install_oracle_model(host, oracle, family, keep)
    host.mle_settings(m, keep=keep, **settings)
    est = h.apop_estimate(d, m)
    rep = h.apop_mle_last_report()
    out = host.parameters_of(est), rep.status, rep.iterations
    h.apop_vtable_drop(b"apop_score", m.contents.log_likelihood)
    return out


@pytest.mark.parametrize("method", ["BFGS cg", "FR cg"])
def test_probit_fit_matches_oracle_driven_fit(b200, host, oracle, method):
    X, y, beta = gen_probit_logit_sample(60000, 6, seed=8)
    raw = X.copy(); raw[:, 0] = y
    settings = dict(method=method, tolerance=1e-7, max_iterations=500)
    est, params, rep = host.estimate(b200.PROBIT, host.data_from_numpy(raw), **settings)
    assert rep["status"] == 0
    want, st, it = _fit_oracle(host, oracle, 0, raw, **settings)
    assert st == 0
    assert np.max(np.abs(params - want) / np.abs(want)) < 1e-8, (params, want)
    assert np.linalg.norm(params - beta) < 0.07          # the reference's own recovery test
    assert est.contents.info.contents.vector.contents.data[0] == 0


def test_logit_fit_multinomial_and_newton(b200, host, oracle):
    X, y, B = gen_probit_logit_sample(30000, 4, seed=3, method="logit", ncat=3)
    raw = X.copy(); raw[:, 0] = y
    settings = dict(method="BFGS cg", tolerance=1e-7, max_iterations=500)
    est, params, rep = host.estimate(b200.LOGIT, host.data_from_numpy(raw), **settings)
    assert rep["status"] == 0
    want, st, it = _fit_oracle(host, oracle, 1, raw, **settings)
    assert st == 0 and np.max(np.abs(params - want) / np.maximum(np.abs(want), 1e-3)) < 1e-7
    assert np.linalg.norm(params.reshape(4, 2) - B) < 0.2


def test_register_installs_and_restores(b200, host):
    from apophenia_b200 import abi
    h, lib = host.lib(), abi.load_library()
    m = h.apop_model_copy(host.model_object(0))
    assert not m.contents.log_likelihood
    assert lib.apop_b200_register(m, 0) == 0
    want = C.cast(lib.apop_b200_probit_ll, C.c_void_p).value
    assert m.contents.log_likelihood == want
    # the score is in the vtable under the new hash, and a second add is a no-op
    assert h.apop_vtable_get(b"apop_score", want) == C.cast(lib.apop_b200_probit_score, C.c_void_p).value
    h.apop_vtable_add(b"apop_score", C.c_void_p(0x1234), want)
    assert h.apop_vtable_get(b"apop_score", want) != 0x1234
    copy = h.apop_model_copy(m)                           # memcpy keeps fn pointers, hence the hash
    assert copy.contents.log_likelihood == want
    assert lib.apop_b200_unregister(m) == 0
    assert not m.contents.log_likelihood
    assert h.apop_vtable_get(b"apop_score", want) is None


def test_metropolis_style_ll_only_calls(b200, oracle):
    """apop_model_metropolis asks for one LL per step and never the score (apop_mcmc.m4.c:113)"""
    ab = b200
    X, y, beta = gen_probit_logit_sample(20000, 5, seed=7, method="logit")
    d = ab.pin(ab.Data(matrix=X, vector=y))
    rng = np.random.default_rng(0)
    b = beta.copy()
    for _ in range(5):
        prop = b + 0.01 * rng.standard_normal(5)
        ll = ab.log_likelihood(ab.LOGIT, d, ab.logit_model(prop, d))
        want = float(oracle.logit_ll(X, y, prop, mode=oracle.EXACT))
        assert abs(ll - want) <= 1e-12 * abs(want)
        b = prop
    ab.unpin(d)


def test_config1_shape_same_optimizer_two_back_ends(b200, host, oracle):
    """BASELINE config 1 in small: apop_estimate(probit, BFGS) by the SAME host optimizer over the
    restated reference CPU path (C entries, oracle/cpu_models.c) and over the device entries:
    same iteration count, parameters to 1e-8 relative"""
    from apophenia_b200 import abi
    h = host.lib()
    om = oracle.models_lib()
    X, y, beta = gen_probit_logit_sample(200000, 10, seed=8)
    raw = X.copy(); raw[:, 0] = y
    settings = dict(method="BFGS cg", tolerance=1e-9, relative_tolerance=True, max_iterations=500)
    d = host.data_from_numpy(raw)
    m = h.apop_model_copy(host.model_object(abi.PROBIT))
    ll_addr = C.cast(om.oracle_model_probit_ll, C.c_void_p).value
    m.contents.log_likelihood = ll_addr
    h.apop_vtable_add(b"apop_score", C.cast(om.oracle_model_probit_score, C.c_void_p), ll_addr)
    keep = []
    host.mle_settings(m, keep=keep, **settings)
    est_c = h.apop_estimate(d, m)
    rep_c = h.apop_mle_last_report()
    p_cpu = host.parameters_of(est_c)
    h.apop_vtable_drop(b"apop_score", ll_addr)
    est_g, p_gpu, rep_g = host.estimate(abi.PROBIT, host.data_from_numpy(raw), **settings)
    assert rep_c.status == 0 and rep_g["status"] == 0
    assert abs(rep_c.iterations - rep_g["iterations"]) <= 2
    assert np.max(np.abs(p_gpu - p_cpu) / np.abs(p_cpu)) < 1e-8


def test_device_prep_matches_host_prep(b200, oracle, host):
    """ols_shuffle, the category table, the recoding and the start point on the
    device (apop_b200_prep_outcome) against the host mirror's probit_prep, and multilogit_expected"""
    ab = b200
    from apophenia_b200 import abi
    X, y, B = gen_probit_logit_sample(30011, 6, seed=21, method="logit", ncat=4)
    cats_true = np.array([-3.0, 2.0, 5.0, 9.5])
    raw = X.copy(); raw[:, 0] = cats_true[y.astype(int)]       # outcome in column 0, as a reference user holds it
    raw[7, 0] = -3.0; raw[8, 0] = 9.5
    d = ab.Data(matrix=raw.copy())
    cats, start = ab.prep_outcome(d, sync_host=False)
    assert np.array_equal(cats, cats_true)
    Xs = raw.copy(); Xs[:, 0] = 1.0
    want_start = np.exp(np.where(Xs != 0, np.log(np.abs(np.where(Xs != 0, Xs, 1.0))), 0.0).mean(0))
    assert np.allclose(start, want_start, rtol=1e-13)
    yi = np.searchsorted(cats_true, raw[:, 0]).astype(float)
    m = ab.logit_model(B, d)
    ll = ab.log_likelihood(ab.LOGIT, d, m)
    want = oracle.logit_ll(Xs, yi, B, mode=oracle.EXACT)
    assert abs(ll - float(want)) <= 1e-12 * abs(float(want))
    assert np.array_equal(d.matrix, raw)                        # sync_host=False leaves the caller's data alone
    probs, best = ab.logit_expected(d, m, raw.shape[0], 4)
    xb = np.concatenate([np.zeros((raw.shape[0], 1)), Xs @ B], axis=1)
    pw = np.exp(xb - xb.max(1, keepdims=True)); pw /= pw.sum(1, keepdims=True)
    assert np.allclose(probs, pw, rtol=1e-12, atol=1e-15) and np.array_equal(best, pw.argmax(1))
    ab.unpin(d)
    # sync_host=True reproduces the reference's in-place mutation
    d2 = ab.Data(matrix=raw.copy())
    ab.prep_outcome(d2, sync_host=True)
    v = d2.struct.vector.contents
    got_y = np.ctypeslib.as_array(v.data, shape=(v.size,))
    assert np.array_equal(got_y, yi) and np.all(d2.matrix[:, 0] == 1.0) and np.array_equal(d2.matrix[:, 1:], raw[:, 1:])
    ab.unpin(d2)
    # apop_estimate through the mirror: device prep and host prep give the same fit
    settings = dict(method="Newton", tolerance=1e-10, relative_tolerance=True, max_iterations=100)
    Xp, yp, bp = gen_probit_logit_sample(40000, 5, seed=22)
    rawp = Xp.copy(); rawp[:, 0] = yp
    fits = []
    for env in (None, "1"):
        if env: os.environ["APOP_B200_HOST_PREP"] = env
        try:
            dd = host.data_from_numpy(rawp.copy())
            est, params, rep = host.estimate(abi.PROBIT, dd, **settings)
            fits.append(params)
        finally:
            os.environ.pop("APOP_B200_HOST_PREP", None)
    assert np.allclose(fits[0], fits[1], rtol=1e-12, atol=0)


def test_amash_logit_real_data_on_device(b200, oracle, host):
    """eg/logit.c on the real 422-row data set: the device-driven fit (device prep included) against
    the oracle-driven fit of the same optimizer, and LL / score / information at the optimum"""
    from apophenia_b200 import abi
    from conftest import amash_design
    ab = b200
    h = host.lib()
    raw = amash_design()
    # (the reference's start point exp(mean ln|x|) saturates every probability on this data -- log
    # contributions are ~9 -- so the same line-search optimizer runs on both sides)
    settings = dict(method="BFGS cg", tolerance=1e-10)
    est, p_dev, rep = host.estimate(abi.LOGIT, host.data_from_numpy(raw.copy()), **settings)
    assert rep["status"] == 0
    # Newton from the same saturated start: the information matrix vanishes there, the first steps fall back to descent
    est_n, p_newton, rep_n = host.estimate(abi.LOGIT, host.data_from_numpy(raw.copy()), method="Newton", tolerance=1e-12,
                                           relative_tolerance=True, max_iterations=200)
    assert rep_n["status"] == 0 and np.max(np.abs(p_newton - p_dev)) < 1e-6
    keep = []
    m = _install_oracle_model(host, oracle, 1, keep)
    host.mle_settings(m, method="BFGS cg", tolerance=1e-10, keep=keep)
    est_o = h.apop_estimate(host.data_from_numpy(raw.copy()), m)
    p_orc = host.parameters_of(est_o)
    h.apop_vtable_drop(b"apop_score", m.contents.log_likelihood)
    assert np.max(np.abs(p_dev - p_orc) / np.abs(p_orc)) < 1e-8
    y = raw[:, 0].copy(); X = raw.copy(); X[:, 0] = 1.0
    d = ab.pin(ab.Data(matrix=X, vector=y))
    ll, g = ab.ll_score(ab.LOGIT, d, p_dev)
    want = float(oracle.logit_ll(X, y, p_dev.reshape(4, 1), mode=oracle.EXACT))
    assert abs(ll - want) <= 1e-12 * abs(want) and np.max(np.abs(g)) < 1e-6
    H = ab.information(ab.LOGIT, d, p_dev)
    assert np.allclose(H, oracle.information(1, X, y, p_dev).astype(float), rtol=1e-11)
    ab.unpin(d)


def test_entries_are_callable_from_any_host_thread(b200, oracle):
    """SURVEY 8(b) threading: the reference's callers may sit in OpenMP regions; every entry takes the
    library mutex and sets the device itself, so concurrent calls from several host threads on different
    data sets must return exactly what serial calls return"""
    import threading
    ab = b200
    sets = []
    for seed, (n, k) in enumerate(((30011, 7), (20000, 12), (9999, 32))):
        X, y, beta = gen_probit_logit_sample(n, k, seed=40 + seed)
        d = ab.pin(ab.Data(matrix=X, vector=y))
        sets.append((d, beta, ab.ll_score(ab.PROBIT, d, beta)))
    errors = []

    def worker(tid):
        try:
            for i in range(25):
                d, beta, (ll0, g0) = sets[(tid + i) % len(sets)]
                ll, g = ab.ll_score(ab.PROBIT, d, beta)
                if ll != ll0 or not np.array_equal(g, g0):
                    errors.append((tid, i))
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for t in threads: t.start()
    for t in threads: t.join()
    assert not errors, errors[:3]
    for d, _, _ in sets:
        ab.unpin(d)
