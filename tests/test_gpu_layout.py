"""The resident copy changes layout between the multinomial logit kernels (swizzled tiles: rows r ^ 8 in odd
columns, so that one bulk copy per tile lands bank-conflict free in shared memory) and every other kernel (plain
tiles).  Each result below is checked against the oracle after the data set has been through the other layout, so a
kernel that read the wrong form -- or a conversion that is not an involution -- shows as a parity failure."""
import numpy as np
import pytest

from conftest import gen_probit_logit_sample

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return abs(float(a) - float(b)) / max(abs(float(b)), 1e-300)


def _check_logit(ab, oracle, d, X, y, B):
    m = ab.logit_model(B, d)
    assert _rel(ab.log_likelihood(ab.LOGIT, d, m), oracle.logit_ll(X, y, B, mode=oracle.EXACT)) <= 1e-12
    ge, sc = oracle.logit_score(X, y, B, mode=oracle.EXACT)
    g = ab.score(ab.LOGIT, d, m, size=B.size)
    assert np.all(np.abs(g - ge.astype(float).ravel()) <= 1e-12 * sc.astype(float).ravel())


@pytest.mark.parametrize("n,k,C,weights", [(20003, 32, 8, False), (4097, 18, 4, False), (30011, 31, 8, True),
                                           (999, 7, 3, True), (31, 8, 3, False), (33, 1, 3, False)])
def test_layout_round_trips_between_kernel_families(b200, oracle, n, k, C, weights):
    ab = b200
    X, y, B = gen_probit_logit_sample(n, k, seed=11 + k, method="logit", ncat=C)
    w = np.random.default_rng(5).uniform(0.5, 2.0, n) if weights else None
    d = ab.pin(ab.Data(matrix=X, vector=y, weights=w))
    # 1. multinomial pass (converts to the swizzled form), twice with different parameters
    _check_logit(ab, oracle, d, X, y, B)
    _check_logit(ab, oracle, d, X, y, 0.5 * B)
    # 2. kernels of the plain layout on the same resident copy: column sums, X.B, the row map, a cell sum
    xb = ab.dot(d, B, n)
    assert np.max(np.abs(xb - X @ B)) <= 1e-12 * max(1.0, np.abs(X @ B).max())
    tot = ab.map_sum(d, ab.MAP_IDENTITY, part="m")
    assert _rel(tot, float(np.sum(X.astype(np.longdouble)))) <= 1e-12 or abs(tot - X.sum()) < 1e-9
    # 3. the information kernel stages either form; take it from the plain form first, then after a logit pass
    H0 = ab.information(ab.LOGIT, d, B.ravel())
    Ho = oracle.logit_information(X, B, C).astype(float)
    assert np.max(np.abs(H0 - Ho)) <= 8e-12 * np.abs(Ho).max()
    _check_logit(ab, oracle, d, X, y, B)
    H1 = ab.information(ab.LOGIT, d, B.ravel())
    assert np.array_equal(H0, H1)
    # 4. back through the plain layout, and the logit once more
    xb2 = ab.dot(d, B, n)
    assert np.array_equal(xb, xb2)
    _check_logit(ab, oracle, d, X, y, B)
    ab.unpin(d)


def test_download_returns_plain_rows_after_a_logit_pass(b200, oracle):
    ab = b200
    k, C, n = 13, 5, 10007
    B = np.random.default_rng(3).uniform(-1, 1, (k, C - 1))
    s = ab.synth(ab.LOGIT, n, k, B.ravel(), ncat=C)
    X0, y0 = s.download()
    ll, g = ab.ll_score(ab.LOGIT, s, B.ravel())
    assert _rel(ll, oracle.logit_ll(X0, y0, B, mode=oracle.EXACT)) <= 1e-12
    X1, y1 = s.download()
    assert np.array_equal(X0, X1) and np.array_equal(y0, y1)
    ll2, g2 = ab.ll_score(ab.LOGIT, s, B.ravel() * 1.0000001)
    ll3, g3 = ab.ll_score(ab.LOGIT, s, B.ravel())
    assert ll3 == ll and np.array_equal(g, g3)
    s.free()


def test_replicate_weighted_multinomial_pass_in_the_swizzled_layout(b200, oracle):
    ab = b200
    n, k, C = 6007, 10, 4
    X, y, B = gen_probit_logit_sample(n, k, seed=9, method="logit", ncat=C)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    ab.bootstrap_counts(d, 4, seed=5)
    c = ab.counts_download(d, 4, n)
    for r in (0, 3):
        ab.set_replicate(d, r)
        idx = np.repeat(np.arange(n), c[:, r])
        m = ab.logit_model(B, d)
        assert _rel(ab.log_likelihood(ab.LOGIT, d, m), oracle.logit_ll(X[idx], y[idx], B, mode=oracle.EXACT)) <= 1e-12
    ab.set_replicate(d, -1)
    ab.unpin(d)
