"""Accuracy of the branch-free FP64 special functions (apophenia_b200/csrc/fastmath.cuh),
compiled for the host and checked against mpmath at 50 digits, and of the probit / logit
row evaluators built on them against the oracle.  The device compiles the same source."""
import ctypes as C
import os
import subprocess
import tempfile

import mpmath as mp
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mp.mp.dps = 50


@pytest.fixture(scope="module")
def fm():
    so = os.path.join(tempfile.mkdtemp(), "fm_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "fastmath_host.cpp")])
    L = C.CDLL(so)
    for f in ("t_fm_exp", "t_fm_exp_fast", "t_fm_log", "t_fm_erfcx"):
        getattr(L, f).restype = C.c_double
        getattr(L, f).argtypes = [C.c_double]
    dp = C.POINTER(C.c_double)
    L.t_fm_gauss_tail.argtypes = [C.c_double, dp, dp]
    L.t_probit_row.argtypes = [C.c_double, C.c_double, dp, dp]
    L.t_logit2_row.argtypes = [C.c_double, C.c_double, dp, dp]
    return L


def _rel(got, want):
    return float(abs(mp.mpf(got) - want) / abs(want))


def test_exp_log_erfcx(fm):
    rng = np.random.default_rng(1)
    for x in np.concatenate([rng.uniform(-700, 709, 3000), rng.uniform(-2, 2, 2000), [0.0, 709.0, -708.0]]):
        assert _rel(fm.t_fm_exp(x), mp.exp(mp.mpf(float(x)))) < 3e-16
    # documented saturation: the argument is clamped to [-708, 709.7]; above ~709.44 -> +inf
    assert 0 < fm.t_fm_exp(-800.0) < 1e-307 and fm.t_fm_exp(711.0) == np.inf
    for z in np.concatenate([10 ** rng.uniform(-307, 300, 3000), rng.uniform(0.5, 2, 3000),
                             [2.2250738585072014e-308, 5e-320, 0.9999999999999999, 1.0000000000000002]]):
        want = mp.log(mp.mpf(float(z)))
        assert abs(mp.mpf(fm.t_fm_log(z)) - want) <= 3e-16 * abs(want) + mp.mpf(2) ** -60
    assert fm.t_fm_log(1.0) == 0.0
    for x in np.concatenate([rng.uniform(0, 30, 5000), 10 ** rng.uniform(-12, 3, 1000), [0.0]]):
        want = mp.erfc(mp.mpf(float(x))) * mp.exp(mp.mpf(float(x)) ** 2)
        assert _rel(fm.t_fm_erfcx(x), want) < 8e-16


def test_exp_fast(fm):
    """the unclamped exponential of the multinomial soft-max: same accuracy as fm_exp wherever its
    guard (|x| < 708, decided on the high word) lets it run"""
    fm.t_fm_exp_fast_ok.argtypes = [C.c_double]
    rng = np.random.default_rng(3)
    for x in np.concatenate([rng.uniform(-708, 708, 4000), rng.uniform(-3, 3, 3000),
                             [0.0, 707.999999, -707.999999, 1e-300, -1e-300, 5e-324]]):
        assert fm.t_fm_exp_fast_ok(x) == 1
        assert _rel(fm.t_fm_exp_fast(x), mp.exp(mp.mpf(float(x)))) < 3e-16
    for x in (708.0, -708.0, 709.5, -745.0, 1e6, -1e300, np.inf, -np.inf, np.nan):
        assert fm.t_fm_exp_fast_ok(x) == 0


def test_gauss_tail(fm):
    rng = np.random.default_rng(2)
    E, T = C.c_double(), C.c_double()
    for a in np.concatenate([rng.uniform(0, 37.519, 8000), [0.0, 8.29, 8.572, 37.519]]):
        fm.t_fm_gauss_tail(a, C.byref(E), C.byref(T))
        am = mp.mpf(float(a))
        assert _rel(E.value, mp.exp(-am * am / 2)) < 4e-16
        assert _rel(T.value, mp.ncdf(-am)) < 9e-16


def test_probit_and_logit_rows_against_oracle(fm, oracle):
    """per row, the fast evaluator agrees with the oracle's reference-shaped arithmetic;
    rows where 1 - cdf cancels are compared at the accuracy that subtraction allows"""
    rng = np.random.default_rng(3)
    ll, d = C.c_double(), C.c_double()
    xs = np.concatenate([rng.uniform(-7, 7, 4000), rng.uniform(-39, 39, 1000),
                         [-40.0, -38.0, -37.6, -37.5, -9.0, -8.6, 0.0, 8.6, 9.0, 37.5, 37.6, 38.0, 40.0]])
    eps = np.finfo(float).eps
    for xb in xs:
        for y in (0.0, 1.0):
            fm.t_probit_row(xb, y, C.byref(ll), C.byref(d))
            X = np.array([[1.0, xb]]); yy = np.array([y]); b = np.array([0.0, 1.0])
            want = float(oracle.probit_ll(X, yy, b, oracle.FAITHFUL))
            n = oracle.gauss_cdf(-xb)
            n = 1e-10 if n == 0 else n
            n = n if n < 1 else 1 - 1e-10
            arg = 1 - n if y else n
            tol = 8 * eps * (n / arg if y else 1.0) + 2 * eps + 4 * eps * abs(want)   # +2 eps: rounding of 1-n itself
            assert abs(ll.value - want) <= tol, (xb, y, ll.value, want)
            # (the faithful pdf exp(-u*u/2) carries u^2/2 ulps of its own: compare with exact)
            gw = float(oracle.probit_score(X, yy, b, oracle.EXACT)[0][0])
            assert abs(d.value - gw) <= (8 * eps * (n / arg if y else 1.0) + 8 * eps) * abs(gw) + 1e-290, (xb, y)   # 1e-290: densities below 2^-1022 flush to 0
    for xb in np.concatenate([rng.uniform(-30, 30, 3000), [-800.0, -720.0, -700.5, 0.0, 700.0, 800.0]]):
        for y in (0.0, 1.0):
            fm.t_logit2_row(xb, y, C.byref(ll), C.byref(d))
            X = np.array([[1.0, xb]]); yy = np.array([y]); B = np.array([[0.0], [1.0]])
            want = float(oracle.logit_ll(X, yy, B, mode=oracle.EXACT))
            assert abs(ll.value - want) <= 4 * eps * max(abs(want), abs(xb), 1.0), (xb, y, ll.value, want)
            gw = float(oracle.logit_score(X, yy, B, mode=oracle.EXACT)[0][0])
            assert abs(d.value - gw) <= 4 * eps, (xb, y)
