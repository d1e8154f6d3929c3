"""The shared-centre expansion of the batched kernel (apophenia_b200/csrc/families.cuh: Taylor<Fam>), compiled for the
host: the degree-6 coefficients of every family's row term against mpmath's Taylor series of the same function at 40
digits, and the truncation error at the kernel's radius (1.5e-2) against the exact function."""
import ctypes as C
import os
import subprocess
import tempfile

import mpmath as mp
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mp.mp.dps = 40
DELTA_MAX = 1.5e-2


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(tempfile.mkdtemp(), "taylor_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "taylor_host.cpp")])
    L = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    for f in ("t_probit_coefs", "t_logit2_coefs", "t_poisson_coefs"):
        getattr(L, f).argtypes = [C.c_double, C.c_double, dp, dp, dp]
    return L


def _coefs(fn, x0, y, D):
    a = (C.c_double * (D + 1))(); ll = C.c_double(); d = C.c_double()
    ok = fn(x0, y, a, C.byref(ll), C.byref(d))
    return bool(ok), np.array(a[:]), ll.value, d.value


FUNCS = {
    "probit": (lambda y: (lambda x: mp.log(mp.ncdf(x if y else -x))), "t_probit_coefs", (-3.4, 3.4), (0.0, 1.0)),
    "logit2": (lambda y: (lambda x: (x if y else 0) - mp.log(1 + mp.exp(x))), "t_logit2_coefs", (-12.0, 12.0), (0.0, 1.0)),
    "poisson": (lambda y: (lambda x: y * x - mp.exp(x)), "t_poisson_coefs", (-3.0, 3.0), (0.0, 2.0, 5.0)),
}


@pytest.mark.parametrize("fam", list(FUNCS))
def test_coefficients_match_the_taylor_series(lib, fam):
    make, sym, (lo, hi), ys = FUNCS[fam]
    D = lib.t_degree()
    assert D == 6
    rng = np.random.default_rng(3)
    for x0 in np.concatenate([rng.uniform(lo, hi, 40), [0.0, lo, hi]]):
        for y in ys:
            ok, a, ll0, d0 = _coefs(getattr(lib, sym), float(x0), float(y), D)
            assert ok
            want = mp.taylor(make(y), mp.mpf(float(x0)), D)
            for n in range(2, D + 1):
                # what matters is the term a_n delta^n at the radius, relative to the row term and its derivative
                scale = max(abs(float(want[0])), abs(float(want[1])), 1e-3)
                err = abs(a[n] - float(want[n])) * DELTA_MAX ** n / scale
                # the probit coefficients are built from lambda = phi/Phi as the reference's formula yields it, i.e. with
                # the rounding of 1 - Phi(-z) in it (relative 2^-53 / Phi(z) on the unlikely side, amplified by the
                # cancellation in lambda + z): the same noise a_0 and a_1 carry, far below 1e-12 of the row term
                tol = 1e-16 if fam != "probit" else 1e-16 + 100 * 2.0 ** -53 / float(mp.ncdf(-abs(float(x0)))) * DELTA_MAX ** 2
                assert err < tol, (fam, x0, y, n, a[n], float(want[n]))


@pytest.mark.parametrize("fam", list(FUNCS))
def test_truncation_at_the_radius(lib, fam):
    make, sym, (lo, hi), ys = FUNCS[fam]
    D = lib.t_degree()
    rng = np.random.default_rng(4)
    worst_ll = worst_d = 0.0
    for x0 in rng.uniform(lo, hi, 60):
        for y in ys:
            f = make(y)
            c = [float(v) for v in mp.taylor(f, mp.mpf(float(x0)), D)]      # exact coefficients: truncation alone
            for dl in (DELTA_MAX, -DELTA_MAX):
                ll = sum(c[n] * dl ** n for n in range(D + 1))
                dv = sum(n * c[n] * dl ** (n - 1) for n in range(1, D + 1))
                wl = f(mp.mpf(float(x0)) + dl); wd = mp.diff(f, mp.mpf(float(x0)) + dl)
                worst_ll = max(worst_ll, float(abs(ll - wl)) / max(abs(float(wl)), 1.0))
                worst_d = max(worst_d, float(abs(dv - wd)) / max(abs(float(wd)), 1.0))
    assert worst_ll < 1e-15 and worst_d < 2e-13, (fam, worst_ll, worst_d)


def test_rows_near_the_clamps_are_not_expanded(lib):
    D = lib.t_degree()
    assert not _coefs(lib.t_probit_coefs, 7.5, 1.0, D)[0] and not _coefs(lib.t_probit_coefs, -8.3, 1.0, D)[0]
    assert _coefs(lib.t_probit_coefs, 7.3, 1.0, D)[0] and _coefs(lib.t_probit_coefs, -7.3, 0.0, D)[0]       # likely side: up to the clamps
    assert not _coefs(lib.t_probit_coefs, -3.6, 1.0, D)[0] and not _coefs(lib.t_probit_coefs, 3.6, 0.0, D)[0]  # unlikely side: formula noise
    assert _coefs(lib.t_probit_coefs, -3.4, 1.0, D)[0]
    assert not _coefs(lib.t_logit2_coefs, 31.0, 1.0, D)[0]
