"""ctypes front end of oracle/liboracle.so (the C restatement in oracle.c).

TEST INFRASTRUCTURE ONLY.  Results come back as numpy longdouble (x87 80-bit),
so the checker keeps more digits than the FP64 device path it checks.
"""
import ctypes as C
import os
import subprocess

import numpy as np

FAITHFUL, EXACT = 0, 1
_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = C.POINTER(C.c_double)
_lp = C.POINTER(C.c_longdouble)


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc -O3 -fopenmp)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = [os.path.join(_HERE, f) for f in ("oracle.c", "oracle.h", "cpu_models.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    return so


def models_lib():
    """oracle/cpu_models.c: the oracle behind reference-signature log_likelihood / score entries"""
    build()
    return C.CDLL(os.path.join(_HERE, "liboracle_models.so"))


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_gauss_cdf.restype = C.c_double
        _LIB.oracle_gauss_cdf.argtypes = [C.c_double]
        _LIB.oracle_gauss_pdf.restype = C.c_double
        _LIB.oracle_gauss_pdf.argtypes = [C.c_double, C.c_double]
        _LIB.oracle_time_probit_ll_score.restype = C.c_double
        _LIB.oracle_time_logit_ll.restype = C.c_double
        _LIB.oracle_max_threads.restype = C.c_int
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _mat(X):
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("X must be 2-d")
    if X.strides[1] != 8 or (X.shape[0] > 1 and X.strides[0] % 8):
        X = np.ascontiguousarray(X)
    tda = X.strides[0] // 8 if X.shape[0] > 1 else max(X.shape[1], 1)
    return X, X.ctypes.data_as(_dp), tda


def _ld(n):
    out = np.zeros(n, dtype=np.longdouble)
    return out, out.ctypes.data_as(_lp)


def _sz(n):
    return C.c_size_t(int(n))


def gauss_cdf(x):
    return lib().oracle_gauss_cdf(float(x))


def gauss_pdf(x, sigma=1.0):
    return lib().oracle_gauss_pdf(float(x), float(sigma))


def probit_ll(X, y, beta, mode=EXACT):
    X, Xp, tda = _mat(X); y, yp = _d(y); b, bp = _d(beta)
    out, op = _ld(1)
    lib().oracle_probit_ll(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(X.shape[1]), bp, C.c_int(mode), op)
    return out[0]


def probit_score(X, y, beta, mode=EXACT):
    """returns (g, scale) with scale_j = sum_i |X_ij d_i|"""
    X, Xp, tda = _mat(X); y, yp = _d(y); b, bp = _d(beta)
    k = X.shape[1]
    g, gp = _ld(k); a, ap = _ld(k)
    lib().oracle_probit_score(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(k), bp, C.c_int(mode), gp, ap)
    return g, a


def probit_start(X):
    X, Xp, tda = _mat(X)
    out = np.zeros(X.shape[1]); 
    lib().oracle_probit_start(Xp, _sz(tda), _sz(X.shape[0]), _sz(X.shape[1]), out.ctypes.data_as(_dp))
    return out


def logit_ll(X, y, B, cats=None, mode=EXACT):
    X, Xp, tda = _mat(X); y, yp = _d(y)
    B = np.ascontiguousarray(np.asarray(B, dtype=np.float64).reshape(X.shape[1], -1))
    Cn = B.shape[1] + 1
    cats, cp = _d(np.arange(Cn, dtype=np.float64) if cats is None else cats)
    out, op = _ld(1)
    lib().oracle_logit_ll(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(X.shape[1]), B.ctypes.data_as(_dp),
                          _sz(Cn), cp, C.c_int(mode), op)
    return out[0]


def logit_score(X, y, B, cats=None, mode=EXACT):
    X, Xp, tda = _mat(X); y, yp = _d(y)
    B = np.ascontiguousarray(np.asarray(B, dtype=np.float64).reshape(X.shape[1], -1))
    Cn = B.shape[1] + 1
    cats, cp = _d(np.arange(Cn, dtype=np.float64) if cats is None else cats)
    P = B.size
    g, gp = _ld(P); a, ap = _ld(P)
    lib().oracle_logit_score(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(X.shape[1]), B.ctypes.data_as(_dp),
                             _sz(Cn), cp, C.c_int(mode), gp, ap)
    return g, a


def _vm(v, M):
    if v is None:
        v_arr, vp, vs = None, None, 0
    else:
        v_arr, vp = _d(v); vs = v_arr.size
    if M is None:
        M_arr, Mp, tda, m1, m2 = None, None, 0, 0, 0
    else:
        M_arr, Mp, tda = _mat(M); m1, m2 = M_arr.shape
    return (v_arr, M_arr), vp, _sz(vs), Mp, _sz(tda), _sz(m1), _sz(m2)


def poisson_ll(v, M, lam, mode=EXACT):
    keep, vp, vs, Mp, tda, m1, m2 = _vm(v, M)
    out, op = _ld(1)
    lib().oracle_poisson_ll(vp, vs, Mp, tda, m1, m2, C.c_double(lam), C.c_int(mode), op)
    return out[0]


def poisson_score(v, M, lam, mode=EXACT):
    keep, vp, vs, Mp, tda, m1, m2 = _vm(v, M)
    out, op = _ld(1)
    lib().oracle_poisson_score(vp, vs, Mp, tda, m1, m2, C.c_double(lam), C.c_int(mode), op)
    return out


def normal_ll(v, M, mu, sigma, mode=EXACT):
    keep, vp, vs, Mp, tda, m1, m2 = _vm(v, M)
    out, op = _ld(1)
    lib().oracle_normal_ll(vp, vs, Mp, tda, m1, m2, C.c_double(mu), C.c_double(sigma), C.c_int(mode), op)
    return out[0]


def normal_score(v, M, mu, sigma, mode=EXACT):
    keep, vp, vs, Mp, tda, m1, m2 = _vm(v, M)
    out, op = _ld(2)
    lib().oracle_normal_score(vp, vs, Mp, tda, m1, m2, C.c_double(mu), C.c_double(sigma), C.c_int(mode), op)
    return out


def _iid(fn, v, M, params, ng, mode):
    keep, vp, vs, Mp, tda, m1, m2 = _vm(v, M)
    ll, lp = _ld(1); g, gp = _ld(ng)
    getattr(lib(), fn)(vp, vs, Mp, tda, m1, m2, *[C.c_double(p) for p in params], C.c_int(mode), lp, gp)
    return ll[0], g


def exponential(v, M, mu, mode=EXACT):
    """(LL, score) of the iid exponential model"""
    return _iid("oracle_exponential", v, M, [mu], 1, mode)


def gamma(v, M, a, b, mode=EXACT):
    return _iid("oracle_gamma", v, M, [a, b], 2, mode)


def lognormal(v, M, mu, sd, mode=EXACT):
    return _iid("oracle_lognormal", v, M, [mu, sd], 2, mode)


def bernoulli(v, M, p, mode=EXACT):
    """(LL, [closed-form score]) of the iid Bernoulli model"""
    return _iid("oracle_bernoulli", v, M, [p], 1, mode)


def beta(v, M, a, b, mode=EXACT):
    return _iid("oracle_beta", v, M, [a, b], 2, mode)


def zipf(v, M, bb, mode=EXACT):
    return _iid("oracle_zipf", v, M, [bb], 1, mode)


def tdist(v, M, mu, sigma, df, mode=EXACT):
    return _iid("oracle_tdist", v, M, [mu, sigma, df], 1, mode)


def yule(v, M, bb, mode=EXACT):
    return _iid("oracle_yule", v, M, [bb], 1, mode)


def dirichlet(M, alpha, mode=EXACT):
    keep, vp, vs, Mp, tda, m1, m2 = _vm(None, M)
    a = np.ascontiguousarray(alpha, dtype=np.float64)
    ll, lp = _ld(1); g, gp = _ld(a.size)
    lib().oracle_dirichlet(Mp, tda, m1, m2, a.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(mode), lp, gp)
    return ll[0], g


def zeta(s):
    lib().oracle_zeta.restype = C.c_double
    return lib().oracle_zeta(C.c_double(s))


def digamma(x):
    lib().oracle_digamma.restype = C.c_double
    return lib().oracle_digamma(C.c_double(x))


def map_sum(v, M, fn, p0=0.0, mode=EXACT):
    keep, vp, vs, Mp, tda, m1, m2 = _vm(v, M)
    out, op = _ld(1)
    lib().oracle_map_sum(vp, vs, Mp, tda, m1, m2, C.c_int(fn), C.c_double(p0), C.c_int(mode), op)
    return out[0]


def ols_ll(X, y, beta, w=None, mode=EXACT):
    X, Xp, tda = _mat(X); y, yp = _d(y); b, bp = _d(beta)
    wk, wp = (None, None) if w is None else _d(w)
    out, op = _ld(1)
    lib().oracle_ols_ll(Xp, _sz(tda), yp, wp, _sz(X.shape[0]), _sz(X.shape[1]), bp, C.c_int(mode), op)
    return out[0]


def ols_score(X, y, beta, w=None, mode=EXACT):
    X, Xp, tda = _mat(X); y, yp = _d(y); b, bp = _d(beta)
    wk, wp = (None, None) if w is None else _d(w)
    k = X.shape[1]
    g, gp = _ld(k); a, ap = _ld(k)
    lib().oracle_ols_score(Xp, _sz(tda), yp, wp, _sz(X.shape[0]), _sz(k), bp, C.c_int(mode), gp, ap)
    return g, a


def ols_gram(X, y, mode=EXACT):
    X, Xp, tda = _mat(X); y, yp = _d(y)
    k = X.shape[1]
    A, Ap = _ld(k * k); b, bp = _ld(k)
    lib().oracle_ols_gram(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(k), C.c_int(mode), Ap, bp)
    return A.reshape(k, k), b


def poisson_reg_ll(X, y, beta, mode=EXACT):
    X, Xp, tda = _mat(X); y, yp = _d(y); b, bp = _d(beta)
    out, op = _ld(1)
    lib().oracle_poisson_reg_ll(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(X.shape[1]), bp, C.c_int(mode), op)
    return out[0]


def poisson_reg_score(X, y, beta, mode=EXACT):
    X, Xp, tda = _mat(X); y, yp = _d(y); b, bp = _d(beta)
    k = X.shape[1]
    g, gp = _ld(k); a, ap = _ld(k)
    lib().oracle_poisson_reg_score(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(k), bp, C.c_int(mode), gp, ap)
    return g, a


def information(family, X, y, beta):
    X, Xp, tda = _mat(X); y, yp = _d(y); b, bp = _d(beta)
    k = X.shape[1]
    H, Hp = _ld(k * k)
    lib().oracle_information(C.c_int(family), Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(k), bp, Hp)
    return H.reshape(k, k)


def time_probit_ll_score(X, y, beta, threads=0):
    """structure-faithful reference CPU path; returns dict of seconds and results"""
    X, Xp, tda = _mat(X); y, yp = _d(y); b, bp = _d(beta)
    k = X.shape[1]
    ll = C.c_double(); g = np.zeros(k); s1 = C.c_double(); s2 = C.c_double()
    tot = lib().oracle_time_probit_ll_score(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(k), bp, C.c_int(threads),
                                            C.byref(ll), g.ctypes.data_as(_dp), C.byref(s1), C.byref(s2))
    return {"seconds": tot, "sec_ll": s1.value, "sec_score": s2.value, "ll": ll.value, "score": g}


def time_logit_ll(X, y, B, threads=0):
    X, Xp, tda = _mat(X); y, yp = _d(y)
    B = np.ascontiguousarray(np.asarray(B, dtype=np.float64).reshape(X.shape[1], -1))
    ll = C.c_double()
    tot = lib().oracle_time_logit_ll(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(X.shape[1]), B.ctypes.data_as(_dp),
                                     _sz(B.shape[1] + 1), C.c_int(threads), C.byref(ll))
    return {"seconds": tot, "ll": ll.value}


def logit_information(X, B, ncat):
    """multinomial-logit observed information, P x P in pack order (long double)"""
    X, Xp, tda = _mat(X); b, bp = _d(np.asarray(B).ravel())
    k = X.shape[1]
    P = k * (ncat - 1)
    H, Hp = _ld(P * P)
    lib().oracle_logit_information(Xp, _sz(tda), _sz(X.shape[0]), _sz(k), bp, _sz(ncat), Hp)
    return H.reshape(P, P)


def probit_ll_score_exact_omp(X, y, beta, threads=0):
    """EXACT-mode probit LL, score and score scale over row blocks in parallel (long double)"""
    X, Xp, tda = _mat(X); y, yp = _d(y); b, bp = _d(beta)
    k = X.shape[1]
    ll, llp = _ld(1); g, gp = _ld(k); a, ap = _ld(k)
    lib().oracle_probit_ll_score_exact_omp(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(k), bp, C.c_int(threads), llp, gp, ap)
    return ll[0], g, a


def time_bootstrap_resample(X, y, seed=8):
    """seconds of the reference's serial row-copy resample of all N rows (apop_bootstrap.m4.c:143-149)"""
    X, Xp, tda = _mat(X); y, yp = _d(y)
    Xo = np.empty_like(X); yo = np.empty_like(y)
    f = lib().oracle_time_bootstrap_resample
    f.restype = C.c_double
    return float(f(Xp, _sz(tda), yp, _sz(X.shape[0]), _sz(X.shape[1]), C.c_ulonglong(seed),
                   Xo.ctypes.data_as(C.POINTER(C.c_double)), yo.ctypes.data_as(C.POINTER(C.c_double))))


def max_threads():
    return lib().oracle_max_threads()
