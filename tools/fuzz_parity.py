#!/usr/bin/env python
"""Randomised parity sweep on a B200: random shapes, parameter scales and a few hostile values, every
family against the oracle (exact mode) at the tolerances of the test suite.  Prints a summary; exits 1
on the first disagreement.  Usage: python tools/fuzz_parity.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import apophenia_b200 as ab  # noqa: E402
from conftest import gen_probit_logit_sample  # noqa: E402
from oracle import oracle as o  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ab.init(0)
t_end = time.time() + budget
count = {}


def _report_last_error():
    print("last_error:", ab.last_error(), flush=True)


def ll_ok(got, want, what, slack=0.0):
    """slack: absolute allowance for terms that vanish below double resolution of their operands
    (x.b - (x.b + log(1 + e^-x.b)) is exactly 0 in the reference's double arithmetic for x.b > 37)"""
    want = float(want)
    if np.isnan(want) and np.isnan(got):
        return
    if np.isinf(want) and got == want:
        return
    if not abs(got - want) <= 1e-12 * abs(want) + slack + 1e-300:
        _report_last_error()
    assert abs(got - want) <= 1e-12 * abs(want) + slack + 1e-300, (what, got, want)


def sc_ok(got, want, scale, what, slack=0.0):
    """slack: per-component absolute allowance (a residual 1 - p with p within an ulp of 1 is 0 in double)"""
    want = want.astype(float).ravel(); scale = scale.astype(float).ravel()
    assert np.all(np.abs(got - want) <= 1e-12 * scale + slack + 1e-300), (what, np.max(np.abs(got - want) / (scale + 1e-300)))


while time.time() < t_end:
    fam = rng.choice(["probit", "logit2", "logitC", "poisson_reg", "ols", "normal", "poisson", "gamma", "beta", "t", "yule", "dirichlet",
                      "batched", "batched", "infoC"])
    n = int(rng.choice([1, 31, 32, 33, 1000, 4097, 30011, 100003]))
    scale = float(rng.choice([0.0, 0.1, 1.0, 4.0, 30.0]))
    count[fam] = count.get(fam, 0) + 1
    if fam == "batched":
        # the batched-beta path: random replicate counts, spreads from far inside the Taylor radius to far outside it,
        # with and without resample counts and a weights column (ignored), every checked replicate against the oracle
        k = int(rng.integers(1, 33)); m = int(rng.choice([1, 7, 8, 9, 63, 64, 65, 130]))
        which = rng.choice(["probit", "logit", "poisson_reg"])
        n = min(n, 30011)
        X, y, beta = gen_probit_logit_sample(n, k, seed=int(rng.integers(1 << 30)), method="probit" if which == "probit" else "logit")
        b = beta * min(scale, 4.0)
        if which == "poisson_reg":
            b = b / (4 * np.sqrt(k)); y = rng.poisson(np.exp(X @ b)).astype(float)
        F = {"probit": ab.PROBIT, "logit": ab.LOGIT, "poisson_reg": ab.POISSON_REG}[which]
        spread = float(rng.choice([0.0, 1e-4, 3e-3, 2e-2, 0.5]))
        Bs = b[None, :] + spread * rng.standard_normal((m, k))
        w = rng.uniform(0.5, 2.0, n) if rng.uniform() < 0.3 else None
        d = ab.pin(ab.Data(matrix=X, vector=y, weights=w))
        use_counts = bool(rng.integers(0, 2))
        if use_counts:
            ab.bootstrap_counts(d, m, seed=int(rng.integers(1 << 30))); cnt = ab.counts_download(d, m, n)
        want_score = bool(rng.integers(0, 2))
        llb, gb = ab.ll_score_batched(F, d, Bs, use_counts=use_counts, want_score=want_score)
        for r in sorted(set([0, m // 2, m - 1])):
            if use_counts:
                idx = np.repeat(np.arange(n), cnt[:, r]); Xr, yr = X[idx], y[idx]
            else:
                Xr, yr = X, y
            if Xr.shape[0] == 0:
                assert llb[r] == 0.0; continue
            eps = np.finfo(float).eps
            if which == "probit":
                from scipy.special import ndtr
                # the reference clamps n only where the rounded double IS 0 or 1 (model/apop_probit.c:85-86); below that, 1 - n
                # keeps its rounding noise of eps / (1 - n)
                xb = Xr @ Bs[r]; nn = ndtr(-xb); nn = np.where(nn == 0, 1e-10, nn); nn = np.where(nn < 1, nn, 1 - 1e-10); arg = np.where(yr != 0, 1 - nn, nn)
                # (absolute noise of log(1 - n) is ulp(1)/2 / (1 - n) per row whatever the size of n: a one-row LL of -3.8e-11
                # carries 1e-16 of it, which is 3e-6 of its value)
                slack = float(np.sum(2 * eps * np.where(yr != 0, 1.0 / arg, 1.0)))
                want = float(o.probit_ll(Xr, yr, Bs[r], o.FAITHFUL))
                # the shared-centre expansion truncates at |a_7| delta^7 <= 5e-5 * 1.7e-13 per row in absolute terms (a row
                # far on the likely side has a term of 1e-6 and below, so relative to a one-row LL that is not 1e-12)
                slack += 3e-17 * Xr.shape[0]
                if not abs(llb[r] - want) <= slack + 1e-12 * abs(want):
                    ll1 = ab.ll_score(ab.PROBIT, d, Bs[r])[0] if not use_counts else float("nan")
                    print("DIAG", which, n, k, m, "spread", spread, "scale", scale, "counts", use_counts, "score", want_score, "r", r, "batched", llb[r], "faithful", want,
                          "exact-oracle", float(o.probit_ll(Xr, yr, Bs[r], o.EXACT)), "single kernel", ll1, "slack", slack, "max|xb|", float(np.max(np.abs(xb))),
                          "exact share", ab.batched_exact_fraction(d), flush=True)
                    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "fuzz_fail.npz"), X=X, y=y, Bs=Bs, r=r, llb=llb)
                assert abs(llb[r] - want) <= slack + 1e-12 * abs(want), (fam, which, n, k, m, spread, use_counts, r, llb[r], want)
                if want_score and np.max(np.abs(xb)) < 8:
                    # the reference's weight pdf / (1 - cdf) (model/apop_probit.c:147-149) inherits the same rounding of the
                    # double cdf: relative eps / (1 - cdf) on a y = 1 row
                    dabs = np.exp(-0.5 * xb * xb) * 0.3989422804014327 / arg
                    noise = dabs * 2 * eps * np.where(yr != 0, 1.0 / arg, 1.0)
                    ge, sc = o.probit_score(Xr, yr, Bs[r], o.EXACT)
                    sc_ok(gb[r], ge, sc, (fam, which, n, k, m, spread, r), slack=8 * eps * np.abs(Xr).sum(0) + (np.abs(Xr) * noise[:, None]).sum(0))
            elif which == "logit":
                ll_ok(llb[r], o.logit_ll(Xr, yr, Bs[r].reshape(k, 1), mode=o.EXACT), (fam, which, n, k, m, spread, r), slack=4 * Xr.shape[0] * eps * max(1.0, float(np.max(np.abs(Xr @ Bs[r])))))
                if want_score:
                    ge, sc = o.logit_score(Xr, yr, Bs[r].reshape(k, 1), mode=o.EXACT); sc_ok(gb[r], ge, sc, (fam, which, n, k, m, spread, r), slack=4 * eps * np.abs(Xr).sum(0))
            else:
                ll_ok(llb[r], o.poisson_reg_ll(Xr, yr, Bs[r], o.EXACT), (fam, which, n, k, m, spread, r))
                if want_score:
                    ge, sc = o.poisson_reg_score(Xr, yr, Bs[r], o.EXACT); sc_ok(gb[r], ge, sc, (fam, which, n, k, m, spread, r))
        ab.unpin(d)
    elif fam == "infoC":
        k = int(rng.integers(1, 13)); C = int(rng.integers(3, 9)); n = min(n, 4097)
        X, y, B = gen_probit_logit_sample(n, k, seed=int(rng.integers(1 << 30)), method="logit", ncat=C)
        Bm = B * min(scale, 4.0)
        d = ab.pin(ab.Data(matrix=X, vector=y))
        H = ab.information(ab.LOGIT, d, Bm.ravel()); want = o.logit_information(X, Bm, C).astype(float)
        assert np.max(np.abs(H - want)) <= 1e-11 * max(np.abs(want).max(), 1e-300), (fam, n, k, C, scale)
        assert np.array_equal(H, H.T)
        ab.unpin(d)
    elif fam in ("probit", "logit2", "poisson_reg", "ols"):
        k = int(rng.integers(1, 41))
        X, y, beta = gen_probit_logit_sample(n, k, seed=int(rng.integers(1 << 30)), method="probit" if fam == "probit" else "logit")
        b = beta * scale
        if fam == "poisson_reg":
            b = b / (4 * np.sqrt(k)); y = rng.poisson(np.exp(X @ b)).astype(float)
        if fam == "ols":
            y = X @ beta + rng.standard_normal(n)
        d = ab.pin(ab.Data(matrix=X, vector=y))
        F = {"probit": ab.PROBIT, "logit2": ab.LOGIT, "poisson_reg": ab.POISSON_REG, "ols": ab.OLS}[fam]
        if fam == "ols" and n < 3:
            ab.unpin(d); continue
        ll, g = ab.ll_score(F, d, b)
        if fam == "probit":
            # |x.beta| beyond ~8: the reference's own double arithmetic rounds 1 - Phi to a handful of bits
            # (arg = 1 - n with n within an ulp of 1), so its result there depends on the last bit of erfc;
            # compare with the faithful restatement at the accuracy that subtraction allows
            from scipy.special import ndtr
            xb = X @ b
            nn = np.clip(ndtr(-xb), 1e-10, 1 - 1e-10)
            arg = np.where(y != 0, 1 - nn, nn)
            slack = float(np.sum(8 * np.finfo(float).eps * np.where(y != 0, nn / arg, 1.0)))
            want = float(o.probit_ll(X, y, b, o.FAITHFUL))
            assert abs(ll - want) <= slack + 1e-12 * abs(want), (fam, n, k, scale, ll, want, slack)
            if scale <= 1.0:      # the score against the exact restatement where nothing saturates
                ge, sc = o.probit_score(X, y, b, o.EXACT); sc_ok(g, ge, sc, (fam, n, k, scale), slack=8 * np.finfo(float).eps * np.abs(X).sum(0))
        elif fam == "logit2":
            ll_ok(ll, o.logit_ll(X, y, b.reshape(k, 1), mode=o.EXACT), (fam, n, k, scale), slack=4 * n * np.finfo(float).eps * max(1.0, float(np.max(np.abs(X @ b))))); ge, sc = o.logit_score(X, y, b.reshape(k, 1), mode=o.EXACT); sc_ok(g, ge, sc, (fam, n, k, scale), slack=4 * np.finfo(float).eps * np.abs(X).sum(0))
        elif fam == "poisson_reg":
            ll_ok(ll, o.poisson_reg_ll(X, y, b, o.EXACT), (fam, n, k)); ge, sc = o.poisson_reg_score(X, y, b, o.EXACT); sc_ok(g, ge, sc, (fam, n, k))
        else:
            want = float(o.ols_ll(X, y, b, None, o.EXACT))
            assert abs(ll - want) <= 1e-11 * max(abs(want), n), (fam, n, k, ll, want)
        ab.unpin(d)
    elif fam == "logitC":
        k = int(rng.integers(1, 33)); C = int(rng.integers(3, 9))
        X, y, B = gen_probit_logit_sample(n, k, seed=int(rng.integers(1 << 30)), method="logit", ncat=C)
        Bm = B * scale
        if rng.uniform() < 0.15 and n > 2:       # a hostile row: the shifted soft-max path
            X[n // 2, -1] = 900.0 * rng.choice([-1, 1])
        d = ab.pin(ab.Data(matrix=X, vector=y))
        ll, g = ab.ll_score(ab.LOGIT, d, Bm.ravel())
        if np.isnan(ll):
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            np.savez_compressed(os.path.join(ROOT, "gpurun_out", "fuzz_fail_logitC.npz"), X=X, y=y, Bm=Bm)
            print("DIAG logitC nan: second call", ab.ll_score(ab.LOGIT, d, Bm.ravel() * (1 + 1e-12))[0], "unpinned copy", ab.ll_score(ab.LOGIT, ab.Data(matrix=X.copy(), vector=y.copy()), Bm.ravel())[0], flush=True)
        ll_ok(ll, o.logit_ll(X, y, Bm, mode=o.EXACT), (fam, n, k, C, scale), slack=4 * n * np.finfo(float).eps * max(1.0, float(np.max(np.abs(X @ Bm)))))
        ge, sc = o.logit_score(X, y, Bm, mode=o.EXACT); sc_ok(g, ge, sc, (fam, n, k, C, scale), slack=np.repeat(4 * np.finfo(float).eps * np.abs(X).sum(0), C - 1))
        ab.unpin(d)
    else:
        k = int(rng.integers(1, 9))
        has_v = bool(rng.integers(0, 2)) and fam != "dirichlet"
        if fam == "normal":
            M = rng.normal(1.0, 2.0, (n, k)); v = rng.normal(1.0, 2.0, n) if has_v else None
            d = ab.pin(ab.Data(matrix=M, vector=v)); m = ab.vector_model([0.7, 1.9], d)
            ll_ok(ab.log_likelihood(ab.NORMAL, d, m), o.normal_ll(v, M, 0.7, 1.9), fam)
        elif fam == "poisson":
            M = rng.poisson(3.1, (n, k)).astype(float); v = rng.poisson(3.1, n).astype(float) if has_v else None
            d = ab.pin(ab.Data(matrix=M, vector=v)); m = ab.vector_model([2.9], d)
            ll_ok(ab.log_likelihood(ab.POISSON, d, m), o.poisson_ll(v, M, 2.9), fam)
        elif fam == "gamma":
            M = rng.gamma(2.5, 1.7, (n, k)); v = rng.gamma(2.5, 1.7, n) if has_v else None
            d = ab.pin(ab.Data(matrix=M, vector=v)); m = ab.vector_model([2.4, 1.8], d)
            ll_ok(ab.log_likelihood(ab.GAMMA, d, m), o.gamma(v, M, 2.4, 1.8)[0], fam)
        elif fam == "beta":
            M = rng.beta(2.0, 3.0, (n, k)); v = rng.beta(2.0, 3.0, n) if has_v else None
            d = ab.pin(ab.Data(matrix=M, vector=v)); m = ab.vector_model([2.2, 3.1], d)
            ll_ok(ab.log_likelihood(ab.BETA, d, m), o.beta(v, M, 2.2, 3.1)[0], fam)
        elif fam == "t":
            M = rng.standard_t(5, (n, k)) * 2 + 1; v = rng.standard_t(5, n) if has_v else None
            d = ab.pin(ab.Data(matrix=M, vector=v)); m = ab.vector_model([0.9, 2.1, 5.5], d)
            ll_ok(ab.log_likelihood(ab.TDIST, d, m), o.tdist(v, M, 0.9, 2.1, 5.5)[0], fam)
        elif fam == "yule":
            M = rng.zipf(2.5, (n, k)).astype(float); v = rng.zipf(2.5, n).astype(float) if has_v else None
            d = ab.pin(ab.Data(matrix=M, vector=v)); m = ab.vector_model([2.4], d)
            ll_ok(ab.log_likelihood(ab.YULE, d, m), o.yule(v, M, 2.4)[0], fam)
        else:
            k = max(k, 2)
            a = rng.uniform(0.5, 4.0, k); M = rng.dirichlet(a, n)
            d = ab.pin(ab.Data(matrix=M)); m = ab.vector_model(a * 1.1, d)
            ll_ok(ab.log_likelihood(ab.DIRICHLET, d, m), o.dirichlet(M, a * 1.1)[0], fam)
        ab.unpin(d)
print("fuzz ok:", sum(count.values()), "cases", dict(sorted(count.items())))
