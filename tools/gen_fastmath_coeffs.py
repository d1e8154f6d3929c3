#!/usr/bin/env python
"""Generates the polynomial coefficients of apophenia_b200/csrc/fastmath.cuh with mpmath
(50 digits): Chebyshev interpolation, converted to the monomial basis, error checked on a
dense grid.  Run: python tools/gen_fastmath_coeffs.py > /tmp/coeffs.txt"""
import mpmath as mp

mp.mp.dps = 60


def cheb_fit(f, n, a=-1, b=1):
    """degree-n interpolant of f on [a,b] at Chebyshev nodes, returned as monomial coeffs in
    the variable u in [-1,1] (u = (2x-a-b)/(b-a))"""
    N = n + 1
    nodes = [mp.cos(mp.pi * (k + mp.mpf(1) / 2) / N) for k in range(N)]
    fv = [f((a + b) / 2 + (b - a) / 2 * t) for t in nodes]
    c = []
    for j in range(N):
        s = sum(fv[k] * mp.cos(mp.pi * j * (k + mp.mpf(1) / 2) / N) for k in range(N))
        c.append(2 * s / N)
    c[0] /= 2
    # Chebyshev -> monomial
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, N):
        t = [mp.mpf(0)] + [2 * v for v in T[j - 1]]
        for i, v in enumerate(T[j - 2]):
            t[i] -= v
        T.append(t)
    mono = [mp.mpf(0)] * N
    for j in range(N):
        for i, v in enumerate(T[j]):
            mono[i] += c[j] * v
    return mono


def horner(co, u):
    r = mp.mpf(0)
    for v in reversed(co):
        r = r * u + v
    return r


def maxerr(f, co, pts, rel=True, a=-1, b=1):
    worst = 0
    for x in pts:
        u = (2 * x - a - b) / (b - a)
        e = abs(horner(co, u) - f(x))
        if rel:
            e /= abs(f(x))
        worst = max(worst, e)
    return worst


def emit(name, co):
    print(f"// {name}: {len(co)} coefficients, c[0] + c[1] u + ...")
    print("static constexpr double " + name + "[] = {")
    for v in co:
        print("    " + mp.nstr(v, 20, min_fixed=0, max_fixed=0) + ",")
    print("};")


if __name__ == "__main__":
    # 1) Mills-type function: Q(q) = erfcx(x) * (x + 4), x = 4 (1+q)/(1-q), q in [-1, 1)
    def Q(q):
        if q >= 1:
            return 1 / mp.sqrt(mp.pi)
        x = 4 * (1 + q) / (1 - q)
        return mp.erfc(x) * mp.exp(x * x) * (x + 4)
    for deg in (22, 24, 26, 28):
        co = cheb_fit(Q, deg)
        pts = [mp.mpf(i) / 2000 * 2 - 1 for i in range(0, 2000)]
        e = maxerr(Q, co, pts)
        print(f"// erfcx*(x+4) in q=(x-4)/(x+4): degree {deg} max rel err {mp.nstr(e, 3)}")
        if e < mp.mpf(3) * mp.mpf(10) ** -17:
            emit("ERFCX_Q", co)
            break
    # 2) exp(r) on |r| <= ln2/2 : exp(r) = 1 + r + r^2 * P(r)
    h = mp.log(2) / 2
    def P(r):
        if abs(r) < mp.mpf(10) ** -20:
            return mp.mpf(1) / 2 + r / 6
        return (mp.exp(r) - 1 - r) / (r * r)
    for deg in (9, 10, 11):
        co = cheb_fit(P, deg, -h, h)
        # convert to coefficients in r: u = r/h
        cor = [v / h ** i for i, v in enumerate(co)]
        pts = [(-h + 2 * h * i / 1000) for i in range(1001)]
        worst = max(abs((1 + r + r * r * horner(cor, r)) - mp.exp(r)) / mp.exp(r) for r in pts)
        print(f"// exp: degree {deg} max rel err {mp.nstr(worst, 3)}")
        if worst < mp.mpf(3) * mp.mpf(10) ** -17:
            emit("EXP_P", cor)
            break
    # 3) log(m), m in [sqrt(1/2), sqrt(2)): s = (m-1)/(m+1), log m = 2 s + s^3 * L(s^2)
    smax = (mp.sqrt(2) - 1) / (mp.sqrt(2) + 1)
    def L(w):  # w = s^2
        if abs(w) < mp.mpf(10) ** -30:
            return mp.mpf(2) / 3 + 2 * w / 5
        s = mp.sqrt(w)
        return (2 * mp.atanh(s) - 2 * s) / (s * w)
    for deg in (5, 6, 7, 8):
        co = cheb_fit(L, deg, 0, smax ** 2)
        a, b = 0, smax ** 2
        # monomial in w: u = (2w - a - b)/(b - a)
        # expand via polynomial composition
        import itertools
        comp = [mp.mpf(0)] * (deg + 1)
        alpha, beta = 2 / (b - a), -(a + b) / (b - a)
        pw = [mp.mpf(1)]
        for i, v in enumerate(co):
            for j, c in enumerate(pw):
                comp[j] += v * c
            nxt = [mp.mpf(0)] * (len(pw) + 1)
            for j, c in enumerate(pw):
                nxt[j] += c * beta
                nxt[j + 1] += c * alpha
            pw = nxt
        pts = [smax * (mp.mpf(i) / 1000 * 2 - 1) for i in range(1, 1001)]
        worst = 0
        for s in pts:
            if s == 0:
                continue
            val = 2 * s + s ** 3 * horner(comp, s * s)
            worst = max(worst, abs(val - 2 * mp.atanh(s)) / abs(2 * mp.atanh(s)))
        print(f"// log: degree {deg} max rel err {mp.nstr(worst, 3)}")
        if worst < mp.mpf(2) * mp.mpf(10) ** -17:
            emit("LOG_L", comp)
            break
