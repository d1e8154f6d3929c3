#!/usr/bin/env python
"""Quick kernel timing + parity for one configuration (development loop).
usage: kbench.py logit8|batched4|batched5|probit|wide:K|info8 [reps]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import apophenia_b200 as ab
from oracle import oracle as o

def timed(fn, reps=20, warm=3):
    for i in range(warm): fn(-i - 1)
    ab.pass_timer_reset()
    t0 = time.perf_counter()
    for i in range(reps): fn(i)
    dt = (time.perf_counter() - t0) / reps
    ms, n = ab.pass_timer()
    return 1e3 * dt, ms / max(n, 1)

def logit8(rows=20_000_000, k=32, C=8):
    rng = np.random.default_rng(8); B = rng.uniform(-1, 1, k * (C - 1))
    w = ab.synth(ab.LOGIT, 20011, k, B, ncat=C); X, y = w.download()
    ll, g = ab.ll_score(ab.LOGIT, w, B); w.free()
    want = o.logit_ll(X, y, B.reshape(k, C - 1), mode=o.EXACT); ge, sc = o.logit_score(X, y, B.reshape(k, C - 1), mode=o.EXACT)
    e1 = abs(ll - float(want)) / abs(float(want)); e2 = float(np.max(np.abs(g - ge.astype(float).ravel()) / sc.astype(float).ravel()))
    s = ab.synth(ab.LOGIT, rows, k, B, ncat=C)
    dt, km = timed(lambda i: ab.ll_score(ab.LOGIT, s, B * (1 + 1e-6 * (i + 1))))
    s.free()
    print(f"logit8 {rows}x{k} C={C}: step {dt:.4f} ms kernel {km:.4f} ms = {rows * (k + 1) * 8 / km / 1e6:.0f} GB/s; parity ll {e1:.1e} score {e2:.1e}", flush=True)

def batched(fam, rows, k, m, score, counts, spread=1e-3):
    rng = np.random.default_rng(8); beta = rng.uniform(-1, 1, k)
    Bs = beta[None, :] + spread * rng.standard_normal((m, k))
    w = ab.synth(fam, 30011, k, beta); X, y = w.download()
    if counts: ab.bootstrap_counts(w, 16, seed=3); c = ab.counts_download(w, 16, 30011)
    llb, gb = ab.ll_score_batched(fam, w, Bs[:16], use_counts=counts, want_score=score); w.free()
    errs = []
    for r in (0, 7, 15):
        if counts:
            idx = np.repeat(np.arange(30011), c[:, r]); Xr, yr = X[idx], y[idx]
        else: Xr, yr = X, y
        if fam == ab.PROBIT: want = o.probit_ll(Xr, yr, Bs[r], o.EXACT); ge, sc = o.probit_score(Xr, yr, Bs[r], o.EXACT)
        else: want = o.logit_ll(Xr, yr, Bs[r].reshape(-1, 1), mode=o.EXACT); ge, sc = o.logit_score(Xr, yr, Bs[r].reshape(-1, 1), mode=o.EXACT); ge, sc = ge.ravel(), sc.ravel()
        errs.append(abs(llb[r] - float(want)) / abs(float(want)))
        if score: errs.append(float(np.max(np.abs(gb[r] - ge.astype(float)) / sc.astype(float))))
    s = ab.synth(fam, rows, k, beta)
    if counts: ab.bootstrap_counts(s, m, seed=8)
    dt, km = timed(lambda i: ab.ll_score_batched(fam, s, Bs * (1 + 1e-6 * (i + 1)), use_counts=counts, want_score=score), reps=3, warm=1)
    s.free()
    print(f"batched fam={fam} {rows}x{k} m={m} score={score} counts={counts} spread={spread}: step {dt:.3f} ms kernel {km:.3f} ms; parity max {max(errs):.1e}", flush=True)

def probit(rows=100_000_000, k=32):
    rng = np.random.default_rng(8); beta = rng.uniform(-1, 1, k)
    s = ab.synth(ab.PROBIT, rows, k, beta)
    dt, km = timed(lambda i: ab.ll_score(ab.PROBIT, s, beta * (1 + 1e-6 * (i + 1))))
    s.free()
    print(f"probit {rows}x{k}: step {dt:.4f} ms kernel {km:.4f} ms = {rows * (k + 1) * 8 / km / 1e6:.0f} GB/s", flush=True)

def wide(k, rows=None):
    rng = np.random.default_rng(8); beta = rng.uniform(-1, 1, k) / np.sqrt(k / 32)
    def make(n):
        if k <= 480: return ab.synth(ab.PROBIT, n, k, beta), None
        X = rng.random((n, k)) * 2 - 1; X[:, 0] = 1   # the data generator stops at 480 parameters: host data, pinned
        y = (rng.standard_normal(n) > -(X @ beta)).astype(float)
        d = ab.pin(ab.Data(matrix=X, vector=y))
        return d, (X, y)
    w, h = make(10007)
    X, y = h if h else w.download()
    ll, g = ab.ll_score(ab.PROBIT, w, beta)
    w.free() if h is None else ab.unpin(w)
    want = o.probit_ll(X, y, beta, o.EXACT); ge, sc = o.probit_score(X, y, beta, o.EXACT)
    e1 = abs(ll - float(want)) / abs(float(want)); e2 = float(np.max(np.abs(g - ge.astype(float)) / sc.astype(float)))
    rows = rows or int((5.2e9 if k <= 480 else 2.0e9) / ((k + 1) * 8))
    s, h = make(rows)
    dt, km = timed(lambda i: ab.ll_score(ab.PROBIT, s, beta * (1 + 1e-6 * (i + 1))))
    s.free() if h is None else ab.unpin(s)
    print(f"probit wide {rows}x{k}: step {dt:.4f} ms kernel {km:.4f} ms = {rows * (k + 1) * 8 / km / 1e6:.0f} GB/s; parity ll {e1:.1e} score {e2:.1e}", flush=True)

def info8(rows=20_000_000, k=32, C=8):
    rng = np.random.default_rng(8); B = rng.uniform(-1, 1, k * (C - 1))
    s = ab.synth(ab.LOGIT, rows, k, B, ncat=C)
    dt, km = timed(lambda i: ab.information(ab.LOGIT, s, B * (1 + 1e-6 * (i + 1))), reps=3, warm=1)
    s.free()
    print(f"info8 {rows}x{k}: call {dt:.3f} ms, kernel per launch {km:.3f} ms", flush=True)

if __name__ == "__main__":
    ab.init(0)
    for nm in sys.argv[1:]:
        if nm == "logit8": logit8()
        elif nm.startswith("batched4"): batched(ab.PROBIT, 5_000_000, 20, 1000, True, True, float(nm.split(":")[1]) if ":" in nm else 1e-3)
        elif nm.startswith("batched5"): batched(ab.LOGIT, 50_000_000, 24, 256, False, False, float(nm.split(":")[1]) if ":" in nm else 1e-4)
        elif nm == "probit": probit()
        elif nm.startswith("wide:"): wide(int(nm.split(":")[1]))
        elif nm == "info8": info8()
