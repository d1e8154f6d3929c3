#!/usr/bin/env python
"""Per-pass host overhead of the fused pass: a data set so small that the kernel is a few microseconds, so the step time
is launch + completion wait + result fetch.  Also the 12.5M x 32 shard of the 8-GPU strong-scaling problem."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import apophenia_b200 as ab
ab.init(0)
rng = np.random.default_rng(8); k = 32; beta = rng.uniform(-1, 1, k)
for rows in (2048, 12_500_000):
    s = ab.synth(ab.PROBIT, rows, k, beta)
    for i in range(20): ab.ll_score(ab.PROBIT, s, beta * (1 + 1e-6 * i))
    ab.set_timing(False)
    n = 2000 if rows < 1e6 else 300
    t0 = time.perf_counter()
    for i in range(n): ab.ll_score(ab.PROBIT, s, beta * (1 + 1e-7 * (i + 30)))
    dt = (time.perf_counter() - t0) / n
    ab.pass_timer_reset()
    for i in range(100): ab.ll_score(ab.PROBIT, s, beta * (1 + 1e-7 * (i + 5000)))
    ms, c = ab.pass_timer(); ab.set_timing(False)
    print(f"rows {rows}: step {1e6 * dt:.1f} us, kernel {1e3 * ms / c:.1f} us, overhead {1e6 * dt - 1e3 * ms / c:.1f} us (python ctypes call included)", flush=True)
    s.free()
