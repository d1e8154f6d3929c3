#!/usr/bin/env python
"""Times apop_b200_pin of a CPU-written host matrix (THP or 4k pages) under the upload modes; APOP_B200_PIN_TRACE=1
prints where the time goes.  usage: pin_modes.py [GB] [thp|4k]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("APOP_B200_PIN_TRACE", "1")
import apophenia_b200 as ab
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 8.0
pages = sys.argv[2] if len(sys.argv) > 2 else "thp"
k = 32
n = int(gb * 1e9 / ((k + 1) * 8)) // 32 * 32
ab.init(0)
X = ab.host_empty((n, k), hugepages=(pages == "thp"))
X[:] = 0.5
y = np.ones(n)
for env in ({"APOP_B200_PIN_MODE": "staged"}, {"APOP_B200_PIN_MODE": "register"}, {"APOP_B200_PIN_MODE": "auto"},
            {"APOP_B200_PIN_MODE": "register", "APOP_B200_PIN_AHEAD": "6"},
            {"APOP_B200_PIN_MODE": "register", "APOP_B200_PIN_WINDOW_MB": "512"},
            {"APOP_B200_PIN_MODE": "register", "APOP_B200_PIN_WINDOW_MB": "32"},
            {"APOP_B200_PIN_MODE": "auto", "APOP_B200_PIN_THREADS_NOTE": "2 gather threads, as one of 8 ranks would have"}):
    for kk in ("APOP_B200_PIN_DEFER_UNREGISTER", "APOP_B200_PIN_AHEAD", "APOP_B200_PIN_WINDOW_MB"):
        os.environ.pop(kk, None)
    os.environ.update(env)
    for rep in range(2):
        d = ab.Data(matrix=X, vector=y)
        t0 = time.perf_counter(); ab.pin(d); dt = time.perf_counter() - t0
        t0 = time.perf_counter(); ab.unpin(d); du = time.perf_counter() - t0
        print(f"{env} {pages}: {n * (k + 1) * 8e-9 / dt:.1f} GB/s ({dt:.3f} s) mode={ab.last_pin_mode()} unpin {du:.3f} s", flush=True)
