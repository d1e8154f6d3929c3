#!/usr/bin/env python
"""Breaks down the cold path (apop_b200_pin of a host apop_data): raw pinned H2D bandwidth of the
box, pageable H2D, and the library's staged upload with APOP_B200_PIN_TRACE=1."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("APOP_B200_PIN_TRACE", "1")
import torch
import apophenia_b200 as ab

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
k = 32
n = int(gb * 1e9 / ((k + 1) * 8)) // 32 * 32
ab.init(0)
torch.cuda.init()
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for name, src in (("pinned", h), ("pageable", torch.empty(1 << 30, dtype=torch.uint8).fill_(1))):
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter(); d.copy_(src, non_blocking=True); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"raw H2D 1 GiB {name}: {1.0737 / best:.1f} GB/s", flush=True)
rng = np.random.default_rng(8)
X = np.empty((n, k)); X[:] = rng.uniform(-1, 1, (1 << 16, k))[np.arange(n) % (1 << 16)] if False else 0.5
y = np.ones(n)
for thr in (os.environ.get("APOP_B200_PIN_THREADS", "16"),):
    for rep in range(3):
        dd = ab.Data(matrix=X, vector=y)
        t0 = time.perf_counter(); ab.pin(dd); dt = time.perf_counter() - t0
        print(f"pin {n}x{k} ({n*(k+1)*8e-9:.2f} GB): {dt:.3f} s = {n*(k+1)*8e-9/dt:.1f} GB/s", flush=True)
        ab.unpin(dd)
