#!/usr/bin/env python
"""A tiny pass through every kernel of the library, for `compute-sanitizer --tool memcheck` /
`--tool racecheck` (smallest case; see B200_PROFILING.md)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import apophenia_b200 as ab  # noqa: E402

ab.init(0)
rng = np.random.default_rng(0)
for n, k in ((1000, 7), (33, 32), (257, 9)):
    X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1
    b = rng.uniform(-1, 1, k)
    y = (rng.standard_normal(n) > -(X @ b)).astype(float)
    d = ab.pin(ab.Data(matrix=X, vector=y))
    ab.ll_score(ab.PROBIT, d, b); ab.ll_score(ab.LOGIT, d, b); ab.ll_score(ab.POISSON_REG, d, b / 4)
    ab.ll_score(ab.OLS, d, b)
    ab.information(ab.PROBIT, d, b); ab.ols_gram(d); ab.probit_start(d, k); ab.dot(d, b, n)
    ab.ll_score_batched(ab.PROBIT, d, b[None, :] + 0.01 * rng.standard_normal((19, k)))
    ab.bootstrap_counts(d, 5, seed=1)
    ab.ll_score_batched(ab.PROBIT, d, b[None, :] + 0.01 * rng.standard_normal((5, k)), use_counts=True)
    ab.map_sum(d, ab.MAP_DEV2, 0.1)
    for C in (3, 8):
        yc = rng.integers(0, C, n).astype(float)
        dc = ab.pin(ab.Data(matrix=X, vector=yc))
        ab.ll_score(ab.LOGIT, dc, rng.uniform(-1, 1, k * (C - 1)))
        ab.unpin(dc)
    ab.unpin(d)
Xw = rng.uniform(-1, 1, (500, 70)); yw = (rng.uniform(size=500) > 0.5).astype(float)
dw = ab.pin(ab.Data(matrix=Xw, vector=yw)); ab.ll_score(ab.PROBIT, dw, rng.uniform(-1, 1, 70) / 8); ab.unpin(dw)
M = rng.gamma(2.0, 1.0, (777, 3))
dm = ab.pin(ab.Data(matrix=M))
for fam, p in ((ab.NORMAL, [1.0, 2.0]), (ab.POISSON, [2.0]), (ab.GAMMA, [2.0, 1.0]), (ab.LOGNORMAL, [0.1, 0.9]), (ab.EXPONENTIAL, [2.0]),
               (ab.BERNOULLI, [0.4]), (ab.ZIPF, [2.0]), (ab.TDIST, [0.5, 1.5, 4.0]), (ab.YULE, [2.5]), (ab.DIRICHLET, [1.0, 2.0, 3.0])):
    ab.log_likelihood(fam, dm, ab.vector_model(p, dm))
ab.unpin(dm)
Mb = rng.beta(2.0, 3.0, (555, 4))
db = ab.pin(ab.Data(matrix=Mb)); ab.log_likelihood(ab.BETA, db, ab.vector_model([2.0, 3.0], db)); ab.score(ab.BETA, db, ab.vector_model([2.0, 3.0], db)); ab.unpin(db)
# the multinomial logit variants (default hybrid ran above)
for var in ("fp64", "dmma"):
    pass
s = ab.synth(ab.PROBIT, 5003, 10, np.linspace(-1, 1, 10)); s.download(); s.free()
ab.finalize()
print("sanitize_smoke done")
