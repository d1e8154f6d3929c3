import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """the CPU restatement (test infrastructure; see oracle/oracle.h)"""
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def b200():
    """the product library bound to cuda:0; fails loudly if it is missing"""
    import apophenia_b200 as ab
    ab.init(0)
    yield ab
    ab.finalize()


def gen_probit_logit_sample(n, k, seed=8, method="probit", ncat=2):
    """the reference's generator (tests/test_apop.c:889-907, :936-939): column 0 = 1,
    other columns U(-1,1), beta_true U(-1,1); probit y = 1{N(0,1) > -xb}, logit
    y = 1{U < e^xb/(1+e^xb)}; multinomial: Categorical(softmax([0, xB]))."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (n, k))
    X[:, 0] = 1.0
    if method == "probit":
        beta = rng.uniform(-1, 1, k)
        y = (rng.standard_normal(n) > -(X @ beta)).astype(np.float64)
    elif ncat == 2:
        beta = rng.uniform(-1, 1, k)
        xb = X @ beta
        y = (rng.uniform(size=n) < np.exp(xb) / (1 + np.exp(xb))).astype(np.float64)
    else:
        beta = rng.uniform(-1, 1, (k, ncat - 1))
        z = np.concatenate([np.zeros((n, 1)), X @ beta], axis=1)
        p = np.exp(z - z.max(1, keepdims=True))
        p /= p.sum(1, keepdims=True)
        y = (rng.uniform(size=(n, 1)) > np.cumsum(p, axis=1)).sum(1).astype(np.float64)
        y = np.minimum(y, ncat - 1)
    return X, y, beta


def amash_design():
    """The design matrix of eg/logit.c over tests/golden/amash.json, outcome in column 0 as the reference
    holds it before prep: [vote factor, ideology, log(contribs+10), party dummy].  apop_data_to_factors
    numbers the sorted distinct texts (Aye = 0, No = 1); apop_data_to_dummies drops the first sorted
    category (Democrat) and appends one column for Republican."""
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "amash.json")))
    rows = g["rows"]
    votes = sorted({r["vote"] for r in rows}); parties = sorted({r["party"] for r in rows})
    assert votes == ["Aye", "No"] and parties == ["Democrat", "Republican"]
    raw = np.array([[votes.index(r["vote"]), r["ideology"], np.log(r["contribs"] + 10), parties.index(r["party"])] for r in rows], dtype=np.float64)
    return raw
