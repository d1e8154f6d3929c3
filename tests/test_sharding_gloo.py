"""The N>1 path on CPU (gloo, world_size 2): rows are sharded contiguously, every rank
reduces its slice, one allreduce(sum) of [LL | score] makes the result global, and the
replicated host optimizer stays in lock-step.  The per-shard reduction is played by the
oracle here; on the GPU box the same plumbing carries the device partials over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import gen_probit_logit_sample


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as o
    from apophenia_b200.shard import shard_range
    X, y, beta = gen_probit_logit_sample(20001, 5, seed=8)
    lo, hi = shard_range(len(y), world, rank)
    Xs, ys = X[lo:hi], y[lo:hi]

    def fdf(b):
        ll = o.probit_ll(Xs, ys, b, o.EXACT)
        g, scale = o.probit_score(Xs, ys, b, o.EXACT)
        t = torch.tensor(np.concatenate([[float(ll)], g.astype(float), scale.astype(float)]))
        dist.all_reduce(t)                    # [LL | score | scale], summed over ranks
        return t[0].item(), t[1:6].numpy(), t[6:].numpy()

    ll, g, scale = fdf(beta)
    # a few replicated steepest-ascent steps: identical on every rank because the
    # allreduced values are bitwise identical on every rank
    b = beta.copy()
    for _ in range(3):
        _, gg, _ = fdf(b)
        b = b + 1e-6 * gg
    q.put((rank, lo, hi, ll, g, scale, b))
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_unsharded(oracle):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=120) for _ in procs])
    [p.join(30) for p in procs]
    X, y, beta = gen_probit_logit_sample(20001, 5, seed=8)
    want_ll = float(oracle.probit_ll(X, y, beta, oracle.EXACT))
    want_g, scale = oracle.probit_score(X, y, beta, oracle.EXACT)
    (r0, lo0, hi0, ll0, g0, s0, b0), (r1, lo1, hi1, ll1, g1, s1, b1) = res
    assert (lo0, hi0, lo1, hi1) == (0, 10001, 10001, 20001)      # contiguous, remainder to the low ranks
    assert ll0 == ll1 and np.array_equal(g0, g1) and np.array_equal(b0, b1)   # lock-step replicas
    assert abs(ll0 - want_ll) <= 1e-12 * abs(want_ll)
    assert np.all(np.abs(g0 - want_g.astype(float)) <= 1e-12 * scale.astype(float))


def test_shard_bounds():
    from apophenia_b200.shard import shard_bounds, shard_range
    assert shard_bounds(10, 3) == [0, 4, 7, 10]
    assert shard_bounds(0, 2) == [0, 0, 0]
    assert shard_range(100_000_000, 8, 7) == (87_500_000, 100_000_000)
    for n in (1, 31, 32, 33, 1000):
        b = shard_bounds(n, 8)
        assert b[0] == 0 and b[-1] == n and all(b[i] <= b[i + 1] for i in range(8))
