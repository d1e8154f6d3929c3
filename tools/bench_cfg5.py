#!/usr/bin/env python
"""BASELINE config 5 across GPUs: one lock-step Metropolis step of 256 chains (batched LL over a binary-logit
data set of 50M x 24 rows in total), rows sharded over the ranks, the m log-likelihoods summed over ranks
(NCCL: m doubles).  Run under torchrun; rank 0 prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    real_stdout = os.dup(1); os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    import apophenia_b200 as ab
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ab.init(local)
    if world > 1:
        ab.comm_init_from_torch()
    total, k, m = 50_000_000, 24, 256
    rows = total // world
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    s = ab.synth(ab.LOGIT, rows, k, beta, seed=8, row_offset=rank * rows)
    B = beta[None, :] + 0.01 * rng.standard_normal((m, k))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(2):
        ll, _ = ab.ll_score_batched(ab.LOGIT, s, B * (1 + 1e-7 * i), want_score=False)
    barrier(); t0 = time.perf_counter()
    reps = 10
    for i in range(reps):
        ll, _ = ab.ll_score_batched(ab.LOGIT, s, B * (1 + 1e-6 * (i + 1)), want_score=False)
    barrier(); dt = (time.perf_counter() - t0) / reps
    if world > 1:
        t = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t[0])
    # a single chain's step for scale
    barrier(); t0 = time.perf_counter()
    for i in range(20):
        ab.ll_score(ab.LOGIT, s, B[0] * (1 + 1e-6 * (i + 1)), want_score=False)
    barrier(); dt1 = (time.perf_counter() - t0) / 20
    out = {"config": f"logit Metropolis step, {m} chains in lock-step over {rows * world} x {k} rows on {world} GPU(s)", "n_gpus": world,
           "ms_per_step_all_chains": 1e3 * dt, "row_chains_per_s": rows * world * m / dt, "chain_steps_per_s": m / dt,
           "single_chain_step_ms": 1e3 * dt1, "ll_finite": bool(np.all(np.isfinite(ll)))}
    if rank == 0:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    s.free()
    if world > 1:
        ab.comm_finalize(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
