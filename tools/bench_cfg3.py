#!/usr/bin/env python
"""BASELINE config 3: Poisson regression MLE on 400M x 16 rows sharded over 8 B200s (50M rows per GPU),
[LL | score | information] summed over ranks once per evaluation.  Run under torchrun:
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/bench_cfg3.py
Every rank runs the same host optimizer on bitwise identical sums; rank 0 prints one JSON line."""
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
    from apophenia_b200 import abi, host
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ab.init(local)
    if world > 1:
        ab.comm_init_from_torch()
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
    k = 16
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k) / np.sqrt(k)
    s = ab.synth(ab.POISSON_REG, rows, k, beta, seed=8, row_offset=rank * rows)
    h, lib = host.lib(), abi.load_library()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ab.ll_score(ab.POISSON_REG, s, beta); ab.information(ab.POISSON_REG, s, beta)   # first-use initialisation
    out = {"config": f"Poisson regression MLE, {rows * world} x {k} rows over {world} GPU(s), start at 0", "n_gpus": world}
    for method in ("Newton", "BFGS cg"):
        m = h.apop_model_copy(host.model_object(abi.POISSON_REG))
        abi.check(lib.apop_b200_register(m, abi.POISSON_REG), "register")
        m.contents.parameters = h.apop_data_alloc(k, 0, 0)
        m.contents.data = s.ptr
        keep = []
        host.mle_settings(m, method=method, tolerance=1e-10, relative_tolerance=True, max_iterations=500, starting_pt=np.zeros(k), keep=keep)
        barrier(); l0 = ab.launch_count(); t0 = time.perf_counter()
        h.apop_maximum_likelihood(s.ptr, m)
        barrier(); dt = time.perf_counter() - t0
        rep = h.apop_mle_last_report()
        est = host.parameters_of(m)
        out[method] = {"seconds": dt, "status": rep.status, "iterations": rep.iterations, "kernel_launches": int(ab.launch_count() - l0),
                       "ll": rep.ll, "gradient_norm": rep.gradient_norm, "max_abs_err_vs_beta_true": float(np.max(np.abs(est - beta)))}
    ab.pass_timer_reset()
    for i in range(20):
        ab.ll_score(ab.POISSON_REG, s, beta * (1 + 1e-6 * i))
    ms, n = ab.pass_timer()
    out["pass_kernel_ms"] = ms / n
    out["rows_per_s"] = rows * world / (ms / n * 1e-3)
    if rank == 0:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    s.free()
    if world > 1:
        ab.comm_finalize(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
