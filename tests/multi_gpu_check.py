#!/usr/bin/env python
"""Multi-GPU parity check, run under torchrun on a box with >= 2 B200s (tests/test_gpu_multi.py
launches it from pytest -m gpu when the box has them):  torchrun --nproc-per-node 2 tests/multi_gpu_check.py
Rows are sharded contiguously; every rank must return the oracle's full-data LL/score, and
bitwise the same numbers as every other rank, with the fused NVLink allreduce (default) and
with APOP_B200_ALLREDUCE=nccl."""
import os
import sys

# the fused allreduce gives up on a missing peer after this many polls (~2 s here instead of ~10 s):
# the timeout leg at the end of this script waits for it
os.environ.setdefault("APOP_B200_P2P_SPIN_LIMIT", "4000000")

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import apophenia_b200 as ab  # noqa: E402
from apophenia_b200 import abi  # noqa: E402
from apophenia_b200.shard import shard_range  # noqa: E402
from conftest import gen_probit_logit_sample  # noqa: E402
from oracle import oracle as o  # noqa: E402


def same_everywhere(arr, world):
    t = torch.tensor(np.asarray(arr, dtype=np.float64).ravel()).cuda()
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    return all(torch.equal(gathered[0], g) for g in gathered)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ab.init(local)
    ab.comm_init_from_torch()
    mode = "p2p-fused" if abi.load_library().apop_b200_comm_p2p_active() else "nccl"
    n, k = 200003, 12
    X, y, beta = gen_probit_logit_sample(n, k, seed=8)
    lo, hi = shard_range(n, world, rank)
    d = ab.pin(ab.Data(matrix=X[lo:hi], vector=y[lo:hi]))
    checks = []
    for b in (beta, np.zeros(k), 0.5 * beta):
        ll, g = ab.ll_score(ab.PROBIT, d, b)
        want = float(o.probit_ll(X, y, b, o.EXACT))
        ge, scale = o.probit_score(X, y, b, o.EXACT)
        checks.append(abs(ll - want) <= 1e-12 * abs(want))
        checks.append(bool(np.all(np.abs(g - ge.astype(float)) <= 1e-12 * scale.astype(float))))
        checks.append(same_everywhere(np.concatenate([[ll], g]), world))
    H = ab.information(ab.PROBIT, d, beta)
    Hw = o.information(0, X, y, beta).astype(float)
    checks.append(bool(np.max(np.abs(H - Hw)) <= 1e-11 * np.abs(Hw).max()))
    checks.append(same_everywhere(H, world))
    # OLS needs the global row count for the variance; Poisson regression the global lgamma constant
    yo = X @ beta + np.sin(np.arange(n))
    do = ab.pin(ab.Data(matrix=X[lo:hi], vector=yo[lo:hi]))
    m = ab.vector_model(beta * 1.1, do)
    checks.append(abs(ab.log_likelihood(ab.OLS, do, m) - float(o.ols_ll(X, yo, beta * 1.1, None, o.EXACT))) <= 1e-11 * n)
    yp = np.floor(np.abs(yo)) % 7
    dp = ab.pin(ab.Data(matrix=X[lo:hi], vector=yp[lo:hi]))
    bp = beta / 4
    checks.append(abs(ab.log_likelihood(ab.POISSON_REG, dp, ab.vector_model(bp, dp)) - float(o.poisson_reg_ll(X, yp, bp, o.EXACT)))
                  <= 1e-12 * abs(float(o.poisson_reg_ll(X, yp, bp, o.EXACT))))
    # multinomial logit and the batched path (NCCL for the large output)
    Xl, yl, B = gen_probit_logit_sample(n, 8, seed=3, method="logit", ncat=5)
    dl = ab.pin(ab.Data(matrix=Xl[lo:hi], vector=yl[lo:hi]))
    ml = ab.logit_model(B, dl)
    checks.append(abs(ab.log_likelihood(ab.LOGIT, dl, ml) - float(o.logit_ll(Xl, yl, B, mode=o.EXACT))) <= 1e-12 * abs(float(o.logit_ll(Xl, yl, B, mode=o.EXACT))))
    # the cfg2 shape of the hybrid kernel (k = 32, C = 8): score too, identical on every rank
    Xh, yh, Bh = gen_probit_logit_sample(50001, 32, seed=4, method="logit", ncat=8)
    lo2, hi2 = shard_range(50001, world, rank)
    dh = ab.pin(ab.Data(matrix=Xh[lo2:hi2], vector=yh[lo2:hi2]))
    llh, gh = ab.ll_score(ab.LOGIT, dh, Bh.ravel())
    geh, sch = o.logit_score(Xh, yh, Bh, mode=o.EXACT)
    checks.append(abs(llh - float(o.logit_ll(Xh, yh, Bh, mode=o.EXACT))) <= 1e-12 * abs(llh))
    checks.append(bool(np.all(np.abs(gh - geh.astype(float).ravel()) <= 1e-12 * sch.astype(float).ravel())))
    checks.append(same_everywhere(np.concatenate([[llh], gh]), world))
    # an iid family with cached parameter-independent sums, and the device prep: the category
    # table is the union over ranks (the last category only occurs in the last rank's shard)
    rngb = np.random.default_rng(5)
    Mb = rngb.beta(2.0, 3.0, (30000, 3))
    lo3, hi3 = shard_range(30000, world, rank)
    db = ab.pin(ab.Data(matrix=Mb[lo3:hi3]))
    wb, gwb = o.beta(None, Mb, 2.2, 3.1)
    mb = ab.vector_model([2.2, 3.1], db)
    checks.append(abs(ab.log_likelihood(ab.BETA, db, mb) - float(wb)) <= 1e-12 * abs(float(wb)))
    checks.append(bool(np.all(np.abs(ab.score(ab.BETA, db, mb) - gwb.astype(float)) <= 1e-11 * Mb.size)))
    raw = Xl.copy(); cat_vals = np.array([-1.0, 0.5, 2.0, 7.0, 11.0]); yl2 = yl.copy(); yl2[: n - 10][yl2[: n - 10] == 4] = 3
    raw[:, 0] = cat_vals[yl2.astype(int)]
    dr = ab.Data(matrix=raw[lo:hi].copy())
    cats, start = ab.prep_outcome(dr)
    checks.append(bool(np.array_equal(cats, cat_vals)))
    Xs = raw.copy(); Xs[:, 0] = 1.0
    checks.append(abs(ab.log_likelihood(ab.LOGIT, dr, ab.logit_model(B, dr)) - float(o.logit_ll(Xs, yl2, B, mode=o.EXACT))) <= 1e-12 * abs(float(o.logit_ll(Xs, yl2, B, mode=o.EXACT))))
    Bs = beta[None, :] + 0.01 * np.arange(5)[:, None]
    llb, gb = ab.ll_score_batched(ab.PROBIT, d, Bs)
    checks.append(all(abs(llb[r] - float(o.probit_ll(X, y, Bs[r], o.EXACT))) <= 1e-12 * abs(llb[r]) for r in range(5)))
    ab.bootstrap_counts(d, 4, seed=5)
    c = ab.counts_download(d, 4, hi - lo)
    tot = torch.tensor(c.sum(0).astype(np.float64)).cuda()
    dist.all_reduce(tot)
    checks.append(bool(torch.all(tot == n)))     # the resample is of the GLOBAL data set
    # multinomial observed information over the shards (P = 32 x 7: wider than the fused exchange slot,
    # so it takes the NCCL path) and identical on every rank
    Hm = ab.information(ab.LOGIT, dh, Bh.ravel())
    Hmw = o.logit_information(Xh, Bh, 8).astype(float)
    checks.append(bool(np.max(np.abs(Hm - Hmw)) <= 1e-11 * np.abs(Hmw).max()))
    checks.append(same_everywhere(Hm, world))
    # local-only mode: this rank's rows alone, and the sum over ranks of those is the fused result
    ab.set_local_only(True)
    lll, lg = ab.ll_score(ab.PROBIT, d, beta)
    ab.set_local_only(False)
    tl = torch.tensor(np.concatenate([[lll], lg])).cuda()
    dist.all_reduce(tl)
    ll_f, g_f = ab.ll_score(ab.PROBIT, d, beta)
    checks.append(bool(np.allclose(tl.cpu().numpy(), np.concatenate([[ll_f], g_f]), rtol=1e-13, atol=1e-9)))
    checks.append(abs(lll - float(o.probit_ll(X[lo:hi], y[lo:hi], beta, o.EXACT))) <= 1e-12 * abs(lll))
    # ---- the peer-timeout path of the fused exchange (common.cuh grid_finalize): rank 0 enters a reduction
    # alone; it must come back with an error (never hang, never a silent sum), stay failed until the
    # communicator is rebuilt, and work again afterwards
    if mode == "p2p-fused":
        dist.barrier()
        if rank == 0:
            ll_t, _ = ab.ll_score(ab.PROBIT, d, beta * 1.01)
            checks.append(bool(np.isnan(ll_t)) and "peer" in ab.last_error())
            ll_t2, _ = ab.ll_score(ab.PROBIT, d, beta * 1.02)        # sticky
            checks.append(bool(np.isnan(ll_t2)) and "communicator failed" in ab.last_error())
        dist.barrier()
        ab.comm_finalize()
        ab.comm_init_from_torch()
        ll_r, g_r = ab.ll_score(ab.PROBIT, d, beta)
        want = float(o.probit_ll(X, y, beta, o.EXACT))
        checks.append(abs(ll_r - want) <= 1e-12 * abs(want))
        checks.append(same_everywhere(np.concatenate([[ll_r], g_r]), world))
    ok = all(checks)
    flag = torch.tensor([1.0 if ok else 0.0]).cuda()
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multi_gpu_check world={world} allreduce={mode}: {'OK' if flag.item() == 1 else 'FAILED'} {checks}")
    ab.comm_finalize()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1 else 1)


if __name__ == "__main__":
    main()
