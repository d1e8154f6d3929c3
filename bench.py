#!/usr/bin/env python
"""bench.py -- likelihood+score rows/s of the fused probit pass (FP64, % of HBM roofline).

Contract: `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line.
A step is one fused log-likelihood + score pass over the resident data set (one
optimizer evaluation: beta in, [LL | score] out, summed over ranks).

  value     rows/s with X, y resident in HBM (the design point named by BASELINE.json: the
            matrix stays resident between optimizer iterations); weak scaling, 100M x 32 per GPU
  e2e       the whole job from HOST memory through the reference-facing calls: a host
            apop_data is pinned (every byte of X and y crosses PCIe and is tiled) and a probit
            MLE runs over the registered model->log_likelihood / apop_score entries
  roofline  algorithmic bytes N*(k+1)*8 per launch / CUDA-event kernel time, against
            MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the restated reference CPU path (oracle/) timed on this box's host cores
  parity    in-run correctness at THIS N: (a) the fused cross-GPU sum of [LL | score] against an
            independent torch.distributed all-reduce of the ranks' local results; (b) a 40k-row
            window of every rank's shard against the oracle (EXACT mode) at 1e-12
  strong    the FIXED 100M x 32 problem split over the N ranks: per-pass ms, BFGS / Newton
            time-to-fit, and the single-GPU pass of the same problem timed in the same run
  configs   the other BASELINE.json configurations at this N (cfg2 logit K=8 20M x 32, cfg3
            Poisson-regression shard 50M x 16 per GPU, cfg4 bootstrap 1000 x 5M x 20, cfg5
            256 Metropolis chains over 50M x 24), each with its roofline and a restated-CPU figure

`--impl reference` times the restated CPU path alone as the reference arm.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "probit likelihood+score rows/sec (FP64)"
UNIT = "rows/s"
FIXED_ROWS = 100_000_000


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=FIXED_ROWS, help="rows per GPU (weak scaling)")
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--family", default="probit", choices=["probit", "logit", "poisson_reg", "ols"])
    ap.add_argument("--ncat", type=int, default=2, help="outcome categories (logit)")
    ap.add_argument("--cpu-rows", type=int, default=3_000_000, help="row sample of the CPU baseline")
    ap.add_argument("--e2e-rows", type=int, default=0, help="rows of the host data set for e2e (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--configs", default="cfg2,cfg3,cfg4,cfg5", help="which BASELINE configurations to run")
    ap.add_argument("--host-pages", default="pinned", choices=["pinned", "thp", "4k"], help="memory behind the e2e host matrix")
    ap.add_argument("--sustain-seconds", type=float, default=2.0, help="length of the back-to-back (sustained) run")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML in a thread,
    every ~5 ms; the recipe's nvidia-smi line needs a process spawn per 100 ms sample)."""

    def __init__(self, index):
        self.index = index
        self.sm, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.t = None

    def _run(self):
        import pynvml as nv
        try:
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            while not self._stop.is_set():
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
                time.sleep(0.005)
        except Exception as e:  # noqa: BLE001
            self.reasons.add(f"nvml unavailable: {e}")

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self._stop.set()
        if self.t:
            self.t.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


FAM = {"probit": 0, "logit": 1, "poisson_reg": 5, "ols": 4}


def host_threads(o):
    # torchrun exports OMP_NUM_THREADS=1: use every host core anyway
    return max(o.max_threads(), os.cpu_count() or 1)


def run_reference(args):
    """the reference arm: restated CPU path, all host threads, a bounded row sample per step"""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as o
    n, k = args.cpu_rows, args.k
    rng = np.random.default_rng(8)
    X = rng.uniform(-1, 1, (n, k)); X[:, 0] = 1.0
    beta = rng.uniform(-1, 1, k)
    y = (rng.standard_normal(n) > -(X @ beta)).astype(np.float64)
    threads = host_threads(o)
    for _ in range(max(args.warmup, 1)):
        o.time_probit_ll_score(X[: n // 8], y[: n // 8], beta, threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        o.time_probit_ll_score(X, y, beta * (1 + 1e-9 * s), threads)
    dt = time.perf_counter() - t0
    val = n * args.steps / dt
    sample = (f"each step is one LL+score pass over a {n}-row x {k}-column SAMPLE of the workload (the full "
              f"{args.rows} rows would take ~{args.rows / val:.0f} s per step on these cores); probit, restated "
              f"reference CPU path, gcc -O3 -fopenmp; rows/s normalises the sizes")
    cfg = workload_config(args, args.gpus)   # identical to the b200 arm's config; the sample size is stated beside it
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cfg, "sample_rows_per_step": n,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world):
    fam = args.family + (f" (C={args.ncat})" if args.family == "logit" else "")
    return {"workload": f"{fam} fused LL+score pass, {args.rows} rows x {args.k} cols per GPU "
                        f"(target 100M x 32 probit MLE of BASELINE.json north_star)",
            "rows_per_gpu": args.rows, "k": args.k, "global_rows": args.rows * world,
            "parallelism": f"row-sharded dp{world}; [LL|score] summed over ranks once per pass, inside the reduction "
                           f"kernel over NVLink peer memory (APOP_B200_ALLREDUCE=nccl: one ncclAllReduce instead)",
            "l2": "inputs (>= 26 GB per pass) far exceed the 126 MB L2; no flush needed"}


class Ctx:
    """what every leg of the benchmark needs"""

    def __init__(self, ab, torch, dist, world, rank, args):
        self.ab, self.torch, self.dist, self.world, self.rank, self.args = ab, torch, dist, world, rank, args

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, *vals):
        if self.dist is None:
            return [float(v) for v in vals]
        t = self.torch.tensor(list(vals), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def sum_over_ranks(self, arr):
        if self.dist is None:
            return np.asarray(arr, dtype=np.float64)
        t = self.torch.tensor(np.asarray(arr, dtype=np.float64), device="cuda", dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def time_passes(self, fn, steps, warmup=3):
        """steps calls of fn(i), barrier + sync on both sides, max over ranks; returns (seconds per call, kernel ms)"""
        ab = self.ab
        for i in range(warmup):
            fn(-i - 2)
        ab.set_timing(False)           # the step time is taken without the per-pass event records
        self.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        self.barrier()
        dt = (time.perf_counter() - t0) / steps
        ab.pass_timer_reset()          # ... and the kernel time in a second loop with them
        for i in range(steps):
            fn(steps + i)
        kms, n = ab.pass_timer()
        ab.set_timing(False)
        dt, kms = self.max_over_ranks(dt, kms / max(n, 1))
        return dt, kms


def roofline_of(bytes_per_launch, kernel_ms, kernel, bound="hbm", extra=None):
    peak, peak_src = measured_peak()
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    r = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
         "peak_source": peak_src, "kernel": kernel, "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": bytes_per_launch,
         "frac_of_nominal_8TBs": achieved / 8000.0}
    if extra:
        r.update(extra)
    return r


def measured_traffic(roofline, kernel, rows, k):
    """DRAM bytes per launch of this kernel at this shape from the committed ncu --set full capture (profiles/traffic.json);
    a run cannot count DRAM bytes itself without a profiler attached, so the figure is looked up, and only for the exact shape"""
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        tj = json.load(open(prof))
    except Exception:
        return
    for e in [tj] + list(tj.get("others", [])):
        if e.get("rows") == rows and e.get("k") == k and str(e.get("kernel", "")).startswith(kernel):
            roofline["traffic"] = e.get("dram_bytes_per_launch")
            roofline["traffic_source"] = e.get("source", "profiles/traffic.json")
            return


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    # stdout carries exactly one JSON line: anything libraries print while initialising
    # (e.g. torch's "NCCL version ..." banner) is sent to stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    # the ranks of one box share its host cores: split them between the ranks' upload pools
    os.environ.setdefault("APOP_B200_PIN_THREADS", str(max(1, (os.cpu_count() or 16) // max(world, 1))))

    import torch
    import apophenia_b200 as ab

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ab.init(local)
    if world > 1:
        ab.comm_init_from_torch()
    cx = Ctx(ab, torch, dist, world, rank, args)

    fam = FAM[args.family]
    n, k = args.rows, args.k
    rng = np.random.default_rng(8)
    beta_true = rng.uniform(-1, 1, k * (args.ncat - 1) if args.family == "logit" else k)
    if args.family == "poisson_reg":
        beta_true /= np.sqrt(k)
    data = ab.synth(fam, n, k, beta_true, ncat=args.ncat, seed=8, row_offset=rank * n)

    def step(i):
        # a new beta every step (an optimizer never evaluates the same point twice)
        b = beta_true * (1.0 + 1e-6 * (i + 1))
        return ab.ll_score(fam, data, b)

    # ---- the headline: K timed steps, per-pass CUDA events on the library stream inside the timed region
    for i in range(args.warmup):
        step(-i - 2)
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    nvml_index = int(vis.split(",")[local]) if vis and vis.split(",")[local].isdigit() else local
    sampler = ClockSampler(nvml_index)
    cx.barrier()
    if rank == 0:
        sampler.start()
    ab.pass_timer_reset()        # switches the per-pass events on: the roofline is measured live in this region
    launches0 = ab.launch_count()
    t0 = time.perf_counter()
    for i in range(args.steps):
        ll, g = step(i)
    cx.barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    kern_ms, n_pass = ab.pass_timer()
    ab.set_timing(False)
    launches = ab.launch_count() - launches0
    assert np.isfinite(ll) and np.all(np.isfinite(g)), "non-finite result"
    dt, kern_ms = cx.max_over_ranks(dt, kern_ms)
    value = n * world * args.steps / dt

    bytes_per_launch = n * (k + 1) * 8
    kern_ms_avg = kern_ms / max(n_pass, 1)
    kernel = "logit_stream_kernel" if (args.family == "logit" and args.ncat > 2) else "glm_fused_kernel"
    roofline = roofline_of(bytes_per_launch, kern_ms_avg, kernel)
    measured_traffic(roofline, kernel, n, k)

    # ---- the same pass back to back for a few seconds (sustained clocks, not the burst of K steps)
    sustained = None
    if args.sustain_seconds > 0:
        reps = max(args.steps, int(args.sustain_seconds / max(dt / args.steps, 1e-6)))
        s2 = ClockSampler(nvml_index)
        cx.barrier()
        if rank == 0:
            s2.start()
        t0 = time.perf_counter()
        for i in range(reps):
            step(args.steps + i)
        cx.barrier()
        sdt, = cx.max_over_ranks(time.perf_counter() - t0)
        sc = s2.stop() if rank == 0 else None
        sustained = {"steps": reps, "seconds": sdt, "ms_per_step": 1e3 * sdt / reps, "value": n * world * reps / sdt,
                     "GBps_per_gpu": bytes_per_launch * reps / sdt / 1e9, "clocks": sc}

    parity = None if args.no_parity else run_parity(cx, fam, data, beta_true)
    e2e = None if args.no_e2e else run_e2e(cx, fam, beta_true)
    fit = None
    if not args.no_fit and args.family == "probit":
        fit = run_time_to_fit(cx, data, k, beta_true)
    strong = None
    if not args.no_strong and args.family == "probit":
        strong = run_strong(cx, data, beta_true)
    cpu = None
    if rank == 0 and not args.no_cpu and args.family == "probit":
        cpu = run_cpu_baseline(ab, args, fam, beta_true)
    data.free()
    configs = None
    if not args.no_configs and args.family == "probit":
        configs = run_configs(cx, [c for c in args.configs.split(",") if c])

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic (device Philox, seed 8)",
               "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
               "gpu_launches": launches, "clocks": clocks, "impl": "b200", "time_to_fit": fit, "sustained": sustained,
               "parity": parity, "strong": strong, "configs": configs}
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if dist is not None:
        ab.comm_finalize()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
#  parity probe (untimed), at every N
# ------------------------------------------------------------------------------------------------
def run_parity(cx, fam, data, beta_true):
    """(a) fused cross-GPU [LL | score] vs an independent all-reduce of the ranks' local results;
    (b) a 40k-row window of this rank's shard vs the oracle in EXACT mode (1e-12)."""
    ab, args = cx.ab, cx.args
    k = args.k
    b = beta_true * (1.0 + 3e-6)
    ll, g = ab.ll_score(fam, data, b)
    fused = np.concatenate([[ll], g])
    ab.set_local_only(True)
    lll, lg = ab.ll_score(fam, data, b)
    ab.set_local_only(False)
    local = np.concatenate([[lll], lg])
    total = cx.sum_over_ranks(local)                       # torch.distributed (NCCL) all-reduce, independent of the library's exchange
    scale = cx.sum_over_ranks(np.abs(local))
    exch = float(np.max(np.abs(fused - total) / np.maximum(scale, 1e-300)))
    # all ranks must hold bitwise the same fused result (rank-order summation)
    spread = cx.max_over_ranks(*fused)
    spread_min = [-v for v in cx.max_over_ranks(*(-fused))]
    bitwise = bool(np.array_equal(np.asarray(spread), np.asarray(spread_min)))
    out = {"n_gpus": cx.world, "exchange_max_rel_err": exch, "exchange_tolerance": 1e-14, "exchange_ok": exch <= 1e-14,
           "fused_result_identical_on_all_ranks": bitwise,
           "exchange_what": "max_j |fused_j - allreduce(local)_j| / allreduce(|local_j|) over [LL | score], one probit pass at 100M rows per rank"}
    # (b) oracle window: the first 40k rows of this rank's shard, regenerated (same Philox counters) and downloaded
    if args.family == "probit":
        from oracle import oracle as o
        nw = 40_000
        win = ab.synth(fam, nw, k, beta_true, seed=8, row_offset=cx.rank * args.rows)
        X, y = win.download()
        ab.set_local_only(True)
        wll, wg = ab.ll_score(fam, win, b)
        ab.set_local_only(False)
        win.free()
        want = o.probit_ll(X, y, b, o.EXACT)
        ge, gs = o.probit_score(X, y, b, o.EXACT)
        e_ll = float(abs(np.longdouble(wll) - want) / abs(want))
        e_g = float(np.max(np.abs(np.asarray(wg, dtype=np.longdouble) - ge) / gs))
        e_ll, e_g = cx.max_over_ranks(e_ll, e_g)
        out.update({"oracle_window_rows": nw, "oracle_ll_rel_err": e_ll, "oracle_score_rel_err": e_g, "oracle_tolerance": 1e-12,
                    "oracle_ok": bool(e_ll <= 1e-12 and e_g <= 1e-12),
                    "oracle_what": "first 40k rows of every rank's shard vs oracle/ EXACT mode; score error relative to sum_i |X_ij d_i|; max over ranks"})
    return out


# ------------------------------------------------------------------------------------------------
#  e2e: host apop_data -> pin -> MLE through the registered entries
# ------------------------------------------------------------------------------------------------
def run_e2e(cx, fam, beta_true):
    import psutil
    ab, args, world, rank, dist, torch = cx.ab, cx.args, cx.world, cx.rank, cx.dist, cx.torch
    k = args.k
    n = args.e2e_rows or args.rows
    need = n * (k + 1) * 8
    avail = psutil.virtual_memory().available
    if need * 2.5 > avail / max(world, 1):
        n = int(avail / max(world, 1) / 2.5 / ((k + 1) * 8))
    n = max(32, n // 32 * 32)
    if dist is not None:   # every rank must hold the same number of rows: take the smallest estimate
        t = torch.tensor([n], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        n = int(t[0])
    # host data set: the same synthetic recipe, generated on the device and copied back
    X, y = host_dataset(ab, fam, n, k, beta_true, args.ncat, rank * n, args.host_pages)
    d = ab.Data(matrix=X, vector=y)
    cx.barrier()
    t0 = time.perf_counter()
    ab.pin(d)                      # one-off: host->device copy of X, y + tiling
    pin_s = time.perf_counter() - t0
    pin_mode, reg_rate = ab.last_pin_mode()
    mk = {0: ab.probit_model, 1: ab.logit_model}.get(fam, ab.vector_model)
    shape = (lambda b: b.reshape(k, -1)) if fam == 1 else (lambda b: b)
    for i in range(2):
        m = mk(shape(beta_true * (1 + 1e-7 * (i + 1))), d)
        ab.log_likelihood(fam, d, m); ab.score(fam, d, m)
    cx.barrier()
    steps = args.steps
    t0 = time.perf_counter()
    for i in range(steps):
        m = mk(shape(beta_true * (1.0 + 1e-6 * (i + 1))), d)
        ll = ab.log_likelihood(fam, d, m)      # model->log_likelihood(d, m)
        g = ab.score(fam, d, m)                # apop_score vtable entry (memo: same pass)
    cx.barrier()
    dt = time.perf_counter() - t0
    dt, pin_s = cx.max_over_ranks(dt, pin_s)
    assert np.isfinite(ll) and np.all(np.isfinite(g))
    ab.unpin(d)
    # the whole job as a user runs it: host apop_data -> pin -> MLE (BFGS) -> estimates, every byte of
    # X and y crossing PCIe inside the timed region
    fit = None
    if fam == 0 and not args.no_fit:
        fit = run_fit_from_host(cx, d, k)
    resident = {"value": n * world * steps / dt, "unit": UNIT, "ms_per_step": 1e3 * dt / steps,
                "h2d_bytes_per_step": beta_true.size * 8 + 32, "d2h_bytes_per_step": (3 + beta_true.size) * 8,
                "note": "X and y already resident (BASELINE.json north_star keeps them in HBM between optimizer "
                        "iterations): per step only beta goes down and [LL|score] comes back"}
    cold = {"pin_seconds": pin_s, "h2d_bytes": n * (k + 1) * 8, "pin_GBps": n * (k + 1) * 8e-9 / pin_s,
            "pin_GBps_all_gpus": world * n * (k + 1) * 8e-9 / pin_s, "pin_mode": pin_mode, "page_lock_GBps_first_window": reg_rate, "host_pages": args.host_pages,
            "pin_mode_env": os.environ.get("APOP_B200_PIN_MODE", "auto"),
            "first_pass_rows_per_s": n * world / (pin_s + dt / steps)}
    api = "apop_b200_pin + apop_maximum_likelihood (host mirror of the reference driver) over the registered " \
          "apop_b200_<family>_ll / _score entries, on a HOST apop_data"
    if fit and fit["status"] == 0:
        passes = max(int(fit["passes"]), 1)
        cold["pin_GBps_in_fit"] = n * (k + 1) * 8e-9 / fit["seconds_pin"]
        return {"value": n * world * passes / fit["seconds_total"], "unit": UNIT,
                "h2d_bytes_per_step": (n * (k + 1) * 8) / passes + beta_true.size * 8 + 32,
                "d2h_bytes_per_step": (3 + beta_true.size) * 8, "rows_per_gpu": n, "steps": passes,
                "ms_per_step": 1e3 * fit["seconds_total"] / passes, "api": api,
                "what": "host apop_data -> pin (H2D of X, y + tiling) -> probit MLE (BFGS) -> estimates; rows x optimizer "
                        "evaluations / total seconds", "fit_from_host": fit, "resident_step": resident, "cold": cold}
    return dict(resident, rows_per_gpu=n, api=api, cold=cold, fit_from_host=fit)


def host_dataset(ab, fam, n, k, beta_true, ncat, row_offset, pages):
    """The e2e host data set: the synthetic recipe generated on the device window by window, copied back, and
    written into the host matrix by the CPU -- as a caller that parsed or computed its data would have.
    pages == "pinned" (default; the bench contract copies its inputs from pinned host memory): the matrix block comes
    from apop_b200_host_alloc (cudaHostAlloc), so apop_b200_pin lets the DMA engine read it in place; "thp" / "4k":
    pageable memory (huge or ordinary pages), which the library gathers through pinned staging buffers."""
    X = ab.host_empty((n, k), hugepages=(pages == "thp"), pinned=(pages == "pinned"))
    y = np.empty(n)
    chunk = 4_000_000
    for lo in range(0, n, chunk):
        rows = min(chunk, n - lo)
        src = ab.synth(fam, rows, k, beta_true, ncat=ncat, seed=8, row_offset=row_offset + lo)
        Xc, yc = src.download()
        src.free()
        X[lo:lo + rows] = Xc
        y[lo:lo + rows] = yc
    return X, y


def run_fit_from_host(cx, d, k):
    """pin + apop_maximum_likelihood (BFGS) on a host apop_data, timed as one job"""
    from apophenia_b200 import abi, host
    ab = cx.ab
    h, lib = host.lib(), abi.load_library()
    cx.barrier()
    t0 = time.perf_counter()
    ab.pin(d)
    t_pin = time.perf_counter() - t0
    start = ab.probit_start(d, k)
    m = h.apop_model_copy(host.model_object(abi.PROBIT))
    abi.check(lib.apop_b200_register(m, abi.PROBIT), "apop_b200_register")
    m.contents.parameters = h.apop_data_alloc(k, 1, 0)
    m.contents.data = d.ptr
    keep = []
    host.mle_settings(m, method="BFGS cg", tolerance=1e-9, relative_tolerance=True, max_iterations=200,
                      starting_pt=start, keep=keep)
    h.apop_maximum_likelihood(d.ptr, m)
    cx.barrier()
    dt = time.perf_counter() - t0
    rep = h.apop_mle_last_report()
    ab.unpin(d)
    dt, t_pin = cx.max_over_ranks(dt, t_pin)
    return {"seconds_total": dt, "seconds_pin": t_pin, "status": rep.status, "passes": rep.f_evals,
            "iterations": rep.iterations}


# ------------------------------------------------------------------------------------------------
#  fits on resident data
# ------------------------------------------------------------------------------------------------
def fit_resident(cx, family, data_ptr, shape, start, method, tol=1e-9, max_iter=300):
    """apop_maximum_likelihood (host mirror, csrc/mle.c) over the device entries installed by apop_b200_register,
    stop rule |gradient| < tol * |LL|.  shape = (vector, rows, cols) of the parameter set."""
    from apophenia_b200 import abi, host
    ab = cx.ab
    h, lib = host.lib(), abi.load_library()
    m = h.apop_model_copy(host.model_object(family))
    abi.check(lib.apop_b200_register(m, family), "apop_b200_register")
    m.contents.parameters = h.apop_data_alloc(*shape)
    m.contents.data = data_ptr
    keep = []
    host.mle_settings(m, method=method, tolerance=tol, relative_tolerance=True, max_iterations=max_iter,
                      starting_pt=start, keep=keep)
    l0 = ab.launch_count()
    cx.barrier()
    t0 = time.perf_counter()
    h.apop_maximum_likelihood(data_ptr, m)
    cx.barrier()
    dt, = cx.max_over_ranks(time.perf_counter() - t0)
    rep = h.apop_mle_last_report()
    est = host.parameters_of(m)
    return m, est, {"seconds": dt, "status": rep.status, "iterations": rep.iterations, "passes": rep.f_evals,
                    "kernel_launches": int(ab.launch_count() - l0), "ll": rep.ll, "gradient_norm": rep.gradient_norm}


def run_time_to_fit(cx, data, k, beta_true):
    """probit MLE on the resident (weak) data, from the probit_prep start point exp(mean ln|x|)"""
    from apophenia_b200 import abi
    ab = cx.ab
    out = {}
    start = ab.probit_start(data, k)   # probit_prep's rule, reduced on the device (one pass over X)
    for method in ("BFGS cg", "Newton"):
        # one untimed pass per kernel the method uses (module load, occupancy queries, buffers)
        ab.ll_score(abi.PROBIT, data, start * 1.0000001)
        if method == "Newton":
            ab.information(abi.PROBIT, data, start * 1.0000001)
        m, est, rep = fit_resident(cx, abi.PROBIT, data.ptr, (k, 1, 0), start, method)
        if rep["status"] != 0 and method == "Newton":
            continue
        rep["max_abs_err_vs_beta_true"] = float(np.max(np.abs(est - beta_true)))
        out[method] = rep
    return out


# ------------------------------------------------------------------------------------------------
#  strong scaling: the FIXED 100M x 32 problem over the N ranks
# ------------------------------------------------------------------------------------------------
def run_strong(cx, weak_data, beta_true):
    from apophenia_b200 import abi
    ab, args, world, rank = cx.ab, cx.args, cx.world, cx.rank
    k = args.k
    total = FIXED_ROWS
    from apophenia_b200.shard import shard_range
    lo, hi = shard_range(total, world, rank)
    rows = hi - lo
    shard = weak_data if world == 1 and args.rows == total else ab.synth(abi.PROBIT, rows, k, beta_true, seed=8, row_offset=lo)
    steps = max(args.steps, 50)
    dt, kms = cx.time_passes(lambda i: ab.ll_score(abi.PROBIT, shard, beta_true * (1.0 + 1e-6 * (i + 7))), steps, warmup=5)
    out = {"problem": f"probit {total} x {k}, fixed, rows split contiguously over {world} rank(s) ({rows} rows on rank {rank})",
           "n_gpus": world, "steps": steps, "ms_per_pass": 1e3 * dt, "kernel_ms": kms, "rows_per_s": total / dt,
           "GBps_per_gpu_kernel": rows * (k + 1) * 8 / (kms * 1e-3) / 1e9 if kms > 0 else None,
           "per_pass_overhead_us": 1e6 * dt - 1e3 * kms,
           "exchange": "in-kernel NVLink peer-store exchange" if ab.abi.load_library().apop_b200_comm_p2p_active() else ("ncclAllReduce" if world > 1 else "none")}
    # the single-GPU time of the same fixed problem, measured in this very run: rank 0's weak shard IS rows
    # [0, 100M) of the global data set when --rows is 100M, evaluated without the exchange
    if args.rows == total:
        ab.set_local_only(True)
        t1 = None
        if rank == 0:
            for i in range(3):
                ab.ll_score(abi.PROBIT, weak_data, beta_true * (1.0 + 1e-6 * (i + 90)))
            t0 = time.perf_counter()
            for i in range(10):
                ab.ll_score(abi.PROBIT, weak_data, beta_true * (1.0 + 1e-6 * (i + 100)))
            t1 = (time.perf_counter() - t0) / 10
        ab.set_local_only(False)
        t1 = cx.max_over_ranks(t1 if t1 is not None else 0.0)[0]
        out["single_gpu_ms_per_pass_same_run"] = 1e3 * t1
        out["speedup_vs_single_gpu_same_run"] = t1 / dt
        out["efficiency_vs_single_gpu_same_run"] = t1 / dt / world
    # time to fit on the fixed problem
    fits = {}
    start = ab.probit_start(shard, k)
    for method in ("BFGS cg", "Newton"):
        ab.ll_score(abi.PROBIT, shard, start * 1.0000001)
        if method == "Newton":
            ab.information(abi.PROBIT, shard, start * 1.0000001)
        m, est, rep = fit_resident(cx, abi.PROBIT, shard.ptr, (k, 1, 0), start, method)
        rep["max_abs_err_vs_beta_true"] = float(np.max(np.abs(est - beta_true)))
        fits[method] = rep
    out["time_to_fit"] = fits
    if shard is not weak_data:
        shard.free()
    return out


# ------------------------------------------------------------------------------------------------
#  CPU baselines (restated reference CPU path, oracle/), rank 0 only
# ------------------------------------------------------------------------------------------------
def run_cpu_baseline(ab, args, fam, beta_true):
    from oracle import oracle as o
    n, k = min(args.cpu_rows, args.rows), args.k
    src = ab.synth(fam, n, k, beta_true, seed=8, row_offset=0)
    X, y = src.download()
    src.free()
    threads = host_threads(o)
    o.time_probit_ll_score(X[: n // 10], y[: n // 10], beta_true, threads)  # warm-up
    best = None
    t_total = 0.0
    reps = 0
    while reps < 3 and t_total < 30.0:
        r = o.time_probit_ll_score(X, y, beta_true, threads)
        t_total += r["seconds"]
        reps += 1
        if best is None or r["seconds"] < best["seconds"]:
            best = r
    return {"value": n / best["seconds"], "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} rows x {k} cols of the same synthetic recipe, best of {reps}; restated reference "
                      f"CPU path (materialised dgemm, OpenMP row map, serial long double sum, serial score)",
            "sec_ll": best["sec_ll"], "sec_score": best["sec_score"]}


# ------------------------------------------------------------------------------------------------
#  the other BASELINE.json configurations at this N
# ------------------------------------------------------------------------------------------------
def run_configs(cx, names):
    out = {}
    table = {"cfg2": cfg2_logit8, "cfg3": cfg3_poisson_shard, "cfg4": cfg4_bootstrap, "cfg5": cfg5_chains}
    for nm in names:
        if nm not in table:
            continue
        try:
            out[nm] = table[nm](cx)
        except Exception as e:  # noqa: BLE001  (a failing side configuration must not lose the headline line)
            out[nm] = {"error": f"{type(e).__name__}: {e}"}
            cx.ab.set_local_only(False)
            cx.ab.set_timing(False)
    return out


def split_rows(total, cx):
    from apophenia_b200.shard import shard_range
    lo, hi = shard_range(total, cx.world, cx.rank)
    return lo, hi - lo


def cfg2_logit8(cx, total=20_000_000, k=32, C=8):
    """cfg2: apop_logit multinomial (K=8) MLE on 20M x 32: fused LL+score pass, BFGS fit from B = 0, the P x P
    observed information (one pass per 4 category pairs) and a Newton fit on it.  At N > 1 the rows are split."""
    from apophenia_b200 import abi
    ab = cx.ab
    P = k * (C - 1)
    rng = np.random.default_rng(8)
    B = rng.uniform(-1, 1, P)
    lo, rows = split_rows(total, cx)
    s = ab.synth(abi.LOGIT, rows, k, B, ncat=C, seed=8, row_offset=lo)
    dt, kms = cx.time_passes(lambda i: ab.ll_score(abi.LOGIT, s, B * (1 + 1e-6 * (i + 1))), 20)
    by = rows * (k + 1) * 8
    out = {"config": f"apop_logit multinomial (K={C}) on {total} x {k}, {cx.world} GPU(s), {rows} rows on this rank",
           "ms_per_pass": 1e3 * dt, "rows_per_s": total / dt,
           "roofline": roofline_of(by, kms, "logit_stream_kernel", extra={
               "note": "both contractions as DMMA from a swizzled resident layout, one TMA bulk copy per tile; 4 k (C-1) = 896 flop/row on the FP64 tensor pipe"})}
    measured_traffic(out["roofline"], "logit_stream_kernel", rows, k)
    m, est, rep = fit_resident(cx, abi.LOGIT, s.ptr, (0, k, C - 1), np.zeros(P), "BFGS cg", max_iter=2000)
    rep["max_abs_err_vs_B_true"] = float(np.max(np.abs(est - B)))
    out["fit_bfgs"] = rep
    # observed information at the estimate: P x P from ceil(28 / 4) = 7 launches on the FP64 tensor path
    ab.information(abi.LOGIT, s, est)
    ab.pass_timer_reset()
    cx.barrier()
    t0 = time.perf_counter()
    H = ab.information(abi.LOGIT, s, est * (1 + 1e-9))
    cx.barrier()
    t_info, = cx.max_over_ranks(time.perf_counter() - t0)
    kms_info, n_info = ab.pass_timer()
    ab.set_timing(False)
    t0 = time.perf_counter()
    cov = np.linalg.inv(H)
    t_inv = time.perf_counter() - t0
    flops = 2.0 * rows * (C - 1) * C / 2 * (k * (k + 8) / 2)      # upper block triangles, 8-wide blocks
    out["information"] = {"seconds": t_info, "launches": n_info, "kernel_ms_total": kms_info, "P": P,
                          "dmma_TFLOPs": flops / (kms_info * 1e-3) / 1e12 if kms_info > 0 else None,
                          "covariance_inverse_seconds_host": t_inv, "min_eig_ok": bool(np.all(np.diag(cov) > 0)),
                          "replaces": "apop_model_hessian: >= 4 P = 896 score passes (the reference has no analytic score: >= 16 P^2 LL passes)"}
    m2, est2, rep2 = fit_resident(cx, abi.LOGIT, s.ptr, (0, k, C - 1), np.zeros(P), "Newton", max_iter=100)
    rep2["max_rel_diff_vs_bfgs"] = float(np.max(np.abs(est2 - est)) / np.max(np.abs(est)))
    out["fit_newton"] = rep2
    if cx.rank == 0 and not cx.args.no_cpu:
        from oracle import oracle as o
        nc = 1_000_000
        w = ab.synth(abi.LOGIT, nc, k, B, ncat=C, seed=8, row_offset=0)
        X, y = w.download()
        w.free()
        thr = host_threads(o)
        o.time_logit_ll(X[:100_000], y[:100_000], B.reshape(k, C - 1), thr)
        r = min((o.time_logit_ll(X, y, B.reshape(k, C - 1), thr) for _ in range(2)), key=lambda q: q["seconds"])
        out["cpu"] = {"kind": "port", "cores": thr, "sample": f"{nc} rows x {k} cols, restated multilogit_log_likelihood "
                      "(materialised single-thread dgemm N x 7, OpenMP row map, serial long double sum), best of 2",
                      "ll_rows_per_s": nc / r["seconds"],
                      "numeric_gradient_ll_passes": 4 * P,
                      "gradient_rows_per_s": nc / r["seconds"] / (4 * P),
                      "note": "the reference has no analytic logit score (model/apop_probit.c:244-289 is commented out): every "
                              "gradient is 4 P = 896 LL passes (apop_mle.m4.c:92-108); extrapolated from the measured LL pass"}
        out["speedup_ll_plus_gradient_vs_cpu"] = (total / dt) / out["cpu"]["gradient_rows_per_s"]
    s.free()
    return out


def cfg3_poisson_shard(cx, rows=50_000_000, k=16):
    """cfg3: apop_poisson regression MLE on 400M x 16 over 8 GPUs = 50M x 16 per GPU (weak: every rank holds
    its shard; at N = 8 this is the configuration as stated), score summed over ranks once per pass."""
    from apophenia_b200 import abi
    ab = cx.ab
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k) / np.sqrt(k)
    s = ab.synth(abi.POISSON_REG, rows, k, beta, seed=8, row_offset=cx.rank * rows)
    dt, kms = cx.time_passes(lambda i: ab.ll_score(abi.POISSON_REG, s, beta * (1 + 1e-6 * (i + 1))), 20)
    by = rows * (k + 1) * 8
    out = {"config": f"apop_poisson regression on {rows * cx.world} x {k}: {rows} rows per GPU, {cx.world} GPU(s)",
           "ms_per_pass": 1e3 * dt, "rows_per_s": rows * cx.world / dt, "roofline": roofline_of(by, kms, "glm_fused_kernel<16, FamPoissonReg>")}
    fits = {}
    for method in ("BFGS cg", "Newton"):
        ab.ll_score(abi.POISSON_REG, s, np.zeros(k) + 1e-7)
        if method == "Newton":
            ab.information(abi.POISSON_REG, s, np.zeros(k) + 1e-7)
        m, est, rep = fit_resident(cx, abi.POISSON_REG, s.ptr, (k, 0, 0), np.zeros(k), method)
        rep["max_abs_err_vs_beta_true"] = float(np.max(np.abs(est - beta)))
        fits[method] = rep
    out["time_to_fit"] = fits
    if cx.rank == 0 and not cx.args.no_cpu:
        from oracle import oracle as o
        nc = 1_000_000
        w = ab.synth(abi.POISSON_REG, nc, k, beta, seed=8, row_offset=0)
        X, y = w.download()
        w.free()
        t0 = time.perf_counter()
        o.poisson_reg_ll(X, y, beta, o.FAITHFUL); o.poisson_reg_score(X, y, beta, o.FAITHFUL)
        tc = time.perf_counter() - t0
        out["cpu"] = {"kind": "port", "cores": 1, "sample": f"{nc} rows x {k} cols, restated serial LL + score pass (the regression "
                      "form has no reference counterpart, SURVEY a16: written in the reference's style, one thread)",
                      "rows_per_s": nc / tc}
        out["speedup_vs_cpu"] = (rows * cx.world / dt) / out["cpu"]["rows_per_s"]
    s.free()
    return out


def cfg4_bootstrap(cx, total=5_000_000, k=20, m=1000):
    """cfg4: apop_bootstrap_cov of probit, 1000 replicates on 5M x 20: the batched-beta DMMA pass with resident
    resample counts, and the driver end to end (counts + lock-step replicate fits)."""
    from apophenia_b200 import abi, host
    ab = cx.ab
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    lo, rows = split_rows(total, cx)
    s = ab.synth(abi.PROBIT, rows, k, beta, seed=8, row_offset=lo)
    t0 = time.perf_counter()
    ab.bootstrap_counts(s, m, seed=8)
    t_counts, = cx.max_over_ranks(time.perf_counter() - t0)
    # the replicate vectors are scattered around beta with the sampling distribution's own spread (one standard error
    # per coordinate, from the observed information), as the replicate estimates of a bootstrap are
    se = np.sqrt(np.diag(np.linalg.inv(ab.information(abi.PROBIT, s, beta))))
    Bm = beta[None, :] + se[None, :] * rng.standard_normal((m, k))
    dt, kms = cx.time_passes(lambda i: ab.ll_score_batched(abi.PROBIT, s, Bm * (1 + 1e-9 * (i + 1)), use_counts=True), 3, warmup=1)
    exact_share = ab.batched_exact_fraction(s)
    # the driver end to end, BEFORE the far-apart experiment below (which sends the next 32 batched passes over this data
    # set to the exact kernel: the library's adaptive choice, not something a bootstrap would meet)
    start = ab.probit_start(s, k)
    mfit, est, rep = fit_resident(cx, abi.PROBIT, s.ptr, (k, 1, 0), start, "Newton")
    l0 = ab.launch_count()
    ab.pass_timer_reset()
    cx.barrier()
    t0 = time.perf_counter()
    cov, boots = host.bootstrap_cov(s.ptr, mfit, iterations=m, seed=8)
    cx.barrier()
    t_boot, = cx.max_over_ranks(time.perf_counter() - t0)
    boot_kms, boot_n = ab.pass_timer()
    ab.set_timing(False)
    boot_launches = int(ab.launch_count() - l0)
    boot_exact_share = ab.batched_exact_fraction(s)
    Bw = beta[None, :] + 0.05 * rng.standard_normal((m, k))       # far apart: the kernel falls back to one link evaluation per pair
    ab.ll_score_batched(abi.PROBIT, s, Bw, use_counts=True)
    dtw, kmsw = cx.time_passes(lambda i: ab.ll_score_batched(abi.PROBIT, s, Bw * (1 + 1e-9 * (i + 1)), use_counts=True), 2, warmup=1)
    flops = 4.0 * rows * k * m
    out = {"config": f"apop_bootstrap_cov of probit, {m} replicates on {total} x {k}, {cx.world} GPU(s)",
           "counts_seconds": t_counts, "batched_pass_ms": 1e3 * dt, "batched_kernel_ms": kms,
           "row_replicates_per_s": total * m / dt, "replicate_spread": "1 standard error per coordinate (mean %.1e)" % float(se.mean()),
           "exact_evaluation_share": exact_share,
           "far_apart_vectors": {"spread": 0.05, "batched_pass_ms": 1e3 * dtw, "kernel_ms": kmsw,
                                 "note": "parameter vectors far from each other: every (row, replicate) pair takes the exact link evaluation"},
           "roofline": {"bound": "fp64 datapath (DMMA + link evaluations share it)", "dmma_TFLOPs": flops / (kms * 1e-3) / 1e12 if kms > 0 else None,
                        "dmma_peak_TFLOPs_measured": 37.0, "link_evaluations_per_s": rows * m / (kms * 1e-3) if kms > 0 else None,
                        "kernel": "batched_taylor_kernel<24, FamProbit, score, counts>", "kernel_ms": kms,
                        "bytes_per_launch": rows * (k + 1) * 8 + rows * m}}
    H = ab.information(abi.PROBIT, s, est)
    want = np.linalg.inv(H)
    out["bootstrap_cov"] = {"seconds": t_boot, "kernel_launches": boot_launches, "timed_passes": boot_n, "kernel_ms_total": boot_kms,
                            "exact_evaluation_share_last_pass": boot_exact_share,
                            "max_rel_dev_of_var_from_inverse_information": float(np.max(np.abs(np.diag(cov) / np.diag(want) - 1))),
                            "full_sample_fit": rep}
    if cx.rank == 0 and not cx.args.no_cpu:
        from oracle import oracle as o
        w = ab.synth(abi.PROBIT, total, k, beta, seed=8, row_offset=0)
        X, y = w.download()
        w.free()
        thr = host_threads(o)
        t_res = o.time_bootstrap_resample(X, y, seed=8)
        r = o.time_probit_ll_score(X, y, beta, thr)
        evals = max(int(rep["passes"]), 1)
        per_rep = t_res + evals * r["seconds"]
        out["cpu"] = {"kind": "port", "cores": thr,
                      "sample": f"ONE replicate's resample of all {total} rows (serial row copies, apop_bootstrap.m4.c:143-149) and ONE "
                                f"LL+score evaluation over {total} x {k}, restated reference CPU path; a replicate's fit is taken to need the "
                                f"{evals} evaluations the full-sample fit needed; x {m} replicates is an EXTRAPOLATION",
                      "resample_seconds": t_res, "ll_score_seconds": r["seconds"], "evaluations_per_fit_assumed": evals,
                      "seconds_per_replicate": per_rep, "seconds_extrapolated_all_replicates": per_rep * m}
        out["speedup_bootstrap_cov_vs_cpu_extrapolated"] = per_rep * m / t_boot
    s.free()
    return out


def cfg5_chains(cx, total=50_000_000, k=24, m=256):
    """cfg5: apop_model_metropolis on (binary) logit, 256 parallel chains over 50M x 24: one lock-step step of all
    chains = one batched LL pass (no score) + one sum over ranks."""
    from apophenia_b200 import abi
    ab = cx.ab
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    lo, rows = split_rows(total, cx)
    s = ab.synth(abi.LOGIT, rows, k, beta, seed=8, row_offset=lo)
    # the chains sit in the posterior, i.e. within about one standard error of each other per coordinate
    se = np.sqrt(np.diag(np.linalg.inv(ab.information(abi.LOGIT, s, beta))))
    Bm = beta[None, :] + se[None, :] * rng.standard_normal((m, k))
    dt, kms = cx.time_passes(lambda i: ab.ll_score_batched(abi.LOGIT, s, Bm * (1 + 1e-9 * (i + 1)), want_score=False), 3, warmup=1)
    exact_share = ab.batched_exact_fraction(s)
    dt1, kms1 = cx.time_passes(lambda i: ab.ll_score(abi.LOGIT, s, Bm[0] * (1 + 1e-6 * (i + 1)), want_score=False), 10)
    out = {"config": f"apop_model_metropolis on logit, {m} parallel chains over {total} x {k}, {cx.world} GPU(s)",
           "ms_per_step_all_chains": 1e3 * dt, "kernel_ms": kms, "chain_steps_per_s": m / dt, "row_chains_per_s": total * m / dt,
           "single_chain_step_ms": 1e3 * dt1, "speedup_vs_m_single_chain_steps": dt1 * m / dt,
           "chain_spread": "1 standard error per coordinate (mean %.1e)" % float(se.mean()), "exact_evaluation_share": exact_share,
           "roofline": {"bound": "fp64 datapath (DMMA + link evaluations share it)", "dmma_TFLOPs": 2.0 * rows * k * m / (kms * 1e-3) / 1e12 if kms > 0 else None,
                        "dmma_peak_TFLOPs_measured": 37.0, "link_evaluations_per_s": rows * m / (kms * 1e-3) if kms > 0 else None,
                        "kernel": "batched_taylor_kernel<24, FamLogit2, LL only>", "kernel_ms": kms, "bytes_per_launch": rows * (k + 1) * 8}}
    if cx.rank == 0 and not cx.args.no_cpu:
        from oracle import oracle as o
        nc = 2_000_000
        w = ab.synth(abi.LOGIT, nc, k, beta, seed=8, row_offset=0)
        X, y = w.download()
        w.free()
        thr = host_threads(o)
        o.time_logit_ll(X[:100_000], y[:100_000], beta.reshape(k, 1), thr)
        r = min((o.time_logit_ll(X, y, beta.reshape(k, 1), thr) for _ in range(3)), key=lambda q: q["seconds"])
        out["cpu"] = {"kind": "port", "cores": thr,
                      "sample": f"{nc} rows x {k} cols: one step of the reference's single chain = one multilogit_log_likelihood pass "
                                "(apop_mcmc.m4.c:104-138), restated CPU path, best of 3",
                      "ll_rows_per_s": nc / r["seconds"], "chain_steps_per_s_at_full_size": nc / r["seconds"] / total}
        out["speedup_chain_steps_vs_cpu"] = (m / dt) / out["cpu"]["chain_steps_per_s_at_full_size"]
    s.free()
    return out


if __name__ == "__main__":
    main()
