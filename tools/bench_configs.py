#!/usr/bin/env python
"""Times the other BASELINE.json configurations on one GPU (bench.py carries the headline
one).  Prints one JSON object per configuration.  Usage: python tools/bench_configs.py [names]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import apophenia_b200 as ab  # noqa: E402


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    ab.pass_timer_reset()
    t0 = time.perf_counter()
    for i in range(reps):
        fn(i)
    dt = (time.perf_counter() - t0) / reps
    ms, n = ab.pass_timer()
    return dt, ms / max(n, 1)


def cfg_logit8(rows=20_000_000, k=32, C=8):
    rng = np.random.default_rng(8)
    B = rng.uniform(-1, 1, k * (C - 1))
    s = ab.synth(ab.LOGIT, rows, k, B, ncat=C)
    dt, kms = timed(lambda i=0: ab.ll_score(ab.LOGIT, s, B * (1 + 1e-6 * (i + 1))), reps=20)
    s.free()
    by = rows * (k + 1) * 8
    return {"config": f"logit multinomial C={C}, {rows}x{k}, fused LL+score", "ms_per_pass": 1e3 * dt, "kernel_ms": kms,
            "rows_per_s": rows / dt, "GBps": by / kms / 1e6, "flops_per_row": 4 * k * (C - 1)}


def cfg_bootstrap(rows=5_000_000, k=20, m=1000, score=True):
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    s = ab.synth(ab.PROBIT, rows, k, beta)
    t0 = time.perf_counter()
    ab.bootstrap_counts(s, m, seed=8)
    t_counts = time.perf_counter() - t0
    B = beta[None, :] + 0.01 * rng.standard_normal((m, k))
    dt, kms = timed(lambda i=0: ab.ll_score_batched(ab.PROBIT, s, B * (1 + 1e-6 * (i + 1)), use_counts=True, want_score=score), reps=3, warm=1)
    # the unbatched alternative: m single-beta passes
    dt1, kms1 = timed(lambda i=0: ab.ll_score(ab.PROBIT, s, B[0] * (1 + 1e-6 * (i + 1))), reps=20)
    s.free()
    return {"config": f"probit bootstrap batched, {rows}x{k}, m={m} replicates with resample counts, score={score}",
            "ms_per_batched_pass": 1e3 * dt, "kernel_ms": kms, "row_replicates_per_s": rows * m / dt,
            "dmma_flops": 4.0 * rows * k * m, "dmma_TFLOPs": 4.0 * rows * k * m / (kms * 1e-3) / 1e12,
            "counts_seconds": t_counts, "single_beta_pass_ms": 1e3 * dt1, "m_single_passes_ms": 1e3 * dt1 * m,
            "speedup_vs_m_single_passes": dt1 * m / dt}


def cfg_chains(rows=50_000_000, k=24, m=256):
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    s = ab.synth(ab.LOGIT, rows, k, beta)
    B = beta[None, :] + 0.01 * rng.standard_normal((m, k))
    dt, kms = timed(lambda i=0: ab.ll_score_batched(ab.LOGIT, s, B * (1 + 1e-6 * (i + 1)), want_score=False), reps=3, warm=1)
    dt1, kms1 = timed(lambda i=0: ab.ll_score(ab.LOGIT, s, B[0] * (1 + 1e-6 * (i + 1)), want_score=False), reps=20)
    s.free()
    return {"config": f"logit metropolis chains batched LL, {rows}x{k}, m={m} chains",
            "ms_per_batched_step": 1e3 * dt, "kernel_ms": kms, "row_chains_per_s": rows * m / dt,
            "single_chain_step_ms": 1e3 * dt1, "speedup_vs_m_single_steps": dt1 * m / dt}


def cfg_poisson(rows=50_000_000, k=16):
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k) / np.sqrt(k)
    s = ab.synth(ab.POISSON_REG, rows, k, beta)
    dt, kms = timed(lambda i=0: ab.ll_score(ab.POISSON_REG, s, beta * (1 + 1e-6 * (i + 1))), reps=20)
    s.free()
    by = rows * (k + 1) * 8
    out = {"config": f"poisson regression, {rows}x{k} (one GPU's shard of 400M x 16), fused LL+score",
           "ms_per_pass": 1e3 * dt, "kernel_ms": kms, "rows_per_s": rows / dt, "GBps": by / kms / 1e6}
    lam = np.array([3.1])
    s = ab.synth(ab.POISSON, rows, k, lam)
    m = ab.vector_model([3.0], None)
    t0 = time.perf_counter(); ll0 = ab.log_likelihood(ab.POISSON, s, m); first = time.perf_counter() - t0
    t0 = time.perf_counter(); ll1 = ab.log_likelihood(ab.POISSON, s, ab.vector_model([3.2], None)); later = time.perf_counter() - t0
    s.free()
    out["iid_poisson_reading"] = {"cells": rows * k, "first_pass_s": first, "later_call_s": later}
    return out


def cfg_newton(rows=100_000_000, k=32):
    """information pass timing at the headline shape"""
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    s = ab.synth(ab.PROBIT, rows, k, beta)
    dt, kms = timed(lambda i=0: ab.information(ab.PROBIT, s, beta * (1 + 1e-6 * (i + 1))), reps=5)
    s.free()
    return {"config": f"probit observed information (k x k), {rows}x{k}", "ms_per_pass": 1e3 * dt, "kernel_ms": kms,
            "GBps": rows * (k + 1) * 8 / kms / 1e6, "dfma_TFLOPs": 2.0 * rows * k * (k + 1) / (kms * 1e-3) / 1e12}


def _fit_on_synth(fam, s, k, start, method="Newton"):
    from apophenia_b200 import abi, host
    h, lib = host.lib(), abi.load_library()
    m = h.apop_model_copy(host.model_object(fam))
    abi.check(lib.apop_b200_register(m, fam), "register")
    m.contents.parameters = h.apop_data_alloc(k, 1, 0)
    m.contents.data = s.ptr
    keep = []
    host.mle_settings(m, method=method, tolerance=1e-10, relative_tolerance=True, max_iterations=200, starting_pt=start, keep=keep)
    h.apop_maximum_likelihood(s.ptr, m)
    return m, host.parameters_of(m), keep


def cfg_logit8_fit(rows=20_000_000, k=32, C=8):
    """BASELINE config 2 as a fit: multinomial logit MLE (P = k (C-1) = 224 parameters) on resident data,
    apop_maximum_likelihood of the host mirror over the device LL and analytic score"""
    from apophenia_b200 import abi, host
    h, lib = host.lib(), abi.load_library()
    rng = np.random.default_rng(8)
    B = rng.uniform(-1, 1, k * (C - 1))
    s = ab.synth(ab.LOGIT, rows, k, B, ncat=C)
    out = {"config": f"multinomial logit MLE, C={C}, {rows}x{k}, P={k * (C - 1)} parameters, start at 0"}
    ab.ll_score(ab.LOGIT, s, B)      # first-use initialisation outside the timings
    for method, tol in (("BFGS cg", 1e-9), ("FR cg", 1e-9)):
        m = h.apop_model_copy(host.model_object(abi.LOGIT))
        abi.check(lib.apop_b200_register(m, abi.LOGIT), "register")
        m.contents.parameters = h.apop_data_alloc(0, k, C - 1)
        m.contents.data = s.ptr
        keep = []
        host.mle_settings(m, method=method, tolerance=tol, relative_tolerance=True, max_iterations=2000, starting_pt=np.zeros(k * (C - 1)), keep=keep)
        l0 = ab.launch_count()
        t0 = time.perf_counter()
        h.apop_maximum_likelihood(s.ptr, m)
        dt = time.perf_counter() - t0
        rep = h.apop_mle_last_report()
        est = host.parameters_of(m)
        out[method] = {"seconds": dt, "status": rep.status, "iterations": rep.iterations, "passes": int(ab.launch_count() - l0),
                       "ll": rep.ll, "gradient_norm": rep.gradient_norm, "max_abs_err_vs_B_true": float(np.max(np.abs(est - B)))}
    s.free()
    return out


def cfg_logit8_info(rows=20_000_000, k=32, C=8):
    """the P x P observed information of the multinomial logit (BASELINE config 2), all launches of one evaluation"""
    rng = np.random.default_rng(8)
    B = rng.uniform(-1, 1, k * (C - 1))
    s = ab.synth(ab.LOGIT, rows, k, B, ncat=C)
    ab.information(ab.LOGIT, s, B)
    t0 = time.perf_counter()
    for i in range(3):
        ab.information(ab.LOGIT, s, B * (1 + 1e-6 * (i + 1)))
    dt = (time.perf_counter() - t0) / 3
    s.free()
    P = k * (C - 1)
    return {"config": f"multinomial logit information, C={C}, {rows}x{k}, P={P}", "ms_per_information": 1e3 * dt,
            "dmma_TFLOPs": 2.0 * rows * (C - 1) * C / 2 * (k * (k + 8) / 2) / dt / 1e12}


def cfg_bootstrap_cov(rows=5_000_000, k=20, m=1000):
    """BASELINE config 4 end to end: apop_bootstrap_cov of a probit fit (counts + lock-step replicate fits)"""
    from apophenia_b200 import host
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    s = ab.synth(ab.PROBIT, rows, k, beta)
    est, mle, keep = _fit_on_synth(ab.PROBIT, s, k, ab.probit_start(s, k))
    l0 = ab.launch_count()
    t0 = time.perf_counter()
    cov, boots = host.bootstrap_cov(s.ptr, est, iterations=m, seed=8)
    dt = time.perf_counter() - t0
    passes = (ab.launch_count() - l0 - 2) // 2
    H = ab.information(ab.PROBIT, s, mle)
    want = np.linalg.inv(H)
    s.free()
    return {"config": f"apop_bootstrap_cov of probit, {m} replicates on {rows}x{k}", "seconds": dt, "batched_passes": int(passes),
            "max_rel_dev_of_var_from_inverse_information": float(np.max(np.abs(np.diag(cov) / np.diag(want) - 1))),
            "mean_abs_bias_of_replicates": float(np.max(np.abs(boots.mean(0) - mle)))}


def cfg_metropolis(rows=50_000_000, k=24, m=256, periods=200):
    """BASELINE config 5: adaptive Metropolis on binary logit, m chains in lock-step"""
    from apophenia_b200 import host
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    s = ab.synth(ab.LOGIT, rows, k, beta)
    est, mle, keep = _fit_on_synth(ab.LOGIT, s, k, np.zeros(k))
    sd = np.sqrt(np.diag(np.linalg.inv(ab.information(ab.LOGIT, s, mle))))
    t0 = time.perf_counter()
    draws = host.metropolis(s.ptr, est, periods=periods, burnin=0.5, initial_scale=float(sd.mean()) / np.sqrt(k), seed=8, n_chains=m)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    one = host.metropolis(s.ptr, est, periods=periods, burnin=0.5, initial_scale=float(sd.mean()) / np.sqrt(k), seed=8, n_chains=1)
    dt1 = time.perf_counter() - t0
    s.free()
    acc = float(np.mean(np.any(np.diff(draws, axis=1) != 0, axis=2)))
    return {"config": f"metropolis on logit, {m} chains x {periods} periods over {rows}x{k}", "seconds": dt,
            "ms_per_step_all_chains": 1e3 * dt / periods, "chain_steps_per_s": m * periods / dt,
            "single_chain_seconds": dt1, "single_chain_steps_per_s": periods / dt1, "acceptance_rate": acc}


def cfg_config1(rows=1_000_000, k=10):
    """BASELINE config 1: apop_estimate(probit, BFGS) on 1M x 10 -- the SAME host optimizer (csrc/mle.c)
    over the restated reference CPU path (oracle/cpu_models.c, all host cores) and over the device entries"""
    import ctypes as C
    from apophenia_b200 import abi, host
    h, lib = host.lib(), abi.load_library()
    om = C.CDLL(os.path.join(ROOT, "oracle", "liboracle_models.so"))
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    X = rng.uniform(-1, 1, (rows, k)); X[:, 0] = 1
    y = (rng.standard_normal(rows) > -(X @ beta)).astype(float)
    raw = X.copy(); raw[:, 0] = y
    settings = dict(method="BFGS cg", tolerance=1e-9, relative_tolerance=True, max_iterations=500)
    # --- CPU arm
    d = host.data_from_numpy(raw)
    m = h.apop_model_copy(host.model_object(abi.PROBIT))
    ll_addr = C.cast(om.oracle_model_probit_ll, C.c_void_p).value
    m.contents.log_likelihood = ll_addr
    h.apop_vtable_add(b"apop_score", C.cast(om.oracle_model_probit_score, C.c_void_p), ll_addr)
    keep = []
    host.mle_settings(m, keep=keep, **settings)
    t0 = time.perf_counter()
    est_c = h.apop_estimate(d, m)
    t_cpu = time.perf_counter() - t0
    rep_c = h.apop_mle_last_report()
    p_cpu = host.parameters_of(est_c)
    cpu = {"seconds": t_cpu, "iterations": rep_c.iterations, "passes": rep_c.f_evals, "status": rep_c.status}
    h.apop_vtable_drop(b"apop_score", ll_addr)
    # --- device arm (includes prep + pin, i.e. the full apop_estimate as a user calls it)
    d2 = host.data_from_numpy(raw)
    t0 = time.perf_counter()
    est_g, p_gpu, rep_g = host.estimate(abi.PROBIT, d2, **settings)
    t_gpu = time.perf_counter() - t0
    return {"config": f"apop_estimate(probit, BFGS), {rows}x{k}: same optimizer, CPU restated path vs B200", "cpu": cpu,
            "cpu_cores": os.cpu_count(), "gpu": {"seconds_incl_prep_and_pin": t_gpu, "mle_seconds": rep_g["seconds"],
                                                 "iterations": rep_g["iterations"], "passes": rep_g["f_evals"], "status": rep_g["status"]},
            "max_rel_param_diff": float(np.max(np.abs(p_gpu - p_cpu) / np.abs(p_cpu))),
            "speedup_mle": (t_cpu) / rep_g["seconds"]}


def cfg_prep(rows=50_000_000, k=16):
    """apop_prep of a probit model (ols_shuffle, category table, recoding, start point)
    on the host (the mirror's restatement of the reference: qsort of the N outcomes) and on the device
    (apop_b200_prep_outcome, which also makes the data resident: the pin is inside its time)"""
    from apophenia_b200 import abi, host
    h, lib = host.lib(), abi.load_library()
    rng = np.random.default_rng(8)
    beta = rng.uniform(-1, 1, k)
    X = rng.uniform(-1, 1, (rows, k))
    X[:, 0] = (rng.standard_normal(rows) > -(X[:, 1:] @ beta[1:] + beta[0]))
    out = {"config": f"probit apop_prep, {rows}x{k} host data (outcome in column 0)"}
    for name, env in (("device", None), ("host", "1")):
        if env: os.environ["APOP_B200_HOST_PREP"] = env
        d = host.data_from_numpy(X)
        m = h.apop_model_copy(host.model_object(abi.PROBIT))
        abi.check(lib.apop_b200_register(m, abi.PROBIT), "register")
        t0 = time.perf_counter(); h.apop_prep(d, m); t_prep = time.perf_counter() - t0
        t0 = time.perf_counter(); abi.check(lib.apop_b200_pin(d), "pin"); t_pin = time.perf_counter() - t0
        out[name] = {"prep_seconds": t_prep, "pin_seconds_after_prep": t_pin, "start": host.parameters_of(m)[:3].tolist()}
        lib.apop_b200_unpin(d)
        os.environ.pop("APOP_B200_HOST_PREP", None)
    out["speedup_prep_plus_pin"] = (out["host"]["prep_seconds"] + out["host"]["pin_seconds_after_prep"]) / (out["device"]["prep_seconds"] + out["device"]["pin_seconds_after_prep"])
    return out


def cfg_ingest(rows=2_000_000, k=16):
    """a numeric CSV straight into the resident layout (apop_b200_text_to_resident)
    against reading it into host memory first (pandas' C parser standing in for apop_text_to_data, which
    parses one character at a time and reallocs once per row) and pinning the result"""
    import tempfile
    import pandas as pd
    rng = np.random.default_rng(8)
    X = np.round(rng.uniform(-1, 1, (rows, k)), 6)
    X[:, 0] = (rng.uniform(size=rows) > 0.5)
    path = os.path.join(tempfile.mkdtemp(), "ingest.csv")
    pd.DataFrame(X, columns=[f"c{j}" for j in range(k)]).to_csv(path, index=False, float_format="%.6f")
    size = os.path.getsize(path)
    with open(path, "rb") as f:            # warm the page cache for both arms
        while f.read(1 << 26): pass
    t0 = time.perf_counter(); d = ab.text_to_resident(path, outcome_col0=True); t_dev = time.perf_counter() - t0
    Xd, yd = d.download(); d.free()
    t0 = time.perf_counter(); H = ab.text_to_host(path); t_host_reader = time.perf_counter() - t0
    t0 = time.perf_counter(); df = pd.read_csv(path).to_numpy(); t_pandas = time.perf_counter() - t0
    dd = ab.Data(matrix=df)
    t0 = time.perf_counter(); ab.pin(dd); t_pin = time.perf_counter() - t0
    ab.unpin(dd)
    ok = bool(np.array_equal(H, X) and np.array_equal(Xd[:, 1:], X[:, 1:]) and np.array_equal(yd, X[:, 0]) and np.array_equal(df, X))
    os.remove(path)
    return {"config": f"text ingest, {rows}x{k} CSV ({size / 1e6:.0f} MB)", "file_to_resident_seconds": t_dev, "text_MBps": size / 1e6 / t_dev,
            "library_reader_to_host_seconds": t_host_reader, "pandas_read_csv_seconds": t_pandas, "pin_seconds": t_pin,
            "speedup_vs_pandas_plus_pin": (t_pandas + t_pin) / t_dev, "identical": ok}


ALL = {"logit8_fit": cfg_logit8_fit, "logit8_info": cfg_logit8_info, "ingest": cfg_ingest, "prep": cfg_prep, "config1": cfg_config1, "bootstrap_cov": cfg_bootstrap_cov, "metropolis": cfg_metropolis, "logit8": cfg_logit8, "bootstrap": cfg_bootstrap, "chains": cfg_chains, "poisson": cfg_poisson, "newton": cfg_newton}

if __name__ == "__main__":
    ab.init(0)
    names = sys.argv[1:] or list(ALL)
    for nm in names:
        kw = {}
        if ":" in nm:  # name:key=value,key=value
            nm, arg = nm.split(":", 1)
            kw = {a.split("=")[0]: int(a.split("=")[1]) for a in arg.split(",")}
        print(json.dumps({nm: ALL[nm](**kw)}), flush=True)
    ab.finalize()
