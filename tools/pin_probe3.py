#!/usr/bin/env python
"""Why does a THP-advised buffer filled by a D2H copy page-lock slowly?  For each way of first-touching a
madvised anonymous mapping (CPU stores, pageable cudaMemcpy D2H) report AnonHugePages and the cudaHostRegister rate."""
import ctypes as C, json, mmap, sys, time
import numpy as np, torch
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gb * 1e9) // 8
rt = torch.cuda.cudart(); torch.cuda.init()
dev = torch.ones(n, dtype=torch.float64, device="cuda")
libc = C.CDLL(None, use_errno=True)
def huge_kb():
    for line in open("/proc/self/smaps_rollup"):
        if line.startswith("AnonHugePages"): return int(line.split()[1])
    return -1
out = {}
for how in ("cpu_touch", "d2h_pageable", "cpu_touch_threads", "d2h_then_nothing_4k"):
    buf = mmap.mmap(-1, n * 8, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    a = np.frombuffer(buf, dtype=np.float64)
    rc = 0
    if how != "d2h_then_nothing_4k":
        rc = libc.madvise(C.c_void_p(a.ctypes.data), C.c_size_t(n * 8), 14)
    h0 = huge_kb(); t0 = time.perf_counter()
    if how == "cpu_touch": a[:] = 1.0
    elif how == "cpu_touch_threads":
        from concurrent.futures import ThreadPoolExecutor
        parts = np.array_split(np.arange(0, n, dtype=np.int64)[:1], 1)
        def fill(i): a[i * (n // 8):(i + 1) * (n // 8) if i < 7 else n] = 1.0
        with ThreadPoolExecutor(8) as ex: list(ex.map(fill, range(8)))
    else:
        torch.from_numpy(a).copy_(dev); torch.cuda.synchronize()
    t_touch = time.perf_counter() - t0
    h1 = huge_kb()
    t0 = time.perf_counter(); e = rt.cudaHostRegister(a.ctypes.data, n * 8, 0); t_reg = time.perf_counter() - t0
    rt.cudaHostUnregister(a.ctypes.data)
    out[how] = {"madvise_rc": rc, "touch_GBps": n * 8e-9 / t_touch, "AnonHugePages_MB": (h1 - h0) / 1024, "register_GBps": n * 8e-9 / t_reg, "register_rc": int(e)}
    del a
for f in ("enabled", "defrag", "khugepaged/defrag"):
    try: out["thp_" + f] = open("/sys/kernel/mm/transparent_hugepage/" + f).read().strip()
    except Exception as ex: out["thp_" + f] = str(ex)
out["meminfo"] = {l.split(":")[0]: l.split(":")[1].strip() for l in open("/proc/meminfo") if l.split(":")[0] in ("MemTotal", "MemFree", "MemAvailable", "AnonHugePages", "HugePages_Total")}
print(json.dumps(out))
