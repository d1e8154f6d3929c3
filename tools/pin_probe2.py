#!/usr/bin/env python
"""What does zero-copy upload cost on this box?  Times, for a pageable numpy buffer of G GB:
cudaHostRegister (plain and read-only, with and without transparent huge pages behind the buffer),
the H2D copy straight out of the registered range, cudaHostUnregister -- next to the library's staged
upload (gather into pinned staging, APOP_B200_PIN_TRACE).  Run once per pool change; the numbers decide
the default of APOP_B200_PIN_MODE."""
import ctypes as C
import json
import mmap
import os
import sys
import time

import numpy as np
import torch

gb = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
n = int(gb * 1e9) // 8
rt = torch.cuda.cudart()
torch.cuda.init()
dev = torch.empty(n, dtype=torch.float64, device="cuda")
libc = C.CDLL(None)
out = {"GB": n * 8e-9}


def alloc(huge):
    buf = mmap.mmap(-1, n * 8, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    a = np.frombuffer(buf, dtype=np.float64)
    if huge:
        MADV_HUGEPAGE = 14
        addr = a.ctypes.data
        libc.madvise(C.c_void_p(addr), C.c_size_t(n * 8), MADV_HUGEPAGE)
    t0 = time.perf_counter()
    a[:] = 1.0          # first touch
    return buf, a, time.perf_counter() - t0


for huge in (False, True):
    for flags, name in ((0, "default"), (8, "readonly")):
        buf, a, t_touch = alloc(huge)
        ptr = a.ctypes.data
        t0 = time.perf_counter()
        rc = rt.cudaHostRegister(ptr, n * 8, flags)
        t_reg = time.perf_counter() - t0
        key = f"{'thp' if huge else '4k'}_{name}"
        if int(rc) != 0:
            out[key] = {"error": int(rc)}
            continue
        src = torch.from_numpy(a)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        dev.copy_(src, non_blocking=True); torch.cuda.synchronize()
        t_copy = time.perf_counter() - t0
        t0 = time.perf_counter(); rt.cudaHostUnregister(ptr); t_unreg = time.perf_counter() - t0
        out[key] = {"touch_GBps": n * 8e-9 / t_touch, "register_GBps": n * 8e-9 / t_reg, "copy_GBps": n * 8e-9 / t_copy,
                    "unregister_GBps": n * 8e-9 / t_unreg, "register_s": t_reg, "unregister_s": t_unreg}
        del src, a
        buf.close() if False else None
try:
    out["thp_enabled"] = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
except Exception as e:  # noqa: BLE001
    out["thp_enabled"] = str(e)
out["cpu_count"] = os.cpu_count()
try:
    out["numa_nodes"] = len([d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")])
except Exception:  # noqa: BLE001
    out["numa_nodes"] = None
print(json.dumps(out))
