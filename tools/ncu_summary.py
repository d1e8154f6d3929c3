#!/usr/bin/env python
"""Prints the handful of ncu metrics the design discussion uses from a .ncu-rep (raw page)."""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = [r"^Kernel Name", r"^gpu__time_duration.sum$", r"^sm__cycles_elapsed.max$", r"^launch__registers_per_thread$", r"^launch__grid_size$",
        r"^sm__warps_active.avg.pct_of_peak_sustained_active$", r"^dram__bytes_read.sum$", r"^dram__bytes_write.sum$",
        r"^dram__throughput.avg.pct_of_peak_sustained_elapsed$", r"^sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active$",
        r"^sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active$", r"smsp__pipe_tensor_subpipe_dmma_cycles_active.avg$",
        r"^smsp__issue_active.avg.pct_of_peak_sustained_active$", r"^smsp__inst_executed.sum$",
        r"^l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$", r"^l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$", r"issue_stalled.*_per_issue_active.ratio$", r"^smsp__warps_eligible.avg.per_cycle_active$",
        r"^sm__inst_executed_pipe_alu.avg.pct", r"^sm__inst_executed_pipe_fma", r"^sm__inst_executed_pipe_lsu.avg.pct", r"^sm__inst_executed_pipe_xu.avg.pct"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if pat and not re.search(pat, d.get("Kernel Name", "")):
        continue
    for k in KEYS:
        for name in hdr:
            if re.search(k, name):
                v = d[name]
                if "issue_stalled" in name:
                    try:
                        if float(v) < 0.05: continue
                    except ValueError: pass
                print(f"{name:95s} {v}")
    print()
