#!/usr/bin/env python
"""One-line digests of the ncu raw CSV exports under profiles/ (or gpurun_out/): duration, DRAM bytes, pipes, stalls."""
import csv, glob, sys
pat = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/r02_ncu_full_*_raw.csv"
for f in sorted(glob.glob(pat)):
    rows = list(csv.reader(open(f)))
    d = dict(zip(rows[0], rows[-1]))
    units = dict(zip(rows[0], rows[1]))
    def gu(k):   # value with its exported unit
        return f"{g(k):.3f} {units.get(k, '')}"
    def g(k):
        try: return float(d[k])
        except Exception: return float("nan")
    st = []
    for h, v in d.items():
        if "pcsamp_warps_issue_stalled" in h and not h.endswith("not_issued"):
            try: st.append((float(v), h.split("stalled_")[1]))
            except Exception: pass
    tot = sum(v for v, _ in st) or 1
    print(f.split("/")[-1])
    print(f"  kernel {d.get('Kernel Name', '')[:70]}")
    print(f"  duration {gu('gpu__time_duration.sum')}  dram read {gu('dram__bytes_read.sum')} write {gu('dram__bytes_write.sum')}  dram % of peak {g('dram__throughput.avg.pct_of_peak_sustained_elapsed'):.1f}")
    print(f"  fp64 pipe {g('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} %  dmma {g('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} %  lsu {g('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f} %  l1 data-pipe wavefronts {g('l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed'):.1f} %  issue active {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} %")
    print(f"  registers {d.get('launch__registers_per_thread')}  warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} % of 64  stalls: " + ", ".join(f"{n} {100 * v / tot:.0f}%" for v, n in sorted(st, reverse=True)[:6]))
