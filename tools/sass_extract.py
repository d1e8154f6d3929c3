#!/usr/bin/env python
"""SASS extracts of the three hot kernels for profiles/sass/: the full listing is tens of thousands of lines, so what is
committed per kernel is (a) the opcode histogram of the whole function and of its main loop (the body of the outermost
backward branch that contains the FP64 tensor instructions, or the DFMAs), (b) the issue order of the DMMA / DFMA
instructions inside that loop (which accumulator each one writes), and (c) the loop itself, one line per instruction.
usage: python tools/sass_extract.py   (needs the built objects under apophenia_b200/csrc/build)"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "apophenia_b200", "csrc", "build")
OUT = os.path.join(ROOT, "profiles", "sass")
KERNELS = [
    ("glm_probit.o", r"glm_fused_kernelILi32ENS_9FamProbitELb0E", "glm_fused_kernel_k32_probit"),
    ("logit_hybrid.o", r"logit_hybrid_kernelILi32ELi7ELi2ELi3E", "logit_hybrid_kernel_k32_c8"),
    ("batched_kernel.o", r"batched_taylor_kernelILi24ENS_9FamProbitELb1ELb1E", "batched_taylor_kernel_k24_probit_score_counts"),
    ("batched_kernel.o", r"batched_kernelILi24ENS_9FamProbitELb1ELb1ELi1E", "batched_kernel_k24_probit_score_counts_exact"),
    ("logit_info.o", r"logit_info_kernelILi32ELi7E", "logit_info_kernel_k32_c8"),
    ("logit_stream.o", r"logit_stream_kernelILi32ELi7ELb0ELb1E", "logit_stream_kernel_k32_c8_paired"),
    ("glm_wide.o", r"glm_cols_kernelILi32ELb0ENS_9FamProbitE", "glm_cols_kernel_rh32_probit"),
    ("glm_wide.o", r"glm_cols_kernelILi16ELb1ENS_9FamProbitE", "glm_cols_kernel_rh16_big_probit"),
]
os.makedirs(OUT, exist_ok=True)
for obj, pat, name in KERNELS:
    syms = subprocess.check_output(["cuobjdump", "-elf", os.path.join(BUILD, obj)], text=True)
    m = re.search(r"(_ZN3apb\w*" + pat + r"\w*)", syms)
    if not m:
        print("not found:", pat); continue
    sym = m.group(1)
    sass = subprocess.check_output(["cuobjdump", "-sass", "-fun", sym, os.path.join(BUILD, obj)], text=True)
    ins = []
    for l in sass.splitlines():
        mm = re.search(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if mm: ins.append((int(mm.group(1), 16), mm.group(2).strip()))
    def op(t): return re.sub(r"^@!?U?P\w+\s+", "", t).split()[0].split(".")[0]
    hist_all = collections.Counter(op(t) for _, t in ins)
    hot = [i for i, (_, t) in enumerate(ins) if "DMMA" in t] or [i for i, (_, t) in enumerate(ins) if t.startswith("DFMA")]
    lo_i, hi_i = hot[0], hot[-1]
    # the outermost backward branch whose body contains all hot instructions
    best = None
    for i, (a, t) in enumerate(ins):
        mm = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", t)
        if mm:
            tgt = int(mm.group(1), 16)
            if tgt < a and tgt <= ins[lo_i][0] and a >= ins[hi_i][0]:
                if best is None or (a - tgt) < (best[1] - best[0]): best = (tgt, a)
    body = [(a, t) for a, t in ins if best and best[0] <= a <= best[1]] or ins
    hist_loop = collections.Counter(op(t) for _, t in body)
    with open(os.path.join(OUT, name + ".txt"), "w") as f:
        f.write(f"# {sym}\n# cuobjdump -sass of {obj} (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a -O3)\n")
        f.write(f"# instructions: {len(ins)} in the function, {len(body)} in the main loop [{best[0]:#x}, {best[1]:#x}]\n" if best else f"# instructions: {len(ins)}\n")
        f.write("# opcode histogram, whole function: " + ", ".join(f"{k} {v}" for k, v in hist_all.most_common(25)) + "\n")
        f.write("# opcode histogram, main loop:      " + ", ".join(f"{k} {v}" for k, v in hist_loop.most_common(25)) + "\n")
        f.write("#\n# issue order of the FP64 tensor / FMA instructions in the main loop (destination register = accumulator chain):\n")
        seq = [(a, t) for a, t in body if "DMMA" in t]
        if seq:
            f.write("# " + " ".join(re.search(r"DMMA\S*\s+(R\d+)", t).group(1) for _, t in seq) + "\n")
        f.write("#\n# main loop:\n")
        for a, t in body: f.write(f"{a:06x}  {t}\n")
    print(name, len(ins), len(body), dict(hist_loop.most_common(6)))
