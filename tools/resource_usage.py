"""Registers / spills / shared memory of every tile-kernel instantiation (ptxas -v), as a table.

    python tools/resource_usage.py [> profiles/resource_usage_rNN.txt]
Template arguments: MODE (0 K_EE, 1 K_EF, 2 K_FF), FAM (0 Dot, 1 RBF, 2 RBF with dK/dl), KS, zeta==2, CHUNKED, BS (B-run split), VK (skips zero k-steps).
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "csp_bo_b200", "csrc")
cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
       "-I", os.path.join(ROOT, "include"), "-I", SRC, "-Xptxas", "-v", "-c", os.path.join(SRC, sys.argv[1] if len(sys.argv) > 1 else "tiles.cu"),
       "-o", "/dev/null"]
err = subprocess.run(cmd, capture_output=True, text=True).stderr
rows, name = [], None
for ln in err.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(cspbo::TileParams\)|void cspbo::", "", name)
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", ln)
    if m:
        stack, st, ld = m.groups()
        continue
    m = re.search(r"Used (\d+) registers", ln)
    if m and name:
        rows.append((name, int(m.group(1)), int(stack), int(st), int(ld)))
        name = None
print("%-60s %5s %6s %8s %8s" % ("kernel", "regs", "stack", "spill_st", "spill_ld"))
for r in sorted(rows):
    print("%-60s %5d %6d %8d %8d" % r)
