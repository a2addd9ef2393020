"""SASS evidence for the tile kernels: instruction-mix counts per instantiation and an excerpt of the headline kernel's
inner loop (cuobjdump -sass of the built library).   python tools/sass_summary.py > profiles/sass_rNN.md"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "csp_bo_b200", "lib", "libcspbo_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for ln in txt.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(cspbo::TileParams\)|void cspbo::", "", cur)
        funcs[cur] = []
    elif cur is not None and re.search(r"/\*[0-9a-f]{4}\*/", ln):
        funcs[cur].append(ln)
KEYS = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDS.128", "LDG.E.128", "LDG", "STG", "STS", "STL", "LDL", "MUFU", "BAR", "SHFL", "CALL"]
T = ", false, false>"      # BS (B-run split), VK (skips zero k-steps)
want = ["block_tile_kernel<2, 0, 6, true, false" + T, "block_tile_kernel<2, 0, 14, true, false" + T, "block_tile_kernel<2, 0, 14, true, false, false, true>",
        "block_tile_kernel<2, 0, 6, true, true" + T,
        "block_tile_kernel<2, 1, 6, true, false" + T, "block_tile_kernel<2, 2, 6, true, false" + T, "block_tile_kernel<2, 2, 14, true, false" + T,
        "block_tile_kernel<1, 0, 6, true, false" + T, "block_tile_kernel<0, 0, 6, true, false" + T, "block_tile_kernel<0, 1, 6, true, false" + T,
        "block_tile_kernel<0, 1, 6, true, false, true, false>"]
print("# SASS instruction mix of the tile kernels (cuobjdump -sass csp_bo_b200/lib/libcspbo_b200.so, sm_100a)\n")
print("Template arguments: MODE (0 K_EE, 1 K_EF, 2 K_FF), FAM (0 Dot, 1 RBF, 2 RBF + dK/dl), KS, zeta == 2, CHUNKED, BS (B-run split), VK (skips zero k-steps).")
print("`STL`/`LDL` = local-memory (spill) instructions; registers and spill bytes per instantiation: profiles/resource_usage_r02.txt.\n")
print("| kernel | instr | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
for name in want:
    body = funcs.get(name)
    if body is None:
        continue
    ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s*(@!?U?P\d+\s+)?", "", l).split(";")[0].strip() for l in body]
    cnt = [sum(1 for o in ops if o.startswith(k)) for k in KEYS]
    print("| `%s` | %d | " % (name, len(ops)) + " | ".join(str(c) for c in cnt) + " |")
head = funcs.get(want[0], [])
ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", l).split(";")[0].rstrip() for l in head]
# the inner loop: the longest window that is dense in DMMA
best, bi = -1, 0
W = 120
for i in range(0, max(1, len(ops) - W)):
    c = sum(1 for o in ops[i:i + W] if "DMMA" in o)
    if c > best:
        best, bi = c, i
print("\n## Headline kernel `%s`: %d-instruction window with the most DMMAs (%d)\n\n```" % (want[0], W, best))
print("\n".join(ops[bi:bi + W]))
print("```\n\n## Bulk-TMA / mbarrier instructions of the same kernel\n\n```")
print("\n".join(o for o in ops if any(k in o for k in ("UBLKCP", "SYNCS", "FENCE", "ELECT"))))
print("```")
