import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib, synthetic, packing
from csp_bo_b200.packing import Pack, build_block
from csp_bo_b200.kernels.dot_kernel import kff_C
X = synthetic.force_set(6667, seed=0)
m = len(X[3])
xh, dxh = torch.from_numpy(X[0]).pin_memory(), torch.from_numpy(X[1]).pin_memory()
Xh = (xh.numpy(), dxh.numpy(), X[2], X[3])
p = Pack(Xh[0], Xh[1], Xh[2], Xh[3])
outh = torch.empty(m * 3 * m * 3, dtype=torch.float64).pin_memory()
o = outh.numpy()
def t(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return round((time.perf_counter() - t0) * 1e3, 1)
print("build->pinned", [t(lambda: build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=3.0, symmetric=True, out=o)) for _ in range(6)])
def full():
    kff_C(Xh, Xh, 18.55, 0.01, 2.0, out=o)
print("kff_C e2e    ", [t(full) for _ in range(6)])
print("build->pinned", [t(lambda: build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=3.0, symmetric=True, out=o)) for _ in range(4)])
pc = Pack(Xh[0], Xh[1], Xh[2], Xh[3], row_classes=6)
outd = torch.empty((m, 3, m * 3), dtype=torch.float64, device="cuda")
print("device build, plain pack  ", [t(lambda: build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=3.0, symmetric=True, out=outd)) for _ in range(3)])
print("device build, 6-class pack", [t(lambda: build_block(_lib.DOT, _lib.KFF, pc, pc, 2.0, scale=3.0, symmetric=True, out=outd)) for _ in range(3)])
print("host build, 6-class pack  ", [t(lambda: build_block(_lib.DOT, _lib.KFF, pc, pc, 2.0, scale=3.0, symmetric=True, out=o)) for _ in range(3)])
print("pack only (pinned host)    ", [t(lambda: Pack(Xh[0], Xh[1], Xh[2], Xh[3], row_classes=6)) for _ in range(3)])
