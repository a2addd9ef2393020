"""Where does the end-to-end time of kff_C(host, host) go?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib, synthetic, packing
from csp_bo_b200.packing import Pack, build_block
from csp_bo_b200.kernels.dot_kernel import kff_C

def T(f, n=3):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), r

X = synthetic.force_set(6667, seed=0)
m = len(X[3])
xh, dxh = torch.from_numpy(X[0]).pin_memory(), torch.from_numpy(X[1]).pin_memory()
Xh = (xh.numpy(), dxh.numpy(), X[2], X[3])
print("pack from pageable host  %.1f ms" % T(lambda: Pack(X[0], X[1], X[2], X[3]))[0])
t, p = T(lambda: Pack(Xh[0], Xh[1], Xh[2], Xh[3])); print("pack from pinned host    %.1f ms" % t)
outd = torch.empty((m, 3, m * 3), dtype=torch.float64, device="cuda")
print("build -> device          %.1f ms" % T(lambda: build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=3.0, symmetric=True, out=outd))[0])
outh = torch.empty(m * 3 * m * 3, dtype=torch.float64).pin_memory()
print("torch D2H 3.2GB pinned   %.1f ms" % T(lambda: outh.copy_(outd.view(-1)))[0])
print("build -> pinned host     %.1f ms" % T(lambda: build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=3.0, symmetric=True, out=outh.numpy()))[0])
outp = np.empty(m * 3 * m * 3)
print("build -> pageable host   %.1f ms" % T(lambda: build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=3.0, symmetric=True, out=outp), 2)[0])
def full():
    return kff_C(Xh, Xh, 18.55, 0.01, 2.0, out=outh.numpy())
print("kff_C(host,host,out=pin) %.1f ms" % T(full)[0])
def alloc():
    t = torch.empty(m * 3 * m * 3, dtype=torch.float64, device="cuda"); del t
import ctypes
lib = _lib.load()
def rawalloc():
    ptr = ctypes.c_void_p(); lib.cspbo_dev_alloc(8 * m * 3 * m * 3, ctypes.byref(ptr)); lib.cspbo_dev_free(ptr)
print("cudaMalloc+Free 3.2GB    %.1f ms" % T(rawalloc)[0])
