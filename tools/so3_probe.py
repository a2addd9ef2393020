import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200.SO3 import SO3
from oracle import so3_reference as so
rng = np.random.default_rng(5)
g = np.array([(i, j, k) for i in range(4) for j in range(6) for k in range(8)], dtype=float)
cell = np.diag([4 * 2.6, 6 * 2.6, 8 * 2.6])
pos = g * 2.6 + rng.normal(0, 0.25, g.shape)
num = rng.choice([1, 8, 78], size=len(pos), p=[0.5, 0.3, 0.2])
st = so.Structure(pos, num, cell)
desc = SO3(nmax=3, lmax=3, rcut=3.5, alpha=2.0, derivative=True, stress=True)
d = desc.calculate(st)
ts = []
for _ in range(5):
    t0 = time.perf_counter(); d = desc.calculate(st); ts.append((time.perf_counter() - t0) * 1e3)
print("GPU SO3.calculate 192 atoms (stress=True): min %.2f ms, n_seq %d" % (min(ts), len(d["seq"])))
t0 = time.perf_counter(); ref = so.so3_calculate(st, nmax=3, lmax=3, rcut=3.5, alpha=2.0, stress=True); print("NumPy restatement of the reference: %.1f ms" % ((time.perf_counter() - t0) * 1e3))
import ctypes
from csp_bo_b200 import _lib
lib = _lib.load(); h = desc._handle()
posc = np.ascontiguousarray(st.positions); numc = np.ascontiguousarray(st.numbers, dtype=np.int32); cellc = np.ascontiguousarray(st.cell); pbc = np.ones(3, dtype=np.int32)
nseq, nc = ctypes.c_int(0), ctypes.c_int(0)
for stress in (1, 0):
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        lib.cspbo_so3_compute(h, len(posc), posc.ctypes.data, numc.ctypes.data, cellc.ctypes.data, pbc.ctypes.data, stress, ctypes.byref(nseq), ctypes.byref(nc))
        ts.append((time.perf_counter() - t0) * 1e3)
    print("compute only stress=%d: %.2f ms" % (stress, min(ts)))
x = np.empty((192, 24)); dx = np.empty((nseq.value, 24, 3)); rd = np.empty((nseq.value, 24, 3, 3)); sq = np.empty((nseq.value, 2), dtype=np.int64)
lib.cspbo_so3_compute(h, len(posc), posc.ctypes.data, numc.ctypes.data, cellc.ctypes.data, pbc.ctypes.data, 1, ctypes.byref(nseq), ctypes.byref(nc))
t0 = time.perf_counter(); lib.cspbo_so3_fetch(h, x.ctypes.data, dx.ctypes.data, rd.ctypes.data, sq.ctypes.data); print("fetch: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
