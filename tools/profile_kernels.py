"""One launch of every kernel of the path at a representative size, for ncu
(`ncu --set full -k regex:"block_tile|predict_gemv|gemv_reduce|pack_rows|so3" ...`).

Order of the launches (each twice: warm-up + the one to read):
  pack_rows (force 6667 atoms, energy 833 structures), K_EF Dot, K_EE Dot (sym), K_EF RBF, K_EE RBF (sym),
  K_FF RBF (sym), K_FF Dot (sym), predict_gemv [20001 x 20001], SO3 of a 192-atom structure, tri_inverse (h = 384).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from csp_bo_b200 import _lib, gp, synthetic
from csp_bo_b200.packing import Pack, build_block
from csp_bo_b200.SO3 import SO3

n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
reps = 2
X = synthetic.force_set(n_atoms, seed=0)
E = synthetic.energy_set(max(8, n_atoms // 8), seed=1)
for _ in range(reps):
    pf = Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
    pe = Pack(torch.from_numpy(E[0]).cuda(), None, E[1], E[2])
kef = torch.empty((pe.m, pf.m, 3), dtype=torch.float64, device="cuda")
kee = torch.empty((pe.m, pe.m), dtype=torch.float64, device="cuda")
kff = torch.empty((pf.m, 3, pf.m * 3), dtype=torch.float64, device="cuda")
dot = dict(sigma2=344.3, sigma02=1e-4)
rbf = dict(sigma2=39.0, l=0.5611029935713949, tol=1e-10)
for _ in range(reps):
    build_block(_lib.DOT, _lib.KEF, pe, pf, 2.0, out=kef, **dot)
for _ in range(reps):
    build_block(_lib.DOT, _lib.KEE, pe, pe, 2.0, symmetric=True, out=kee, **dot)
for _ in range(reps):
    build_block(_lib.RBF, _lib.KEF, pe, pf, 2.0, out=kef, **rbf)
for _ in range(reps):
    build_block(_lib.RBF, _lib.KEE, pe, pe, 2.0, symmetric=True, out=kee, **rbf)
for _ in range(reps):
    build_block(_lib.RBF, _lib.KFF, pf, pf, 2.0, symmetric=True, use_tol=True, out=kff, **rbf)
for _ in range(reps):
    build_block(_lib.DOT, _lib.KFF, pf, pf, 2.0, symmetric=True, out=kff, **dot)
N = 3 * pf.m
al = torch.randn(N, dtype=torch.float64, device="cuda")
for _ in range(reps):
    p = gp.predict_mean(kff.view(N, N), al)
torch.cuda.synchronize()


class Struct:
    def __init__(self, pos, num, cell):
        self.positions, self.numbers, self._cell, self.pbc = pos, num, cell, np.array([True] * 3)
        self.symbols = [{1: "H", 8: "O", 78: "Pt"}[int(z)] for z in num]

    def get_cell(self):
        return self._cell

    def __len__(self):
        return len(self.positions)


rng = np.random.default_rng(5)
g = np.array([(i, j, k) for i in range(4) for j in range(6) for k in range(8)], dtype=float)
st = Struct(g * 2.6 + rng.normal(0, 0.25, g.shape), rng.choice([1, 8, 78], size=192, p=[0.5, 0.3, 0.2]),
            np.diag([4 * 2.6, 6 * 2.6, 8 * 2.6]))
desc = SO3(nmax=3, lmax=3, rcut=3.5, alpha=2.0, derivative=True, stress=True)
for _ in range(reps):
    desc.calculate(st)
torch.cuda.synchronize()
# the hand-written triangular inverse of the sharded Cholesky's diagonal chain (h = 384)
from csp_bo_b200.distributed import _DeviceOps
A = torch.randn(384, 384, dtype=torch.float64, device="cuda")
D0 = A @ A.T / 384 + torch.eye(384, dtype=torch.float64, device="cuda")
inv = torch.zeros(384, 384, dtype=torch.float64, device="cuda")
info = torch.zeros(1, dtype=torch.int32, device="cuda")
for _ in range(reps):
    Dw = D0.clone()
    _DeviceOps().diag(Dw, inv, info, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("ok", float(p[0]))
