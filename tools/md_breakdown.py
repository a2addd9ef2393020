"""Where one predict_structure call (192-atom Pt/H/O structure, 16 E + 409 F model) spends its time."""
import os, sys, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from csp_bo_b200 import _lib, gp, synthetic, packing
from csp_bo_b200.SO3 import SO3
from csp_bo_b200.gaussianprocess import GaussianProcess, _Z
from csp_bo_b200.kernels import Dot_mb, device as kdev
from csp_bo_b200.utilities import tuple_to_list

class Struct:
    def __init__(self, pos, num, cell):
        self.positions, self.numbers, self._cell, self.pbc = pos, num, cell, np.array([True] * 3)
        self.symbols = [{1: "H", 8: "O", 78: "Pt"}[int(z)] for z in num]
    def get_cell(self): return self._cell
    def __len__(self): return len(self.positions)

rng = np.random.default_rng(5)
g = np.array([(i, j, k) for i in range(4) for j in range(6) for k in range(8)], dtype=float)
st = Struct(g * 2.6 + rng.normal(0, 0.25, g.shape), rng.choice([1, 8, 78], size=192, p=[0.5, 0.3, 0.2]), np.diag([4 * 2.6, 6 * 2.6, 8 * 2.6]))
trainE, trainF = synthetic.energy_set(16, seed=1), synthetic.force_set(409, seed=2)
E = [(x, float(rng.normal()), ele) for x, ele in tuple_to_list(trainE, mode='energy')]
F = [(x, dx, rng.normal(size=3), ele) for x, dx, ele in tuple_to_list(trainF)]
desc = SO3(nmax=3, lmax=3, rcut=3.5, alpha=2.0, derivative=True, stress=True)
model = GaussianProcess(Dot_mb(para=[18.555, 0.01], zeta=2), descriptor=desc, f_coef=30, noise_e=[0.005, 0.002, 0.1])
with contextlib.redirect_stdout(sys.stderr):
    model.fit({"energy": E, "force": F}, show=False, opt=False)

def T(f, n=9):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return round(float(np.median(ts)), 3), r

for stress in (False, True):
    print("---- stress =", stress)
    t, _ = T(lambda: desc._compute(st, stress)); print("so3 _compute (host nbr search + H2D + 4 kernels + sync)", t)
    t, d = T(lambda: desc.calculate_device(st, stress=stress)); print("calculate_device", t)
    def gather():
        ele = np.array([_Z[e] for e in d['elements']])
        seq = d['seq']; order = np.argsort(seq[:, 1], kind='stable'); centres = seq[order, 0]
        counts = np.bincount(seq[:, 1], minlength=192)
        od = torch.as_tensor(order, device="cuda")
        dX = d['dxdr'].index_select(0, od)
        if stress:
            rd = d['rdxdr'].index_select(0, od).reshape(len(order), dX.shape[1], 9)[:, :, [0, 4, 8, 1, 2, 5]]
            dX = torch.cat((dX, rd), dim=2)
        return {"energy": (d['x'], ele, [len(ele)]), "force": (d['x'].index_select(0, torch.as_tensor(centres, device="cuda")), dX, ele[centres], [int(c) for c in counts])}
    t, data = T(gather); print("host bookkeeping + device gather", t)
    t, pe = T(lambda: packing.spread_energy(data["energy"])); print("spread_energy pack", t)
    t, pf = T(lambda: packing.pack_force(data["force"])); print("force pack", t)
    train = model._fitted()
    pk = {"energy": pe, "force": pf}
    if stress:
        t, K = T(lambda: kdev.k_total_with_stress_device(model.kernel, pk, train, 1e-8)); print("k_total_with_stress_device (packs given)", t)
    else:
        t, K = T(lambda: kdev.k_total_device(model.kernel, pk, train, 1e-8)); print("k_total_device (packs given)", t)
    Kt = K[0] if stress else K
    t, _ = T(lambda: gp.predict_mean(Kt, model._alpha_dev).cpu().numpy()); print("gemv + D2H", t)
    t, _ = T(lambda: model.predict_structure(st, stress=stress)); print("predict_structure total", t)
