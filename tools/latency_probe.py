"""Prediction-path latency (config 1 shape: one 192-atom Pt/H/O structure against a trained model)
and d=50 throughput."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib, synthetic, packing
from csp_bo_b200.kernels import Dot_mb, RBF_mb
from csp_bo_b200.gaussianprocess import GaussianProcess
from csp_bo_b200.utilities import tuple_to_list
from csp_bo_b200.packing import Pack, build_block

def T(f, n=5):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return round(min(ts), 3), round(float(np.median(ts)), 3)

for nE, nF in ((16, 409), (100, 3000)):
    trainE = synthetic.energy_set(nE, seed=1)
    trainF = synthetic.force_set(nF, seed=2)
    rng = np.random.default_rng(0)
    E = [(x, float(rng.normal()), ele) for x, ele in tuple_to_list(trainE, mode='energy')]
    F = [(x, dx, rng.normal(size=3), ele) for x, dx, ele in tuple_to_list(trainF)]
    for kern in (Dot_mb(para=[18.55, 0.01], zeta=2), RBF_mb(para=[6.24, 0.56], zeta=2)):
        model = GaussianProcess(kern, f_coef=30, noise_e=[0.005, 0.002, 0.1])
        t0 = time.perf_counter(); model.fit({"energy": E, "force": F}, show=False, opt=False); torch.cuda.synchronize()
        tfit = (time.perf_counter() - t0) * 1e3
        testE = synthetic.energy_set(1, seed=7)
        testF = synthetic.force_set(192, seed=8)
        data = {"energy": tuple_to_list(testE, mode='energy'), "force": tuple_to_list(testF)}
        datat = {"energy": testE, "force": testF}
        print(kern.name, "train %dE+%dF: fit(opt=False) %.1f ms | predict(list) min/med %s | predict(tuple) %s | k_total(list,train)->numpy %s | return_std %s"
              % (nE, nF, tfit, T(lambda: model.predict(data)), T(lambda: model.predict(datat)),
                 T(lambda: kern.k_total(data, model.train_x)), T(lambda: model.predict(datat, return_std=True))), flush=True)

# d = 50 (AgI.json: nmax = lmax = 4) throughput, single species
rng = np.random.default_rng(3)
m, d = 2000, 50
counts = rng.integers(11, 35, size=m)
n = int(counts.sum())
x = (np.abs(rng.standard_normal((n, d))) + 0.05) * 30
dx = rng.standard_normal((n, d, 3)) * 10
p = Pack(torch.from_numpy(x).cuda(), torch.from_numpy(dx).cuda(), np.full(n, 47, np.int32), list(counts))
out = torch.empty((m, 3, m * 3), dtype=torch.float64, device="cuda")
pairs = p.pair_count(p)
for fam in (_lib.DOT, _lib.RBF):
    ms = T(lambda: build_block(fam, _lib.KFF, p, p, 2.0, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True, symmetric=True, out=out), 3)[0]
    print("d=50 fam=%d: %.2f ms  %.2f TFLOP/s (2*16*d+100 = %d flop/pair)" % (fam, ms, (pairs + n) / 2 * (32 * d + 100) / (ms * 1e-3) / 1e12, 32 * d + 100))
