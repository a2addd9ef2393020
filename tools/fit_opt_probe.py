"""GaussianProcess.fit(opt=True) -- L-BFGS-B over (sigma, sigma0 | l, noise), gaussianprocess.py:84-127 -- on the headline force set
(N = 20 001), one GPU: wall time, number of LML evaluations, resulting hyper-parameters.   python tools/fit_opt_probe.py [force_atoms]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import synthetic
from csp_bo_b200.gaussianprocess import GaussianProcess
from csp_bo_b200.kernels import Dot_mb, RBF_mb

nF = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
X = synthetic.force_set(nF, seed=0)
y = np.random.default_rng(0).standard_normal(3 * nF)
for kern in (Dot_mb(para=[18.555440586011372, 0.01], bounds=[[1e-2, 5e+1], [1e-2, 1e+1]], zeta=2),
             RBF_mb(para=[6.244955524643115, 0.5611029935713949], bounds=[[1e-2, 5e+1], [1e-1, 1e+1]], zeta=2)):
    model = GaussianProcess(kern, f_coef=30, noise_e=[0.005, 0.002, 0.1])
    model.train_x = {"energy": [], "force": X}
    model.train_y = {"energy": [], "force": list(y.reshape(nF, 3))}
    model.update_y_train()
    calls = {"n": 0}
    orig = model.log_marginal_likelihood

    def counted(*a, **k):
        calls["n"] += 1
        return orig(*a, **k)
    model.log_marginal_likelihood = counted
    torch.cuda.synchronize(); t0 = time.perf_counter()
    model.fit(show=False, opt=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps(dict(kernel=kern.name, N=3 * nF, fit_opt_s=round(dt, 3), lml_evaluations=calls["n"],
                          params=[float(p) for p in kern.parameters()], noise_e=float(model.noise_e))), flush=True)
