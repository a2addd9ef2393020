"""Phase times of one log_marginal_likelihood(eval_gradient=True) evaluation at N = 20 001 on one GPU
(gaussianprocess.log_marginal_likelihood): K build, potrf + potrs, trtri / potri, RBF dK/dl trace.
    python tools/lml_breakdown.py [force_atoms]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import gp, synthetic
from csp_bo_b200.kernels import Dot_mb, RBF_mb
from csp_bo_b200.kernels import device as kdev
from csp_bo_b200.packing import resident

nF = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
F = synthetic.force_set(nF, seed=0)
data = resident({"force": F})
N = 3 * nF
y = torch.from_numpy(np.random.default_rng(3).standard_normal(N)).cuda()


def t(fn, reps=1):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    torch.cuda.synchronize()
    return r, (time.perf_counter() - t0) * 1e3 / reps


for kern, name in ((Dot_mb(para=[18.555440586011372, 0.01], zeta=2), "Dot"), (RBF_mb(para=[6.244955524643115, 0.5611029935713949], zeta=2), "RBF")):
    for rep in range(2):
        out = dict(kernel=name, N=N)
        K, out["build_ms"] = t(lambda: kdev.k_total_device(kern, data, grad_semantics=True))
        (alpha, logdet), out["potrf_potrs_ms"] = t(lambda: gp.fit_alpha(K, y, n_energy=0, noise_e=0.005, f_coef=30.0, return_logdet=True, overwrite=True))
        if name == "Dot":
            dinv, out["inverse_diag_trtri_ms"] = t(lambda: gp.inverse_diag(K))
        else:
            _, out["potri_ms"] = t(lambda: gp.cholesky_inverse_inplace(K))
            _, out["rbf_dl_trace_ms"] = t(lambda: kdev.rbf_dl_trace(kern, data, alpha.clone(), K))
        del K
        if rep:
            print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)
