"""Which fit is closer to the truth?  alpha from (a) one cuSOLVER potrf/potrs and (b) the sharded panel Cholesky
(world 1), both compared with an iteratively refined solution of the same system (residuals in fp64, 4 rounds)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib, gp, synthetic
from csp_bo_b200.distributed import ShardedSymmetricBlock, fit_alpha_sharded
from csp_bo_b200.packing import Pack, build_block
n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
X = synthetic.force_set(n_atoms, seed=0)
p = Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
N = 3 * n_atoms
s2 = 18.555440586011372 ** 2 * 2.0
nf2 = (30.0 * 0.005) ** 2
y = torch.from_numpy(np.random.default_rng(0).standard_normal(N)).cuda()
K = torch.empty((p.m, 3, p.m * 3), dtype=torch.float64, device="cuda")
build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=s2, symmetric=True, out=K)
K = K.view(N, N)
L = K.clone()
a_cus = gp.fit_alpha(L, y, n_energy=0, noise_e=0.005, f_coef=30.0, overwrite=True)     # L now holds the factor
truth = a_cus.clone()
for it in range(4):
    r = y - (K @ truth + nf2 * truth)
    truth += gp.cho_solve(L, r.reshape(N, 1)).reshape(-1)
res = {}
for G in (8, 16):
    sh = ShardedSymmetricBlock(p, _lib.KFF, sb_blocks=G, zigzag=True, packed_cols=True)
    sh.build(_lib.DOT, 2.0, scale=s2)
    a_sh = fit_alpha_sharded(sh, y, nf2)
    res["sharded_G%d" % G] = a_sh
    sh.close()
Kt = K[:512]
pt = Kt @ truth
def e(a):
    return dict(alpha_rel=float((a - truth).abs().max() / truth.abs().max()), pred_rel=float((Kt @ a - pt).abs().max() / pt.abs().max()))
print(json.dumps(dict(N=N, cusolver=e(a_cus), **{k: e(v) for k, v in res.items()},
                      residual_truth=float((y - (K @ truth + nf2 * truth)).abs().max()))))

# ---- the reference's own solver on the same matrix: SciPy cholesky + cho_solve (gaussianprocess.py:116,127) on the host
if len(sys.argv) > 2 and sys.argv[2] == "scipy":
    import time
    from scipy.linalg import cho_factor, cho_solve
    Kh = K.cpu().numpy()
    Kh[np.diag_indices(N)] += nf2
    t0 = time.perf_counter()
    a_sp = cho_solve(cho_factor(Kh, lower=True, overwrite_a=True, check_finite=False), y.cpu().numpy(), check_finite=False)
    t_sp = time.perf_counter() - t0
    a_sp = torch.from_numpy(a_sp).cuda()
    psp = Kt @ a_sp
    cmp_ = lambda a: float((Kt @ a - psp).abs().max() / psp.abs().max())
    print(json.dumps(dict(scipy_seconds=round(t_sp, 1), scipy_vs_truth=e(a_sp), pred_rel_vs_scipy=dict(cusolver=cmp_(a_cus), refined=cmp_(truth),
                                                                                                     **{k: cmp_(v) for k, v in res.items()}))))
