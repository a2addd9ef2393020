import sys, os, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from csp_bo_b200 import _lib, packing, synthetic
X = synthetic.force_set(3334, seed=0, d=50)
p = packing.Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
out = torch.empty((p.m, 3, p.m * 3), dtype=torch.float64, device="cuda")
og = torch.empty_like(out)
res = {}
for name, kw in (("rbf", dict(sigma2=39.0, l=0.56, tol=1e-10, use_tol=True)), ("rbf_grad", dict(sigma2=39.0, l=0.56, with_grad=True)),):
    fn = lambda: packing.build_block(_lib.RBF, _lib.KFF, p, p, 2.0, symmetric=True, out=out, out_grad=og, **kw)
    for _ in range(2): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    res[name] = round(float(np.median(ts)), 2)
fn = lambda: packing.build_block(_lib.DOT, _lib.KFF, p, p, 2.5, symmetric=True, out=out, scale=3.0)
for _ in range(2): fn()
torch.cuda.synchronize(); ts = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
res["dot_z2.5"] = round(float(np.median(ts)), 2)
print(os.environ.get("CSPBO_KS14_CHUNKED"), json.dumps(res))
