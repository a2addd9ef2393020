"""RBF K_EE / K_EF / K_FF build times at the bench's leg sizes (the exp of the RBF epilogue shares the FP64 pipe with the DMMAs)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib, packing, synthetic
def dev_pack(T):
    if len(T) == 4:
        return packing.Pack(torch.from_numpy(T[0]).cuda(), torch.from_numpy(T[1]).cuda(), T[2], T[3])
    return packing.Pack(torch.from_numpy(T[0]).cuda(), None, T[1], T[2])
pf = dev_pack(synthetic.force_set(6667, seed=0))
pe = dev_pack(synthetic.energy_set(833, seed=1))
ps = dev_pack(synthetic.energy_set(5352, seed=1, single_element=True, rows_per_structure=64))
kw = dict(sigma2=6.244955524643115 ** 2, l=0.5611029935713949)
res = {}
def timed(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    res[name] = round(float(np.median(ts)), 3)
o1 = torch.empty((pe.m, pe.m), dtype=torch.float64, device="cuda")
timed("kee_833x192", lambda: packing.build_block(_lib.RBF, _lib.KEE, pe, pe, 2.0, symmetric=True, out=o1, **kw))
o4 = torch.empty((ps.m, ps.m), dtype=torch.float64, device="cuda")
timed("kee_si5352x64", lambda: packing.build_block(_lib.RBF, _lib.KEE, ps, ps, 2.0, symmetric=True, out=o4, **kw))
o2 = torch.empty((pe.m, pf.m, 3), dtype=torch.float64, device="cuda")
timed("kef_833x20001", lambda: packing.build_block(_lib.RBF, _lib.KEF, pe, pf, 2.0, out=o2, **kw))
o3 = torch.empty((pf.m, 3, pf.m * 3), dtype=torch.float64, device="cuda")
timed("kff_20001", lambda: packing.build_block(_lib.RBF, _lib.KFF, pf, pf, 2.0, symmetric=True, out=o3, tol=1e-10, use_tol=True, **kw))
timed("kee_grad_833", lambda: packing.build_block(_lib.RBF, _lib.KEE, pe, pe, 2.0, symmetric=True, with_grad=True, device=True, **kw))
print(json.dumps(res))
