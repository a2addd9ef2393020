import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from csp_bo_b200 import _lib, synthetic
from csp_bo_b200.packing import Pack, build_block
X = synthetic.force_set(6667, seed=0)
p = Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
out = torch.empty((p.m, 3, p.m * 3), dtype=torch.float64, device="cuda")
for _ in range(2):
    build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=3.0, symmetric=True, out=out)
torch.cuda.synchronize()
