"""Quick K_FF timing at the bench size (Pt/H/O mix + single species, Dot + RBF, symmetric)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib, synthetic
from csp_bo_b200.packing import Pack, build_block
n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
for single in (False, True):
    X = synthetic.force_set(n_atoms, seed=0, single_element=single)
    p = Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
    pairs = p.pair_count(p)
    out = torch.empty((p.m, 3, p.m * 3), dtype=torch.float64, device="cuda")
    for fam in (_lib.DOT, _lib.RBF):
        ms = []
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            build_block(fam, _lib.KFF, p, p, 2.0, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True, symmetric=True, out=out)
            e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        best = min(ms)
        print(json.dumps(dict(single=single, fam=fam, ms=round(best, 3), tflops=round((pairs + p.n) / 2 * 868 / (best * 1e-3) / 1e12, 2))), flush=True)
