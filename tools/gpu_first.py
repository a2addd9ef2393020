"""First contact with the GPU: per-case parity errors + K_FF timings at a few sizes."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tests._cases import ProductImpl, golden_cases, load_fixtures, load_golden, max_norm_err
from csp_bo_b200 import _lib, synthetic
from csp_bo_b200.packing import Pack, build_block

fx, gold = load_fixtures(), load_golden()
t = time.time()
try:
    out = golden_cases(fx, ProductImpl())
    errs = {k: max_norm_err(out[k], gold[k]) for k in gold}
    for k, v in sorted(errs.items(), key=lambda kv: -kv[1])[:15]:
        print("%-28s %.3e" % (k, v))
    print("worst", max(errs.values()), "bad", [k for k, v in errs.items() if not v <= 1e-10], "time", time.time() - t)
except Exception as e:
    import traceback; traceback.print_exc()

lib = _lib.load()
res = []
for single in (False, True):
    for n_atoms in (192, 1667, 6667):
        X = synthetic.force_set(n_atoms, seed=0, single_element=single)
        xd, dxd = torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda()
        t0 = time.time(); p = Pack(xd, dxd, X[2], X[3]); torch.cuda.synchronize(); tpack = time.time() - t0
        pairs = p.pair_count(p)
        outb = torch.empty((p.m, 3, p.m * 3), dtype=torch.float64, device="cuda")
        for sym in (False, True):
            for fam in (_lib.DOT, _lib.RBF):
                ms = []
                for it in range(3):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    torch.cuda.synchronize(); e0.record()
                    build_block(fam, _lib.KFF, p, p, 2.0, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True, scale=1.0, symmetric=sym, out=outb)
                    e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
                best = min(ms)
                eff_pairs = pairs / 2 if sym else pairs
                r = dict(single=single, n_atoms=n_atoms, rows=p.n, sym=sym, fam=fam, ms=best, pack_s=tpack,
                         entries_per_s=9.0 * p.m * p.m / (best * 1e-3), tflops=eff_pairs * 868 / (best * 1e-3) / 1e12)
                res.append(r); print(json.dumps(r), flush=True)
        del outb, p
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/first_timings.json", "w"), indent=1)
