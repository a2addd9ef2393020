"""K_FF / K_EF / K_EE at descriptor lengths that use the KS = 14 kernels (24 < d <= 56): Dot zeta = 2 / 2.5 and RBF, 10 002 force rows.
CSPBO_LIB selects an alternative library build.      python tools/ks14_probe.py [d ...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib, packing, synthetic


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return round(float(np.median(ts)), 2)


for d in [int(v) for v in sys.argv[1:]] or [30, 40, 50, 56]:
    X = synthetic.force_set(3334, seed=0, d=d)
    p = packing.Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
    E = synthetic.energy_set(416, seed=1, d=d)
    pe = packing.Pack(torch.from_numpy(E[0]).cuda(), None, E[1], E[2])
    out = torch.empty((p.m, 3, p.m * 3), dtype=torch.float64, device="cuda")
    oef = torch.empty((pe.m, p.m, 3), dtype=torch.float64, device="cuda")
    oee = torch.empty((pe.m, pe.m), dtype=torch.float64, device="cuda")
    res = {"d": d, "lib": os.path.basename(os.environ.get("CSPBO_LIB", "default"))}
    res["dot_z2_kff"] = timed(lambda: packing.build_block(_lib.DOT, _lib.KFF, p, p, 2.0, symmetric=True, out=out, sigma2=344.0))
    res["dot_z2.5_kff"] = timed(lambda: packing.build_block(_lib.DOT, _lib.KFF, p, p, 2.5, symmetric=True, out=out, scale=3.0))
    res["dot_z3_kff"] = timed(lambda: packing.build_block(_lib.DOT, _lib.KFF, p, p, 3.0, symmetric=True, out=out, scale=3.0))
    res["rbf_z3_kff"] = timed(lambda: packing.build_block(_lib.RBF, _lib.KFF, p, p, 3.0, symmetric=True, out=out, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True))
    res["rbf_kff"] = timed(lambda: packing.build_block(_lib.RBF, _lib.KFF, p, p, 2.0, symmetric=True, out=out, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True))
    res["dot_z2_kef"] = timed(lambda: packing.build_block(_lib.DOT, _lib.KEF, pe, p, 2.0, out=oef, sigma2=344.0))
    res["rbf_kee"] = timed(lambda: packing.build_block(_lib.RBF, _lib.KEE, pe, pe, 2.0, symmetric=True, out=oee, sigma2=39.0, l=0.56))
    res["check"] = float(out[5, 1, 100]) + float(oef[3, 7, 1]) + float(oee[2, 9])
    print(json.dumps(res), flush=True)
    del p, pe, out, oef, oee
