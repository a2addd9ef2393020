import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import gp
for n, N in ((5770, 20001), (20001, 20001), (20000, 20000), (40002, 40002)):
    Ks = torch.randn((n, N), dtype=torch.float64, device="cuda"); al = torch.randn(N, dtype=torch.float64, device="cuda")
    ref = Ks @ al
    got = gp.predict_mean(Ks, al)
    for _ in range(3): gp.predict_mean(Ks, al)
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gp.predict_mean(Ks, al); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = float(np.median(ts))
    tt = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); Ks @ al; e1.record(); torch.cuda.synchronize(); tt.append(e0.elapsed_time(e1))
    print(json.dumps(dict(variant=os.environ.get("CSPBO_GEMV", "0"), n=n, N=N, ms=round(t, 4), gbs=round(8.0 * n * N / t / 1e6, 1),
                          cublas_gbs=round(8.0 * n * N / float(np.median(tt)) / 1e6, 1), err=float((got - ref).abs().max() / ref.abs().max()))), flush=True)
    del Ks, al
