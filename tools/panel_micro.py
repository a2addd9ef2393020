"""Isolated timings of the sharded-Cholesky building blocks (C ABI, csrc/gp.cu) on one GPU."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200.distributed import _DeviceOps
ops = _DeviceOps()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20001
st = torch.cuda.Stream()
info = torch.zeros(4, dtype=torch.int32, device="cuda")

def timed(fn, reps=7):
    with torch.cuda.stream(st):
        fn(); fn()
        st.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st); st.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

for h in (192, 384):
    A = torch.randn(h, h, dtype=torch.float64, device="cuda")
    D0 = A @ A.T / h + torch.eye(h, dtype=torch.float64, device="cuda")
    rows0 = torch.randn(h, N, dtype=torch.float64, device="cuda")
    rows0[:, :h] = D0
    rows = rows0.clone()
    inv = torch.zeros(h, h, dtype=torch.float64, device="cuda")
    def diag():
        rows[:, :h].copy_(D0)
        ops.diag(rows[:, :h], inv, info[:1], st.cuda_stream)
    def copy_only():
        rows[:, :h].copy_(D0)
    t_c, t_d = timed(copy_only), timed(diag)
    U = torch.triu(rows[:, :h])
    err_f = float((U.T @ U - D0).abs().max())
    err_i = float((inv @ U - torch.eye(h, dtype=torch.float64, device="cuda")).abs().max())     # inv (row-major) = U^-1
    print(json.dumps(dict(op="chol_diag (potrf + triangular inverse)", h=h, ms=round(t_d - t_c, 4), factor_err=err_f, inverse_err=err_i)), flush=True)
    out = torch.zeros(h, N, dtype=torch.float64, device="cuda")
    for w in (2 * h, N // 2, N - h):
        t = timed(lambda: ops.apply_inv(inv, rows[:, h:h + w], out[:, h:h + w], st.cuda_stream))
        print(json.dumps(dict(op="chol_apply_inv (trmm)", h=h, w=w, ms=round(t, 4), tflops=round(1.0 * h * h * w / (t * 1e-3) / 1e12, 2))), flush=True)
    for nb, w in ((h, 2 * h), (h, N), (h, N // 2), (4 * h, N)):
        C = torch.zeros(nb, N, dtype=torch.float64, device="cuda")
        t = timed(lambda: ops.update(C[:, :w], out[:, :nb], out[:, :w], st.cuda_stream))
        print(json.dumps(dict(op="chol_update (dgemm)", h=h, nb=nb, w=w, ms=round(t, 4), tflops=round(2.0 * h * nb * w / (t * 1e-3) / 1e12, 2))), flush=True)
