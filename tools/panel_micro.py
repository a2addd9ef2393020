"""Isolated timings of the sharded-Cholesky panel operations (C ABI, csrc/gp.cu) on one GPU."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib
from csp_bo_b200._lib import c_vp
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20001
st = torch.cuda.Stream()
info = torch.zeros(4, dtype=torch.int32, device="cuda")

def timed(fn, reps=5):
    with torch.cuda.stream(st):
        fn(); fn()
        st.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); fn(); e1.record(st); st.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))

for h in (192, 384, 768):
    A = torch.randn(h, h, dtype=torch.float64, device="cuda")
    D0 = A @ A.T / h + torch.eye(h, dtype=torch.float64, device="cuda")
    for ncols in (N, N // 2, h):
        rows0 = torch.randn(h, N, dtype=torch.float64, device="cuda")
        rows0[:, :h] = D0
        rows = rows0.clone()
        def factor():
            rows.copy_(rows0)
            _lib.check(lib.cspbo_gp_panel_factor(c_vp(rows.data_ptr()), N, 0, h, ncols, c_vp(info.data_ptr()), c_vp(st.cuda_stream)))
        def copy_only():
            rows.copy_(rows0)
        t_f, t_c = timed(factor), timed(copy_only)
        print(json.dumps(dict(op="panel_factor(potrf+trsm)", h=h, ncols=ncols, ms=round(t_f - t_c, 4), copy_ms=round(t_c, 4))), flush=True)
    panel = torch.randn(h, N, dtype=torch.float64, device="cuda")
    for nb, ncols in ((h, N), (h, N // 2), (4 * h, N), (N // 4 // h * h, N)):
        C = torch.zeros(nb, N, dtype=torch.float64, device="cuda")
        def upd():
            _lib.check(lib.cspbo_gp_panel_update(c_vp(panel.data_ptr()), N, h, c_vp(C.data_ptr()), N, 0, nb, ncols, c_vp(st.cuda_stream)))
        t = timed(upd)
        print(json.dumps(dict(op="panel_update(dgemm)", h=h, nb=nb, ncols=ncols, ms=round(t, 4), tflops=round(2.0 * h * nb * ncols / (t * 1e-3) / 1e12, 2))), flush=True)
    # alternatives for the rest-of-panel solve: explicit inverse of the diagonal factor + dgemm
    U = torch.linalg.cholesky(D0, upper=True)
    R = torch.randn(h, N - h, dtype=torch.float64, device="cuda")
    eye = torch.eye(h, dtype=torch.float64, device="cuda")
    def trsm_torch():
        torch.linalg.solve_triangular(U.T, R, upper=False)
    def inv_gemm():
        Ui = torch.linalg.solve_triangular(U, eye, upper=True)
        torch.mm(Ui.T, R)
    def potrf_torch():
        torch.linalg.cholesky(D0, upper=True)
    with torch.cuda.stream(st):
        print(json.dumps(dict(op="alternatives", h=h, potrf_torch_ms=round(timed(potrf_torch), 4), trsm_torch_ms=round(timed(trsm_torch), 4),
                              inv_plus_gemm_ms=round(timed(inv_gemm), 4))), flush=True)
