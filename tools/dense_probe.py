"""dense.cu against cuSOLVER at N = 20 001 (one process per setting: the switches are environment variables), plus the
effect of an odd leading dimension on cuBLAS dgemm.   python tools/dense_probe.py"""
import json, os, subprocess, sys, time
if len(sys.argv) > 1:
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch
    from csp_bo_b200 import gp
    N = int(sys.argv[1])
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(N, 2048, dtype=torch.float64, device="cuda", generator=g)
    K = A @ A.T / 2048 + torch.eye(N, dtype=torch.float64, device="cuda")
    del A
    y = torch.randn(N, dtype=torch.float64, device="cuda", generator=g)
    out = dict(env={k: v for k, v in os.environ.items() if k.startswith("CSPBO_DENSE")}, N=N)
    for name, fn in (("inverse_diag_ms", gp.inverse_diag), ("cholesky_inverse_ms", gp.cholesky_inverse_inplace)):
        ts = []
        for _ in range(3):
            F = K.clone()
            gp.fit_alpha(F, y, n_energy=0, noise_e=0.0, f_coef=1.0, overwrite=True)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = fn(F)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        out[name] = round(min(ts), 2)
        out[name + "_check"] = float(r.sum()) if r.dim() == 1 else float(torch.triu(r).sum())
    if "gemm" in sys.argv:
        for ld in (N, (N + 3) // 4 * 4):
            B = torch.randn(N, ld, dtype=torch.float64, device="cuda", generator=g)
            h = N // 2
            a, b = B[:h, :h], B[h:2 * h, :h]
            for _ in range(2):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                c = a @ b
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
            out["dgemm_%d_ld%d_tflops" % (h, ld)] = round(2 * h ** 3 / dt / 1e12, 2)
            del B, c
    print(json.dumps(out), flush=True)
else:
    N = "20001"
    for env, extra in (({"CSPBO_DENSE": "0"}, ["gemm"]), ({}, []), ({"CSPBO_DENSE_WS_MB": "96"}, []),
                       ({"CSPBO_DENSE_WS_MB": "800"}, []), ({"CSPBO_DENSE_NB": "256"}, [])):
        subprocess.run([sys.executable, __file__, N] + extra, env=dict(os.environ, **env))
