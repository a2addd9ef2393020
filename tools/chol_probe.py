"""Timing of the sharded fit (packed-order K_FF build -> panel Cholesky with look-ahead -> solve) against the
single-call cuSOLVER fit; run directly (1 GPU) or under torchrun (N GPUs).
    python tools/chol_probe.py [n_atoms] [sb_blocks,...]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from csp_bo_b200 import _lib, gp, synthetic
from csp_bo_b200.distributed import ShardedSymmetricBlock, ShardedCholesky
from csp_bo_b200.packing import Pack, build_block

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); _lib.check(_lib.load().cspbo_set_device(local))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
sbs = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [8, 16, 32]
X = synthetic.force_set(n_atoms, seed=0)
p = Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
N = 3 * n_atoms
s2 = 18.555440586011372 ** 2 * 2.0
y = torch.from_numpy(np.random.default_rng(0).standard_normal(N)).cuda()
nf2 = (30.0 * 0.005) ** 2

def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()

ref = None
nobase = len(sys.argv) > 3 and sys.argv[3] == "nobase"
if not nobase:
    Kd = torch.empty((p.m, 3, p.m * 3), dtype=torch.float64, device="cuda")
    ts = []
    for _ in range(2):
        build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=s2, symmetric=True, out=Kd)
        sync(); t0 = time.perf_counter()
        ref = gp.fit_alpha(Kd.view(N, N), y, n_energy=0, noise_e=0.005, f_coef=30.0, overwrite=True)
        sync(); ts.append((time.perf_counter() - t0) * 1e3)
    if rank == 0:
        print(json.dumps(dict(what="cuSOLVER potrf+potrs (replicated)", N=N, ms=min(ts), tflops=N ** 3 / 3 / (min(ts) * 1e-3) / 1e12)), flush=True)
    build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=s2, symmetric=True, out=Kd)
    Ktest = Kd.view(N, N)[:64].clone()
    del Kd
    torch.cuda.empty_cache()

for G in sbs:
    sh = ShardedSymmetricBlock(p, _lib.KFF, sb_blocks=G, zigzag=True, packed_cols=True)
    rec = []
    for it in range(3):
        sync(); t0 = time.perf_counter()
        sh.build(_lib.DOT, 2.0, scale=s2, sync=True)
        sync(); t1 = time.perf_counter()
        chol = ShardedCholesky.from_block(sh)
        chol.add_noise(nf2)
        sync(); t2 = time.perf_counter()
        chol.factor()
        sync(); t3 = time.perf_counter()
        chol.check()
        order = torch.as_tensor(np.asarray(sh.order[: sh.m], dtype=np.int64), device="cuda")
        ap = chol.solve(y.reshape(sh.m, 3)[order].reshape(-1))
        alpha = torch.empty((sh.m, 3), dtype=torch.float64, device="cuda"); alpha[order] = ap.reshape(sh.m, 3)
        sync(); t4 = time.perf_counter()
        if os.environ.get("CSPBO_CHOL_TRACE") and it == 2:
            print(json.dumps(dict(rank=rank, G=G, trace={k: round(v, 2) for k, v in chol.trace_summary().items()})), flush=True)
            if rank <= 1:
                for row in chol.trace_timeline(range(8, 14)):
                    print("TL rank%d %s" % (rank, json.dumps(row)), flush=True)
        rec.append(dict(host_issue=chol.host_ms, build=(t1 - t0) * 1e3, alloc_noise=(t2 - t1) * 1e3, factor=(t3 - t2) * 1e3, solve=(t4 - t3) * 1e3, total=(t4 - t0) * 1e3))
        del chol
    err = float(((Ktest @ ref) - (Ktest @ alpha.reshape(-1))).abs().max() / (Ktest @ ref).abs().max()) if ref is not None else None
    if rank == 0:
        best = min(rec, key=lambda r: r["total"])
        print(json.dumps(dict(what="sharded fit", world=world, N=N, sb_blocks=G, panel_rows=24 * G, pred_err_vs_cusolver=err,
                              factor_tflops=N ** 3 / 3 / (best["factor"] * 1e-3) / 1e12, **{k: round(v, 2) for k, v in best.items()})), flush=True)
    sh.close()
if world > 1:
    dist.destroy_process_group()
