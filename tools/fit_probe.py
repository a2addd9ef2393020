"""Energy + force GP fit of a Si_5352-shaped training set (BASELINE.json config 3: 5352 structures of 64 atoms +
3334 force atoms, RBF_mb of examples/models/Si.json, N = 15 354) through fit_sharded on 1..8 GPUs.
    python tools/fit_probe.py            |   torchrun --nproc-per-node W tools/fit_probe.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from csp_bo_b200 import _lib, gp, synthetic
from csp_bo_b200.distributed import ShardedCovariance, ShardedCholesky, ShardedSolver
from csp_bo_b200.kernels import RBF_mb
from csp_bo_b200.kernels import device as kdev

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); _lib.check(_lib.load().cspbo_set_device(local))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nE, nF = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (5352, 3334)
E = synthetic.energy_set(nE, seed=1, single_element=True, rows_per_structure=64)
F = synthetic.force_set(nF, seed=2, single_element=True)
data = {"energy": E, "force": F}
y = np.random.default_rng(3).standard_normal(nE + 3 * nF)
kernel = RBF_mb(para=[6.244955524643115, 0.5611029935713949], zeta=2)
noise_e, f_coef = 0.005, 30.0

def sync():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()

ref = None
if world == 1:
    ts = []
    for _ in range(3):
        sync(); t0 = time.perf_counter()
        K = kdev.k_total_device(kernel, data)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        ref = gp.fit_alpha(K, torch.from_numpy(y).cuda(), n_energy=nE, noise_e=noise_e, f_coef=f_coef, overwrite=True)
        sync(); ts.append(((time.perf_counter() - t0) * 1e3, (t1 - t0) * 1e3))
        del K
    print(json.dumps(dict(what="dense single-GPU fit (k_total_device + cuSOLVER)", N=nE + 3 * nF, ms=min(ts)[0], build_ms=min(ts)[1])), flush=True)
rec = []
for it in range(3):
    sync(); t0 = time.perf_counter()
    cov = ShardedCovariance(kernel, data, nbw=384)
    sync(); t1 = time.perf_counter()
    cov.build(noise_e, f_coef)
    sync(); t2 = time.perf_counter()
    chol = ShardedCholesky(cov.local, cov.N, 384, cov.rank, cov.world, True, None)
    chol.factor(); chol.check()
    sync(); t3 = time.perf_counter()
    alpha = ShardedSolver(chol, cov.permutation()).solve(torch.from_numpy(y).cuda())
    sync(); t4 = time.perf_counter()
    rec.append(dict(pack_alloc=(t1 - t0) * 1e3, build=(t2 - t1) * 1e3, factor=(t3 - t2) * 1e3, solve=(t4 - t3) * 1e3, total=(t4 - t0) * 1e3))
    chol.local = None; cov.close(); del chol, cov
best = min(rec, key=lambda r: r["total"])
if rank == 0:
    err = float((alpha - ref).abs().max() / ref.abs().max()) if ref is not None else None
    print(json.dumps(dict(what="fit_sharded", world=world, N=nE + 3 * nF, alpha_vs_dense=err, **{k: round(v, 2) for k, v in best.items()})), flush=True)
if world > 1:
    dist.destroy_process_group()
