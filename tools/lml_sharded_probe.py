"""distributed.lml_sharded (LML + gradient with K distributed) at N = 20 001, Dot and RBF, with its phase times.
    python tools/lml_sharded_probe.py     |     torchrun --nproc-per-node W tools/lml_sharded_probe.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.distributed as dist
from csp_bo_b200 import _lib, synthetic
from csp_bo_b200.distributed import lml_sharded
from csp_bo_b200.kernels import Dot_mb, RBF_mb
from csp_bo_b200.packing import resident

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); _lib.check(_lib.load().cspbo_set_device(local))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nF = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
full = resident({"force": synthetic.force_set(nF, seed=0)})["force"]
y = np.random.default_rng(0).standard_normal(3 * nF)


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier(); torch.cuda.synchronize()


for kern in (Dot_mb(para=[18.555440586011372, 0.01], zeta=2), RBF_mb(para=[6.244955524643115, 0.5611029935713949], zeta=2)):
    ts = []
    for _ in range(3):
        barrier(); t0 = time.perf_counter()
        v = lml_sharded(kern, {"force": full}, y, 0.005, 30.0, eval_gradient=False)
        barrier(); t1 = time.perf_counter()
        lml, grad = lml_sharded(kern, {"force": full}, y, 0.005, 30.0, eval_gradient=True)
        barrier(); ts.append(((t1 - t0) * 1e3, (time.perf_counter() - t1) * 1e3))
    if rank == 0:
        print(json.dumps(dict(kernel=kern.name, world=world, N=3 * nF, lml_only_ms=round(min(t[0] for t in ts), 2),
                              lml_grad_ms=round(min(t[1] for t in ts), 2), lml=lml, grad=[float(g) for g in grad])), flush=True)
if world > 1:
    dist.destroy_process_group()
