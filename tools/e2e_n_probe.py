"""e2e of the sharded host-bound build at N ranks: variants of the slab schedule.  torchrun --nproc-per-node N tools/e2e_n_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if os.environ.get("HP", "1") == "1":
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
from csp_bo_b200 import _lib, packing, synthetic
from csp_bo_b200.distributed import ShardedSymmetricBlock
_lib.check(_lib.load().cspbo_set_device(local))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X = synthetic.force_set(6667, seed=0)
full = packing.Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
sh = ShardedSymmetricBlock(full, _lib.KFF, zigzag=True, packed_cols=True)
outh = torch.empty(tuple(sh.local.shape), dtype=torch.float64).pin_memory()
S2 = 18.555440586011372 ** 2 * 2.0
def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
res = {}
for name, fr in (("default", None), ("16 equal", tuple(i / 16 for i in range(17))), ("4", (0, 0.05, 0.3, 0.6, 1.0)),
                 ("front-loaded", (0, 0.02, 0.06, 0.12, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0))):
    ts = []
    for i in range(4):
        barrier(); t0 = time.perf_counter()
        sh.build_to_host(_lib.DOT, 2.0, outh, scale=S2, fractions=fr)
        barrier(); ts.append((time.perf_counter() - t0) * 1e3)
    res[name] = round(min(ts[1:]), 2)
# references: build only, D2H only
ts = []
for i in range(3):
    barrier(); t0 = time.perf_counter(); sh.build(_lib.DOT, 2.0, scale=S2, sync=False); barrier(); ts.append((time.perf_counter() - t0) * 1e3)
res["build_only"] = round(min(ts[1:]), 2)
ts = []
for i in range(3):
    barrier(); t0 = time.perf_counter(); outh.copy_(sh.local); barrier(); ts.append((time.perf_counter() - t0) * 1e3)
res["d2h_only"] = round(min(ts[1:]), 2)
if rank == 0:
    print(json.dumps({"world": world, "hp": os.environ.get("HP", "1"), **res}))
if world > 1:
    dist.destroy_process_group()
