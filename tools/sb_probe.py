"""Sharded K_FF build time against the superblock size (blocks per ownership unit).  torchrun --nproc-per-node N tools/sb_probe.py"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
from csp_bo_b200 import _lib, packing, synthetic
from csp_bo_b200.distributed import ShardedSymmetricBlock
_lib.check(_lib.load().cspbo_set_device(local))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
X = synthetic.force_set(6667, seed=0)
full = packing.Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
S2 = 18.555440586011372 ** 2 * 2.0
res = {}
for sb, bal in ((1, False), (4, False), (1, True), (2, True), (4, True), (8, True)):
    sh = ShardedSymmetricBlock(full, _lib.KFF, sb_blocks=sb, zigzag=True, packed_cols=True, balance=bal)
    for _ in range(3):
        sh.build(_lib.DOT, 2.0, scale=S2, sync=False)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        sh.build(_lib.DOT, 2.0, scale=S2, sync=False)
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / 10], dtype=torch.float64, device="cuda")
    tmin = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    res["%d%s" % (sb, "b" if bal else "")] = [round(float(t), 3), round(float(tmin), 3)]
    sh.close()
if rank == 0:
    print(json.dumps({"world": world, "ms_max_min_over_ranks_by_sb_blocks": res}))
if world > 1:
    dist.destroy_process_group()
