"""Host-side D2H ceiling with N ranks copying at once, with and without pinning each rank (and the first touch of its
pinned buffer) to the CPUs NVML reports as local to its GPU.   torchrun --nproc-per-node N tools/d2h_probe.py"""
import os, sys, time, json
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB = int(sys.argv[1]) if len(sys.argv) > 1 else 400
src = torch.empty(MB << 17, dtype=torch.float64, device="cuda").normal_()
res = {}


def run(tag):
    host = torch.empty(MB << 17, dtype=torch.float64).pin_memory()
    host.zero_()
    ts = []
    for i in range(4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        host.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    t = torch.tensor([min(ts[1:])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res[tag] = {"ms_max_over_ranks": float(t) * 1e3, "aggregate_GBs": world * MB * 1.048576e-3 / float(t)}
    del host


run("default")
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    res["affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]
    run("nvml_affinity")
except Exception as exc:
    res["affinity_error"] = str(exc)
all_res = [None] * world
if world > 1:
    dist.all_gather_object(all_res, res)
else:
    all_res = [res]
if rank == 0:
    print(json.dumps({"world": world, "MB_per_rank": MB, "rank0": all_res[0], "affinities": [r.get("affinity") for r in all_res]}))
if world > 1:
    dist.destroy_process_group()
