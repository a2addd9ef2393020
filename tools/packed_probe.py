"""K_FF(X,X) at 20 001 force rows, one GPU: caller-order build (24-byte output pieces) against the packed-column
build (full rows staged through shared memory), CUDA-event timed.  `ncu` over this script gives the DRAM traffic of both."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib, packing, synthetic
from csp_bo_b200.distributed import ShardedSymmetricBlock

n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 6667
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
X = synthetic.force_set(n_atoms, seed=0)
p = packing.Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])
S2 = 18.555440586011372 ** 2 * 2.0
out = torch.empty((n_atoms, 3, n_atoms * 3), dtype=torch.float64, device="cuda")
res = {}

def timed(name, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    res[name] = float(np.median(ts))

timed("caller_order", lambda: packing.build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=S2, symmetric=True, out=out))
ref = out.reshape(3 * n_atoms, 3 * n_atoms)
for name, kw in (("sharded_w1_caller_cols", dict(zigzag=True)), ("sharded_w1_packed_sb1", dict(zigzag=True, packed_cols=True)),
                 ("sharded_w1_packed_sb16", dict(sb_blocks=16, zigzag=True, packed_cols=True))):
    sh = ShardedSymmetricBlock(p, _lib.KFF, **kw)
    timed(name, lambda: sh.build(_lib.DOT, 2.0, scale=S2, sync=False))
    if n_atoms <= 2000:
        K = sh.gather_full()
        res[name + "_err"] = float((K - ref).abs().max() / ref.abs().max())
    sh.close()
print(json.dumps(res))
