import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import _lib, synthetic, packing, gp
from csp_bo_b200.kernels import Dot_mb, device as kdev
from csp_bo_b200.utilities import tuple_to_list, list_to_tuple
from csp_bo_b200.packing import Pack, build_block
def T(f, n=7):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return round(min(ts), 3)
trainE = synthetic.energy_set(16, seed=1); trainF = synthetic.force_set(409, seed=2)
testE = synthetic.energy_set(1, seed=7); testF = synthetic.force_set(192, seed=8)
kern = Dot_mb(para=[18.55, 0.01], zeta=2)
pe2, pf2 = Pack(trainE[0], None, trainE[1], trainE[2]), Pack(*trainF)
print("list_to_tuple(192 force groups)", T(lambda: list_to_tuple(tuple_to_list(testF))))
print("Pack(test force 4.2k rows)     ", T(lambda: Pack(*testF)))
print("Pack(test energy 192 rows)     ", T(lambda: Pack(testE[0], None, testE[1], testE[2])))
pe1, pf1 = Pack(testE[0], None, testE[1], testE[2]), Pack(*testF)
h = kdev._hyper(kern, True)
K = torch.empty((1 + 576, 16 + 1227), dtype=torch.float64, device="cuda")
print("_into KFF 576x1227             ", T(lambda: kdev._into(h, _lib.KFF, 2.0, 1.0, pf1, pf2, K, 1, 16)))
print("_into KEF 1x1227               ", T(lambda: kdev._into(h, _lib.KEF, 2.0, 1.0, pe1, pf2, K, 0, 16)))
print("_into KEE 1x16                 ", T(lambda: kdev._into(h, _lib.KEE, 2.0, 1.0, pe1, pe2, K, 0, 0)))
tmp = torch.empty((16, 576), dtype=torch.float64, device="cuda")
print("_into KEF 16x576 (FE)          ", T(lambda: kdev._into(h, _lib.KEF, 2.0, 1.0, pe2, pf1, tmp, 0, 0)))
alpha = torch.randn(1243, dtype=torch.float64, device="cuda")
print("predict_mean + D2H             ", T(lambda: gp.predict_mean(K, alpha).cpu().numpy()))
print("torch small op (K[:1,:16] /= c)", T(lambda: K[:1, :16].div_(3.0)))
print("torch.as_tensor(counts).cuda   ", T(lambda: torch.as_tensor(pe1.counts, dtype=torch.float64, device="cuda")))
data = {"energy": testE, "force": testF}; train = {"energy": trainE, "force": trainF}
print("k_total_device(test, train)    ", T(lambda: kdev.k_total_device(kern, data, train)))
