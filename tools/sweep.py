"""Per-kernel roofline sweep (BASELINE.json config 5: synthetic K_EE / K_EF / K_FF scaling sweep on
X1_big-shaped descriptors, 1k-40k force rows, Dot and RBF) plus the HBM-bound K*.alpha GEMV and the
full energy+force fit of a Si_5352-shaped training set.

    python tools/sweep.py [--quick] [--out gpurun_out/sweep.json]

Every line: the block, its size, the device time (CUDA events, mean of 3 after 2 warm-ups), the
algorithmic flops (F_EE = 2d+12 (+20 RBF), F_EF = 8d+30, F_FF = 32d+100 per same-species
row pair; unique pairs for symmetric builds) or bytes, and the fraction of the measured peak
(FP64 DMMA 37.03 TFLOP/s from profiles/fp64_peak_r01.json; HBM from MEASURED_PEAKS.json).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from csp_bo_b200 import _lib, gp, synthetic
from csp_bo_b200.packing import Pack, build_block

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
D = 24
F = {_lib.KEE: 2 * D + 12, _lib.KEF: 8 * D + 30, _lib.KFF: 32 * D + 100}
NAMES = {_lib.KEE: "K_EE", _lib.KEF: "K_EF", _lib.KFF: "K_FF"}
DOT = dict(sigma=18.555440586011372, sigma0=0.01, zeta=2.0)          # examples/models/test_2.json
RBF = dict(sigma=6.244955524643115, l=0.5611029935713949, zeta=2.0)  # examples/models/Si.json


def peaks():
    fp64 = 37.03
    try:
        fp64 = float(json.load(open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")))["dmma_peak_tflops"])
    except Exception:
        pass
    hbm = 6547.2
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    return fp64, hbm


def timed(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms)), float(np.min(ms))


def fpack(n_atoms, single):
    X = synthetic.force_set(n_atoms, seed=0, single_element=single)
    return Pack(torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda(), X[2], X[3])


def epack(n_struct, single, rows=192):
    X = synthetic.energy_set(n_struct, seed=1, single_element=single, rows_per_structure=rows)
    return Pack(torch.from_numpy(X[0]).cuda(), None, X[1], X[2])


def block_line(fam, block, a, b, sym, what, fp64):
    famname = "Dot" if fam == _lib.DOT else "RBF"
    kw = dict(sigma2=DOT["sigma"] ** 2, sigma02=DOT["sigma0"] ** 2) if fam == _lib.DOT else \
        dict(sigma2=RBF["sigma"] ** 2, l=RBF["l"], tol=1e-10, use_tol=(block == _lib.KFF))
    shp = {_lib.KEE: (a.m, b.m), _lib.KEF: (a.m, b.m, 3), _lib.KFF: (a.m, 3, b.m * 3)}[block]
    out = torch.empty(shp, dtype=torch.float64, device="cuda")
    mean, best = timed(lambda: build_block(fam, block, a, b, 2.0, symmetric=sym, out=out, **kw))
    pairs = a.pair_count(b)
    algo_pairs = (pairs + (a.n if sym else 0)) / 2 if sym else pairs
    flop_pp = F[block] + (20 if fam == _lib.RBF else 0)
    tf = algo_pairs * flop_pp / (mean * 1e-3) / 1e12
    r = dict(kernel="block_tile_kernel", block=NAMES[block], family=famname, what=what, symmetric=bool(sym),
             groups=[a.m, b.m], stacked_rows=[a.n, b.n], entries=int(np.prod(shp)), same_species_pairs=int(pairs),
             ms=mean, ms_best=best, entries_per_s=float(np.prod(shp)) / (mean * 1e-3), flop_per_pair=flop_pp,
             tflops=tf, bound="fp64 tensor", frac=tf / fp64,
             out_gbs=8.0 * float(np.prod(shp)) / (mean * 1e-3) / 1e9)
    del out
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.json"))
    args = ap.parse_args()
    fp64, hbm = peaks()
    res = []

    def emit(r):
        res.append(r)
        print(json.dumps(r), flush=True)

    sizes = [334, 1667] if args.quick else [334, 667, 1667, 3334, 6667, 13334]
    for single in (False, True):
        tag = "single species (Si-like)" if single else "Pt/H/O mix"
        for n_atoms in sizes:
            pf = fpack(n_atoms, single)
            n_struct = max(8, n_atoms // 8)          # energy side: 192-atom structures, ~as many rows as the force side
            pe = epack(n_struct, single)
            for fam in (_lib.DOT, _lib.RBF):
                emit(block_line(fam, _lib.KFF, pf, pf, True, "%d force rows, %s" % (3 * n_atoms, tag), fp64))
                emit(block_line(fam, _lib.KEF, pe, pf, False, "%d structures x %d force rows, %s" % (n_struct, 3 * n_atoms, tag), fp64))
                emit(block_line(fam, _lib.KEE, pe, pe, True, "%d structures of 192 atoms, %s" % (n_struct, tag), fp64))
            if n_atoms == 6667 and not single:
                emit(block_line(_lib.DOT, _lib.KFF, pf, pf, False, "non-symmetric K_FF(X1,X2) %d x %d force rows" % (3 * n_atoms, 3 * n_atoms), fp64))
            del pf, pe
            torch.cuda.empty_cache()

    # Si_5352-shaped energy set: 5352 structures x 64 atoms, one species
    if not args.quick:
        pe = epack(5352, True, rows=64)
        for fam in (_lib.DOT, _lib.RBF):
            emit(block_line(fam, _lib.KEE, pe, pe, True, "Si_5352-shaped: 5352 structures of 64 atoms", fp64))
        pf = fpack(3334, True)
        for fam in (_lib.DOT, _lib.RBF):
            emit(block_line(fam, _lib.KEF, pe, pf, False, "Si_5352-shaped: 5352 structures x 10002 force rows", fp64))
        del pe, pf

    # K*.alpha GEMV: HBM bound, 8 bytes per matrix element read once
    shapes = [(577, 1243), (577, 20001), (5770, 20001), (20001, 20001)] + ([] if args.quick else [(40002, 40002)])
    for n, N in shapes:
        Ks = torch.randn((n, N), dtype=torch.float64, device="cuda")
        al = torch.randn(N, dtype=torch.float64, device="cuda")
        ref = (Ks @ al)
        got = gp.predict_mean(Ks, al)
        err = float((got - ref).abs().max() / ref.abs().max())
        mean, best = timed(lambda: gp.predict_mean(Ks, al), reps=5)
        gbs = 8.0 * n * N / (mean * 1e-3) / 1e9
        emit(dict(kernel="predict_gemv_kernel", what="K*[%d,%d] . alpha" % (n, N), ms=mean, ms_best=best, bytes=8 * n * N,
                  gbs=gbs, gbs_best=8.0 * n * N / (best * 1e-3) / 1e9, bound="hbm", frac=gbs / hbm,
                  fits_l2=bool(8 * n * N < 100e6), max_norm_err_vs_torch=err))
        del Ks, al, ref, got
        torch.cuda.empty_cache()

    # full energy + force fit (gaussianprocess.py:105-127, opt=False) on a Si_5352-shaped set
    if not args.quick:
        from csp_bo_b200.gaussianprocess import GaussianProcess
        from csp_bo_b200.kernels import RBF_mb
        from csp_bo_b200.utilities import tuple_to_list
        rng = np.random.default_rng(3)
        E = synthetic.energy_set(5352, seed=1, single_element=True, rows_per_structure=64)
        Fs = synthetic.force_set(3334, seed=2, single_element=True)
        data = {"energy": [(x, float(rng.normal()), ele) for x, ele in tuple_to_list(E, mode='energy')],
                "force": [(x, dx, rng.normal(size=3), ele) for x, dx, ele in tuple_to_list(Fs)]}
        model = GaussianProcess(RBF_mb(para=[RBF["sigma"], RBF["l"]], zeta=2), f_coef=30, noise_e=[0.005, 0.002, 0.1])
        model.set_train_pts(data)
        ts = []
        for _ in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            model.fit(show=False, opt=False)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        emit(dict(kernel="fit", what="GaussianProcess.fit(opt=False), RBF_mb, 5352 energies (64-atom structures) + 3334 force atoms: N = %d" % (5352 + 10002),
                  ms=float(np.min(ts[1:])), ms_all=ts))

    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(dict(fp64_peak_tflops=fp64, hbm_gbs=hbm, lines=res), open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
