#!/usr/bin/env python
"""bench.py -- headline benchmark of the GP kernel-matrix hot path (BASELINE.json metric:
"K_FF entries/s (fp64) + GP fit wall-time at 1/2/4/8 B200 vs host CPU").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one K_FF(X, X) build for a synthetic training set shaped like the reference's
Cpp_code/X1_big.npy (csp_bo_b200/synthetic.py): 6667 force atoms = 20 001 force rows, 146 976
stacked descriptor rows, d = 24, Pt/H/O species mix, Dot kernel sigma=18.555 zeta=2
(examples/models/test_2.json).  `value` = entries of the [3m,3m] K_FF matrix delivered per
second with the packed descriptors already resident in HBM; `e2e` = the same through the public
kff_C() call with pinned HOST buffers in and the K_FF matrix back in pinned HOST memory.

N > 1 (torchrun): the symmetric build is sharded over ranks by 8-atom row blocks dealt in zigzag order, rows and
columns in the packed order of the descriptor set (P K P^T -- what the sharded Cholesky factorises; full-row stores
locally and over NVLink); strong scaling, fixed total work, no data-path collective; the max over ranks of the
device-timed region is reported.  Every N > 1 line carries `parity` (rank 0's first row blocks against the reference
C++ and the residual of the sharded fit).  The N = 1 line carries `legs`: driver-timed builds of the other configs
(RBF, single species, K_EE / K_EF, d = 50, the LML gradient) each with its own roofline; every line carries `fit_ef`,
the energy + force fit of a Si_5352-shaped set on the N GPUs.

--impl reference times the reference's own CPU implementation (oracle/_ref = the unmodified
cspbo/kernels/dot_kernel.cpp compiled with g++ -O2, threaded like its MPI split) on a bounded
slab of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ATOMS = 6667
SIGMA, SIGMA0, ZETA = 18.555440586011372, 0.01, 2.0
FLOP_PER_PAIR = 32 * 24 + 100     # F_FF(d) = 2*16*d + 100 at d = 24
FIT_SB_BLOCKS = 16                # panel width of the sharded Cholesky fit leg: 16 blocks x 8 atoms x 3 = 384 rows
FP64_PEAK_FALLBACK = 37.03        # TFLOP/s, DMMA.8x8x4 issue-bound peak measured on this pool (profiles/fp64_peak_r01.json)


_JSON_FD = None


def quiet_stdout():
    """Everything except the one JSON line goes to stderr: NCCL / library banners and the reference-style prints of
    the regressor would otherwise share stdout with it."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)
        sys.stdout = sys.stderr


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def fp64_peak():
    """MEASURED_PEAKS.json carries no FP64 entry; the denominator is the DMMA peak measured on this
    pool's B200 with tools/fp64_peak.cu and committed under profiles/."""
    for name in sorted(os.listdir(os.path.join(ROOT, "profiles")), reverse=True):
        if name.startswith("fp64_peak_r") and name.endswith(".json"):
            try:
                d = json.load(open(os.path.join(ROOT, "profiles", name)))
                return float(d["dmma_peak_tflops"]), "profiles/" + name + " (tools/fp64_peak.cu, DMMA.8x8x4)"
            except Exception:
                pass
    return FP64_PEAK_FALLBACK, "fallback constant"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload(single_element=False, n_atoms=N_ATOMS):
    from csp_bo_b200 import synthetic
    X = synthetic.force_set(n_atoms, seed=0, single_element=single_element)
    pairs = synthetic.same_species_pairs(X[2], X[2])
    return X, pairs


# ------------------------------------------------------------------------------------------
def cpu_reference_rate(X, n_x1_atoms, threads, backend="auto"):
    """K_FF slab (first n_x1_atoms x1 atoms against all x2) with the reference CPU code; returns
    (entries/s, seconds, kind)."""
    from oracle import cpu_reference as cr
    be = cr.get_backend(backend)
    x, dx, ele, counts = X
    rows = int(np.sum(counts[:n_x1_atoms]))
    X1 = (x[:rows], dx[:rows], ele[:rows], counts[:n_x1_atoms])
    t0 = time.perf_counter()
    C = cr.dot_kff_C(X1, X, SIGMA, SIGMA0, ZETA, backend=be.kind, threads=threads)
    dt = time.perf_counter() - t0
    return C.size / dt, dt, ("reference" if be.kind == "ref" else "port"), C


def native_cpu_rate(X, n_x1_atoms, threads):
    """the CPU number again with -O3 -march=native (the C restatement, bit-identical to the reference on
    the goldens, compiled on this host); None when the compiler is unavailable."""
    try:
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port_native"], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL)
        rate, dt, _, _ = cpu_reference_rate(X, n_x1_atoms, threads, backend="port_native")
        return {"value": rate, "unit": "entries/s", "kind": "port", "flags": "-O3 -march=native", "seconds": dt}
    except Exception:
        return None


def run_reference(args):
    """--impl reference: the reference's CPU path on the host cores, bounded slab per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    X, pairs = workload()
    cores = os.cpu_count() or 1
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=False, stdout=subprocess.DEVNULL)
    n_x1 = max(1, cores // 2)     # ~0.25 s of one core per x1 atom => a few seconds per step
    for _ in range(args.warmup):
        cpu_reference_rate(X, 1, cores)
    times, entries = [], 0
    for _ in range(args.steps):
        rate, dt, kind, C = cpu_reference_rate(X, n_x1, cores)
        times.append(dt); entries = C.size
    ms = 1e3 * float(np.mean(times))
    value = entries / (ms * 1e-3)
    sample = "K_FF slab: first %d of %d x1 atoms x all %d x2 atoms per step (%d entries)" % (n_x1, len(X[3]), len(X[3]), entries)
    line = {"impl": "reference", "metric": "K_FF entries/s (fp64)", "value": value, "unit": "entries/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(X, pairs), "gpu_launches": 0,
            "cpu_baseline": {"value": value, "unit": "entries/s", "cores": cores, "kind": kind, "sample": sample,
                             "flags": "-O2 (what the reference's setup.py builds)", "o3_native": native_cpu_rate(X, n_x1, cores)},
            "e2e": {"value": value, "unit": "entries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def config_dict(X, pairs):
    return {"workload": "K_FF(X,X) Dot zeta=2 d=24, synthetic X1_big-shaped Pt/H/O training set, "
                        "%d force atoms = %d force rows, %d stacked rows" % (len(X[3]), 3 * len(X[3]), len(X[0])),
            "force_rows": 3 * len(X[3]), "stacked_rows": int(len(X[0])), "d": 24, "kernel": "Dot_mb",
            "sigma": SIGMA, "zeta": ZETA, "same_species_row_pairs": int(pairs),
            "l2": "no explicit flush: each step streams a 3.2 GB output + 113 MB of packed rows, >> 126 MB L2"}


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from csp_bo_b200 import _lib, packing
    from csp_bo_b200.kernels.dot_kernel import kff_C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; csp_bo_b200 has no CPU path")
    torch.cuda.set_device(local)
    lib = _lib.load()
    _lib.check(lib.cspbo_set_device(local))
    if world > 1:
        # NCCL kernels on high-priority streams: the 1-element "slab k is finished everywhere" all-reduces of the
        # host-bound build must not queue behind the pending CTAs of the next slab's tile kernel
        os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    X, pairs = workload()
    m = len(X[3])
    xd, dxd = torch.from_numpy(X[0]).cuda(), torch.from_numpy(X[1]).cuda()
    full = packing.Pack(xd, dxd, X[2], X[3])
    sh = None
    if world == 1:
        out = torch.empty((m, 3, m * 3), dtype=torch.float64, device="cuda")
        algo_pairs = (pairs + len(X[0])) / 2      # unique same-species row pairs of the symmetric build

        def step():
            packing.build_block(_lib.DOT, _lib.KFF, full, full, ZETA, scale=SIGMA * SIGMA * ZETA, symmetric=True, out=out)
    else:
        # strong scaling: the symmetric build is sharded by block-cyclic row blocks; mirrored tiles go
        # to the owner's HBM over NVLink peer memory inside the tile kernel (csp_bo_b200/distributed.py)
        from csp_bo_b200.distributed import ShardedSymmetricBlock
        # zigzag dealing of 4-block superblocks: every rank gets the same triangular work and the four A blocks of a CTA are
        # neighbours in packed order (with single blocks dealt cyclically they are ~8 N blocks apart, and the CTAs that straddle
        # the diagonal idle half their warps: measured 14.01 -> 13.71 ms at N = 8, tools/sb_probe.py); packed columns:
        # full-row stores (local + NVLink)
        # balance=True: inside every cycle of N superblocks the heaviest goes to the rank with the least work so far (tile pairs)
        sh = ShardedSymmetricBlock(full, _lib.KFF, sb_blocks=4, zigzag=True, packed_cols=True, balance=True)
        algo_pairs = (pairs + len(X[0])) / 2 / world

        def step():
            sh.build(_lib.DOT, ZETA, scale=SIGMA * SIGMA * ZETA, sync=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_all0.record()
    for e0, e1 in evs:
        e0.record(); step(); e1.record()
    t_all1.record()
    barrier()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    tl = torch.tensor([float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tl, op=dist.ReduceOp.SUM)
    launches = int(tl.item())
    total_ms = t_all0.elapsed_time(t_all1)
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    entries = 9.0 * m * m
    value = entries / (ms_per_step * 1e-3)

    # ---- end-to-end through the public API: pinned host descriptors in, K_FF rows in pinned host memory out
    xh = torch.from_numpy(X[0]).pin_memory(); dxh = torch.from_numpy(X[1]).pin_memory()
    Xh = (xh.numpy(), dxh.numpy(), X[2], X[3])
    n_e2e = min(args.steps, 3)
    if world == 1:
        outh = torch.empty(m * 3 * m * 3, dtype=torch.float64).pin_memory()
        d2h = int(outh.numel() * 8)

        def e2e_step():
            K = kff_C(Xh, Xh, SIGMA, SIGMA0, ZETA, out=outh.numpy())   # H2D + pack + build + D2H inside
            return float(K[0, 0])
    else:
        outh = torch.empty(tuple(sh.local.shape), dtype=torch.float64).pin_memory()
        d2h = int(outh.numel() * 8) * world
        xd_e, dxd_e = torch.empty_like(xd), torch.empty_like(dxd)
        e2e_parts = []

        def e2e_step():
            # descriptors: ONE upload (rank 0, pinned host -> HBM) + NCCL broadcast over NVLink, then every rank packs
            # from HBM; the rows are built in slabs whose copy-back to pinned host memory overlaps the next slab
            t0 = time.perf_counter()
            if rank == 0:
                xd_e.copy_(xh, non_blocking=True); dxd_e.copy_(dxh, non_blocking=True)
            dist.broadcast(xd_e, src=0); dist.broadcast(dxd_e, src=0)
            pk = packing.Pack(xd_e, dxd_e, Xh[2], Xh[3])
            torch.cuda.synchronize(); t1 = time.perf_counter()
            sh.pack = pk
            sh.build_to_host(_lib.DOT, ZETA, outh, scale=SIGMA * SIGMA * ZETA)
            torch.cuda.synchronize(); t2 = time.perf_counter()
            sh.pack = full
            e2e_parts.append([round((b - a) * 1e3, 2) for a, b in ((t0, t1), (t1, t2))])
            return float(outh.view(-1)[0])
    e2e_ms, chk = [], 0.0
    for i in range(1 + n_e2e):
        barrier(); t0 = time.perf_counter()
        chk = e2e_step()
        barrier(); dt = time.perf_counter() - t0
        if i > 0:
            e2e_ms.append(dt * 1e3)
    te = torch.tensor([float(np.mean(e2e_ms))], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": entries / (float(te.item()) * 1e-3), "unit": "entries/s", "ms_per_step": float(te.item()),
           "h2d_bytes_per_step": int(X[0].nbytes + X[1].nbytes) + 2 * 4 * len(X[0]) * world,
           "d2h_bytes_per_step": d2h, "check": chk, "ms_all": [round(v, 2) for v in e2e_ms],
           "rank0_upload_bcast_pack__build_with_overlapped_d2h_ms": (e2e_parts[-1] if world > 1 else None),
           "host_d2h_ceiling": "measured on this pool (tools/d2h_probe.py, profiles/d2h_probe_r02.json): 91-94 GB/s aggregate with 8 "
                               "GPUs copying at once, 71 GB/s with 2, ~55 GB/s for one (a 32-vCPU single-NUMA-node VM): the 3.2 GB "
                               "result cannot reach host memory in less than ~34 ms however many GPUs build it",
           "api": "kernels.dot_kernel.kff_C(host tuple, host tuple, out=pinned)" if world == 1 else
                  "rank-0 upload + NCCL broadcast of the descriptors, packing.Pack(device), "
                  "distributed.ShardedSymmetricBlock.build_to_host (row slabs, copy-back overlapped) -> this rank's rows of "
                  "P K P^T in pinned host memory"}
    del outh

    # ---- GP fit wall time: K_FF(X,X) + noise -> Cholesky -> alpha  (gaussianprocess.py:105-127, opt=False)
    # N = 1: dense build + one cuSOLVER potrf/potrs.  N > 1: the build is sharded by row panels in packed
    # order and factorised in place by the sharded Cholesky (panel potrf/trsm on the owner, NCCL broadcast of the
    # panel rows, trailing dgemm updates on every rank's own rows; csp_bo_b200/distributed.py)
    fit = None
    try:
        from csp_bo_b200 import gp
        y = torch.from_numpy(np.random.default_rng(0).standard_normal(3 * m)).cuda()
        fit_ms = []
        nf2 = (30.0 * 0.005) ** 2
        if world == 1:
            for i in range(4):
                barrier(); t0 = time.perf_counter()
                step(); Kfull = out.view(3 * m, 3 * m)
                torch.cuda.synchronize(); t1 = time.perf_counter()
                alpha = gp.fit_alpha(Kfull, y, n_energy=0, noise_e=0.005, f_coef=30.0, overwrite=True)
                torch.cuda.synchronize(); t2 = time.perf_counter()
                fit_ms.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
                del Kfull
            how = "K_FF + noise -> cuSOLVER potrf/potrs -> alpha"
        else:
            from csp_bo_b200.distributed import ShardedSymmetricBlock, fit_alpha_sharded
            shf = ShardedSymmetricBlock(full, _lib.KFF, sb_blocks=FIT_SB_BLOCKS, zigzag=True, packed_cols=True, balance=True)
            for i in range(3):
                barrier(); t0 = time.perf_counter()
                shf.build(_lib.DOT, ZETA, scale=SIGMA * SIGMA * ZETA, sync=True)
                torch.cuda.synchronize(); t1 = time.perf_counter()
                alpha = fit_alpha_sharded(shf, y, nf2)
                barrier(); t2 = time.perf_counter()
                fit_ms.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
            how = ("K_FF sharded by %d-row panels (packed order, zigzag owners) -> sharded Cholesky (diagonal chain: potrf + "
                   "inverse + narrow piece on its own SMs and communicator; wide trmm, NCCL panel broadcast, look-ahead and "
                   "per-rank dgemm updates behind it) -> potrs on the replicated factor" % (24 * FIT_SB_BLOCKS))
        best = min(fit_ms[1:], key=lambda v: v[0])
        tl = torch.tensor(list(best), dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tl, op=dist.ReduceOp.MAX)
        # residual of the fit through an independent path: K rebuilt, (K + noise) alpha - y with the streaming GEMV kernel
        # on every rank's own rows, squared norms summed over the ranks
        if world == 1:
            step(); torch.cuda.synchronize()
            r = gp.predict_mean(out.view(3 * m, 3 * m), alpha) + nf2 * alpha - y
            rr = torch.stack(((r * r).sum(), (y * y).sum()))
        else:
            shf.build(_lib.DOT, ZETA, scale=SIGMA * SIGMA * ZETA, sync=True)
            order = torch.as_tensor(np.asarray(shf.order[:m], dtype=np.int64), device="cuda")
            ap, yp = alpha.reshape(m, 3)[order].reshape(-1), y.reshape(m, 3)[order].reshape(-1)      # packed order
            pos = torch.as_tensor(shf.positions(rank), device="cuda")
            keep = (pos >= 0) & (pos < m)
            rows3 = (pos.clamp(min=0)[:, None] * 3 + torch.arange(3, device="cuda")[None, :]).reshape(-1)
            keep3 = keep[:, None].expand(-1, 3).reshape(-1)
            Ka = gp.predict_mean(shf.local2d().contiguous(), ap)
            r = (Ka + nf2 * ap[rows3.clamp(max=3 * m - 1)] - yp[rows3.clamp(max=3 * m - 1)])[keep3]
            rr = torch.stack(((r * r).sum(), (yp[rows3.clamp(max=3 * m - 1)][keep3] ** 2).sum()))
            dist.all_reduce(rr, op=dist.ReduceOp.SUM)
            shf.close()
        fit = {"wall_ms": float(tl[0]), "k_build_ms": float(tl[1]), "cholesky_solve_ms": float(tl[2]), "N": 3 * m, "what": how,
               "alpha_check": float(alpha.abs().sum()),
               "residual": float(torch.sqrt(rr[0] / rr[1])),
               "residual_what": "||(K + noise) alpha - y|| / ||y|| with K rebuilt and applied by the GEMV kernel on every rank's own rows"}
    except Exception as exc:
        fit = {"error": str(exc)[:300]}

    # ---- N > 1: parity of what the ranks built against the reference C++ (rank 0's first two row blocks = 16 atoms)
    parity = None
    if world > 1:
        try:
            sh.build(_lib.DOT, ZETA, scale=SIGMA * SIGMA * ZETA, sync=True)
            if rank == 0:
                from oracle import cpu_reference as cr
                subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=False, stdout=subprocess.DEVNULL)
                pos = sh.positions(rank)[:16]
                groups = sh.order[pos]
                off = np.concatenate(([0], np.cumsum(X[3])))
                rows = np.concatenate([np.arange(off[g], off[g + 1]) for g in groups])
                X1 = (X[0][rows], X[1][rows], X[2][rows], [int(X[3][g]) for g in groups])
                be = cr.get_backend("auto")
                C = cr.dot_kff_C(X1, X, SIGMA, SIGMA0, ZETA, backend=be.kind, threads=os.cpu_count() or 1)
                loc = sh.local[:16].cpu().numpy().reshape(16, 3, m, 3)
                caller = np.empty_like(loc)
                caller[:, :, sh.order[:m], :] = loc          # packed column position p holds group order[p]
                err = float(np.abs(caller.reshape(48, 3 * m) - C).max() / np.abs(C).max())
                parity = {"max_norm_err_vs_reference": err, "tolerance": 1e-10, "kind": "reference" if be.kind == "ref" else "port",
                          "what": "rank 0's first 16 atoms (48 rows x %d columns of P K P^T, un-permuted) against the reference "
                                  "C++ kff_many on the same descriptors" % (3 * m)}
        except Exception as exc:
            parity = {"error": str(exc)[:300]}

    # ---- the energy + force fit of a Si_5352-shaped training set (BASELINE.json config 3) on the N GPUs
    fit_ef = None
    try:
        fit_ef = fit_ef_leg(world, rank, barrier)
    except Exception as exc:
        fit_ef = {"error": str(exc)[:300]}

    # ---- N = 1: the other configs of BASELINE.json, each a short driver-timed leg with its own roofline
    legs = None
    if world == 1 and not args.no_legs:
        try:
            legs = extra_legs(full, X, fp64_peak()[0])
        except Exception as exc:
            legs = {"error": str(exc)[:300]}
    elif world > 1 and not args.no_legs:
        # N > 1: one LML + gradient evaluation per kernel family over the ranks (what fit(opt=True) evaluates per L-BFGS step)
        try:
            legs = lml_sharded_legs(full, m, world, barrier)
        except Exception as exc:
            legs = {"error": str(exc)[:300]}

    # ---- one MD / validation step (config 1 of BASELINE.json): SO3 descriptor of a 192-atom Pt/H/O structure
    # + K* against a 16-energy + 409-force model + K*.alpha, through GaussianProcess.predict_structure.
    # The reference's README quotes 35.767 s for 10 such structures (README.md:28).
    import contextlib
    md = None
    if world == 1:
        try:
            with contextlib.redirect_stdout(sys.stderr):     # stdout carries the one JSON line only
                md = md_step_leg()
        except Exception as exc:
            md = {"error": str(exc)[:200]}

    # ---- real data (BASELINE.json configs 2 and 4): the reference's shipped Si model on Si_244_DFT, and a
    # Cu-EMT-300K model driven like an MD loop (descriptor + K* + alpha per step)
    real = None
    if world == 1:
        try:
            with contextlib.redirect_stdout(sys.stderr):
                real = real_data_leg()
        except Exception as exc:
            real = {"error": str(exc)[:200]}

    if sh is not None:
        sh.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = fp64_peak()
    kern_ms = float(np.mean(step_ms))
    achieved = algo_pairs * FLOP_PER_PAIR / (kern_ms * 1e-3) / 1e12
    # DRAM bytes per launch are an ncu figure (profiles/kff_traffic_r02.md, same kernel and workload), not measured in
    # this run: caller-order output (N = 1) 16.25 GB, packed-column output (N > 1, the whole matrix) 3.60 GB
    traffic = 16.48e9 if world == 1 else 3.60e9 / world
    tensor_flops = algo_pairs * 32 * 24          # the 16 length-d dot products only: what the DMMA pipe executes
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": "ncu dram__bytes_read+write per launch: profiles/kff_r02_final_ncu.csv (N = 1, caller-order output: 24-byte pieces, inherent to the layout kff_C must return), profiles/kff_traffic_r02.md (N > 1, packed columns); not measured in this run",
                "algorithmic_bytes_per_step": (8.0 * 9 * m * m + 113e6) / world,
                "peak_source": "builder-measured DMMA peak: " + peak_src,
                "algorithmic_flops_per_step": algo_pairs * FLOP_PER_PAIR,
                "achieved_tensor_flops_only": tensor_flops / (kern_ms * 1e-3) / 1e12,
                "frac_tensor_flops_only": tensor_flops / (kern_ms * 1e-3) / 1e12 / peak,
                "note": "FP64 tensor (DMMA) pipe; `frac` counts SURVEY 8d's 868 flop per same-species row pair (768 of "
                        "dot products + 100 charged for the epilogue, of which the zeta = 2 chain rule spends 36); "
                        "`frac_tensor_flops_only` counts the 768 the DMMA pipe executes; unique pairs of the symmetric "
                        "build; device time of the block kernels per step"}

    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "port"], check=False, stdout=subprocess.DEVNULL)
            cores = os.cpu_count() or 1
            n_x1 = max(2, 3 * cores)   # ~0.25 core-s per x1 atom => ~10-20 s on all cores
            rate, dt, kind, C = cpu_reference_rate(X, n_x1, cores)
            # parity of the slab actually computed (max-normalised)
            step(); torch.cuda.synchronize()
            Kd = out.reshape(m * 3, m * 3)[: C.shape[0]].cpu().numpy()
            err = float(np.abs(Kd - C).max() / np.abs(C).max())
            cpu = {"value": rate, "unit": "entries/s", "cores": cores, "kind": kind, "seconds": dt,
                   "sample": "K_FF slab: first %d of %d x1 atoms x all x2 atoms (%d entries), reference C++ "
                             "threaded over x2 like its MPI split" % (n_x1, m, C.size),
                   "max_norm_err_vs_gpu": err, "flags": "-O2 (what the reference's setup.py builds)",
                   "o3_native": native_cpu_rate(X, max(2, n_x1 // 2), cores)}
        except Exception as exc:  # the baseline is reported, never required for the GPU number
            cpu = {"value": None, "unit": "entries/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(exc)}

    line = {"metric": "K_FF entries/s (fp64)", "value": value, "unit": "entries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(X, pairs), "clocks": clocks, "gpu_launches": int(launches),
            "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "fit": fit, "fit_ef": fit_ef, "legs": legs,
            "md_step": md, "real_data": real}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def _events_ms(fn, reps=3, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.mean(ms))


def extra_legs(full, X, peak):
    """BASELINE.json config 5 beyond the headline (RBF, single species, K_EE / K_EF, Dot and RBF), d = 50 (AgI.json) and
    d = 90 (the chunked kernels), and one LML + gradient evaluation per kernel family.  Flop counts: SURVEY 8(d)."""
    import torch
    from csp_bo_b200 import _lib, packing, synthetic
    from csp_bo_b200.gaussianprocess import GaussianProcess
    from csp_bo_b200.kernels import Dot_mb, RBF_mb
    hbm = 6547.2
    try:
        hbm = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    RBF_P = dict(sigma=6.244955524643115, l=0.5611029935713949)           # examples/models/Si.json
    NAMES = {_lib.KEE: "K_EE", _lib.KEF: "K_EF", _lib.KFF: "K_FF"}
    out = {}

    def dev_pack(T):
        if len(T) == 4:
            return packing.Pack(torch.from_numpy(T[0]).cuda(), torch.from_numpy(T[1]).cuda(), T[2], T[3])
        return packing.Pack(torch.from_numpy(T[0]).cuda(), None, T[1], T[2])

    def leg(name, fam, block, a, b, sym, what):
        d = a.d
        kw = dict(sigma2=SIGMA ** 2, sigma02=SIGMA0 ** 2, scale=(SIGMA ** 2 * ZETA if block == _lib.KFF else SIGMA ** 2)) if fam == _lib.DOT else \
            dict(sigma2=RBF_P["sigma"] ** 2, l=RBF_P["l"], tol=1e-10, use_tol=(block == _lib.KFF))
        shp = {_lib.KEE: (a.m, b.m), _lib.KEF: (a.m, b.m, 3), _lib.KFF: (a.m, 3, b.m * 3)}[block]
        buf = torch.empty(shp, dtype=torch.float64, device="cuda")
        ms = _events_ms(lambda: packing.build_block(fam, block, a, b, 2.0, symmetric=sym, out=buf, **kw))
        pairs = a.pair_count(b)
        algo = (pairs + a.n) / 2 if sym else pairs
        fpp = {_lib.KEE: 2 * d + 12, _lib.KEF: 8 * d + 30, _lib.KFF: 32 * d + 100}[block] + (20 if fam == _lib.RBF else 0)
        tpp = {_lib.KEE: 2 * d, _lib.KEF: 8 * d, _lib.KFF: 32 * d}[block]
        tf = algo * fpp / (ms * 1e-3) / 1e12
        out[name] = {"what": what, "block": NAMES[block], "family": "Dot" if fam == _lib.DOT else "RBF", "d": d, "ms": ms,
                     "entries_per_s": float(np.prod(shp)) / (ms * 1e-3), "same_species_row_pairs": int(pairs),
                     "roofline": {"bound": "tensor", "achieved": tf, "peak": peak, "unit": "TFLOP/s", "frac": tf / peak,
                                  "flop_per_pair": fpp, "frac_tensor_flops_only": algo * tpp / (ms * 1e-3) / 1e12 / peak}}
        del buf

    m = full.m
    leg("rbf_kff_mix", _lib.RBF, _lib.KFF, full, full, True, "K_FF(X,X) RBF (exp + tol test), the headline set: %d force rows, Pt/H/O" % (3 * m))
    pe = dev_pack(synthetic.energy_set(833, seed=1))
    for fam, tag in ((_lib.DOT, "dot"), (_lib.RBF, "rbf")):
        leg(tag + "_kee_mix", fam, _lib.KEE, pe, pe, True, "K_EE, 833 structures of 192 atoms, Pt/H/O")
        leg(tag + "_kef_mix", fam, _lib.KEF, pe, full, False, "K_EF, 833 structures x %d force rows, Pt/H/O" % (3 * m))
    del pe
    ps = dev_pack(synthetic.force_set(m, seed=0, single_element=True))
    leg("dot_kff_single", _lib.DOT, _lib.KFF, ps, ps, True, "K_FF(X,X) Dot, %d force rows, one species (Si / Cu-like: every row pair counts)" % (3 * m))
    del ps
    torch.cuda.empty_cache()
    p50 = dev_pack(synthetic.force_set(m, seed=0, d=50))
    leg("dot_kff_d50", _lib.DOT, _lib.KFF, p50, p50, True, "K_FF(X,X) Dot, d = 50 (AgI.json: nmax = lmax = 4), %d force rows, Pt/H/O" % (3 * m))
    del p50
    p90 = dev_pack(synthetic.force_set(3334, seed=0, d=90))
    leg("dot_kff_d90", _lib.DOT, _lib.KFF, p90, p90, True, "K_FF(X,X) Dot, d = 90 (nmax = lmax = 5, chunked kernels), 10002 force rows, Pt/H/O")
    del p90
    torch.cuda.empty_cache()

    # K*.alpha GEMV: HBM bound, every K* byte read once
    N = 3 * m
    Ks = torch.empty((N, N), dtype=torch.float64, device="cuda").normal_()
    al = torch.empty(N, dtype=torch.float64, device="cuda").normal_()
    from csp_bo_b200 import gp
    ms = _events_ms(lambda: gp.predict_mean(Ks, al), reps=5)
    gbs = 8.0 * N * N / (ms * 1e-3) / 1e9
    out["gemv"] = {"what": "K*[%d,%d] . alpha" % (N, N), "ms": ms,
                   "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}}
    del Ks, al
    torch.cuda.empty_cache()

    # one LML + gradient evaluation (gaussianprocess.py:453-503) on the headline force set: fused traces, one N x N array
    y = np.random.default_rng(0).standard_normal(N)
    for name, kern in (("lml_grad_dot", Dot_mb(para=[SIGMA, SIGMA0], zeta=2)), ("lml_grad_rbf", RBF_mb(para=[RBF_P["sigma"], RBF_P["l"]], zeta=2))):
        model = GaussianProcess(kern, f_coef=30, noise_e=[0.005, 0.002, 0.1])
        model.train_x = {"energy": [], "force": X}
        model.train_y = {"energy": [], "force": list(y.reshape(m, 3))}
        model.update_y_train()
        model._packs = {"force": full}
        params = np.array(kern.parameters() + [0.005])
        model.log_marginal_likelihood(params, eval_gradient=True)
        torch.cuda.synchronize()
        base = torch.cuda.memory_allocated()
        torch.cuda.reset_peak_memory_stats()
        t0 = time.perf_counter()
        lml, grad = model.log_marginal_likelihood(params, eval_gradient=True)
        torch.cuda.synchronize()
        out[name] = {"what": "log_marginal_likelihood(eval_gradient=True), N = %d, %s" % (N, kern.name), "ms": (time.perf_counter() - t0) * 1e3,
                     "peak_extra_memory_in_N2_doubles": (torch.cuda.max_memory_allocated() - base + gp.inverse_workspace_bytes(N)) / (8.0 * N * N),
                     "peak_memory_note": "torch allocator peak (the one N x N array) + the factor-inverse workspace of csrc/dense.cu (leaf results + chunk buffer, about N^2/16 + N*384 doubles)",
                     "resident_cache_in_N2_doubles": 1.0 if getattr(model, "_lml_cache", None) is not None else 0.0,
                     "resident_cache_note": "Dot_mb keeps K(sigma = 1, sigma0 = 0) across the evaluations of one hyper-optimisation (GaussianProcess._dot_scaled_K; CSPBO_LML_CACHE=0 turns it off: 285 ms per evaluation instead of 183)",
                     "lml": float(lml), "grad": [float(g) for g in grad]}
        del model
        torch.cuda.empty_cache()
    return out


def lml_sharded_legs(full, m, world, barrier):
    """distributed.lml_sharded on the headline force set (N = 3m), Dot and RBF: panel-layout build + sharded Cholesky +
    per-panel solves for diag(K^-1) (+ row slabs of K^-1 contracted with dK/dl tiles for RBF)."""
    import torch
    import torch.distributed as dist
    from csp_bo_b200.distributed import lml_sharded
    from csp_bo_b200.kernels import Dot_mb, RBF_mb
    y = np.random.default_rng(0).standard_normal(3 * m)
    out = {}
    for name, kern in (("lml_grad_dot", Dot_mb(para=[SIGMA, SIGMA0], zeta=2)),
                       ("lml_grad_rbf", RBF_mb(para=[6.244955524643115, 0.5611029935713949], zeta=2))):
        ts = []
        cache = {} if kern.name == "Dot_mb" else None      # what GaussianProcess keeps while L-BFGS-B runs (Dot_mb: rows of K(1, 0))
        for _ in range(2):
            barrier(); t0 = time.perf_counter()
            lml, grad = lml_sharded(kern, {"force": full}, y, 0.005, 30.0, eval_gradient=True, cache=cache)
            barrier(); ts.append((time.perf_counter() - t0) * 1e3)
        t = torch.tensor([ts[0], min(ts)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = {"what": "lml_sharded(eval_gradient=True), N = %d, %s, %d GPUs" % (3 * m, kern.name, world), "ms": float(t[1]),
                     "ms_first_evaluation": float(t[0]),
                     "note": "Dot_mb: evaluations after the first reuse each rank's rows of K(sigma = 1, sigma0 = 0) (ShardedCovariance.build cache)" if cache is not None else "",
                     "lml": float(lml), "grad": [float(g) for g in grad]}
    return out


def fit_ef_leg(world, rank, barrier):
    """BASELINE.json config 3 shape: RBF_mb (examples/models/Si.json), 5352 energies of 64-atom structures + 3334 force
    atoms, one species, N = 15 354: K_EE, K_EF (+K_FE), K_FF + noise -> Cholesky -> alpha on the N GPUs."""
    import torch
    import torch.distributed as dist
    from csp_bo_b200 import gp, synthetic
    from csp_bo_b200.kernels import RBF_mb
    from csp_bo_b200.kernels import device as kdev
    from csp_bo_b200.packing import resident
    nE, nF = 5352, 3334
    data = resident({"energy": synthetic.energy_set(nE, seed=1, single_element=True, rows_per_structure=64),
                     "force": synthetic.force_set(nF, seed=2, single_element=True)})
    y = torch.from_numpy(np.random.default_rng(3).standard_normal(nE + 3 * nF)).cuda()
    kernel = RBF_mb(para=[6.244955524643115, 0.5611029935713949], zeta=2)
    ts = []
    for _ in range(3):
        barrier(); t0 = time.perf_counter()
        if world == 1:
            K = kdev.k_total_device(kernel, data)
            alpha = gp.fit_alpha(K, y, n_energy=nE, noise_e=0.005, f_coef=30.0, overwrite=True)
            del K
        else:
            from csp_bo_b200.distributed import fit_sharded
            alpha = fit_sharded(kernel, data, y, 0.005, 30.0)
        barrier(); ts.append((time.perf_counter() - t0) * 1e3)
    t = torch.tensor([min(ts[1:])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # residual through an independent dense build on this rank (N = 15 354: 1.9 GB)
    K = kdev.k_total_device(kernel, data)
    r = gp.predict_mean(K, alpha) - y
    d = torch.full_like(y, (30.0 * 0.005) ** 2)
    d[:nE] = 0.005 ** 2
    r += d * alpha
    res = float(torch.sqrt((r * r).sum() / (y * y).sum()))
    del K
    return {"wall_ms": float(t), "N": nE + 3 * nF, "residual": res,
            "what": "RBF_mb, 5352 energies (64-atom structures) + 3334 force atoms: K_EE / K_EF + K_FE / K_FF + noise -> "
                    + ("cuSOLVER potrf/potrs" if world == 1 else "panel-layout sharded build + sharded Cholesky (fit_sharded)") + " -> alpha"}


def md_step_leg():
    import torch
    from csp_bo_b200 import synthetic
    from csp_bo_b200.SO3 import SO3
    from csp_bo_b200.gaussianprocess import GaussianProcess
    from csp_bo_b200.kernels import Dot_mb
    from csp_bo_b200.utilities import tuple_to_list

    class Struct:                      # the few ase.Atoms members the descriptor reads
        def __init__(self, pos, num, cell):
            self.positions, self.numbers, self._cell, self.pbc = pos, num, cell, np.array([True] * 3)
            self.symbols = [{1: "H", 8: "O", 78: "Pt"}[int(z)] for z in num]

        def get_cell(self):
            return self._cell

        def __len__(self):
            return len(self.positions)

    rng = np.random.default_rng(5)
    g = np.array([(i, j, k) for i in range(4) for j in range(6) for k in range(8)], dtype=float)
    st = Struct(g * 2.6 + rng.normal(0, 0.25, g.shape), rng.choice([1, 8, 78], size=192, p=[0.5, 0.3, 0.2]),
                np.diag([4 * 2.6, 6 * 2.6, 8 * 2.6]))
    trainE, trainF = synthetic.energy_set(16, seed=1), synthetic.force_set(409, seed=2)
    E = [(x, float(rng.normal()), ele) for x, ele in tuple_to_list(trainE, mode='energy')]
    F = [(x, dx, rng.normal(size=3), ele) for x, dx, ele in tuple_to_list(trainF)]
    desc = SO3(nmax=3, lmax=3, rcut=3.5, alpha=2.0, derivative=True, stress=True)
    model = GaussianProcess(Dot_mb(para=[SIGMA, SIGMA0], zeta=2), descriptor=desc, f_coef=30, noise_e=[0.005, 0.002, 0.1])
    model.fit({"energy": E, "force": F}, show=False, opt=False)
    out = {}
    for stress in (False, True):
        ts = []
        for _ in range(6):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            model.predict_structure(st, stress=stress)
            torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        out["predict_structure_ms_stress_%d" % int(stress)] = float(np.median(ts[1:]))
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); desc.calculate(st); ts.append((time.perf_counter() - t0) * 1e3)
    out["so3_descriptor_ms"] = float(np.median(ts[1:]))
    out["what"] = ("192-atom Pt/H/O structure, SO3(nmax=3,lmax=3,rcut=3.5) + K*[577(+1152 stress) x 1243] + K*.alpha; "
                   "reference README.md:28: 35.767 s for 10 structures")
    return out


def real_data_leg():
    import torch
    from csp_bo_b200 import asedb
    from csp_bo_b200.SO3 import SO3
    from csp_bo_b200.gaussianprocess import GaussianProcess
    from csp_bo_b200.kernels import Dot_mb
    from csp_bo_b200.utilities import mae, r2
    G = os.path.join(ROOT, "tests", "golden")
    d = dict(np.load(os.path.join(G, "datasets.npz")))
    out = {}
    # config 2: examples/models/Si.json (RBF_mb, SO3 rcut 4.5) refitted from its database, tested on Si_244_DFT
    t0 = time.perf_counter()
    model = GaussianProcess(kernel=None)
    model.load(os.path.join(G, "si_model.json"), db_rows=asedb.rows_from_npz_dict(d, "si_model"))
    torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
    test = asedb.rows_from_npz_dict(d, "si244")
    t0 = time.perf_counter()
    E, E1, F, F1 = [], [], [], []
    for r in test:
        e, f, _ = model.predict_structure(r["atoms"], stress=True)
        E.append(r["energy"] / len(r["atoms"])); E1.append(e / len(r["atoms"])); F.append(r["force"].ravel()); F1.append(f.ravel())
    torch.cuda.synchronize(); t_test = time.perf_counter() - t0
    F, F1 = np.concatenate(F), np.concatenate(F1)
    out["si244"] = {"model": "examples/models/Si.json (RBF_mb zeta=2, SO3 nmax=lmax=3 rcut=4.5), %d energies + %d forces" % (model.N_energy, model.N_forces),
                    "load_descriptors_fit_s": t_fit, "test_structures": len(test), "test_atoms": int(sum(len(r["atoms"]) for r in test)),
                    "predict_ms_per_structure": 1e3 * t_test / len(test),
                    "test_energy_r2": r2(np.array(E), np.array(E1)), "test_energy_mae": mae(np.array(E), np.array(E1)),
                    "test_force_r2": r2(F, F1), "test_force_mae": mae(F, F1)}
    # config 4: Cu-EMT-300K, 100 x 32 atoms: train on 40 structures (energy + 4 forces each), then step through the rest
    rows = asedb.rows_from_npz_dict(d, "cu300")
    desc = SO3(nmax=3, lmax=3, rcut=4.0, alpha=2.0, derivative=True, stress=True)
    gp2 = GaussianProcess(Dot_mb(para=[5.0, 0.01], zeta=2), descriptor=desc, f_coef=10, noise_e=[0.005, 0.002, 0.1])
    for r in rows[:40]:
        r["energy_in"], r["force_in"] = True, [0, 8, 16, 24]
    t0 = time.perf_counter()
    gp2.extract_db(rows[:40])
    gp2.fit(show=False, opt=True)
    torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
    ts, F, F1 = [], [], []
    for r in rows[40:]:
        torch.cuda.synchronize(); t0 = time.perf_counter()
        e, f, _ = gp2.predict_structure(r["atoms"], stress=False)
        torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        F.append(r["force"].ravel()); F1.append(f.ravel())
    F, F1 = np.concatenate(F), np.concatenate(F1)
    out["cu_emt_300k"] = {"train": "40 structures: 40 energies + 160 forces, Dot_mb zeta=2, L-BFGS-B hyper-optimisation",
                          "descriptors_fit_opt_s": t_fit, "steps": len(ts), "ms_per_step_median": float(np.median(ts)),
                          "force_r2": r2(F, F1), "force_mae": mae(F, F1)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the extra per-config legs of the N = 1 line")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
