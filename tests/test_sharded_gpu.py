"""CUDA side of the sharded symmetric build.  Single process: exercises the packed-order row layout
and the write path of the sharded kernel with world = 1.  The world-2 tests spawn two ranks: on a box with >= 2 GPUs
they are NCCL ranks on two devices that store mirrored tiles into each other's HBM over NVLink peer memory; on a
ONE-GPU box both ranks sit on cuda:0 (gloo for the control plane and the panel broadcasts, CUDA IPC for the peer
buffers), so the `sharded` store paths of the tile kernel and the CUDA ShardedCholesky choreography still execute."""
import os

import numpy as np
import pytest

from tests._cases import force_tuple, max_norm_err

pytestmark = pytest.mark.gpu


def test_sharded_world1_matches_golden(fixtures, golden):
    import torch
    from csp_bo_b200 import _lib
    from csp_bo_b200.distributed import ShardedSymmetricBlock
    from csp_bo_b200.packing import Pack
    x2 = force_tuple(fixtures, "x2")
    p = Pack(x2[0], x2[1], x2[2], x2[3])
    sh = ShardedSymmetricBlock(p, _lib.KFF)
    sh.build(_lib.DOT, 2.0, scale=28.835 ** 2 * 2.0)
    K = sh.gather_full().cpu().numpy()
    assert max_norm_err(K, golden["dot_kff_x2x2"]) < 1e-10
    sh.close()


def _init_rank(rank, world, port):
    """NCCL on `world` devices when the box has them, otherwise every rank on cuda:0 over gloo."""
    import torch
    import torch.distributed as dist
    from csp_bo_b200 import _lib
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    multi = torch.cuda.device_count() >= world
    dev = rank if multi else 0
    torch.cuda.set_device(dev)
    _lib.check(_lib.load().cspbo_set_device(dev))
    if multi:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)


def _spawn2(fn, port, tol):
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(fn, args=(2, port, out), nprocs=2, join=True)
        assert len(out) == 2 and all(v < tol for v in out.values()), dict(out)


def _rank_main(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from csp_bo_b200 import _lib, synthetic
    from csp_bo_b200.distributed import ShardedSymmetricBlock
    from csp_bo_b200.packing import Pack, build_block
    _init_rank(rank, world, port)
    try:
        X = synthetic.force_set(300, seed=11)
        p = Pack(X[0], X[1], X[2], X[3])
        ref, _ = build_block(_lib.RBF, _lib.KFF, p, p, 2.0, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True, device=True)
        err = 0.0
        # plain block-cyclic rows / caller-order columns, and the superblock + zigzag + packed-column layout of the fit
        for kw in (dict(), dict(sb_blocks=4, zigzag=True, packed_cols=True), dict(sb_blocks=2, zigzag=True, packed_cols=True, balance=True)):
            sh = ShardedSymmetricBlock(p, _lib.KFF, **kw)
            sh._full.zero_()             # rows of padding slots are never written: give them a defined value
            torch.cuda.synchronize()
            dist.barrier()
            for _ in range(2):
                sh.build(_lib.RBF, 2.0, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True)
            K = sh.gather_full()
            err = max(err, float((K - ref.reshape(900, 900)).abs().max() / ref.abs().max()))
            # host-bound build in row slabs (copy-back of slab k overlaps slab k+1; slabs are final once every rank
            # has finished them): bit-identical to the one-shot build
            want = sh.local.cpu()
            sh.local.zero_()
            dist.barrier()
            host = torch.empty(tuple(sh.local.shape), dtype=torch.float64).pin_memory()
            sh.build_to_host(_lib.RBF, 2.0, host, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True, fractions=(0.0, 0.1, 0.4, 1.0))
            torch.cuda.synchronize()
            dist.barrier()
            err = max(err, float((host - want).abs().max()))
            sh.close()
        out[rank] = err
    finally:
        dist.destroy_process_group()


def test_sharded_two_ranks_peer_stores():
    _spawn2(_rank_main, 29611, 1e-12)


def test_build_to_host_world1():
    import torch
    from csp_bo_b200 import _lib, synthetic
    from csp_bo_b200.distributed import ShardedSymmetricBlock
    from csp_bo_b200.packing import Pack
    X = synthetic.force_set(333, seed=2)
    p = Pack(X[0], X[1], X[2], X[3])
    for kw in (dict(), dict(zigzag=True, packed_cols=True), dict(sb_blocks=4, zigzag=True, packed_cols=True)):
        sh = ShardedSymmetricBlock(p, _lib.KFF, **kw)
        sh._full.zero_()                 # rows of padding slots are never written: give them a defined value
        sh.build(_lib.DOT, 2.0, scale=700.0)
        want = sh.local.cpu()
        sh.local.zero_()
        host = torch.empty(tuple(sh.local.shape), dtype=torch.float64).pin_memory()
        sh.build_to_host(_lib.DOT, 2.0, host, scale=700.0)
        torch.cuda.synchronize()
        assert torch.equal(host, want), kw
        sh.close()


@pytest.mark.parametrize("sb_blocks,zigzag", [(1, False), (2, True), (4, True)])
def test_sharded_layouts_world1(fixtures, golden, sb_blocks, zigzag):
    """superblock ownership + packed column order of cspbo_block_build_sharded_ex (the layout the sharded
    Cholesky factorises in place) reproduce the golden K_FF after un-permuting."""
    from csp_bo_b200 import _lib
    from csp_bo_b200.distributed import ShardedSymmetricBlock
    from csp_bo_b200.packing import Pack
    x2 = force_tuple(fixtures, "x2")
    p = Pack(x2[0], x2[1], x2[2], x2[3])
    for packed in (False, True):
        sh = ShardedSymmetricBlock(p, _lib.KFF, sb_blocks=sb_blocks, zigzag=zigzag, packed_cols=packed)
        sh.build(_lib.DOT, 2.0, scale=28.835 ** 2 * 2.0)
        K = sh.gather_full().cpu().numpy()
        assert max_norm_err(K, golden["dot_kff_x2x2"]) < 1e-10
        sh.close()


@pytest.mark.parametrize("n_atoms,sb_blocks", [(27, 1), (300, 4), (301, 8)])
def test_sharded_cholesky_world1_matches_cusolver(n_atoms, sb_blocks):
    """fit through the sharded path (packed-order build -> panel Cholesky with look-ahead -> solve) against the
    dense single-call cuSOLVER fit of the same K; 1e-8 is the north_star tolerance on predictions."""
    import torch
    from csp_bo_b200 import _lib, gp, synthetic
    from csp_bo_b200.distributed import ShardedSymmetricBlock, fit_alpha_sharded
    from csp_bo_b200.packing import Pack, build_block
    X = synthetic.force_set(n_atoms, seed=3)
    p = Pack(X[0], X[1], X[2], X[3])
    s2 = 18.555440586011372 ** 2 * 2.0
    Kd, _ = build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=s2, symmetric=True, device=True)
    N = 3 * n_atoms
    y = torch.from_numpy(np.random.default_rng(0).standard_normal(N)).cuda()
    nf2 = (30.0 * 0.005) ** 2
    ref, logdet = gp.fit_alpha(Kd.reshape(N, N).clone(), y, n_energy=0, noise_e=0.005, f_coef=30.0, return_logdet=True)
    sh = ShardedSymmetricBlock(p, _lib.KFF, sb_blocks=sb_blocks, zigzag=True, packed_cols=True)
    sh.build(_lib.DOT, 2.0, scale=s2)
    alpha, chol = fit_alpha_sharded(sh, y, nf2, return_solver=True)
    Ktest = Kd.reshape(N, N)[: min(N, 64)]
    pr, pa = Ktest @ ref, Ktest @ alpha
    assert float((pr - pa).abs().max() / pr.abs().max()) < 1e-8
    assert abs(chol.logdet() - logdet) < 1e-8 * abs(logdet)
    sh.close()


def _chol_rank_main(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from csp_bo_b200 import _lib, gp, synthetic
    from csp_bo_b200.distributed import ShardedSymmetricBlock, fit_alpha_sharded
    from csp_bo_b200.packing import Pack, build_block
    _init_rank(rank, world, port)
    try:
        X = synthetic.force_set(500, seed=12)
        p = Pack(X[0], X[1], X[2], X[3])
        N = 1500
        Kd, _ = build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=700.0, symmetric=True, device=True)
        y = torch.from_numpy(np.random.default_rng(0).standard_normal(N)).cuda()
        ref = gp.fit_alpha(Kd.reshape(N, N).clone(), y, n_energy=0, noise_e=0.005, f_coef=30.0)
        err = 0.0
        for balance in (False, True):          # positional zigzag panels / panels dealt by their build work
            sh = ShardedSymmetricBlock(p, _lib.KFF, sb_blocks=4, zigzag=True, packed_cols=True, balance=balance)
            sh.build(_lib.DOT, 2.0, scale=700.0)
            alpha = fit_alpha_sharded(sh, y, (30.0 * 0.005) ** 2)
            pr, pa = Kd.reshape(N, N)[:64] @ ref, Kd.reshape(N, N)[:64] @ alpha
            err = max(err, float((pr - pa).abs().max() / pr.abs().max()))
            sh.close()
        out[rank] = err
    finally:
        dist.destroy_process_group()


def test_sharded_cholesky_two_ranks():
    _spawn2(_chol_rank_main, 29631, 1e-8)


def _mixed_problem(n_struct, n_atoms, single=False, rows=64):
    from csp_bo_b200 import synthetic
    E = synthetic.energy_set(n_struct, seed=21, single_element=single, rows_per_structure=rows)
    F = synthetic.force_set(n_atoms, seed=22, single_element=single)
    y = np.random.default_rng(5).standard_normal(n_struct + 3 * n_atoms)
    return {"energy": E, "force": F}, y


@pytest.mark.parametrize("fam", ["dot", "rbf"])
@pytest.mark.parametrize("n_struct,n_atoms,nbw", [(5, 40, 24), (30, 100, 48), (50, 0, 24), (0, 70, 48)])
def test_sharded_covariance_world1(fam, n_struct, n_atoms, nbw):
    """whole [[EE, EF], [FE, FF]] in the panel layout + sharded Cholesky against the dense single-GPU fit."""
    import torch
    from csp_bo_b200.distributed import fit_sharded
    from csp_bo_b200.kernels import Dot_mb, RBF_mb
    kernel = Dot_mb(para=[18.555440586011372, 0.01], zeta=2) if fam == "dot" else RBF_mb(para=[6.244955524643115, 0.5611029935713949], zeta=2)
    data, y = _mixed_problem(max(n_struct, 1), max(n_atoms, 1))
    if n_struct == 0:
        data.pop("energy"); y = y[1:]
    if n_atoms == 0:
        data.pop("force"); y = y[:n_struct]
    from csp_bo_b200 import gp
    from csp_bo_b200.kernels import device as kdev
    K = kdev.k_total_device(kernel, data)
    Kc = K.clone()
    ne = len(data["energy"][2]) if "energy" in data else 0
    ref = gp.fit_alpha(K, torch.from_numpy(y).cuda(), n_energy=ne, noise_e=0.01, f_coef=20.0, overwrite=True)
    alpha = fit_sharded(kernel, data, y, 0.01, 20.0, nbw=nbw)
    pr, pa = Kc @ ref, Kc @ alpha
    assert float((pr - pa).abs().max() / pr.abs().max()) < 1e-8
    assert float((ref - alpha).abs().max() / ref.abs().max()) < 1e-6


@pytest.mark.parametrize("fam,n_struct,n_atoms,nbw", [("dot", 5, 40, 24), ("rbf", 30, 100, 48), ("rbf", 50, 0, 24), ("dot", 0, 70, 48),
                                                      ("rbf", 0, 70, 48)])
def test_lml_sharded_world1_matches_dense(fam, n_struct, n_atoms, nbw):
    """distributed.lml_sharded (panel-layout K, sharded Cholesky, per-panel triangular solves for diag(K^-1), dK/dl traces
    over row slabs of K^-1) against the dense single-GPU evaluation of GaussianProcess.log_marginal_likelihood."""
    from csp_bo_b200.distributed import lml_sharded
    from csp_bo_b200.gaussianprocess import GaussianProcess
    from csp_bo_b200.kernels import Dot_mb, RBF_mb
    from csp_bo_b200.utilities import tuple_to_list
    kernel = Dot_mb(para=[18.555440586011372, 0.05], zeta=2) if fam == "dot" else RBF_mb(para=[6.244955524643115, 0.5611029935713949], zeta=2)
    data, y = _mixed_problem(max(n_struct, 1), max(n_atoms, 1))
    lists = {}
    if n_struct:
        lists["energy"] = [(x, float(y[i]), ele) for i, (x, ele) in enumerate(tuple_to_list(data["energy"], mode='energy'))]
    else:
        data.pop("energy"); y = y[1:]
    if n_atoms:
        ne = n_struct
        lists["force"] = [(x, dx, y[ne + 3 * i: ne + 3 * i + 3], ele) for i, (x, dx, ele) in enumerate(tuple_to_list(data["force"]))]
    else:
        data.pop("force"); y = y[:n_struct]
    model = GaussianProcess(kernel, f_coef=20.0, noise_e=[0.01, 0.002, 0.1])
    model.set_train_pts(lists)
    params = np.array(kernel.parameters() + [0.01])
    lml, grad = model.log_marginal_likelihood(params, eval_gradient=True)
    lml_s, grad_s = lml_sharded(kernel, data, y, 0.01, 20.0, eval_gradient=True, nbw=nbw)
    assert abs(lml - lml_s) <= 1e-9 * abs(lml)
    assert max_norm_err(grad_s, grad) < 1e-7, (grad_s, grad)
    assert abs(lml_sharded(kernel, data, y, 0.01, 20.0, nbw=nbw) - model.log_marginal_likelihood(params)) <= 1e-9 * abs(lml)
    if fam == "dot":
        # Dot_mb: the rows of K(sigma = 1, sigma0 = 0) cached across evaluations (ShardedCovariance.build): same numbers
        # while (sigma, sigma0, noise) move
        cache = {}
        for sig, sig0, ne in ((18.5, 0.05, 0.01), (7.0, 0.7, 0.02), (31.0, 0.02, 0.005)):
            kernel.update([sig, sig0])
            a = lml_sharded(kernel, data, y, ne, 20.0, eval_gradient=True, nbw=nbw, cache=cache)
            b = lml_sharded(kernel, data, y, ne, 20.0, eval_gradient=True, nbw=nbw)
            assert "K1" in cache
            assert abs(a[0] - b[0]) <= 1e-11 * abs(b[0]) and max_norm_err(a[1], b[1]) < 1e-9


def _cov_rank_main(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from csp_bo_b200 import _lib, gp
    from csp_bo_b200.distributed import fit_sharded
    from csp_bo_b200.kernels import RBF_mb
    from csp_bo_b200.kernels import device as kdev
    _init_rank(rank, world, port)
    try:
        kernel = RBF_mb(para=[6.244955524643115, 0.5611029935713949], zeta=2)
        data, y = _mixed_problem(130, 300)
        K = kdev.k_total_device(kernel, data)
        Kc = K.clone()
        ref = gp.fit_alpha(K, torch.from_numpy(y).cuda(), n_energy=130, noise_e=0.01, f_coef=20.0, overwrite=True)
        alpha = fit_sharded(kernel, data, y, 0.01, 20.0, nbw=96)
        pr, pa = Kc @ ref, Kc @ alpha
        err = float((pr - pa).abs().max() / pr.abs().max())
        # the regressor picks the sharded fit up by itself under torch.distributed: same predictions and
        # predictive std as the dense single-GPU algebra
        from csp_bo_b200.gaussianprocess import GaussianProcess
        from csp_bo_b200.utilities import tuple_to_list
        E = [(x, float(y[i]), ele) for i, (x, ele) in enumerate(tuple_to_list(data["energy"], mode='energy'))]
        F = [(x, dx, y[130 + 3 * i: 133 + 3 * i], ele) for i, (x, dx, ele) in enumerate(tuple_to_list(data["force"]))]
        model = GaussianProcess(kernel, f_coef=20.0, noise_e=[0.01, 0.002, 0.1])
        model.fit({"energy": E, "force": F}, show=False, opt=False)
        assert model._solver is not None
        test, _ = _mixed_problem(6, 10)
        test = {"energy": [(x, ele) for x, ele in tuple_to_list(test["energy"], mode='energy')],
                "force": [(x, dx, ele) for x, dx, ele in tuple_to_list(test["force"])]}
        mean, std = model.predict(test, return_std=True)
        Ks = kdev.k_total_device(kernel, test, data)
        mean_ref = (Ks @ ref).cpu().numpy()
        var_ref = kernel.diag(test) - (Ks * gp.cho_solve(K, Ks.T.contiguous()).T).sum(dim=1).cpu().numpy()
        err2 = float(np.abs(mean - mean_ref).max() / np.abs(mean_ref).max())
        err3 = float(np.abs(std - np.sqrt(np.maximum(var_ref, 0))).max() / np.sqrt(np.maximum(var_ref, 0)).max())
        # test groups sharded over the ranks, predictions all-gathered: identical to the replicated prediction
        from csp_bo_b200.distributed import predict_sharded
        m2, s2 = predict_sharded(model, test, return_std=True)
        err4 = float(max(np.abs(m2 - mean).max() / np.abs(mean).max(), np.abs(s2 - std).max() / np.abs(std).max()))
        # LML + gradient over the ranks (what fit(opt=True) evaluates under torchrun) against the dense algebra
        from csp_bo_b200.distributed import lml_sharded
        params = np.array(kernel.parameters() + [0.01])
        lml_s, grad_s = model.log_marginal_likelihood(params, eval_gradient=True)          # sharded: world > 1
        dense = GaussianProcess(kernel, f_coef=20.0, noise_e=[0.01, 0.002, 0.1])
        dense.set_train_pts({"energy": E, "force": F})
        import csp_bo_b200.gaussianprocess as gpm
        ws, gpm._world_size = gpm._world_size, (lambda: 1)
        try:
            lml_d, grad_d = dense.log_marginal_likelihood(params, eval_gradient=True)
        finally:
            gpm._world_size = ws
        err5 = max(abs(lml_s - lml_d) / abs(lml_d), float(np.abs(grad_s - grad_d).max() / np.abs(grad_d).max()) * 1e-1)
        # Dot_mb under world 2: the regressor keeps every rank's rows of K(1, 0) across evaluations (second call: cached)
        from csp_bo_b200.kernels import Dot_mb
        dk = Dot_mb(para=[18.5, 0.05], zeta=2)
        md = GaussianProcess(dk, f_coef=20.0, noise_e=[0.01, 0.002, 0.1])
        md.set_train_pts({"energy": E, "force": F})
        md.log_marginal_likelihood(np.array([18.5, 0.05, 0.01]), eval_gradient=True)
        l_c, g_c = md.log_marginal_likelihood(np.array([9.0, 0.4, 0.02]), eval_gradient=True)
        assert isinstance(md._lml_cache[1], dict) and "K1" in md._lml_cache[1]
        l_f, g_f = lml_sharded(dk, md._packs, md.y_train[:, 0], 0.02, 20.0, eval_gradient=True)
        err5 = max(err5, abs(l_c - l_f) / abs(l_f) * 1e2, float(np.abs(g_c - g_f).max() / np.abs(g_f).max()))     # rounding of the rescale x conditioning
        out[rank] = max(err, err2, err3 * 1e-2, err4 * 1e2, err5)
    finally:
        dist.destroy_process_group()


def test_sharded_covariance_two_ranks():
    _spawn2(_cov_rank_main, 29651, 1e-8)


@pytest.mark.parametrize("N,nbk,j0,hj", [(1500, 256, 0, 384), (1500, 256, 288, 600), (2100, 1100, 96, 96), (700, 2048, 192, 508), (700, 2048, 0, 96), (1030, 128, 1000, 30)])
def test_blocked_substitution_matches_triangular_solves(N, nbk, j0, hj):
    """distributed._BlockedSubstitution (the slab solves of lml_sharded): explicit diagonal-block inverses + dgemm sweeps
    against torch's triangular solves; ragged first and last blocks, slabs wider than a block, junk below U's diagonal."""
    import torch
    from csp_bo_b200.distributed import _BlockedSubstitution
    g = torch.Generator(device="cuda").manual_seed(N + j0)
    A = torch.randn(N, N + 30, dtype=torch.float64, device="cuda", generator=g)
    U = torch.linalg.cholesky(A @ A.T / N + 0.5 * torch.eye(N, dtype=torch.float64, device="cuda"), upper=True)
    Ujunk = U + torch.tril(torch.randn(N, N, dtype=torch.float64, device="cuda", generator=g), -1)   # the sharded factor's lower part is scratch
    sub = _BlockedSubstitution(Ujunk, N, nbk=nbk)
    E = torch.zeros((N - j0, hj), dtype=torch.float64, device="cuda")
    E[:hj].fill_diagonal_(1.0)
    Yref = torch.linalg.solve_triangular(U[j0:, j0:].T, E, upper=False)
    Y = sub.forward(j0, hj)
    assert float((Y - Yref).abs().max()) <= 1e-11 * float(Yref.abs().max())
    Zref = torch.linalg.solve_triangular(U[j0:, j0:], Yref, upper=True)
    Z = sub.backward(j0, Y)
    assert float((Z - Zref).abs().max()) <= 1e-11 * float(Zref.abs().max())
    assert torch.equal(torch.triu(Ujunk), U)          # the factor itself is left alone
