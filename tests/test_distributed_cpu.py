"""world_size-2 gloo test of the N>1 host logic (block-cyclic ownership, gather of unequal row
blocks, undoing the packed order).  The tile kernel is replaced by the oracle's K_FF so the test
runs without a GPU; the CUDA side of the same path is covered by tests/test_sharded_gpu.py."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from csp_bo_b200 import distributed as D


def test_ownership_partition():
    for nb in (1, 2, 7, 24, 834):
        for world in (1, 2, 4, 8):
            seen = sorted(b for r in range(world) for b in D.owned_blocks(nb, r, world))
            assert seen == list(range(nb))
            assert sum(D.sharded_rows(nb, r, world) for r in range(world)) == 8 * nb
            # triangular work (blocks b >= a) is balanced to within one block row
            work = [sum(nb - a for a in D.owned_blocks(nb, r, world)) for r in range(world)]
            assert max(work) - min(work) <= nb


def _worker(rank, world, port, K, order, m, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        nb = len(order) // 8
        pos = D.local_positions(nb, rank, world)
        g = order[pos]
        local = np.zeros((len(pos), 3, m * 3))
        K4 = K.reshape(m, 3, m * 3)
        local[g >= 0] = K4[g[g >= 0]]                      # what the sharded kernel leaves in this rank's buffer
        parts = D.gather_row_blocks(torch.from_numpy(local), nb, rank, world)
        full = D.assemble_full(parts, order, m, 3)
        out[rank] = float((full - torch.from_numpy(K)).abs().max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("m", [27, 40])
def test_gloo_world2_gather_and_unpermute(m, golden):
    K = golden["dot_kff_x2x2"] if m == 27 else None
    rng = np.random.default_rng(m)
    if K is None:
        A = rng.standard_normal((3 * m, 3 * m))
        K = A + A.T
    nb = (m + 7) // 8
    order = np.full(nb * 8, -1, dtype=np.int64)
    order[:m] = rng.permutation(m)
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        port = 29500 + (os.getpid() % 2000)
        mp.spawn(_worker, args=(world, port, np.ascontiguousarray(K), order, m, out), nprocs=world, join=True)
        assert len(out) == world and all(v == 0.0 for v in out.values()), dict(out)


# ------------------------------------------------------------------------------------------
# sharded Cholesky: the rank choreography (ownership, look-ahead, broadcasts, triangular updates) with a
# torch-CPU twin of the device panel operations, world 1 in-process and world 2 over gloo
class TorchOps:
    """CPU twin of distributed._DeviceOps (csrc/gp.cu chol_diag / chol_apply_inv / chol_update), test-only."""
    cuda = False

    def diag(self, D, inv, info, stream):
        Uss, inf = torch.linalg.cholesky_ex(D, upper=True)
        info.copy_(inf.to(info.dtype).reshape(1))
        D.copy_(torch.triu(Uss) + torch.tril(D, -1))
        inv.copy_(torch.linalg.solve_triangular(torch.triu(Uss), torch.eye(D.shape[0], dtype=D.dtype), upper=True))

    def apply_inv(self, inv, src, dst, stream):
        dst.copy_(inv.T @ src)

    def update(self, C, A, B, stream):
        C -= A.T @ B

    def solve(self, U, B):
        Ut = torch.triu(U)
        return torch.cholesky_solve(B, Ut, upper=True)


def _spd(N, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((N, N))
    return A @ A.T / N + np.eye(N)


def _rows_of(K, N, nbw, rank, world, zigzag):
    S = (N + nbw - 1) // nbw
    mine = [s for s in range(S) if D.superblock_owner(s, world, zigzag) == rank]
    local = torch.zeros((len(mine) * nbw, N), dtype=torch.float64)
    for i, s in enumerate(mine):
        h = min(nbw, N - s * nbw)
        local[i * nbw:i * nbw + h] = torch.from_numpy(K[s * nbw:s * nbw + h])
    return local


@pytest.mark.parametrize("N,nbw,zigzag", [(24, 24, False), (30, 24, False), (49, 24, True), (73, 24, True), (100, 24, True),
                                          (193, 48, True), (96, 48, False), (240, 24, False)])
def test_sharded_cholesky_world1(N, nbw, zigzag):
    K = _spd(N, N)
    ch = D.ShardedCholesky(_rows_of(K, N, nbw, 0, 1, zigzag), N, nbw, 0, 1, zigzag, ops=TorchOps())
    ch.add_noise(0.25)
    ch.factor().check()
    ref = np.linalg.cholesky(K + 0.25 * np.eye(N)).T
    assert np.abs(np.triu(ch.U.numpy()) - ref).max() < 1e-12
    y = np.random.default_rng(1).standard_normal(N)
    assert np.abs(ch.solve(torch.from_numpy(y)).numpy() - np.linalg.solve(K + 0.25 * np.eye(N), y)).max() < 1e-10
    assert abs(ch.logdet() - np.linalg.slogdet(K + 0.25 * np.eye(N))[1]) < 1e-10


def test_sharded_cholesky_not_positive_definite():
    K = _spd(60, 3)
    K[30, 30] = -5.0
    ch = D.ShardedCholesky(_rows_of(K, 60, 24, 0, 1, False), 60, 24, 0, 1, ops=TorchOps())
    from csp_bo_b200._lib import CspboError
    with pytest.raises(CspboError):
        ch.factor().check()


def _chol_worker(rank, world, port, K, N, nbw, zigzag, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ch = D.ShardedCholesky(_rows_of(K, N, nbw, rank, world, zigzag), N, nbw, rank, world, zigzag, ops=TorchOps())
        ch.add_noise(0.5)
        ch.factor().check()
        ref = np.linalg.cholesky(K + 0.5 * np.eye(N)).T
        y = np.random.default_rng(7).standard_normal(N)
        a = ch.solve(torch.from_numpy(y)).numpy()
        out[rank] = (float(np.abs(np.triu(ch.U.numpy()) - ref).max()),
                     float(np.abs(a - np.linalg.solve(K + 0.5 * np.eye(N), y)).max()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("N,nbw,zigzag,world", [(100, 24, True, 2), (170, 24, False, 2), (130, 24, True, 3)])
def test_sharded_cholesky_gloo(N, nbw, zigzag, world):
    K = _spd(N, 11)
    with mp.Manager() as mgr:
        out = mgr.dict()
        port = 31500 + (os.getpid() % 2000) + N
        mp.spawn(_chol_worker, args=(world, port, K, N, nbw, zigzag, out), nprocs=world, join=True)
        assert len(out) == world and all(v[0] < 1e-12 and v[1] < 1e-10 for v in out.values()), dict(out)


def test_layout_roundtrip_superblocks_packed_cols():
    """assemble_full undoes every (sb_blocks, zigzag, packed_cols) layout of cspbo_block_build_sharded_ex."""
    m = 61
    rng = np.random.default_rng(5)
    A = rng.standard_normal((3 * m, 3 * m)); K = A + A.T
    nb = (m + 7) // 8
    order = np.full(nb * 8, -1, dtype=np.int64)
    order[:m] = rng.permutation(m)
    K4 = K.reshape(m, 3, m, 3)
    for world in (1, 2, 3):
        for G in (1, 2, 4):
            for zz in (False, True):
                for packed in (False, True):
                    parts = []
                    for r in range(world):
                        pos = D.local_positions(nb, r, world, G, zz)
                        g = np.where(pos >= 0, order[np.maximum(pos, 0)], -1)
                        loc = np.zeros((len(pos), 3, m, 3))
                        src = K4[g[g >= 0]]
                        loc[g >= 0] = src[:, :, order[:m], :] if packed else src
                        parts.append(loc.reshape(len(pos), 3, m * 3))
                    full = D.assemble_full(parts, order, m, 3, G, zz, packed)
                    assert np.array_equal(full, K), (world, G, zz, packed)


def test_group_slices_cover_everything():
    for n in (0, 1, 5, 17, 192):
        for world in (1, 2, 3, 8):
            parts = [D.group_slice(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def test_predict_sharded_assembles_in_reference_order(fixtures):
    """predict_sharded over 3 simulated ranks returns model.predict's [E...; F...] vector: the model here is a
    stand-in whose prediction is a checksum of each group's rows, so the test runs without a GPU."""
    from tests._cases import energy_tuple, force_tuple, take_groups, take_groups_e

    class Model:
        def predict(self, X, total_E=False, return_std=False):
            out = []
            if "energy" in X:
                x, ele, ind = X["energy"]
                off = np.concatenate(([0], np.cumsum(ind)))
                out += [x[off[i]:off[i + 1]].sum() for i in range(len(ind))]
            if "force" in X:
                x, dx, ele, ind = X["force"]
                off = np.concatenate(([0], np.cumsum(ind)))
                for i in range(len(ind)):
                    out += list(dx[off[i]:off[i + 1]].sum(axis=(0, 1)))
            out = np.asarray(out, dtype=np.float64)
            return (out, 2 * out) if return_std else out

    X = {"energy": take_groups_e(energy_tuple(fixtures, "ee2"), 0, 7), "force": take_groups(force_tuple(fixtures, "big"), 0, 23)}
    ref, ref_std = Model().predict(X, return_std=True)
    world = 3
    import torch.distributed as dist_mod
    # simulate the ranks one after another: first collect every rank's local vectors, then replay the gathers
    locals_ = []
    for r in range(world):
        e0, e1 = D.group_slice(7, r, world)
        f0, f1 = D.group_slice(23, r, world)
        mine = {"energy": D.slice_groups(X["energy"], e0, e1, 'energy'), "force": D.slice_groups(X["force"], f0, f1, 'force')}
        locals_.append((Model().predict(mine), e1 - e0))
    E = np.concatenate([v[:ne] for v, ne in locals_])
    F = np.concatenate([v[ne:] for v, ne in locals_])
    assert np.array_equal(np.concatenate((E, F)), ref)
    # and the function itself at world 1
    got, got_std = D.predict_sharded(Model(), X, return_std=True)
    assert np.array_equal(got, ref) and np.array_equal(got_std, ref_std)


class _SumModel:
    def predict(self, X, total_E=False, return_std=False):
        out = []
        if "energy" in X:
            out += [float(np.sum(x)) for x, _ in X["energy"]]
        if "force" in X:
            for x, dx, _ in X["force"]:
                out += list(np.sum(dx, axis=(0, 1)))
        out = np.asarray(out, dtype=np.float64)
        return (out, 2 * out) if return_std else out


def _pred_worker(rank, world, port, X, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        got, std = D.predict_sharded(_SumModel(), X, return_std=True)
        ref, ref_std = _SumModel().predict(X, return_std=True)
        out[rank] = bool(np.array_equal(got, ref) and np.array_equal(std, ref_std))
    finally:
        dist.destroy_process_group()


def test_predict_sharded_gloo_world2():
    """unequal slices (3 structures, 5 force atoms over 2 ranks) through the variable-length all_gather"""
    rng = np.random.default_rng(0)
    X = {"energy": [(rng.standard_normal((4 + i, 6)), np.ones(4 + i, dtype=int)) for i in range(3)],
         "force": [(rng.standard_normal((3 + i, 6)), rng.standard_normal((3 + i, 6, 3)), np.ones(3 + i, dtype=int)) for i in range(5)]}
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_pred_worker, args=(2, 33500 + (os.getpid() % 2000), X, out), nprocs=2, join=True)
        assert len(out) == 2 and all(out.values()), dict(out)


def test_panel_layout_host_logic():
    """panel_local_rows / panel_permutation: the rows a rank holds cover every global row exactly once, local rows follow
    the kernel's formula (s // world) * nbw + r % nbw, and the permutation sends the caller's [E; F] rows to
    [energies packed | padding | forces packed x 3]."""
    nbw = 24
    for world in (1, 2, 3):
        for N in (24, 50, 97, 240):
            S = (N + nbw - 1) // nbw
            seen = np.zeros(N, dtype=int)
            for r in range(world):
                mine = [s for s in range(S) if D.superblock_owner(s, world, True) == r]
                loc, glob = D.panel_local_rows(mine, nbw, 0, N)
                seen[glob] += 1
                for l, g in zip(loc, glob):
                    s = g // nbw
                    assert l == (s // world) * nbw + g % nbw
                # sub-ranges partition the rows of the rank
                l1, g1 = D.panel_local_rows(mine, nbw, 0, N // 2)
                l2, g2 = D.panel_local_rows(mine, nbw, N // 2, N)
                assert sorted(np.concatenate((g1, g2)).tolist()) == sorted(glob.tolist())
            assert np.all(seen == 1)
    rng = np.random.default_rng(0)
    mE, mF, NE_pad = 5, 7, 24
    oe, of = rng.permutation(mE), rng.permutation(mF)
    perm = D.panel_permutation(oe, of, NE_pad)
    assert len(perm) == mE + 3 * mF and len(set(perm.tolist())) == len(perm)
    y = np.arange(mE + 3 * mF, dtype=float)
    packed = np.zeros(NE_pad + 3 * mF)
    packed[perm] = y
    assert np.array_equal(packed[:mE], y[:mE][oe])                                   # energies in packed order
    assert np.all(packed[mE:NE_pad] == 0)                                            # identity padding rows
    assert np.array_equal(packed[NE_pad:].reshape(mF, 3), y[mE:].reshape(mF, 3)[of])  # forces, 3 rows per atom
    assert np.array_equal(packed[perm], y)


def test_balanced_superblock_perm_properties():
    """Work-balanced ownership of the sharded build: a permutation of the superblocks that only moves them inside their
    cycle of `world` (so row slabs stay contiguous ranges of packed blocks), with inverse, and loads no worse than the
    positional zigzag dealing."""
    from csp_bo_b200.distributed import balanced_superblock_perm, superblock_owner
    rng = np.random.default_rng(0)
    for nsb, W, z in ((26, 8, True), (209, 8, True), (13, 4, False), (5, 8, True), (1, 2, True), (16, 2, True)):
        w = np.sort(rng.uniform(1, 3, nsb))[::-1] * np.linspace(2, 0.1, nsb)
        perm, inv = balanced_superblock_perm(w, W, z)
        assert len(perm) == -(-nsb // W) * W and sorted(int(t) for t in perm if t >= 0) == list(range(nsb))
        load, plain = np.zeros(W), np.zeros(W)
        for v, t in enumerate(perm):
            if t >= 0:
                assert inv[t] == v and t // W == v // W
                load[superblock_owner(v, W, z)] += w[t]
        for t in range(nsb):
            plain[superblock_owner(t, W, z)] += w[t]
        assert load.max() <= plain.max() * (1 + 1e-12)


@pytest.mark.parametrize("Ntot,NE_pad,nbw", [(20001, 0, 384), (15354 + 24, 5376, 384), (1000, 96, 48), (96, 96, 96), (50, 0, 384)])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_inverse_slabs_partition_the_columns(Ntot, NE_pad, nbw, world):
    """distributed._inverse_slabs: the ranks' slabs tile [0, Ntot) exactly once, in order, never straddle the energy /
    force boundary, stay block-aligned and under the column cap; costs are balanced at production sizes."""
    from csp_bo_b200.distributed import _inverse_slabs, _SLAB_MAX
    for rbf in (False, True):
        cover, cost = [], []
        for r in range(world):
            sl = _inverse_slabs(Ntot, NE_pad, r, world, rbf, nbw)
            cost.append(sum((Ntot - (a + b) / 2.0) ** 2 * (b - a) for a, b in sl))
            cover += sl
        assert cover[0][0] == 0 and cover[-1][1] == Ntot
        unit = 96 if nbw % 96 == 0 else nbw
        for (a, b), nxt in zip(cover, cover[1:] + [(Ntot, Ntot)]):
            assert a < b <= Ntot and b == nxt[0] and b - a <= _SLAB_MAX
            assert a % unit == 0 and (b % unit == 0 or b == Ntot)
            assert not (a < NE_pad < b)
        if Ntot >= 15000 and not rbf and world > 1:
            assert max(cost) <= 1.12 * (sum(cost) / world)
