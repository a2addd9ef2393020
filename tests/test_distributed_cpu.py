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
