"""CUDA side of the sharded symmetric build.  Single process: exercises the packed-order row layout
and the write path of the sharded kernel with world = 1.  With >= 2 GPUs it also spawns two NCCL
ranks that store mirrored tiles into each other's HBM over NVLink peer memory."""
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


def _rank_main(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from csp_bo_b200 import _lib, synthetic
    from csp_bo_b200.distributed import ShardedSymmetricBlock
    from csp_bo_b200.packing import Pack, build_block
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    _lib.check(_lib.load().cspbo_set_device(rank))
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        X = synthetic.force_set(300, seed=11)
        p = Pack(X[0], X[1], X[2], X[3])
        ref, _ = build_block(_lib.RBF, _lib.KFF, p, p, 2.0, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True, device=True)
        sh = ShardedSymmetricBlock(p, _lib.KFF)
        for _ in range(2):
            sh.build(_lib.RBF, 2.0, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True)
        K = sh.gather_full()
        out[rank] = float((K - ref.reshape(900, 900)).abs().max() / ref.abs().max())
        sh.close()
    finally:
        dist.destroy_process_group()


def test_sharded_two_gpus_peer_stores():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_rank_main, args=(2, 29611, out), nprocs=2, join=True)
        assert len(out) == 2 and all(v < 1e-12 for v in out.values()), dict(out)
