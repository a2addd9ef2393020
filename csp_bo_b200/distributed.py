"""Multi-GPU K(X,X): one process per GPU (torch.distributed), the symmetric covariance block
sharded by row blocks.

The reference parallelises K_FF only, by cutting the x2 rows into nprocs slices and
`Allreduce(SUM)`-ing the whole dense matrix (cspbo/kernels/dot_kernel.py:188-247).  Here the
descriptors are replicated (every rank packs the same training set), the groups are taken in the
pack's order and cut into blocks of 8, and rank r owns the row blocks r, r+world, r+2*world, ...
(block-cyclic: the triangular work is balanced).  A rank computes only the tiles (a owned, b >= a)
and the tile kernel stores the mirrored tile straight into the owner's HBM through NVLink peer
memory (CUDA IPC mapped pointers), so after one kernel + barrier every rank holds complete rows and
no collective has moved the matrix.  NCCL is used for the plumbing only (handle exchange, barriers,
the optional gather of the row blocks for a replicated Cholesky, the alpha broadcast).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import BlockDesc, c_vp


# ----------------------------------------------------------------------------- pure host logic
def owned_blocks(n_blocks, rank, world):
    """Block ids owned by `rank` (block-cyclic)."""
    return list(range(rank, n_blocks, world))


def sharded_rows(n_blocks, rank, world):
    """Number of group rows (incl. padding slots) in `rank`'s buffer."""
    return 8 * len(owned_blocks(n_blocks, rank, world))


def local_positions(n_blocks, rank, world):
    """Packed positions (block*8+slot) of the rows of `rank`'s buffer, in buffer order."""
    blocks = np.asarray(owned_blocks(n_blocks, rank, world), dtype=np.int64)
    return (blocks[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)


def assemble_full(parts, order, m, ncomp=3):
    """Rebuild the dense matrix in the caller's group order from per-rank row blocks.

    parts[r]: array [sharded_rows(r), ncomp, m*ncomp] (K_FF) or [sharded_rows(r), m] (K_EE, ncomp=1)
    order:    packed position -> group id (-1 = padding slot), length 8*n_blocks
    Works on numpy arrays and torch tensors (CPU or CUDA)."""
    world = len(parts)
    n_blocks = len(order) // 8
    order = np.asarray(order)
    is_torch = type(parts[0]).__module__.startswith("torch")
    if is_torch:
        import torch
        full = torch.empty((m,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype, device=parts[0].device)
    else:
        full = np.empty((m,) + tuple(parts[0].shape[1:]), dtype=parts[0].dtype)
    for r in range(world):
        pos = local_positions(n_blocks, r, world)
        g = order[pos]
        keep = np.nonzero(g >= 0)[0]
        if len(keep) == 0:
            continue
        if is_torch:
            import torch
            idx = torch.as_tensor(g[keep], device=parts[r].device, dtype=torch.long)
            src = torch.as_tensor(keep, device=parts[r].device, dtype=torch.long)
            full[idx] = parts[r][src]
        else:
            full[g[keep]] = parts[r][keep]
    if ncomp == 1:
        return full.reshape(m, m)
    return full.reshape(m * ncomp, m * ncomp)


def gather_row_blocks(local, n_blocks, rank, world, group=None):
    """all_gather of the (unequal) per-rank row blocks; device-agnostic (NCCL on CUDA tensors, gloo on
    CPU tensors).  Returns the list of per-rank arrays, trimmed to their true row counts."""
    import torch
    import torch.distributed as dist
    max_rows = 8 * ((n_blocks + world - 1) // world)
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return [p[: sharded_rows(n_blocks, r, world)] for r, p in enumerate(parts)]


# ----------------------------------------------------------------------------- device side
class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can view a cudaMalloc'd buffer."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class ShardedSymmetricBlock:
    """K_FF(X,X) or K_EE(X,X) of one packed training set, distributed over the ranks of `group`."""

    def __init__(self, pack, block=_lib.KFF, group=None):
        import torch
        import torch.distributed as dist
        self.lib = _lib.load()
        self.pack, self.block, self.group = pack, block, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.m = pack.m
        self.n_blocks = (self.m + 7) // 8
        self.ncomp = 3 if block == _lib.KFF else 1
        self.row_shape = (3, self.m * 3) if block == _lib.KFF else (self.m,)
        self.rows = int(self.lib.cspbo_sharded_rows(pack.handle, self.rank, self.world))
        assert self.rows == sharded_rows(self.n_blocks, self.rank, self.world)
        order = np.empty(self.n_blocks * 8, dtype=np.int32)
        _lib.check(self.lib.cspbo_pack_order(pack.handle, c_vp(order.ctypes.data), len(order)))
        self.order = order
        nbytes = 8 * self.rows * int(np.prod(self.row_shape))
        ptr = c_vp()
        _lib.check(self.lib.cspbo_dev_alloc(nbytes, ctypes.byref(ptr)))
        self._own = ptr
        self._peers = [None] * self.world
        self._ptrs = (c_vp * self.world)()
        self._ptrs[self.rank] = ptr
        if self.world > 1:
            handle = (ctypes.c_ubyte * 64)()
            _lib.check(self.lib.cspbo_ipc_export(ptr, ctypes.cast(handle, c_vp)))
            mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device="cuda")
            allh = [torch.empty(64, dtype=torch.uint8, device="cuda") for _ in range(self.world)]
            dist.all_gather(allh, mine, group=group)
            for r in range(self.world):
                if r == self.rank:
                    continue
                hb = (ctypes.c_ubyte * 64)(*allh[r].cpu().tolist())
                pp = c_vp()
                _lib.check(self.lib.cspbo_ipc_open(ctypes.cast(hb, c_vp), ctypes.byref(pp)))
                self._peers[r] = pp
                self._ptrs[r] = pp
        self.local = torch.as_tensor(_RawCuda(ptr.value, (max(self.rows, 1),) + self.row_shape), device="cuda")[: self.rows]

    def _barrier(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.group)

    def build(self, family, zeta, sigma2=1.0, sigma02=0.0, l=1.0, tol=0.0, use_tol=False, scale=1.0, sync=True):
        """One sharded build.  With sync=True the call is bracketed by cross-rank barriers (peer
        buffers ready before, peer stores landed after)."""
        if sync:
            self._barrier()
        desc = BlockDesc(family, self.block, float(zeta), float(sigma2), float(sigma02), float(l), float(tol),
                         int(bool(use_tol)), 0, 0, float(scale), 1.0, 1)
        _lib.check(self.lib.cspbo_block_build_sharded(ctypes.byref(desc), self.pack.handle, self.rank, self.world,
                                                      self._ptrs))
        if sync:
            self._barrier()
        return self.local

    def gather_full(self):
        """All-gather the row blocks (NCCL) and undo the packed order: the dense [3m,3m] (or [m,m])
        matrix in the caller's group order on every rank, e.g. for a replicated Cholesky."""
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return assemble_full([self.local], self.order, self.m, self.ncomp)
        parts = gather_row_blocks(self.local, self.n_blocks, self.rank, self.world, self.group)
        return assemble_full(parts, self.order, self.m, self.ncomp)

    def close(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
        self.local = None
        for r, pp in enumerate(self._peers):
            if pp is not None:
                self.lib.cspbo_ipc_close(pp)
                self._peers[r] = None
        if self._own is not None:
            self.lib.cspbo_dev_free(self._own)
            self._own = None

    def __del__(self):
        try:
            if self._own is not None:
                self.close()
        except Exception:
            pass
