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
import os

import numpy as np

from . import _lib
from ._lib import BlockDesc, c_vp


# ----------------------------------------------------------------------------- pure host logic
def superblock_owner(s, world, zigzag=False):
    """Owner of superblock s: s % world, mirrored on odd cycles when zigzag (balances triangular work)."""
    cyc, pos = divmod(s, world)
    return world - 1 - pos if (zigzag and cyc & 1) else pos


def owned_superblocks(n_blocks, rank, world, sb_blocks=1, zigzag=False):
    nsb = (n_blocks + sb_blocks - 1) // sb_blocks
    return [s for s in range(nsb) if superblock_owner(s, world, zigzag) == rank]


def owned_blocks(n_blocks, rank, world, sb_blocks=1, zigzag=False):
    """Block ids owned by `rank`, in the order of its buffer (block-cyclic by superblocks of `sb_blocks`)."""
    return [b for s in owned_superblocks(n_blocks, rank, world, sb_blocks, zigzag)
            for b in range(s * sb_blocks, min((s + 1) * sb_blocks, n_blocks))]


def sharded_rows(n_blocks, rank, world, sb_blocks=1, zigzag=False):
    """Number of group rows (incl. padding slots and the ragged tail of the last superblock) in `rank`'s buffer."""
    return 8 * sb_blocks * len(owned_superblocks(n_blocks, rank, world, sb_blocks, zigzag))


def local_positions(n_blocks, rank, world, sb_blocks=1, zigzag=False):
    """Packed positions (block*8+slot) of the rows of `rank`'s buffer, in buffer order; -1 for the rows
    of blocks past the end of a ragged last superblock."""
    pos = []
    for s in owned_superblocks(n_blocks, rank, world, sb_blocks, zigzag):
        for b in range(s * sb_blocks, (s + 1) * sb_blocks):
            pos.extend(range(b * 8, b * 8 + 8) if b < n_blocks else [-1] * 8)
    return np.asarray(pos, dtype=np.int64)


def balanced_superblock_perm(weights, world, zigzag=False):
    """Work-balanced ownership for the sharded symmetric build.  The dealing formula (superblock_owner) is kept but applied
    to VIRTUAL indices: inside every cycle of `world` consecutive superblocks the heaviest one goes to the rank with the
    least work so far, and so on.  Returns (perm, inv): perm[v] = real superblock at virtual index v (-1 = none, length
    cycles * world), inv[s] = virtual index of superblock s.  Only superblocks of one cycle change places, so slab k of all
    ranks still covers a contiguous range of packed row blocks (what build_to_host relies on)."""
    w = np.asarray(weights, dtype=np.float64)
    nsb = len(w)
    ncyc = (nsb + world - 1) // world
    perm = np.full(ncyc * world, -1, dtype=np.int32)
    inv = np.zeros(max(nsb, 1), dtype=np.int32)
    load = np.zeros(world)
    for c in range(ncyc):
        sbs = sorted(range(c * world, min((c + 1) * world, nsb)), key=lambda t: (-w[t], t))
        free = list(range(world))
        for t in sbs:
            r = min(free, key=lambda q: (load[q], q))
            free.remove(r)
            load[r] += w[t]
            pos = world - 1 - r if (zigzag and c & 1) else r
            perm[c * world + pos] = t
            inv[t] = c * world + pos
    return perm, inv


def assemble_full(parts, order, m, ncomp=3, sb_blocks=1, zigzag=False, packed_cols=False, positions=None):
    """Rebuild the dense matrix in the caller's group order from per-rank row blocks.

    parts[r]: array [sharded_rows(r), ncomp, m*ncomp] (K_FF) or [sharded_rows(r), m] (K_EE, ncomp=1)
    order:    packed position -> group id (-1 = padding slot), length 8*n_blocks
    packed_cols: the columns of `parts` are in packed order too and are un-permuted as well.
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
        pos = local_positions(n_blocks, r, world, sb_blocks, zigzag) if positions is None else np.asarray(positions[r])
        g = np.where(pos >= 0, order[np.maximum(pos, 0)], -1)
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
    if packed_cols:
        # column (p, l) of `full` belongs to group order[p]: gather so that column (g, l) comes first
        inv = np.empty(m, dtype=np.int64)
        inv[order[:m]] = np.arange(m)
        if is_torch:
            import torch
            invt = torch.as_tensor(inv, device=full.device, dtype=torch.long)
            full = full.reshape(full.shape[:-1] + (m, ncomp)).index_select(full.dim() - 1, invt).reshape(full.shape)
        else:
            full = full.reshape(full.shape[:-1] + (m, ncomp))[..., inv, :].reshape(full.shape)
    if ncomp == 1:
        return full.reshape(m, m)
    return full.reshape(m * ncomp, m * ncomp)


def _exchange_bytes(blob, group, world):
    """Every rank's `blob` (bytes) on every rank.  Control-plane only (IPC handles); works on any backend, so the
    peer-memory paths also run with two processes on ONE GPU over gloo (the single-GPU test mode of
    tests/test_sharded_gpu.py -- NCCL refuses two ranks on one device)."""
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, bytes(blob), group=group)
    return out


def _open_peers(lib, own, group, world, rank):
    """Export `own` (a cudaMalloc'd pointer), open every peer's buffer through CUDA IPC.
    Returns (ctypes array of the world pointers valid in this process, list of opened peer pointers)."""
    ptrs = (c_vp * world)()
    ptrs[rank] = own
    peers = [None] * world
    if world > 1:
        handle = (ctypes.c_ubyte * 64)()
        _lib.check(lib.cspbo_ipc_export(own, ctypes.cast(handle, c_vp)))
        for r, blob in enumerate(_exchange_bytes(bytes(handle), group, world)):
            if r == rank:
                continue
            hb = (ctypes.c_ubyte * 64)(*blob)
            pp = c_vp()
            _lib.check(lib.cspbo_ipc_open(ctypes.cast(hb, c_vp), ctypes.byref(pp)))
            peers[r] = pp
            ptrs[r] = pp
    return ptrs, peers


def gather_row_blocks(local, n_blocks, rank, world, group=None, sb_blocks=1, zigzag=False, rows=None):
    """all_gather of the (unequal) per-rank row blocks; device-agnostic (NCCL on CUDA tensors, gloo on
    CPU tensors; CUDA tensors under gloo are staged through the host).  Returns the list of per-rank arrays,
    trimmed to their true row counts."""
    import torch
    import torch.distributed as dist
    rows = [sharded_rows(n_blocks, r, world, sb_blocks, zigzag) for r in range(world)] if rows is None else list(rows)
    max_rows = max(rows)
    via_host = local.is_cuda and dist.get_backend(group) != "nccl"
    dev = "cpu" if via_host else local.device
    pad = torch.zeros((max_rows,) + tuple(local.shape[1:]), dtype=local.dtype, device=dev)
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    if via_host:
        parts = [p.to(local.device) for p in parts]
    return [p[: rows[r]] for r, p in enumerate(parts)]


# ----------------------------------------------------------------------------- device side
class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can view a cudaMalloc'd buffer."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class ShardedSymmetricBlock:
    """K_FF(X,X) or K_EE(X,X) of one packed training set, distributed over the ranks of `group`."""

    def __init__(self, pack, block=_lib.KFF, group=None, sb_blocks=1, zigzag=False, packed_cols=False, balance=False):
        """sb_blocks / zigzag / packed_cols: the ownership and column order of cspbo_block_build_sharded_ex
        (include/cspbo_b200.h); the defaults are the plain block-cyclic layout, ShardedCholesky.from_block
        needs packed_cols=True and wide superblocks (the panels of the factorisation).  balance=True deals the
        superblocks by their actual work (tile pairs, cspbo_pack_block_work) instead of by position: see
        balanced_superblock_perm (ShardedCholesky.from_block follows the same ownership)."""
        import torch
        import torch.distributed as dist
        self.lib = _lib.load()
        self.pack, self.block, self.group = pack, block, group
        self.sb_blocks, self.zigzag, self.packed_cols = int(sb_blocks), bool(zigzag), bool(packed_cols)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.m = pack.m
        self.n_blocks = (self.m + 7) // 8
        self.ncomp = 3 if block == _lib.KFF else 1
        self.ncols = self.m * self.ncomp
        # packed columns: pad the row pitch to 16 doubles so that the kernel's staged full-row stores (local and NVLink
        # peer) are whole 32-byte sectors / 128-byte lines; caller-order columns keep the dense pitch
        self.ld = ((self.ncols + 15) // 16) * 16 if self.packed_cols else self.ncols
        self.row_shape = (3, self.ld) if block == _lib.KFF else (self.ld,)
        self.balance = bool(balance) and self.world > 1
        self._perm = self._inv = self._perm_dev = self._inv_dev = None
        if self.balance:
            work = np.zeros(self.n_blocks)
            _lib.check(self.lib.cspbo_pack_block_work(pack.handle, c_vp(work.ctypes.data), len(work)))
            nsb = (self.n_blocks + self.sb_blocks - 1) // self.sb_blocks
            wsb = np.add.reduceat(work, np.arange(0, self.n_blocks, self.sb_blocks))[:nsb]
            self._perm, self._inv = balanced_superblock_perm(wsb, self.world, self.zigzag)
            self._perm_dev = torch.as_tensor(self._perm, device="cuda")
            self._inv_dev = torch.as_tensor(self._inv, device="cuda")
            self.rows = 8 * self.sb_blocks * (len(self._perm) // self.world)      # every cycle has a slot on every rank
        else:
            self.rows = int(self.lib.cspbo_sharded_rows_ex(pack.handle, self.rank, self.world, self.sb_blocks, int(self.zigzag)))
            assert self.rows == sharded_rows(self.n_blocks, self.rank, self.world, self.sb_blocks, self.zigzag)
        order = np.empty(self.n_blocks * 8, dtype=np.int32)
        _lib.check(self.lib.cspbo_pack_order(pack.handle, c_vp(order.ctypes.data), len(order)))
        self.order = order
        nbytes = 8 * self.rows * int(np.prod(self.row_shape))
        ptr = c_vp()
        _lib.check(self.lib.cspbo_dev_alloc(nbytes, ctypes.byref(ptr)))
        self._own = ptr
        self._ptrs, self._peers = _open_peers(self.lib, ptr, group, self.world, self.rank)
        self._full = torch.as_tensor(_RawCuda(ptr.value, (max(self.rows, 1),) + self.row_shape), device="cuda")[: self.rows]
        self.local = self._full[..., : self.ncols]        # [rows, 3, 3m] / [rows, m] view of the pitched buffer

    def positions(self, rank=None):
        """Packed positions (block*8 + slot) of the group rows of `rank`'s buffer, in buffer order; -1 where a slot holds no
        group row (ragged tail, or a cycle that gave this rank no superblock)."""
        rank = self.rank if rank is None else rank
        if not self.balance:
            return local_positions(self.n_blocks, rank, self.world, self.sb_blocks, self.zigzag)
        pos = []
        W, sb = self.world, self.sb_blocks
        for c in range(len(self._perm) // W):
            p = W - 1 - rank if (self.zigzag and c & 1) else rank
            t = int(self._perm[c * W + p])
            for b in range(t * sb, (t + 1) * sb):
                pos.extend(range(b * 8, b * 8 + 8) if (t >= 0 and b < self.n_blocks) else [-1] * 8)
        return np.asarray(pos, dtype=np.int64)

    def _perm_ptrs(self):
        if not self.balance:
            return None, None
        return c_vp(self._perm_dev.data_ptr()), c_vp(self._inv_dev.data_ptr())

    def local2d(self):
        """The local rows as a [rows * ncomp, N] matrix view (row pitch self.ld)."""
        return self._full.reshape(self.rows * self.ncomp, self.ld)[:, : self.ncols]

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
        pp, pi = self._perm_ptrs()
        _lib.check(self.lib.cspbo_block_build_sharded_slab(ctypes.byref(desc), self.pack.handle, self.rank, self.world,
                                                           self._ptrs, self.sb_blocks, int(self.zigzag), int(self.packed_cols),
                                                           self.ld, 0, -1, None, pp, pi))
        if sync:
            self._barrier()
        return self.local

    # 16 equal slabs measured best of (8 graded, 16 equal, 4, 12 front-loaded) at N = 2: 75.9 ms against 79-83 (tools/e2e_n_probe.py);
    # the symmetric build finishes its first rows last-but-cheapest, so a host-bound build cannot beat
    # T (1 - u^2) + u D with u = min(1, D / 2T) (T: build, D: copy-back time): 64 ms at N = 2, D itself at N = 8
    SLAB_FRACTIONS = tuple(i / 16 for i in range(17))

    def slab_slots(self, fractions=None):
        """Block-slot boundaries of the row slabs of build_to_host: the same on every rank and aligned to superblock
        cycles, so that slab k of all ranks together covers a contiguous range of packed row blocks."""
        fractions = self.SLAB_FRACTIONS if fractions is None else fractions
        cycles = (self.n_blocks + self.sb_blocks * self.world - 1) // (self.sb_blocks * self.world)
        cuts = sorted(set(int(round(f * cycles)) for f in fractions) | {0, cycles})
        return [c * self.sb_blocks for c in cuts]

    def build_to_host(self, family, zeta, out_host, sigma2=1.0, sigma02=0.0, l=1.0, tol=0.0, use_tol=False, scale=1.0,
                      fractions=None):
        """Host-bound sharded build: this rank's rows [rows, ncomp, ncols] (packed row order) land in the PINNED CPU
        tensor `out_host` (dense).  The local block slots are cut into slabs that every rank builds in ascending order
        on a compute stream; once all ranks have finished slab k (a 1-element all-reduce on a side stream -- mirrored
        tiles come from peers) its rows are final and their copy-back overlaps the slabs still being computed.  The
        reference instead Allreduces the whole matrix to every rank's host memory (dot_kernel.py:242-247)."""
        import torch
        import torch.distributed as dist
        assert out_host.is_pinned() and out_host.dtype == torch.float64 and out_host.numel() == self.rows * self.ncomp * self.ncols
        desc = BlockDesc(family, self.block, float(zeta), float(sigma2), float(sigma02), float(l), float(tol),
                         int(bool(use_tol)), 0, 0, float(scale), 1.0, 1)
        if not hasattr(self, "_s_build"):
            self._s_build, self._s_copy = torch.cuda.Stream(), torch.cuda.Stream()
            self._flag = torch.zeros(1, dtype=torch.float32, device="cuda")
        sb, sc = self._s_build, self._s_copy
        cur = torch.cuda.current_stream()
        sb.wait_stream(cur)
        sc.wait_stream(cur)
        cuts = self.slab_slots(fractions)
        my_slots = self.rows // 8
        rowb = self.ncomp * self.ld * 8                     # bytes per group row of the pitched device buffer
        width = self.ncols * 8
        pp, pi = self._perm_ptrs()
        for s0, s1 in zip(cuts[:-1], cuts[1:]):
            _lib.check(self.lib.cspbo_block_build_sharded_slab(ctypes.byref(desc), self.pack.handle, self.rank, self.world,
                                                               self._ptrs, self.sb_blocks, int(self.zigzag), int(self.packed_cols),
                                                               self.ld, int(s0), int(s1), c_vp(sb.cuda_stream), pp, pi))
            ev = torch.cuda.Event()
            ev.record(sb)
            with torch.cuda.stream(sc):
                sc.wait_event(ev)
                if self.world > 1:
                    w = dist.all_reduce(self._flag, group=self.group, async_op=True)   # every rank has finished slab k
                    w.wait()
                r0, r1 = min(s0, my_slots) * 8, min(s1, my_slots) * 8              # group rows of this slab held here
                if r1 > r0:
                    _lib.check(self.lib.cspbo_copy_rows_to_host(
                        c_vp(out_host.data_ptr() + r0 * self.ncomp * width), width,
                        c_vp(self._own.value + r0 * rowb), self.ld * 8, width, (r1 - r0) * self.ncomp, c_vp(sc.cuda_stream)))
        cur.wait_stream(sb)
        cur.wait_stream(sc)
        return out_host

    def gather_full(self):
        """All-gather the row blocks (NCCL) and undo the packed order: the dense [3m,3m] (or [m,m])
        matrix in the caller's group order on every rank, e.g. for a replicated Cholesky."""
        import torch
        import torch.distributed as dist
        lay = dict(sb_blocks=self.sb_blocks, zigzag=self.zigzag)
        if self.world == 1:
            return assemble_full([self.local], self.order, self.m, self.ncomp, packed_cols=self.packed_cols, **lay)
        allpos = [self.positions(r) for r in range(self.world)]
        parts = gather_row_blocks(self.local, self.n_blocks, self.rank, self.world, self.group, rows=[len(q) for q in allpos], **lay)
        return assemble_full(parts, self.order, self.m, self.ncomp, packed_cols=self.packed_cols, positions=allpos, **lay)

    def close(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
        self.local = self._full = None
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


# ----------------------------------------------------------------------------- sharded Cholesky
def _view(t):
    """(pointer, leading dimension) of a 2-D row-major view with unit column stride."""
    assert t.dim() == 2 and (t.shape[1] <= 1 or t.stride(1) == 1), "row-major view expected"
    return c_vp(t.data_ptr()), int(t.stride(0))


class _DeviceOps:
    """Building blocks of the factorisation on the GPU through the C ABI (cuSOLVER potrf, cuBLAS trsm / trmm / dgemm
    on explicit streams, csrc/gp.cu).  The only implementation the product has; tests/test_distributed_cpu.py
    injects a torch-CPU twin to exercise the rank choreography over gloo without a GPU."""
    cuda = True

    def __init__(self):
        self.lib = _lib.load()

    def diag(self, D, inv, info, stream):
        """D[h,h] -> its upper factor in place; inv[h,h] (contiguous) = inverse of the factor."""
        dp, ld = _view(D)
        _lib.check(self.lib.cspbo_gp_chol_diag(dp, ld, int(D.shape[0]), c_vp(inv.data_ptr()), c_vp(info.data_ptr()), c_vp(stream)))

    def apply_inv(self, inv, src, dst, stream):
        """dst[h,w] = U_ss^-T src[h,w]"""
        sp, lds = _view(src)
        dp, ldd = _view(dst)
        _lib.check(self.lib.cspbo_gp_chol_apply_inv(c_vp(inv.data_ptr()), int(src.shape[0]), sp, lds, int(src.shape[1]), dp, ldd,
                                                    c_vp(stream)))

    def update(self, C, A, B, stream):
        """C[nb,w] -= A[h,nb]^T B[h,w]"""
        cp, ldc = _view(C)
        ap, lda = _view(A)
        bp, ldb = _view(B)
        _lib.check(self.lib.cspbo_gp_chol_update(cp, ldc, int(C.shape[0]), int(C.shape[1]), ap, lda, bp, ldb, int(A.shape[0]),
                                                 c_vp(stream)))

    def sm_target(self, stream, count):
        _lib.check(self.lib.cspbo_gp_stream_sm_target(c_vp(stream), int(count)))


_STREAMS = {}
_NARROW_GROUPS = {}


def _chol_streams():
    """(diagonal, panel, look-ahead, update) CUDA streams of the current device, shared by every ShardedCholesky of the process.
    Preferred: green-context streams from the library (the diagonal chain gets CSPBO_CHOL_DIAG_SMS SMs of its own,
    default 16); fallback: plain priority streams."""
    import torch
    dev = torch.cuda.current_device()
    if dev not in _STREAMS:
        made = None
        want = int(os.environ.get("CSPBO_CHOL_DIAG_SMS", "16"))
        if want > 0:
            lib = _lib.load()
            a, b, c, d = c_vp(), c_vp(), c_vp(), c_vp()
            n1, n2 = ctypes.c_int(0), ctypes.c_int(0)
            rc = lib.cspbo_gp_chol_streams(want, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d), ctypes.byref(n1),
                                           ctypes.byref(n2))
            if rc == 0:
                made = tuple(torch.cuda.ExternalStream(p.value, device=dev) for p in (a, b, c, d)) + ((n1.value, n2.value),)
        if made is None:
            lo, hi = torch.cuda.Stream.priority_range()
            made = (torch.cuda.Stream(priority=hi), torch.cuda.Stream(priority=hi), torch.cuda.Stream(priority=min(hi + 1, lo)),
                    torch.cuda.Stream(priority=lo), None)
        _STREAMS[dev] = made
    return _STREAMS[dev]


def _narrow_group(group, world):
    """A second communicator over the same ranks: the small diagonal-chain broadcasts must not queue behind the
    wide panel broadcasts on one NCCL stream."""
    import torch.distributed as dist
    key = id(group)
    if key not in _NARROW_GROUPS:
        ranks = dist.get_process_group_ranks(group) if group is not None else list(range(world))
        _NARROW_GROUPS[key] = dist.new_group(ranks=ranks)
    return _NARROW_GROUPS[key]


class ShardedCholesky:
    """In-place Cholesky K = U^T U of a symmetric matrix whose rows are distributed over the ranks by
    superblocks (panels) of `nbw` rows -- exactly what ShardedSymmetricBlock(..., packed_cols=True) leaves in
    HBM -- followed by the solves for alpha.  Replaces the reference's replicated scipy `cholesky` /
    `cho_solve` (gaussianprocess.py:116,127: every MPI rank factorises the whole K redundantly).

    Right-looking, one process per GPU, three streams per rank.  The only sequential chain of a blocked Cholesky
    is diagonal block -> diagonal block, so it is kept as short as the data dependencies allow:

      diagonal stream (panel s, owner): potrf of the nbw x nbw diagonal block + its explicit inverse, then only
          the next TWO blocks of the panel rows ("narrow piece", trmm with the inverse); these few MB are broadcast on
          their own communicator and the owner of panel s+1 applies them to its diagonal block and the block
          right of it, which is all the next potrf and the next narrow piece need.
      panel stream: the rest of the panel rows (wide trmm), their broadcast (NCCL over NVLink: the only bulk
          collective), and the look-ahead update of the remaining columns of panel s+1.
      update stream: every rank applies panel s to its own superblocks t >= s+2 (one dgemm per owned superblock,
          triangular: only the columns from the diagonal block on).

    Every rank keeps the broadcast rows, so after the factorisation U is replicated ([N,N], upper triangle valid)
    and the O(N^2) solves, the predictive variance and the LML run locally.
    """
    RING = 4

    def __init__(self, local2d, N, nbw, rank, world, zigzag=False, group=None, ops=None, owner_inv=None):
        """owner_inv: optional inverse of a within-cycle ownership permutation (balanced_superblock_perm): panel s then
        belongs to the rank the dealing formula gives for virtual index owner_inv[s]; its place in the owner's buffer,
        (s // world) * nbw, is the same because the permutation never leaves the cycle."""
        import torch
        self.torch = torch
        self.local, self.N, self.nbw = local2d, int(N), int(nbw)
        self.rank, self.world, self.zigzag, self.group = rank, world, bool(zigzag), group
        self.ops = ops if ops is not None else _DeviceOps()
        self.owner_inv = None if owner_inv is None else [int(v) for v in owner_inv]
        self.S = (self.N + self.nbw - 1) // self.nbw
        self.max_chunk = 4      # neighbouring superblocks updated by one dgemm (world 1 only)
        self.bulk_sms = int(os.environ.get("CSPBO_BULK_SMS", "0"))
        self.depth = max(1, int(os.environ.get("CSPBO_CHOL_DEPTH", "3")))    # panels of look-ahead
        # trailing updates two panels at a time (k = 2 nbw): pays where the factorisation is bound by the bulk dgemms
        # (world <= 4); at world 8 the diagonal chain is the limit and the one-panel delay of the bulk would only hurt
        pair = os.environ.get("CSPBO_CHOL_PAIR")
        self.pair_updates = ((world <= 4) if pair is None else bool(int(pair))) and self.depth >= 2
        self.cuda = bool(getattr(self.ops, "cuda", False))
        dev = local2d.device
        self.U = torch.empty((self.N, self.N), dtype=torch.float64, device=dev)
        self.info = torch.zeros(max(self.S, 1), dtype=torch.int32, device=dev)
        self.inv = torch.zeros((self.RING, self.nbw * self.nbw), dtype=torch.float64, device=dev)
        self.stage = torch.zeros((self.RING, self.nbw * 3 * self.nbw), dtype=torch.float64, device=dev)
        if self.cuda:
            self.s_diag, self.s_panel, self.s_look, self.s_update, self.sm_split = _chol_streams()
        self._src = [torch.distributed.get_global_rank(group, r) if (group is not None and world > 1) else r
                     for r in range(world)]
        self.g_narrow = _narrow_group(group, world) if world > 1 else None
        self.trace = None

    @classmethod
    def from_block(cls, sh, ops=None):
        """Factorise the matrix a ShardedSymmetricBlock(packed_cols=True) has built, in place."""
        if not sh.packed_cols:
            raise ValueError("ShardedCholesky needs a block built with packed_cols=True")
        N = sh.m * sh.ncomp
        return cls(sh.local2d(), N, 8 * sh.sb_blocks * sh.ncomp, sh.rank, sh.world, sh.zigzag, sh.group, ops,
                   owner_inv=sh._inv if sh.balance else None)

    def _owner(self, s):
        if self.owner_inv is not None:
            assert self.owner_inv[s] // self.world == s // self.world
            s = self.owner_inv[s]
        return superblock_owner(s, self.world, self.zigzag)

    def add_noise(self, noise2):
        """K[i,i] += noise2 on the rows this rank owns (gaussianprocess.py:109-113)."""
        for s in range(self.S):
            if self._owner(s) != self.rank:
                continue
            r0, l0 = s * self.nbw, (s // self.world) * self.nbw
            h = min(self.nbw, self.N - r0)
            self.local[l0:l0 + h, r0:r0 + h].diagonal().add_(float(noise2))

    def factor(self):
        torch, ops, dist = self.torch, self.ops, self.torch.distributed
        N, nbw, S, W, me, R = self.N, self.nbw, self.S, self.world, self.rank, self.RING
        U, loc, cuda = self.U, self.local, self.cuda
        D = self.depth
        sd = self.s_diag.cuda_stream if cuda else None
        sp = self.s_panel.cuda_stream if cuda else None
        sl = self.s_look.cuda_stream if cuda else None
        su = self.s_update.cuda_stream if cuda else None
        if cuda:
            cur = torch.cuda.current_stream()
            for st in (self.s_diag, self.s_panel, self.s_look, self.s_update):
                st.wait_stream(cur)
            if self.bulk_sms:
                ops.sm_target(su, self.bulk_sms)

        class _Null:
            def __enter__(self): return self
            def __exit__(self, *a): return False
        on = (lambda st: torch.cuda.stream(st)) if cuda else (lambda st: _Null())
        SD, SP, SL, SU = (self.s_diag, self.s_panel, self.s_look, self.s_update) if cuda else (None, None, None, None)

        def record(st):
            if not cuda:
                return None
            e = torch.cuda.Event()
            e.record(st)
            return e

        def wait(st, ev):
            if cuda and ev is not None:
                st.wait_event(ev)

        trace = self.trace = [] if (cuda and os.environ.get("CSPBO_CHOL_TRACE")) else None
        if trace is not None:
            self.t_start = torch.cuda.Event(enable_timing=True)
            self.t_start.record(torch.cuda.current_stream())

        def mark(st):
            if trace is None:
                return None
            e = torch.cuda.Event(enable_timing=True)
            e.record(st)
            return e

        import time as _time
        host_t0 = _time.perf_counter()
        mine = [s for s in range(S) if self._owner(s) == me]
        # e_first[k]: the first bulk dgemm of panel k is done (rows k+D+1 carry panel k); e_sl[t]: the latest look-ahead
        # update of rows t is done; e_look[k]: rows k+1 carry panel k in the columns right of the narrow piece;
        # e_wtrmm[k]: the wide trmm of panel k has consumed its inverse slot
        e_first, e_look, e_wtrmm, e_sl = {}, {}, {}, {}
        bounds = lambda s: (s * nbw, min((s + 1) * nbw, N), min((s + 2) * nbw, N), min((s + 3) * nbw, N))

        for s in range(S):
            o = self._owner(s)
            own = me == o
            r0, r1, r2, r3 = bounds(s)
            h, wn = r1 - r0, r3 - r1                   # panel height, width of the narrow piece (next two blocks)
            l0 = (s // W) * nbw                        # first local row of panel s on its owner
            slot = s % R
            msg = self.stage[slot][: h * (h + wn)].view(h, h + wn)     # [diag block | narrow piece], contiguous
            inv = self.inv[slot][: h * h].view(h, h)
            nxt = s + 1 < S and self._owner(s + 1) == me
            lt = ((s + 1) // W) * nbw
            # ------------------------------------------------------------------ diagonal chain
            with on(SD):
                t0 = mark(SD)
                if own:
                    wait(SD, e_wtrmm.get(s - R))      # the wide trmm that used this inverse slot is done
                    # (every earlier panel has reached [diag | next block]: the narrow update below, in stream order,
                    # waited for the bulk and look-ahead updates of these rows)
                    t0 = mark(SD)
                    ops.diag(loc[l0:l0 + h, r0:r1], inv, self.info[s:s + 1], sd)
                    msg[:, :h].copy_(loc[l0:l0 + h, r0:r1], non_blocking=True)
                    ta = mark(SD)
                    if wn:
                        wait(SD, e_look.get(s - 1))   # panel s-1 has reached the second block through the look-ahead
                        tb = mark(SD)
                        ops.apply_inv(inv, loc[l0:l0 + h, r1:r3], msg[:, h:], sd)
                        if trace is not None:
                            trace.append(("wait_look", s, ta, tb))
                    U[r0:r1, r0:r3].copy_(msg, non_blocking=True)
                    if trace is not None:
                        trace.append(("potrf+inv", s, t0, ta))
                t1 = mark(SD)
                if W > 1 and wn:
                    dist.broadcast(msg, src=self._src[o], group=self.g_narrow)
                t2 = mark(SD)
                e_diag = record(SD) if own else None
                if nxt:
                    # rows s+1: [diag block | next block] -= P_s[:, cols(s+1)]^T P_s[:, cols(s+1) u cols(s+2)]
                    wait(SD, e_first.get(s - D))      # bulk updates (panels <= s-D) and look-ahead updates (panels
                    wait(SD, e_sl.get(s + 1))         # s-D+1 .. s-1) have reached rows s+1
                    ops.update(loc[lt:lt + (r2 - r1), r1:r3], msg[:, h:h + (r2 - r1)], msg[:, h:], sd)
                if trace is not None:
                    if own:
                        trace.append(("diag+narrow", s, t0, t1))
                    trace.append(("narrow_bcast" if own else "narrow_wait", s, t1, t2))
            # ------------------------------------------------------------------ wide panel
            with on(SP):
                if own:
                    wait(SP, e_diag)
                    t3 = mark(SP)
                    if r3 < N:
                        ops.apply_inv(inv, loc[l0:l0 + h, r3:N], U[r0:r1, r3:N], sp)
                    e_wtrmm[s] = record(SP)
                    if trace is not None:
                        trace.append(("wide_trmm", s, t3, mark(SP)))
                t4 = mark(SP)
                if W > 1:
                    dist.broadcast(U[r0:r1], src=self._src[o], group=self.group)
                e_wide = record(SP)
                if trace is not None:
                    trace.append(("wide_bcast" if own else "wide_wait", s, t4, mark(SP)))
                if nxt:
                    wait(SP, e_first.get(s - D))
                    wait(SP, e_sl.get(s + 1))
                    if r3 < N:
                        t5 = mark(SP)
                        ops.update(loc[lt:lt + (r2 - r1), r3:N], U[r0:r1, r1:r2], U[r0:r1, r3:N], sp)
                        if trace is not None:
                            trace.append(("lookahead", s, t5, mark(SP)))
                    e_look[s] = record(SP)
            # ------------------------------------------------------------------ bulk updates of this rank's rows
            # (issued before the deeper look-ahead below: with paired updates the look-ahead of THIS iteration may wait for an
            # e_first event that is recorded here)
            with on(SU):
                wait(SU, e_wide)

                def bulk(todo, p0, p1, key):
                    """rows of the superblocks `todo` -= U[p0:p1, cols(t)]^T U[p0:p1, cols(t):]; records e_first[key] after the
                    first dgemm (or at once when there is nothing to do)"""
                    i, first = 0, True
                    while i < len(todo):
                        # superblocks that are neighbours in the matrix AND in this rank's buffer (world 1, zigzag turns)
                        # share one dgemm: the few extra nbw x nbw blocks left of their diagonals are never read again
                        j = i + 1
                        while j < len(todo) and j - i < self.max_chunk and todo[j] == todo[j - 1] + 1 and \
                                todo[j] // W == todo[j - 1] // W + 1:
                            j += 1
                        t = todo[i]
                        lr, c0 = (t // W) * nbw, t * nbw
                        ht = min((todo[j - 1] + 1) * nbw, N) - c0
                        ops.update(loc[lr:lr + ht, c0:N], U[p0:p1, c0:c0 + ht], U[p0:p1, c0:N], su)
                        if first:
                            e_first[key] = record(SU)
                            first = False
                        i = j
                    if first:
                        e_first[key] = record(SU)

                if not self.pair_updates:
                    bulk([t for t in mine if t >= s + D + 1], r0, r1, s)      # rows s+D+1 (if mine) now carry panel s
                elif s % 2 == 0 and s + 1 < S:
                    pass    # first panel of a pair: its bulk update waits for its partner (one dgemm with k = 2 nbw)
                elif s % 2 == 1:
                    # panel s-1 alone on the one superblock that needs it before the pair reaches it (rows s+D), then both
                    # panels together on everything below: k = 2 nbw dgemms run at 34 instead of 28-29 TFLOP/s
                    pa0 = (s - 1) * nbw
                    bulk([t for t in mine if t == s + D], pa0, r0, s - 1)
                    bulk([t for t in mine if t >= s + D + 1], pa0, r1, s)
                else:
                    bulk([t for t in mine if t >= s + D + 1], r0, r1, s)      # unpaired last panel
            # ------------------------------------------------------------------ deeper look-ahead: rows s+2 .. s+D
            with on(SL):
                for d in range(2, D + 1):
                    t = s + d
                    if t >= S or self._owner(t) != me:
                        continue
                    wait(SL, e_wide)
                    wait(SL, e_first.get(t - D - 1))  # the bulk updates of rows t are over
                    lr, c0 = (t // W) * nbw, t * nbw
                    ht = min(c0 + nbw, N) - c0
                    ops.update(loc[lr:lr + ht, c0:N], U[r0:r1, c0:c0 + ht], U[r0:r1, c0:N], sl)
                    e_sl[t] = record(SL)
            for d in (e_first, e_look, e_wtrmm, e_sl):
                d.pop(s - R - D - 2, None)
        self.host_ms = (_time.perf_counter() - host_t0) * 1e3     # host time spent issuing the whole factorisation
        if cuda:
            cur = torch.cuda.current_stream()
            for st in (self.s_diag, self.s_panel, self.s_look, self.s_update):
                cur.wait_stream(st)
        return self

    def trace_summary(self):
        """Per-phase device time (CSPBO_CHOL_TRACE=1), ms summed over panels."""
        out = {}
        for name, s, a, b in (self.trace or []):
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out

    def trace_timeline(self, panels):
        """[(panel, phase, begin_ms, end_ms)] relative to the start of factor() for the given panels."""
        return [(s, name, round(self.t_start.elapsed_time(a), 3), round(self.t_start.elapsed_time(b), 3))
                for name, s, a, b in (self.trace or []) if s in panels]

    def check(self):
        """Raise CspboError(CSPBO_E_NOTPD) if a diagonal block was not positive definite (one host sync)."""
        bad = self.torch.nonzero(self.info > 0)
        if self.world > 1:
            flag = self.torch.tensor([float(bad.numel() > 0)], dtype=self.torch.float64, device=self.local.device)
            self.torch.distributed.all_reduce(flag, op=self.torch.distributed.ReduceOp.MAX, group=self.group)
            failed = flag.item() > 0
        else:
            failed = bad.numel() > 0
        if failed:
            raise _lib.CspboError(5, "K is not positive definite (sharded Cholesky, panel %s)"
                                  % (int(bad[0]) if bad.numel() else "on another rank"))

    def logdet(self):
        return float(2.0 * self.torch.log(self.U.diagonal()).sum())

    def solve(self, B):
        """(U^T U)^-1 B with the replicated factor; B: [N] or [N, nrhs] in the matrix's (packed) order."""
        from . import gp
        Bt = B.reshape(self.N, -1)
        if not self.cuda:
            return self.ops.solve(self.U, Bt).reshape(B.shape)
        return gp.cho_solve(self.U, Bt).reshape(B.shape)


def fit_alpha_sharded(sh, y, noise2, ops=None, return_solver=False):
    """alpha = (K + noise2*I)^-1 y for the K a ShardedSymmetricBlock(packed_cols=True) has just built.
    y and alpha are in the caller's row order (group-major, ncomp rows per group); every rank gets alpha."""
    import torch
    chol = ShardedCholesky.from_block(sh, ops=ops)
    chol.add_noise(noise2)
    chol.factor()
    chol.check()
    order = torch.as_tensor(np.asarray(sh.order[: sh.m], dtype=np.int64), device=chol.U.device)
    yt = torch.as_tensor(np.asarray(y, dtype=np.float64) if not isinstance(y, torch.Tensor) else y,
                         dtype=torch.float64, device=chol.U.device).reshape(sh.m, sh.ncomp)
    ap = chol.solve(yt[order].reshape(-1))
    alpha = torch.empty_like(yt)
    alpha[order] = ap.reshape(sh.m, sh.ncomp)
    alpha = alpha.reshape(-1)
    return (alpha, chol) if return_solver else alpha


# ----------------------------------------------------------------------------- peer-mapped buffers
_IPC_POOL = {}


def _shared_buffer(nbytes, group, world, rank):
    """A device buffer of >= nbytes on every rank, each mapped into every other rank's address space (CUDA IPC):
    (own pointer, ctypes array of the world pointers valid in THIS process).  Grow-only and cached per process group,
    because cudaIpcOpenMemHandle costs milliseconds per peer (80 ms of set-up per fit at 8 GPUs when done every time).
    `nbytes` must be the same on all ranks."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    key = (id(group), world)
    ent = _IPC_POOL.get(key)
    if ent is not None and ent["bytes"] >= nbytes:
        return ent["own"], ent["ptrs"]
    if ent is not None:
        _release_shared(ent, group, world)
    want = int(nbytes) + int(nbytes) // 8
    own = c_vp()
    _lib.check(lib.cspbo_dev_alloc(want, ctypes.byref(own)))
    ptrs, peers = _open_peers(lib, own, group, world, rank)
    _IPC_POOL[key] = dict(bytes=want, own=own, ptrs=ptrs, peers=peers)
    return own, ptrs


def _release_shared(ent, group, world):
    import torch
    lib = _lib.load()
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier(group=group)
    for pp in ent["peers"]:
        if pp is not None:
            lib.cspbo_ipc_close(pp)
    lib.cspbo_dev_free(ent["own"])


def release_shared_buffers():
    """Free the cached peer-mapped buffers (collective: call on every rank)."""
    for (gid, world), ent in list(_IPC_POOL.items()):
        _release_shared(ent, None, world)
    _IPC_POOL.clear()


# ----------------------------------------------------------------------------- whole covariance, sharded
def panel_local_rows(mine, nbw, g0, g1):
    """(local row indices, global row indices) of the rows a rank holds inside the global range [g0, g1), for a rank
    that owns the superblocks `mine` (ascending; the i-th owned superblock is stored at local rows [i*nbw, +nbw))."""
    loc, glob = [], []
    for i, s in enumerate(mine):
        a, b = max(s * nbw, g0), min((s + 1) * nbw, g1)
        if b > a:
            glob.append(np.arange(a, b))
            loc.append(np.arange(a, b) - s * nbw + i * nbw)
    if not loc:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    return np.concatenate(loc), np.concatenate(glob)


def panel_permutation(order_e, order_f, NE_pad):
    """perm[i] = row of the panel-layout matrix that holds row i of the caller's [E; F] ordering (energies, then forces
    atom-major); order_x[p] = group id at packed position p (cspbo_pack_order)."""
    order_e, order_f = np.asarray(order_e, dtype=np.int64), np.asarray(order_f, dtype=np.int64)
    inv_e = np.empty(len(order_e), dtype=np.int64)
    inv_e[order_e] = np.arange(len(order_e))
    inv_f = np.empty(len(order_f), dtype=np.int64)
    inv_f[order_f] = np.arange(len(order_f))
    pf = (NE_pad + inv_f[:, None] * 3 + np.arange(3)[None, :]).reshape(-1)
    return np.concatenate((inv_e, pf))


class ShardedCovariance:
    """[[K_EE, K_EF], [K_FE, K_FF]] (reference base.py:13-14, Dot_mb.py:87-119) of a training set as ONE matrix
    distributed over the ranks in the panel layout of cspbo_block_build_panel, ready for ShardedCholesky:

        global rows = columns:  [ energies in packed order | identity padding up to a panel boundary | forces in
                                  packed order, 3 rows per atom ]
        superblock s = rows [s*nbw, (s+1)*nbw), owner zigzag(s), stored at local rows [(s // world)*nbw, +nbw).

    Every rank packs the same (replicated) descriptors; K_EE and K_FF are built as symmetric sharded blocks, K_EF by
    the owners of the energy rows, each tile also stored transposed (K_FE) into the force rows of its owner over
    NVLink peer memory.  The reference's post-scalings (/N_i N_j, /N_i; dot_kernel.py:45,129-130) are applied to
    the local rows afterwards.
    """

    def __init__(self, kernel, data, group=None, nbw=384, grad_semantics=False):
        """grad_semantics: build the K of `k_total_with_grad` instead of `k_total`'s (kernels/device.py: k_total_device)."""
        import torch
        import torch.distributed as dist
        from .kernels import device as kdev
        self.torch, self.lib, self.group, self.nbw = torch, _lib.load(), group, int(nbw)
        if nbw % 24:
            raise ValueError("nbw must be a multiple of 24")
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.grad_semantics = bool(grad_semantics)
        self.h = kdev._hyper(kernel, quirk=not grad_semantics)
        self.pe, self.pf = kdev._tuples(data)
        self.mE = self.pe.m if self.pe else 0
        self.mF = self.pf.m if self.pf else 0
        self.NE_pad = ((self.mE + nbw - 1) // nbw) * nbw
        self.N = self.NE_pad + 3 * self.mF
        self.S = (self.N + nbw - 1) // nbw
        self.mine = [s for s in range(self.S) if superblock_owner(s, self.world, True) == self.rank]
        self.rows = len(self.mine) * nbw
        self.order_e = self._order(self.pe) if self.pe else np.zeros(0, dtype=np.int64)
        self.order_f = self._order(self.pf) if self.pf else np.zeros(0, dtype=np.int64)
        max_rows = ((self.S + self.world - 1) // self.world) * nbw          # the same on every rank
        self.ld = ((self.N + 15) // 16) * 16       # row pitch: the kernel's full-row stores are whole sectors / lines
        ptr, self._ptrs = _shared_buffer(8 * max(max_rows, 1) * self.ld, group, self.world, self.rank)
        self._own = ptr
        self.local = torch.as_tensor(_RawCuda(ptr.value, (max(self.rows, 1), self.ld)), device="cuda")[: self.rows, : self.N]

    def _order(self, pack):
        nb = (pack.m + 7) // 8
        order = np.empty(nb * 8, dtype=np.int32)
        _lib.check(self.lib.cspbo_pack_order(pack.handle, c_vp(order.ctypes.data), len(order)))
        return order[: pack.m].astype(np.int64)

    def _barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.torch.distributed.barrier(group=self.group)

    def _panel(self, block, zeta, scale, a, b, row0_a, row0_b, tol=0.0, use_tol=False):
        h = self.h
        desc = BlockDesc(h["family"], block, float(zeta), float(h["sigma2"]), float(h["sigma02"]), float(h["l"]), float(tol),
                         int(bool(use_tol)), 0, 0, float(scale), 1.0, 0)
        _lib.check(self.lib.cspbo_block_build_panel(ctypes.byref(desc), a.handle, b.handle, self.rank, self.world, self._ptrs,
                                                    self.ld, self.nbw, int(row0_a), int(row0_b)))

    def _local_rows(self, g0, g1):
        return panel_local_rows(self.mine, self.nbw, g0, g1)

    def build(self, noise_e, f_coef, tol=1e-10, cache=None):
        """Build K + noise into the distributed rows (gaussianprocess.py:105-113).

        cache (a dict the caller keeps across calls on ONE training set; Dot_mb only): the covariance of a Dot kernel is
        K(sigma, sigma0) = sigma^2 (K(1, 0) + sigma0^2 P), P on the energy block only, so this rank's rows of K(1, 0) and
        of P are built once and every later call is a local rescale -- no tile kernels, no peer stores (the sharded twin
        of GaussianProcess._dot_scaled_K; costs N^2 / world more doubles per rank while the cache lives)."""
        torch = self.torch
        if cache is not None and self.h["family"] == _lib.DOT:
            sig = (self.N, self.mE, self.rows, float(self.h["zee"]), float(self.h["zff"]), self.grad_semantics)
            if cache.get("sig") != sig:
                real = self.h
                self.h = dict(real, sigma2=1.0, sigma02=0.0, see=1.0, sef=-1.0, sff=float(real["zff"]))
                self._build_raw(tol)
                K1, P, le = self.local.clone(), None, None
                if self.pe:
                    self.h = dict(self.h, sigma02=1.0)
                    self._build_raw(tol, energy_only=True)
                    le, _ = self._local_rows(0, self.mE)
                    le = torch.as_tensor(le, dtype=torch.long, device=K1.device)
                    P = self.local[le][:, : self.mE] - K1[le][:, : self.mE]
                self.h = real
                cache.clear()
                cache.update(sig=sig, K1=K1, P=P, le=le)
            s2 = float(self.h["see"])                                   # sigma^2
            torch.mul(cache["K1"], s2, out=self.local)
            if cache["P"] is not None and len(cache["le"]):
                rows = self.local[cache["le"]]
                rows[:, : self.mE] += (s2 * float(self.h["sigma02"])) * cache["P"]
                self.local[cache["le"]] = rows
        else:
            self._build_raw(tol)
        # noise on the diagonal, identity on the padding rows
        dev = self.local.device
        for rng, val in (((0, self.mE), float(noise_e) ** 2), ((self.mE, self.NE_pad), 1.0),
                         ((self.NE_pad, self.N), (float(f_coef) * float(noise_e)) ** 2)):
            lr, gr = self._local_rows(*rng)
            if len(lr):
                self.local[torch.as_tensor(lr, device=dev), torch.as_tensor(gr, device=dev)] += val
        return self.local

    def _build_raw(self, tol=1e-10, energy_only=False):
        """The blocks of K (no noise) into the distributed rows, 1/N_i normalisations of the energy rows / columns applied."""
        torch, h = self.torch, self.h
        self.local.zero_()
        self._barrier()                    # every buffer is zeroed before peers store mirrored tiles into it
        if self.pe:
            self._panel(_lib.KEE, h["zee"], h["see"], self.pe, self.pe, 0, 0)
        if self.pe and self.pf and not energy_only:
            self._panel(_lib.KEF, h["zef"], h["sef"], self.pe, self.pf, 0, self.NE_pad)
        if self.pf and not energy_only:
            self._panel(_lib.KFF, h["zff"], h["sff"], self.pf, self.pf, self.NE_pad, self.NE_pad, tol=tol,
                        use_tol=(h["family"] == _lib.RBF and not self.grad_semantics))
        self._barrier()                    # peer stores have landed
        dev = self.local.device
        if self.pe:
            ce = torch.as_tensor(self.pe.counts[self.order_e], dtype=torch.float64, device=dev)      # packed order
            le, ge = self._local_rows(0, self.mE)
            if len(le):
                lt = torch.as_tensor(le, device=dev)
                rows = self.local[lt]
                rows[:, : self.mE] /= ce[None, :]
                rows /= ce[torch.as_tensor(ge, device=dev)][:, None]
                self.local[lt] = rows
            lf, _ = self._local_rows(self.NE_pad, self.N)
            if len(lf):
                lt = torch.as_tensor(lf, device=dev)
                self.local[lt, : self.mE] = self.local[lt, : self.mE] / ce[None, :]

    def permutation(self):
        return panel_permutation(self.order_e, self.order_f, self.NE_pad)

    def close(self):
        """The buffer itself is pooled (_shared_buffer); only drop the view, after every rank is done with it."""
        if self._own is not None:
            self._barrier()
        self.local = None
        self._own = None


class ShardedSolver:
    """K^-1 applied through the replicated factor of a ShardedCholesky, in the caller's [E; F] row order."""

    def __init__(self, chol, perm):
        self.chol = chol
        self.perm = chol.torch.as_tensor(perm, device=chol.U.device)

    def solve(self, B):
        """B: [N] or [N, nrhs] CUDA tensor in the caller's order -> K^-1 B, same shape."""
        torch = self.chol.torch
        B2 = B.reshape(B.shape[0], -1)
        P = torch.zeros((self.chol.N, B2.shape[1]), dtype=torch.float64, device=B.device)
        P[self.perm] = B2
        return self.chol.solve(P)[self.perm].reshape(B.shape)

    def logdet(self):
        return self.chol.logdet()        # the identity padding rows contribute log 1 = 0


def fit_sharded(kernel, data, y, noise_e, f_coef, group=None, nbw=384, tol=1e-10, return_solver=False):
    """The fit of gaussianprocess.py:105-127 (opt=False) on all ranks at once: K + noise built in the panel layout,
    factorised in place by the sharded Cholesky, alpha = K^-1 y returned in the caller's [E; F] order on every rank
    (with return_solver=True also a ShardedSolver for the predictive variance and the LML)."""
    import torch
    cov = ShardedCovariance(kernel, data, group=group, nbw=nbw)
    cov.build(noise_e, f_coef, tol=tol)
    chol = ShardedCholesky(cov.local, cov.N, nbw, cov.rank, cov.world, True, group)
    chol.factor()
    chol.check()
    solver = ShardedSolver(chol, cov.permutation())
    yt = torch.as_tensor(np.asarray(y, dtype=np.float64).reshape(-1) if not isinstance(y, torch.Tensor) else y.reshape(-1),
                         dtype=torch.float64, device=chol.U.device)
    alpha = solver.solve(yt)
    chol.local = None
    cov.close()                          # the distributed rows are no longer needed: U is replicated
    return (alpha, solver) if return_solver else alpha


# ----------------------------------------------------------------------------- sharded LML + gradient
_SLAB_UNIT = 96          # rows: a multiple of the 8-row energy blocks and the 24-row force blocks of the packs
_SLAB_MAX = int(os.environ.get('CSPBO_SLAB_MAX', '2304'))         # columns of one triangular solve (wide right-hand sides run at dgemm-like rates, thin ones do not)
_TRACE_ROWS = 8000.0     # RBF: the dK/dl trace of a slab row costs about as much as this many rows of a triangular solve


def _inverse_slabs(Ntot, NE_pad, rank, world, rbf, nbw=384, max_cols=_SLAB_MAX):
    """Column ranges [j0, j1) of K^-1 that `rank` works out in lml_sharded.  Every rank holds the whole factor after
    ShardedCholesky.factor(), so the inverse work is dealt independently of who factorised what: CONTIGUOUS ranges whose
    cost -- (Ntot - j)^2 per column for the triangular solves against the trailing triangle, plus a term linear in
    (Ntot - j) for the RBF dK/dl traces -- is equal over the ranks, cut into slabs of at most `max_cols` columns that do
    not straddle the energy / force boundary NE_pad.  Boundaries are multiples of 96 rows (of nbw when that is not a
    multiple of 96), i.e. aligned to the packs' blocks for the trace kernels; the last slab may be ragged."""
    unit = _SLAB_UNIT if nbw % _SLAB_UNIT == 0 else nbw
    assert NE_pad % unit == 0 and unit % 24 == 0, (NE_pad, unit)
    nU = (Ntot + unit - 1) // unit
    rem = (Ntot - np.arange(nU, dtype=np.float64) * unit)
    cum = np.concatenate(([0.0], np.cumsum(rem * rem + (_TRACE_ROWS * rem if rbf else 0.0))))
    cut = [int(np.searchsorted(cum, cum[-1] * r / world, side='left')) for r in range(world + 1)]
    cut[0], cut[-1] = 0, nU
    lo, hi = cut[rank] * unit, min(cut[rank + 1] * unit, Ntot)
    out = []
    j0 = lo
    while j0 < hi:
        j1 = min(j0 + max_cols, hi)
        if j0 < NE_pad < j1:
            j1 = NE_pad
        out.append((j0, j1))
        j0 = j1
    return out


class _BlockedSubstitution:
    """Triangular solves of lml_sharded against the trailing triangle U[j0:, j0:] of the factor (Ky = U^T U) with a slab of
    right-hand sides, as blocked substitutions over row blocks of `nbk` rows whose diagonal blocks are inverted
    EXPLICITLY once (gp.tri_inverse_inplace, cached per block): every step is then one large dgemm plus one product
    with a small triangle, which runs about twice as fast as cuBLAS' trsm at these shapes (tools/lml_sharded_probe.py).
    Blocks sit on a global grid (multiples of nbk) so that all slabs of a rank share the inverses; a slab that starts
    inside a block gets a ragged first block of its own."""

    def __init__(self, U, Ntot, nbk=None, group=None, world=1, rank=0):
        self.U, self.N = U, Ntot
        self.nbk = nbk or int(os.environ.get("CSPBO_SUB_BLOCK", "1024"))
        self._inv = {}
        if world > 1:
            self._share_grid_inverses(group, world, rank)

    def _share_grid_inverses(self, group, world, rank):
        """world > 1: the rank whose columns start at 0 needs the inverse of EVERY grid block (20 at N = 20 001: two thirds
        of its share of the substitution work at world = 8), the last rank only a few.  The factor is replicated, so the
        grid blocks are inverted round-robin and summed into place by one all-reduce of a zero-initialised buffer
        (N x nbk doubles; gloo, the one-GPU test mode, has all-reduce for CUDA tensors but no all-gather)."""
        import torch
        import torch.distributed as dist
        nb = (self.N + self.nbk - 1) // self.nbk
        buf = torch.zeros((nb, self.nbk, self.nbk), dtype=self.U.dtype, device=self.U.device)
        for b in range(rank, nb, world):
            a, e = b * self.nbk, min((b + 1) * self.nbk, self.N)
            buf[b, : e - a, : e - a] = self._invert(a, e)
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        for b in range(nb):
            a, e = b * self.nbk, min((b + 1) * self.nbk, self.N)
            self._inv[(a, e)] = buf[b, : e - a, : e - a]

    def _blocks(self, j0):
        out, a = [], j0
        while a < self.N:
            b = min((a // self.nbk + 1) * self.nbk, self.N)
            out.append((a, b))
            a = b
        return out

    def _invert(self, a, b):
        import torch
        from . import gp
        blk = self.U[a:b, a:b].clone(memory_format=torch.contiguous_format)            # a copy even when the slice IS the factor
        return gp.tri_inverse_inplace(blk)                                            # U_kk^-1 (upper)

    def _inverse(self, a, b):
        key = (a, b)
        if key not in self._inv:
            self._inv[key] = self._invert(a, b)
        return self._inv[key]

    def forward(self, j0, hj):
        """Y = (U[j0:, j0:]^T)^-1 E, E = the first hj columns of the identity: [N - j0, hj]."""
        import torch
        U = self.U
        Y = torch.empty((self.N - j0, hj), dtype=U.dtype, device=U.device)
        for a, b in self._blocks(j0):
            if a == j0:
                T = torch.zeros((b - a, hj), dtype=U.dtype, device=U.device)
            else:
                T = torch.mm(U[j0:a, a:b].T, Y[: a - j0])          # L[a:b, j0:a] Y[:a]
                T.neg_()
            if a - j0 < hj:                                        # rows of E inside this block
                n1 = min(b - j0, hj) - (a - j0)
                T[:n1, a - j0: a - j0 + n1].diagonal().add_(1.0)
            torch.mm(self._inverse(a, b).T, T, out=Y[a - j0: b - j0])
        return Y

    def backward(self, j0, Y):
        """Z = U[j0:, j0:]^-1 Y, overwriting Y."""
        import torch
        U = self.U
        for a, b in reversed(self._blocks(j0)):
            T = Y[a - j0: b - j0]
            T = T - torch.mm(U[a:b, b:], Y[b - j0:]) if b < self.N else T.clone()
            torch.mm(self._inverse(a, b), T, out=Y[a - j0: b - j0])
        return Y


def lml_sharded(kernel, data, y, noise_e, f_coef, eval_gradient=False, group=None, nbw=384, cache=None):
    """log marginal likelihood (and its gradient w.r.t. the kernel parameters and the noise) of
    gaussianprocess.py:453-503 with K distributed over the ranks: panel-layout build, sharded Cholesky, and the
    gradient as fused traces --

      * diag(Ky^-1): every rank inverts the factor only against the identity columns of ITS column range (contiguous
        ranges of equal cost, _inverse_slabs; one triangular solve of the trailing triangle per slab of up to 2304
        columns -- wide right-hand sides, which cuBLAS' trsm runs several times faster than one 384-column panel at a
        time), the pieces are summed by one all-reduce of an N-vector;
      * Dot: dK/dsigma = 2K/sigma and the block-constant dK/dsigma0 need nothing else (two scalars from alpha and one
        extra solve for 1_E);
      * RBF: the same slab solves continued to full rows of Ky^-1, contracted with the dK/dl tiles of those rows inside
        the tile-kernel epilogue (cspbo_block_trace_slab), partial traces summed by an all-reduce.

    Neither K_gradient [N,N,3] nor Ky^-1 exists anywhere: per rank one slab's rows of Ky^-1 (<= 2304 x N) at a time.
    cache: a dict kept by the caller across the evaluations of one hyper-optimisation on one training set (Dot_mb only,
    ShardedCovariance.build): the tile kernels then run for the first evaluation only.
    y in the caller's [E; F] order.  Returns MLL or (MLL, grad) -- identical numbers on every rank."""
    import torch
    import torch.distributed as dist
    from .kernels import device as kdev
    cov = ShardedCovariance(kernel, data, group=group, nbw=nbw, grad_semantics=eval_gradient)
    npar = len(kernel.parameters()) + 1
    cov.build(noise_e, f_coef, cache=cache if eval_gradient else None)
    chol = ShardedCholesky(cov.local, cov.N, nbw, cov.rank, cov.world, True, group)
    chol.factor()
    try:
        chol.check()
    except _lib.CspboError as exc:
        cov.close()
        if exc.code == 5:
            return (-np.inf, np.zeros(npar)) if eval_gradient else -np.inf
        raise
    dev = chol.U.device
    perm = torch.as_tensor(cov.permutation(), device=dev)
    yt = torch.as_tensor(np.asarray(y, dtype=np.float64).reshape(-1) if not isinstance(y, torch.Tensor) else y.reshape(-1),
                         dtype=torch.float64, device=dev)
    Nreal, NE, Ntot = yt.numel(), cov.mE, cov.N
    yp = torch.zeros(Ntot, dtype=torch.float64, device=dev)
    yp[perm] = yt
    ap = chol.solve(yp)                                    # alpha in panel order (0 on the identity padding rows)
    MLL = float(-0.5 * torch.dot(yp, ap) - 0.5 * chol.logdet() - Nreal / 2 * np.log(2 * np.pi))
    chol.local = None
    if not eval_gradient:
        cov.close()
        return MLL
    U, W, me = chol.U, cov.world, cov.rank
    ne, nf = float(noise_e), float(f_coef * noise_e)
    is_e = torch.zeros(Ntot, dtype=torch.bool, device=dev)
    is_e[: cov.mE] = True
    is_f = torch.zeros(Ntot, dtype=torch.bool, device=dev)
    is_f[cov.NE_pad:] = True
    is_e, is_f = is_e.to(torch.float64), is_f.to(torch.float64)
    d = is_e * (ne * ne) + is_f * (nf * nf)                                        # noise on the real rows (pads: 0)
    b = is_e * (2 * ne) + is_f * (2 * f_coef ** 2 * ne)
    rbf = kernel.name != 'Dot_mb'
    dinv = torch.zeros(Ntot, dtype=torch.float64, device=dev)
    tr = torch.zeros(1, dtype=torch.float64, device=dev)
    if rbf:
        h = kdev._hyper(kernel, quirk=False)
        l = kernel.l
        fscale = torch.ones(Ntot, dtype=torch.float64, device=dev)       # 1/N_i of the energy rows (rbf_kernel.py:55-60,157)
        if cov.pe:
            fscale[: cov.mE] = 1.0 / torch.as_tensor(cov.pe.counts[cov.order_e], dtype=torch.float64, device=dev)
        aps = ap * fscale
        lib = _lib.load()

        def trace(block, a, bp, row0, col0, slab0, blk0, blk1, scale_grad, wptr, ldw):
            desc = BlockDesc(_lib.RBF, block, float(h["zee"]), float(h["sigma2"]), 0.0, float(l), 0.0, 0, 1, 0, 1.0, float(scale_grad), 0)
            _lib.check(lib.cspbo_block_trace_slab(ctypes.byref(desc), a.handle, bp.handle, c_vp(aps.data_ptr()), c_vp(wptr),
                                                  int(ldw), int(row0), int(col0), int(slab0), int(blk0), int(blk1),
                                                  c_vp(tr.data_ptr()), 1))
    sub = _BlockedSubstitution(U, Ntot, group=group, world=W, rank=me)
    for j0, j1 in _inverse_slabs(Ntot, cov.NE_pad, me, W, rbf, nbw):
        hj = j1 - j0
        # Ky = U^T U: Y = (U^T)^-1 E_J has zeros above row j0, so only the trailing triangle is solved
        Y = sub.forward(j0, hj)
        dinv[j0:j1] = (Y * Y).sum(dim=0)
        if rbf and (j0 < cov.mE or j0 >= cov.NE_pad):
            # rows J of Ky^-1 = (U^-1 Y)^T; only the columns from j0 on are needed (the traces below walk the blocks
            # b >= a, off-diagonal tiles twice), and those depend on the trailing triangle alone
            slab = sub.backward(j0, Y).T.contiguous()      # [hj, Ntot - j0]
            del Y
            slab *= fscale[None, j0:]
            slab *= fscale[j0:j1, None]
            wptr = slab.data_ptr() - 8 * j0             # so that column c of the matrix is slab[:, c - j0]
            ldw = slab.shape[1]
            if j0 < cov.mE:                     # energy rows: E-E (symmetric), and E-F twice (K_FE = K_EF^T contributes the same)
                b0, b1 = j0 // 8, (j1 + 7) // 8
                trace(_lib.KEE, cov.pe, cov.pe, 0, 0, j0, b0, b1, 1.0 / (l * l * l), wptr, ldw)
                if cov.pf:
                    trace(_lib.KEF, cov.pe, cov.pf, 0, cov.NE_pad, j0, b0, b1, 2.0, wptr, ldw)
            else:                               # force rows: F-F (symmetric)
                b0, b1 = (j0 - cov.NE_pad) // 24, (j1 - cov.NE_pad + 23) // 24
                trace(_lib.KFF, cov.pf, cov.pf, cov.NE_pad, cov.NE_pad, j0, b0, b1, 1.0, wptr, ldw)
            del slab
        else:
            del Y
    red = torch.cat((dinv, tr))
    if W > 1:
        dist.all_reduce(red, op=dist.ReduceOp.SUM, group=group)
    dinv, trv = red[:Ntot], float(red[Ntot])
    a2 = ap * ap
    sigma = kernel.sigma
    g_s = (1.0 / sigma) * float(torch.dot(ap, yp) - torch.dot(d, a2) - Nreal + torch.dot(d, dinv))
    g_noise = 0.5 * float(torch.dot(b, a2 - dinv))
    if rbf:
        # sum_ij (alpha alpha^T - Ky^-1)_ij dK/dl_ij: the alpha alpha^T part is in the slab traces already
        grad = np.array([g_s, 0.5 * trv, g_noise])
    else:
        g_s0 = 0.0
        if NE > 0:
            ones = torch.zeros(Ntot, dtype=torch.float64, device=dev)
            ones[: cov.mE] = 1.0
            u = chol.solve(ones)
            c0 = 0.8 * 2 * kernel.sigma ** 2 * kernel.sigma0
            g_s0 = 0.5 * c0 * float(ap[: cov.mE].sum() ** 2 - u[: cov.mE].sum())
        grad = np.array([g_s, g_s0, g_noise])
    cov.close()
    return MLL, grad


# ----------------------------------------------------------------------------- sharded prediction
def group_slice(n, rank, world):
    """[lo, hi) of the n groups (structures / force atoms) handled by `rank`: contiguous, sizes differ by at most 1."""
    return n * rank // world, n * (rank + 1) // world


def slice_groups(data, lo, hi, mode='force'):
    """Groups [lo, hi) of a stacked tuple or of a list, in the same format (cspbo/utilities.py:350-416)."""
    if isinstance(data, list):
        return data[lo:hi]
    counts = np.asarray(data[-1], dtype=np.int64)
    off = np.concatenate(([0], np.cumsum(counts)))
    r0, r1 = int(off[lo]), int(off[hi])
    if mode == 'force':
        X, dXdR, ELE, indices = data
        return (X[r0:r1], dXdR[r0:r1], ELE[r0:r1], list(indices[lo:hi]))
    X, ELE, indices = data
    return (X[r0:r1], ELE[r0:r1], list(indices[lo:hi]))


def _n_groups(d):
    return len(d) if isinstance(d, list) else len(d[-1])


def predict_sharded(model, X, group=None, return_std=False, total_E=False, gather=None):
    """K* . alpha (gaussianprocess.py:133-184) with the TEST groups sharded over the ranks: every rank builds the K*
    rows of its slice of the energy structures and force atoms against the (replicated) training set and the
    predictions are all-gathered -- K* blocks are independent, no other exchange.  Returns what
    model.predict(X, ...) returns, on every rank.  `gather`: test hook replacing the all_gather of a list of arrays."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    nE = _n_groups(X["energy"]) if "energy" in X and len(X["energy"]) > 0 else 0
    nF = _n_groups(X["force"]) if "force" in X and len(X["force"]) > 0 else 0
    e0, e1 = group_slice(nE, rank, world)
    f0, f1 = group_slice(nF, rank, world)
    mine = {}
    if e1 > e0:
        mine["energy"] = slice_groups(X["energy"], e0, e1, mode='energy')
    if f1 > f0:
        mine["force"] = slice_groups(X["force"], f0, f1, mode='force')
    if mine:
        res = model.predict(mine, total_E=total_E, return_std=return_std)
        mean, std = (res if return_std else (res, None))
    else:
        mean, std = np.zeros(0), (np.zeros(0) if return_std else None)

    def all_gather(vec):
        """variable-length float64 vectors -> list of per-rank numpy arrays"""
        if gather is not None:
            return gather(vec)
        if world == 1:
            return [np.asarray(vec)]
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        n = torch.tensor([len(vec)], dtype=torch.int64, device=dev)
        ns = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
        dist.all_gather(ns, n, group=group)
        mx = max(int(v.item()) for v in ns)
        buf = torch.zeros(max(mx, 1), dtype=torch.float64, device=dev)
        buf[: len(vec)] = torch.as_tensor(np.asarray(vec, dtype=np.float64), device=dev)
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf, group=group)
        return [p[: int(k.item())].cpu().numpy() for p, k in zip(parts, ns)]

    def assemble(vec):
        ne = e1 - e0
        E = np.concatenate(all_gather(vec[:ne]))
        F = np.concatenate(all_gather(vec[ne:]))
        return np.concatenate((E, F))

    out = assemble(mean)
    return (out, assemble(std)) if return_std else out
