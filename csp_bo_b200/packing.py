"""Host-side handles on the resident layer of the C ABI: packed descriptor sets and block builds.

A `Pack` is a stacked "energy" tuple (X, ELE, indices) or "force" tuple (X, dXdR, ELE, indices)
(reference cspbo/utilities.py:350-400) normalised, projected and tiled in HBM once
(csrc/pack.cu).  It replaces the per-call `ffi.new('double[n]', list(arr.ravel()))` deep copies of
the reference wrappers (cspbo/kernels/dot_kernel.py:207-216).
"""
import ctypes
import os

import numpy as np

from . import _lib
from ._lib import BlockDesc, MEM_DEVICE, MEM_HOST, c_vp


def _is_torch(a):
    return type(a).__module__.startswith("torch")


def _ptr(a):
    if a is None:
        return None
    if _is_torch(a):
        return c_vp(a.data_ptr())
    return c_vp(a.ctypes.data)


def group_ids(indices):
    """x_inds of the reference wrappers (dot_kernel.py:179-184): row -> group id."""
    counts = np.asarray(indices, dtype=np.int64)
    return np.repeat(np.arange(len(counts), dtype=np.int32), counts)


class Pack:
    """Device-resident, tiled form of one stacked descriptor set."""

    def __init__(self, x, dxdr, ele, indices, row_range=None, norm_shift=0.0, row_classes=1):
        lib = _lib.load()
        on_dev = _is_torch(x)
        if on_dev:
            import torch
            assert x.is_cuda and x.dtype == torch.float64
            x = x.contiguous()
            if dxdr is not None:
                assert dxdr.is_cuda and dxdr.dtype == torch.float64
                dxdr = dxdr.contiguous()
        else:
            x = np.ascontiguousarray(x, dtype=np.float64)
            if dxdr is not None:
                dxdr = np.ascontiguousarray(dxdr, dtype=np.float64)
        n, d = int(x.shape[0]), int(x.shape[1])
        nc = 0 if dxdr is None else int(dxdr.shape[2])
        if dxdr is not None and (int(dxdr.shape[0]) != n or int(dxdr.shape[1]) != d):
            raise ValueError("dxdr shape %s does not match x shape %s" % (tuple(dxdr.shape), tuple(x.shape)))
        ele = np.ascontiguousarray(np.asarray(ele).ravel(), dtype=np.int32)
        self.counts = np.asarray(indices, dtype=np.int64)
        inds = group_ids(self.counts)
        if len(ele) != n or len(inds) != n:
            raise ValueError("ele (%d) / indices (sum %d) do not match the %d stacked rows" % (len(ele), len(inds), n))
        self.m, self.n, self.d, self.nc = len(self.counts), n, d, nc
        r0, r1 = (0, n) if row_range is None else row_range
        h = c_vp()
        _lib.check(lib.cspbo_pack_create_ex(n, d, nc, self.m, int(r0), int(r1), _ptr(x), _ptr(dxdr), _ptr(ele),
                                            _ptr(inds), MEM_DEVICE if on_dev else MEM_HOST, float(norm_shift),
                                            int(row_classes), ctypes.byref(h)))
        self.row_classes = int(row_classes)
        self._h = h
        # spread packs (spread_energy): `spread` consecutive pseudo-groups make one real group
        self.spread, self.m_real, self.real_counts = 1, self.m, self.counts

    def __len__(self):
        """Number of groups, so that the `len(d) > 0` tests of k_total treat a Pack like a stacked tuple."""
        return self.m

    @property
    def handle(self):
        return self._h

    @property
    def nbytes(self):
        return int(_lib.load().cspbo_pack_bytes(self._h))

    def pair_count(self, other):
        return int(_lib.load().cspbo_pack_pair_count(self._h, other._h))

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.load().cspbo_pack_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------
# Packing is STATELESS, like the reference wrappers (cspbo/kernels/dot_kernel.py:207-258 copy their inputs on every
# call): a stacked tuple handed to kee_C / kef_C / kff_C / k_total is uploaded and tiled again each time, so an array
# that was updated in place (MD loops, finite differences, NEB images reusing buffers) can never be served from a
# stale copy.  Residency is explicit: build the packs once with `resident()` (or `Pack(...)`) and pass THEM wherever a
# stacked tuple is accepted -- that is what GaussianProcess does with its training set.
def pack_energy(X, spread_small=False):
    """Stacked energy tuple (X, ELE, indices) -> Pack; a Pack passes through.  spread_small: sets of up to
    SPREAD_MAX_GROUPS structures are packed spread over the tile slots (spread_energy), which build_block understands."""
    if isinstance(X, Pack):
        return X
    x, ele, indices = X
    if spread_small and len(indices) <= SPREAD_MAX_GROUPS:
        return spread_energy(X)
    return Pack(x, None, ele, indices)


HOST_ROW_CLASSES = int(os.environ.get("CSPBO_HOST_ROW_CLASSES", "6"))     # row classes of packs that feed large host-bound K_FF builds (see cspbo_pack_create_ex)


def pack_force(X, row_range=None, row_classes=1):
    """Stacked force tuple (X, dXdR, ELE, indices) -> Pack; a Pack passes through."""
    if isinstance(X, Pack):
        return X
    x, dxdr, ele, indices = X
    return Pack(x, dxdr, ele, indices, row_range=row_range, row_classes=row_classes)


SPREAD_WAYS = 8
SPREAD_MAX_GROUPS = 64      # larger sets: the host-side regrouping (two argsorts + a row gather) costs more than the kernel gains


def spread_energy(X, ways=SPREAD_WAYS):
    """Cut every structure of a stacked energy tuple into `ways` pseudo-structures with (nearly) equal per-species row
    counts.  A covariance block is a plain sum over the rows of a group (dot_kernel.cpp:45-47), so the block of the real
    structure is the sum of its pseudo-structures' rows / columns -- but a pack with ONE structure fills one slot of
    every 8-slot tile (1/8 of the DMMA work useful, one CTA column), while its 8 pseudo-structures fill them all: the
    K_EE / K_EF / K_FE builds of an MD step (one 192-atom structure against the training set) get 8x shorter.
    Returns a Pack whose .spread / .m_real / .real_counts describe the real groups."""
    x, ele, indices = X
    counts = np.asarray(indices, dtype=np.int64)
    ele_np = np.asarray(ele).ravel()
    n = len(ele_np)
    struct = np.repeat(np.arange(len(counts)), counts)
    # rank of every row among the rows of its (structure, species), in row order
    key = struct * (int(ele_np.max(initial=0)) - int(ele_np.min(initial=0)) + 1) + (ele_np - int(ele_np.min(initial=0)))
    order = np.argsort(key, kind='stable')
    sk = key[order]
    first = np.concatenate(([0], np.nonzero(sk[1:] != sk[:-1])[0] + 1)) if n else np.zeros(0, dtype=np.int64)
    rank = np.arange(n) - np.repeat(first, np.diff(np.concatenate((first, [n])))) if n else np.zeros(0, dtype=np.int64)
    pseudo = np.empty(n, dtype=np.int64)
    pseudo[order] = struct[order] * ways + rank % ways
    perm = np.argsort(pseudo, kind='stable')
    new_counts = np.bincount(pseudo, minlength=len(counts) * ways)
    if _is_torch(x):
        import torch
        xs = x[torch.as_tensor(perm, device=x.device)]
    else:
        xs = np.asarray(x)[perm]
    p = Pack(xs, None, ele_np[perm], new_counts)
    p.spread, p.m_real, p.real_counts = ways, len(counts), counts
    return p


def resident(data, stress=False, spread_small=False):
    """{'energy': tuple|list, 'force': tuple|list} -> {'energy': Pack, 'force': Pack}: the caller-held handles that
    keep a descriptor set in HBM across calls.  The packs are snapshots: later changes of the source arrays do not
    reach them (build new ones).  spread_small: energy sets of up to 64 structures become spread packs (spread_energy),
    which the K* assembly of kernels/device.py understands; plain packs are what everything else expects."""
    from .utilities import list_to_tuple
    out = {}
    E = data.get("energy")
    if E is not None and len(E) > 0:
        E = list_to_tuple(E, mode='energy') if isinstance(E, list) else E
        if spread_small and not isinstance(E, Pack) and len(E[2]) <= SPREAD_MAX_GROUPS:
            out["energy"] = spread_energy(E)      # prediction-side packs only: see spread_energy
        else:
            out["energy"] = pack_energy(E)
    F = data.get("force")
    if F is not None and len(F) > 0:
        out["force"] = pack_force(list_to_tuple(F, stress=stress) if isinstance(F, list) else F)
    return out


# ---------------------------------------------------------------------------------------------
def out_shapes(block, m1, m2, stress):
    if block == _lib.KEE:
        return (m1, m2), (m1, m2)
    if block == _lib.KEF:
        return (m1, m2, 9 if stress else 3), (m1, m2, 3)
    return (m1, 9 if stress else 3, m2 * 3), (m1, 3, m2 * 3)


def build_block(family, block, a, b, zeta, sigma2=1.0, sigma02=0.0, l=1.0, tol=0.0, use_tol=False,
                with_grad=False, stress=False, scale=1.0, scale_grad=1.0, symmetric=False,
                out=None, out_grad=None, device=False):
    """cspbo_block_build.  Returns (out, out_grad|None): numpy arrays, or CUDA torch tensors when
    device=True (or when `out` is a CUDA tensor)."""
    lib = _lib.load()
    grad = bool(with_grad) and family == _lib.RBF
    if a.spread > 1 or b.spread > 1:
        return _build_block_spread(family, block, a, b, zeta, sigma2, sigma02, l, tol, use_tol, with_grad, stress, scale,
                                   scale_grad, symmetric, out, out_grad, device)
    shp, shp_g = out_shapes(block, a.m, b.m, stress)
    if out is not None and _is_torch(out) and out.is_cuda:
        device = True
    if device:
        import torch
        if out is None:
            out = torch.empty(shp, dtype=torch.float64, device="cuda")
        if grad and out_grad is None:
            out_grad = torch.empty(shp_g, dtype=torch.float64, device="cuda")
        mem = MEM_DEVICE
    else:
        if out is None:
            out = np.empty(shp, dtype=np.float64)
        if grad and out_grad is None:
            out_grad = np.empty(shp_g, dtype=np.float64)
        mem = MEM_HOST
    n_out = int(np.prod(shp))
    got = out.numel() if _is_torch(out) else out.size
    if got != n_out:
        raise ValueError("out has %d elements, block needs %d" % (got, n_out))
    desc = BlockDesc(family, block, float(zeta), float(sigma2), float(sigma02), float(l), float(tol),
                     int(bool(use_tol)), int(grad), int(bool(stress)), float(scale), float(scale_grad),
                     int(bool(symmetric)))
    _lib.check(lib.cspbo_block_build(ctypes.byref(desc), a.handle, b.handle, _ptr(out),
                                     _ptr(out_grad) if grad else None, mem, 0))
    return out, (out_grad if grad else None)


def _build_block_spread(family, block, a, b, zeta, sigma2, sigma02, l, tol, use_tol, with_grad, stress, scale, scale_grad,
                        symmetric, out, out_grad, device):
    """build_block for spread energy packs (spread_energy): the block of the pseudo-structures is built on the device and
    its rows / columns are added up `spread` by `spread` into the real structures' entries; results have the shapes of the
    real groups.  K_EE and K_EF only (force packs are never spread)."""
    import torch
    if block == _lib.KFF:
        raise ValueError("spread packs are energy packs: K_EE / K_EF only")
    pa, pb = _plain_view(a), _plain_view(b)
    tmp, tmp_g = build_block(family, block, pa, pb, zeta, sigma2, sigma02, l, tol, use_tol, with_grad, stress, scale, scale_grad,
                             symmetric, device=True)

    def fold(t):
        if t is None:
            return None
        if block == _lib.KEE:
            return t.view(a.m_real, a.spread, b.m_real, b.spread).sum(dim=(1, 3))
        return t.view(a.m_real, a.spread, b.m, t.shape[-1]).sum(dim=1)        # K_EF: side a is the energy pack
    res, res_g = fold(tmp), fold(tmp_g)
    want_dev = device or (out is not None and _is_torch(out) and out.is_cuda)

    def deliver(r, dst):
        if r is None:
            return None
        if dst is None:
            return r if want_dev else r.cpu().numpy()
        if _is_torch(dst):
            dst.view(r.shape).copy_(r)
        else:
            dst.reshape(r.shape)[...] = r.cpu().numpy()
        return dst
    return deliver(res, out), deliver(res_g, out_grad)


class _PlainView:
    """A spread pack seen as the plain pack of its pseudo-structures (what the C side holds)."""

    def __init__(self, p):
        self._p, self.spread, self.m, self.handle = p, 1, p.m, p.handle
        self.n, self.d, self.nc = p.n, p.d, p.nc


def _plain_view(p):
    return _PlainView(p) if p.spread > 1 else p


def block_diag(family, block, a, zeta, sigma2=1.0, sigma02=0.0, l=1.0, scale=1.0):
    """cspbo_block_diag: per-group self covariance, [m] (K_EE) or [m,3,3] (K_FF), as a numpy array."""
    lib = _lib.load()
    out = np.zeros((a.m,) if block == _lib.KEE else (a.m, 3, 3))
    desc = BlockDesc(family, block, float(zeta), float(sigma2), float(sigma02), float(l), 0.0, 0, 0, 0, float(scale), 1.0, 0)
    _lib.check(lib.cspbo_block_diag(ctypes.byref(desc), a.handle, _ptr(out), MEM_HOST))
    return out
