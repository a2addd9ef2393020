"""Device-resident assembly of the covariance matrices: the same blocks, scalings and layout as
`Dot_mb.k_total*` / `RBF_mb.k_total*` (reference Dot_mb.py:87-173, RBF_mb.py:87-172, base.py:3-26), but
built straight into one CUDA matrix (`cspbo_block_build_into`) so that the GP fit and prediction
(csp_bo_b200/gaussianprocess.py) never round-trip K through the host.
"""
import ctypes

from .. import _lib
from .._lib import BlockDesc, c_vp
from ..packing import SPREAD_MAX_GROUPS, Pack, build_block, pack_energy, pack_force, spread_energy
from ..utilities import list_to_tuple


def _tuples(data, stress=False, spread_small=False):
    """{'energy': tuple|list|Pack, 'force': tuple|list|Pack} -> packs (None when absent/empty).  Tuples and lists are
    packed afresh on every call (stateless, like the reference wrappers); caller-held Packs (packing.resident) pass
    through."""
    pe = pf = None
    E = data.get("energy")
    if E is not None and len(E) > 0:
        if isinstance(E, list):
            E = list_to_tuple(E, mode='energy')
        if spread_small and not isinstance(E, Pack) and len(E[2]) <= SPREAD_MAX_GROUPS:
            pe = spread_energy(E)         # few structures (an MD step has one): fill all 8 slots of the tiles
        else:
            pe = pack_energy(E)
    F = data.get("force")
    if F is not None and len(F) > 0:
        if isinstance(F, list):
            F = list_to_tuple(F, stress=stress)
        pf = pack_force(F)
    return pe, pf


def _hyper(kernel, quirk):
    """family, (zeta_ee, zeta_ef, zeta_ff), block scales.  `quirk`: Dot_mb.k_total passes zeta in the
    sigma0 slot, so K_EF / K_FF run with the wrappers' default zeta = 2.0 (Dot_mb.py:108-117)."""
    if kernel.name == 'Dot_mb':
        z = float(kernel.zeta)
        zf = 2.0 if quirk else z
        s2 = kernel.sigma ** 2
        return dict(family=_lib.DOT, zee=z, zef=zf, zff=zf, sigma2=s2, sigma02=kernel.sigma0 ** 2, l=1.0,
                    see=s2, sef=-s2, sff=s2 * zf)
    z = float(kernel.zeta)
    return dict(family=_lib.RBF, zee=z, zef=z, zff=z, sigma2=kernel.sigma ** 2, sigma02=0.0, l=kernel.l,
                see=1.0, sef=1.0, sff=1.0)


def _into(h, block, zeta, scale, a, b, K, row0, col0, tol=0.0, use_tol=False, symmetric=False):
    if a.spread > 1 or b.spread > 1:
        # spread energy packs (packing.spread_energy): build the block of the pseudo-structures, then add their rows /
        # columns up into the real structures' entries
        import torch
        rows = a.m * (3 if block == _lib.KFF else 1)
        cols = b.m * (1 if block == _lib.KEE else 3)
        tmp = torch.empty((rows, cols), dtype=torch.float64, device="cuda")
        _into_plain(h, block, zeta, scale, a, b, tmp, 0, 0, tol, use_tol, symmetric)
        ra, cb = rows // a.spread, cols // b.spread
        red = tmp.view(ra, a.spread, cb, b.spread).sum(dim=(1, 3))
        K[row0:row0 + ra, col0:col0 + cb] = red
        return
    _into_plain(h, block, zeta, scale, a, b, K, row0, col0, tol, use_tol, symmetric)


def _into_plain(h, block, zeta, scale, a, b, K, row0, col0, tol=0.0, use_tol=False, symmetric=False):
    desc = BlockDesc(h["family"], block, float(zeta), float(h["sigma2"]), float(h["sigma02"]), float(h["l"]),
                     float(tol), int(bool(use_tol)), 0, 0, float(scale), 1.0, int(bool(symmetric)))
    _lib.check(_lib.load().cspbo_block_build_into(ctypes.byref(desc), a.handle, b.handle, c_vp(K.data_ptr()),
                                                  K.shape[1], int(row0), int(col0)))


def _counts(p):
    import torch
    return torch.as_tensor(p.real_counts, dtype=torch.float64, device="cuda")


def k_total_device(kernel, data1, data2=None, tol=1e-10, grad_semantics=False):
    """[[K_EE, K_EF], [K_FE, K_FF]] between data1 (rows) and data2 (columns) as a CUDA tensor.

    grad_semantics=True gives the K that `k_total_with_grad` returns instead of `k_total`'s: the Dot K_EF / K_FF blocks
    use the kernel's own zeta (Dot_mb.py:138-143 vs the zeta-in-sigma0-slot calls of :108-117) and the RBF K_FF has no
    `tol` test (rbf_kernel.cpp `kff_many_with_grad` vs :394)."""
    import torch
    h = _hyper(kernel, quirk=not grad_semantics)
    same = data2 is None
    # K* (data2 given): energy sets of a few structures are packed spread over the 8 tile slots (see _tuples)
    e1, f1 = _tuples(data1, spread_small=not same)
    e2, f2 = (e1, f1) if same else _tuples(data2, spread_small=True)
    ne1, nf1 = (e1.m_real if e1 else 0), (3 * f1.m if f1 else 0)
    ne2, nf2 = (e2.m_real if e2 else 0), (3 * f2.m if f2 else 0)
    K = torch.empty((ne1 + nf1, ne2 + nf2), dtype=torch.float64, device="cuda")
    if e1 and e2:
        _into(h, _lib.KEE, h["zee"], h["see"], e1, e2, K, 0, 0, symmetric=same)
        K[:ne1, :ne2] /= (_counts(e1)[:, None] * _counts(e2)[None, :])
    if e1 and f2:
        _into(h, _lib.KEF, h["zef"], h["sef"], e1, f2, K, 0, ne2)
        K[:ne1, ne2:] /= _counts(e1)[:, None]
    if f1 and e2:
        if same:
            K[ne1:, :ne2] = K[:ne1, ne2:].T
        else:
            tmp = torch.empty((ne2, nf1), dtype=torch.float64, device="cuda")
            _into(h, _lib.KEF, h["zef"], h["sef"], e2, f1, tmp, 0, 0)
            tmp /= _counts(e2)[:, None]
            K[ne1:, :ne2] = tmp.T
    if f1 and f2:
        _into(h, _lib.KFF, h["zff"], h["sff"], f1, f2, K, ne1, ne2, tol=tol,
              use_tol=(h["family"] == _lib.RBF and not grad_semantics), symmetric=same)
    return K


def k_total_with_grad_device(kernel, data1):
    """(K, [dK/dtheta1, dK/dtheta2]) for the training set, all CUDA tensors (Dot_mb.py:121-148,
    RBF_mb.py:122-148).  theta = (sigma, sigma0) for Dot, (sigma, l) for RBF."""
    import torch
    h = _hyper(kernel, quirk=False)
    e1, f1 = _tuples(data1)
    ne, nf = (e1.m if e1 else 0), (3 * f1.m if f1 else 0)
    N = ne + nf
    K = torch.empty((N, N), dtype=torch.float64, device="cuda")
    dK2 = torch.zeros((N, N), dtype=torch.float64, device="cuda")
    ce = torch.as_tensor(e1.counts, dtype=torch.float64, device="cuda") if e1 else None
    if h["family"] == _lib.DOT:
        if e1:
            _into(h, _lib.KEE, h["zee"], h["see"], e1, e1, K, 0, 0, symmetric=True)
            K[:ne, :ne] /= (ce[:, None] * ce[None, :])
            dK2[:ne, :ne] = 0.8 * 2 * kernel.sigma ** 2 * kernel.sigma0      # dot_kernel.py:57, as is
        if e1 and f1:
            _into(h, _lib.KEF, h["zef"], h["sef"], e1, f1, K, 0, ne)
            K[:ne, ne:] /= ce[:, None]
            K[ne:, :ne] = K[:ne, ne:].T
        if f1:
            _into(h, _lib.KFF, h["zff"], h["sff"], f1, f1, K, ne, ne, symmetric=True)
    else:
        l = kernel.l
        if e1:
            C, Cl = build_block(_lib.RBF, _lib.KEE, e1, e1, h["zee"], sigma2=h["sigma2"], l=l, with_grad=True,
                                scale_grad=1.0 / (l * l * l), device=True)
            norm = ce[:, None] * ce[None, :]
            K[:ne, :ne] = C / norm
            dK2[:ne, :ne] = Cl / norm
        if e1 and f1:
            C, Cl = build_block(_lib.RBF, _lib.KEF, e1, f1, h["zef"], sigma2=h["sigma2"], l=l, with_grad=True, device=True)
            K[:ne, ne:] = C.reshape(ne, nf) / ce[:, None]
            dK2[:ne, ne:] = Cl.reshape(ne, nf) / ce[:, None]
            K[ne:, :ne] = K[:ne, ne:].T
            dK2[ne:, :ne] = dK2[:ne, ne:].T
        if f1:
            C, Cl = build_block(_lib.RBF, _lib.KFF, f1, f1, h["zff"], sigma2=h["sigma2"], l=l, with_grad=True,
                                symmetric=True, device=True)
            K[ne:, ne:] = C.reshape(nf, nf)
            dK2[ne:, ne:] = Cl.reshape(nf, nf)
    dK1 = K * (2.0 / kernel.sigma)
    return K, [dK1, dK2]


def k_total_with_stress_device(kernel, data1, data2, tol=1e-10):
    """(K*, K*_stress) as CUDA tensors (Dot_mb.py:150-173); data1 force entries carry 9 columns."""
    import torch
    h = _hyper(kernel, quirk=True)
    e1, f1 = _tuples(data1, stress=True, spread_small=True)
    e2, f2 = _tuples(data2, spread_small=True)
    ne1, m1 = (e1.m_real if e1 else 0), (f1.m if f1 else 0)
    ne2, nf2 = (e2.m_real if e2 else 0), (3 * f2.m if f2 else 0)
    K = torch.empty((ne1 + 3 * m1, ne2 + nf2), dtype=torch.float64, device="cuda")
    Ks = torch.empty((6 * m1, ne2 + nf2), dtype=torch.float64, device="cuda")
    fam = h["family"]
    if e1 and e2:
        _into(h, _lib.KEE, h["zee"], h["see"], e1, e2, K, 0, 0)
        K[:ne1, :ne2] /= (_counts(e1)[:, None] * _counts(e2)[None, :])
    if e1 and f2:
        _into(h, _lib.KEF, h["zef"], h["sef"], e1, f2, K, 0, ne2)
        K[:ne1, ne2:] /= _counts(e1)[:, None]
    if f1 and e2:
        out, _ = build_block(fam, _lib.KEF, e2, f1, h["zef"], sigma2=h["sigma2"], l=h["l"], stress=True,
                             scale=h["sef"], device=True)                    # [ne2, m1, 9] (spread packs are folded by build_block)
        out = out / _counts(e2)[:, None, None]
        K[ne1:, :ne2] = out[:, :, :3].reshape(ne2, 3 * m1).T
        Ks[:, :ne2] = out[:, :, 3:].reshape(ne2, 6 * m1).T
    if f1 and f2:
        out, _ = build_block(fam, _lib.KFF, f1, f2, h["zff"], sigma2=h["sigma2"], l=h["l"], stress=True, tol=tol,
                             use_tol=(fam == _lib.RBF), scale=h["sff"], device=True)   # [m1, 9, nf2]
        K[ne1:, ne2:] = out[:, :3, :].reshape(3 * m1, nf2)
        Ks[:, ne2:] = out[:, 3:, :].reshape(6 * m1, nf2)
    return K, Ks


def rbf_dl_trace(kernel, data1, alpha, W):
    """sum_ij (alpha alpha^T - W)_ij (dK/dl)_ij over the whole training covariance of an RBF_mb kernel, with dK/dl
    evaluated tile by tile inside the kernel epilogue and never stored (cspbo_block_trace; the reference materialises
    it as one plane of the [N,N,3] K_gradient, gaussianprocess.py:462,492-499).  alpha: [N] CUDA tensor; W: [N,N] CUDA
    tensor whose row-major UPPER triangle holds K^-1.  Both are modified: the 1/N_i normalisations of the energy rows
    (rbf_kernel.py:55-60,157) are folded into them.  Returns a 0-dim CUDA tensor."""
    import torch
    h = _hyper(kernel, quirk=False)
    assert h["family"] == _lib.RBF
    e1, f1 = _tuples(data1)
    ne = e1.m if e1 else 0
    l = kernel.l
    if e1:
        f = 1.0 / torch.as_tensor(e1.counts, dtype=torch.float64, device="cuda")
        alpha[:ne] *= f
        W[:ne, :] *= f[:, None]           # rows of the energy part (upper triangle: E-E and E-F blocks)
        W[:ne, :ne] *= f[None, :]
    res = torch.zeros(1, dtype=torch.float64, device="cuda")
    lib = _lib.load()

    def trace(block, a, b, row0, col0, scale_grad, symmetric):
        desc = BlockDesc(_lib.RBF, block, float(h["zee"]), float(h["sigma2"]), 0.0, float(l), 0.0, 0, 1, 0, 1.0, float(scale_grad),
                         int(bool(symmetric)))
        _lib.check(lib.cspbo_block_trace(ctypes.byref(desc), a.handle, b.handle, c_vp(alpha.data_ptr()), c_vp(W.data_ptr()),
                                         W.shape[1], int(row0), int(col0), c_vp(res.data_ptr()), 1))
    if e1:
        trace(_lib.KEE, e1, e1, 0, 0, 1.0 / (l * l * l), True)
    if e1 and f1:
        # the E-F block appears twice in the symmetric matrix (K_EF and K_FE = K_EF^T)
        trace(_lib.KEF, e1, f1, 0, ne, 2.0, False)
    if f1:
        trace(_lib.KFF, f1, f1, ne, ne, 1.0, True)
    return res[0]
