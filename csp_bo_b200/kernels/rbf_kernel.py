"""RBF many-body kernel blocks on the GPU: kee_C / kef_C / kff_C.

Same names, arguments, return values and scalings as the reference wrappers
(cspbo/kernels/rbf_kernel.py:8-84, :86-185, :187-333); the arithmetic runs in the sm_100a tile
kernels of libcspbo_b200.so (csrc/tiles.cu) instead of rbf_kernel.cpp.
"""
import numpy as np

from .. import _lib
from ..packing import build_block, pack_energy, pack_force
from .dot_kernel import _energy, _force, _force_packs


def kee_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False):
    """Energy-energy block; with grad=True also (dC/dsigma, dC/dl)."""
    X1, X2 = _energy(X1), _energy(X2)
    a, b = pack_energy(X1, spread_small=True), pack_energy(X2, spread_small=True)
    norm = a.real_counts[:, None] * b.real_counts[None, :]
    C, C_l = build_block(_lib.RBF, _lib.KEE, a, b, zeta, sigma2=sigma * sigma, l=l, with_grad=grad)
    C /= norm
    if grad:
        C_l /= norm
        C_l *= 1 / (l * (l * l))      # rbf_kernel.py:47,59
        return C, (2 / sigma) * C, C_l
    return C


def kef_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False, stress=False, transpose=False):
    """Energy-force block; K_FE with transpose=True; stress block or (dC/dsigma, dC/dl) on request."""
    X1, X2 = _energy(X1), _force(X2, stress)
    a, b = pack_energy(X1, spread_small=True), pack_force(X2)
    m1, m2 = a.m_real, b.m
    with_grad = grad and not stress   # the reference gives stress priority (rbf_kernel.py:127-145)
    out, out_l = build_block(_lib.RBF, _lib.KEF, a, b, zeta, sigma2=sigma * sigma, l=l, stress=stress,
                             with_grad=with_grad)
    out /= a.real_counts[:, None, None]
    C = out[:, :, :3].reshape([m1, m2 * 3])
    Cs = np.zeros([m1, m2 * 6])
    C_s = C_l = None
    if stress:
        Cs = out[:, :, 3:].reshape([m1, m2 * 6])
    elif grad:
        out_l /= a.real_counts[:, None, None]
        C_l = out_l.reshape([m1, m2 * 3])
        C_s = (2 / sigma) * C
    if transpose:
        C, Cs = C.T, Cs.T
    if grad:
        return C, C_s, C_l
    if stress:
        return C, Cs
    return C


def kff_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False, stress=False, tol=1e-12, out=None):
    """Force-force block; pairs with dK/dD <= tol are skipped unless grad=True (rbf_kernel.cpp:394)."""
    X1, X2 = _force(X1, stress), _force(X2)
    on_dev = out is not None and type(out).__module__.startswith('torch') and out.is_cuda
    a, b, same = _force_packs(X1, X2, stress, on_dev)
    m1, m2 = a.m, b.m
    with_grad = grad and not stress
    res, res_l = build_block(_lib.RBF, _lib.KFF, a, b, zeta, sigma2=sigma * sigma, l=l, tol=tol,
                             use_tol=not with_grad, with_grad=with_grad, stress=stress,
                             symmetric=(same and not stress), out=out)
    nc1 = 9 if stress else 3
    res = res.reshape(m1, nc1, m2 * 3)
    C = res[:, :3, :].reshape([m1 * 3, m2 * 3])
    if stress:
        return C, res[:, 3:, :].reshape([m1 * 6, m2 * 3])
    if grad:
        return C, (2 / sigma) * C, res_l.reshape([m1 * 3, m2 * 3])
    return C
