"""Dot many-body kernel blocks on the GPU: kee_C / kef_C / kff_C.

Same names, arguments, return values and scalings as the reference wrappers
(cspbo/kernels/dot_kernel.py:8-62, :65-159, :161-274); the arithmetic runs in the sm_100a tile
kernels of libcspbo_b200.so (csrc/tiles.cu) instead of dot_kernel.cpp.  No MPI: the reference's
row split + Allreduce (dot_kernel.py:188-247) is replaced by the GPU grid.
"""
import numpy as np

from .. import _lib
from ..packing import HOST_ROW_CLASSES, Pack, build_block, pack_energy, pack_force
from ..utilities import list_to_tuple


def _energy(X):
    """list | stacked tuple | caller-held Pack (packing.resident) -> tuple | Pack"""
    if isinstance(X, list):
        X = list_to_tuple(X, mode='energy')
    return X


def _force(X, stress=False):
    if isinstance(X, list):
        X = list_to_tuple(X, stress=stress)
    return X


def _same(X1, X2):
    """Both sides are the same stacked set (K(X, X): one pack, symmetric build)."""
    if X1 is X2:
        return True
    if isinstance(X1, Pack) or isinstance(X2, Pack):
        return False
    return X1[0] is X2[0] and X1[1] is X2[1]


def _force_packs(X1, X2, stress, on_dev):
    """Packs of the two force sets of a K_FF build.  A host-bound result above 64 MB packs side 1 in row classes so
    that finished classes stream back while the next one is computed (cspbo_block_build, pipelined path)."""
    m1 = len(X1) if isinstance(X1, Pack) else len(X1[3])
    m2 = len(X2) if isinstance(X2, Pack) else len(X2[3])
    big = (not on_dev) and 72.0 * m1 * m2 * (3 if stress else 1) >= (64 << 20) and m1 >= 512
    a = pack_force(X1, row_classes=HOST_ROW_CLASSES if big else 1)
    same = _same(X1, X2)
    b = a if same else pack_force(X2)
    return a, b, same


def kee_C(X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False):
    """Energy-energy block [m1, m2] (and dC/dsigma, dC/dsigma0 when grad=True)."""
    X1, X2 = _energy(X1), _energy(X2)
    # small energy sets are packed spread over the tile slots (packing.spread_energy): 8x the CTAs for a K* row
    a, b = pack_energy(X1, spread_small=True), pack_energy(X2, spread_small=True)
    C, _ = build_block(_lib.DOT, _lib.KEE, a, b, zeta, sigma2=sigma ** 2, sigma02=sigma0 ** 2, scale=sigma ** 2)
    C /= (a.real_counts[:, None] * b.real_counts[None, :])
    if grad:
        C1 = 2 * C / sigma
        # the reference's d/dsigma0 is this constant (dot_kernel.py:57), reproduced as is
        C2 = 0.8 * 2 * sigma ** 2 * sigma0 * np.ones([a.m_real, b.m_real])
        return C, C1, C2
    return C


def kef_C(X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False, stress=False, transpose=False):
    """Energy-force block [m1, 3*m2] (K_FE when transpose=True; + stress block when stress=True)."""
    X1, X2 = _energy(X1), _force(X2, stress)
    a, b = pack_energy(X1, spread_small=True), pack_force(X2)
    m1, m2 = a.m_real, b.m
    out, _ = build_block(_lib.DOT, _lib.KEF, a, b, zeta, stress=stress, scale=1.0)
    out /= a.real_counts[:, None, None]
    out *= -sigma * sigma
    C = out[:, :, :3].reshape([m1, m2 * 3])
    Cs = out[:, :, 3:].reshape([m1, m2 * 6]) if stress else np.zeros([m1, m2 * 6])
    if transpose:
        C, Cs = C.T, Cs.T
    if grad:
        return C, 2 * C / sigma, np.zeros([m1, m2 * 3])
    if stress:
        return C, Cs
    return C


def kff_C(X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False, stress=False, out=None):
    """Force-force block [3*m1, 3*m2] (+ [6*m1, 3*m2] stress rows when stress=True).

    `out` (extension): optional preallocated (pinned) float64 buffer of m1*nc1*m2*3 elements."""
    X1, X2 = _force(X1, stress), _force(X2)
    on_dev = out is not None and type(out).__module__.startswith('torch') and out.is_cuda
    a, b, same = _force_packs(X1, X2, stress, on_dev)
    m1, m2 = a.m, b.m
    res, _ = build_block(_lib.DOT, _lib.KFF, a, b, zeta, stress=stress, scale=sigma * sigma * zeta,
                         symmetric=(same and not stress), out=out)
    nc1 = 9 if stress else 3
    res = res.reshape(m1, nc1, m2 * 3)
    C = res[:, :3, :].reshape([m1 * 3, m2 * 3])
    if grad:
        return C, 2 * C / sigma, np.zeros([m1 * 3, m2 * 3])
    if stress:
        return C, res[:, 3:, :].reshape([m1 * 6, m2 * 3])
    return C
