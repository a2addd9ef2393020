"""Synthetic descriptor sets shaped like the reference's Cpp_code/X1_big.npy.

Force-atom groups are drawn with replacement from the 192 groups of the X1_big fixture
(tests/golden/fixtures.npz, converted from the reference file by tests/golden/make_golden.py); each
group's (x, dxdr, ele) block is kept intact and every row is multiplied by exp(N(0, 0.05^2)) to
break exact duplicates.  Preserved: d = 24, 11-34 rows per atom (mean 22), species {1, 8, 78}
(41 % of row pairs same-species), |x| in [1e2, 1.2e5].  `single_element=True` relabels every row
as species 14 (the Si / Cu case: 100 % same-species pairs).
"""
import os

import numpy as np

_FIX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "fixtures.npz")
_cache = {}


def _base():
    if "big" not in _cache:
        fx = np.load(_FIX)
        _cache["big"] = (fx["big_x"], fx["big_dxdr"], fx["big_ele"], fx["big_counts"].astype(np.int64))
        _cache["ee2"] = (fx["ee2_x"], fx["ee2_ele"], fx["ee2_counts"].astype(np.int64))
    return _cache["big"]


def _mix(d):
    """Fixed non-negative [24, d] mixing matrix: descriptors of another length d (e.g. 50 = AgI.json's nmax = lmax = 4,
    90 = nmax = lmax = 5; SO3.py:205) with the group sizes, species and positive cosines of the X1_big rows."""
    key = ("mix", d)
    if key not in _cache:
        _cache[key] = np.random.default_rng(1234 + d).uniform(0.0, 1.0, size=(24, d)) / np.sqrt(24.0)
    return _cache[key]


def force_set(n_atoms, seed=0, single_element=False, d=24):
    """Stacked force tuple (X, dXdR, ELE, indices) with `n_atoms` groups (3*n_atoms force rows)."""
    x, dx, ele, counts = _base()
    off = np.concatenate(([0], np.cumsum(counts)))
    rng = np.random.default_rng(seed)
    pick = rng.integers(0, len(counts), size=n_atoms)
    rows = np.concatenate([np.arange(off[g], off[g + 1]) for g in pick])
    fac = np.exp(rng.normal(0.0, 0.05, size=len(rows)))
    X = x[rows] * fac[:, None]
    dX = dx[rows] * fac[:, None, None]
    if d != 24:
        M = _mix(d)
        X, dX = X @ M, np.einsum("nfk,fd->ndk", dX, M)
    E = np.full(len(rows), 14, dtype=np.int32) if single_element else ele[rows].astype(np.int32)
    return (np.ascontiguousarray(X), np.ascontiguousarray(dX), E, [int(c) for c in counts[pick]])


def energy_set(n_structures, seed=1, single_element=False, rows_per_structure=192, d=24):
    """Stacked energy tuple (X, ELE, indices): structures of `rows_per_structure` atoms drawn from the
    X2_EE fixture rows (Cpp_code/Kff_QZ/X2_EE.npy: 16 structures x 192 atoms)."""
    _base()
    x, ele, _ = _cache["ee2"]
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, len(x), size=n_structures * rows_per_structure)
    fac = np.exp(rng.normal(0.0, 0.05, size=len(rows)))
    X = x[rows] * fac[:, None]
    if d != 24:
        X = X @ _mix(d)
    E = np.full(len(rows), 14, dtype=np.int32) if single_element else ele[rows].astype(np.int32)
    return (np.ascontiguousarray(X), E, [rows_per_structure] * n_structures)


def same_species_pairs(ele1, ele2):
    """Number of same-species row pairs: the unit SURVEY 8(d) counts algorithmic flops in."""
    u1, c1 = np.unique(ele1, return_counts=True)
    u2, c2 = np.unique(ele2, return_counts=True)
    d2 = dict(zip(u2.tolist(), c2.tolist()))
    return int(sum(int(c) * d2.get(int(u), 0) for u, c in zip(u1, c1)))
