"""Runs the reference's OWN cspbo/SO3.py (exec'd from source; container only) with a stub `ase` whose
NeighborList is oracle.so3_reference.neighbor_list and a shim for scipy's removed sph_harm."""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = os.environ.get("CSPBO_REFERENCE", "/root/reference")

from oracle import so3_reference as so  # noqa: E402


def load_reference_so3():
    import scipy.special as sp
    ase = types.ModuleType("ase")
    ase.Atoms = so.Structure
    nlmod = types.ModuleType("ase.neighborlist")
    nlmod.NeighborList = so.StubNeighborList
    nlmod.PrimitiveNeighborList = so.StubNeighborList
    sys.modules["ase"], sys.modules["ase.neighborlist"] = ase, nlmod
    if not hasattr(sp, "sph_harm"):
        sp.sph_harm = lambda m, l, az, pol: sp.sph_harm_y(l, m, pol, az)
    src = open(os.path.join(REF, "cspbo", "SO3.py")).read()
    ns = {"__name__": "ref_so3"}
    exec(compile(src, "SO3.py", "exec"), ns)
    return ns["SO3"]


def test_structures():
    rng = np.random.default_rng(7)
    a = 5.43
    base = np.array([[0, 0, 0], [.25, .25, .25], [0, .5, .5], [.25, .75, .75], [.5, 0, .5], [.75, .25, .75], [.5, .5, 0], [.75, .75, .25]])
    si = so.Structure(base * a + rng.normal(0, 0.08, (8, 3)), [14] * 8, np.eye(3) * a)
    cell = np.array([[6.2, 0.0, 0.0], [0.4, 6.0, 0.0], [0.3, -0.5, 7.1]])
    frac = rng.random((12, 3))
    # keep atoms apart: jittered 2x2x3 grid
    grid = np.array([(i, j, k) for i in range(2) for j in range(2) for k in range(3)]) / np.array([2, 2, 3]) + 0.1
    ptho = so.Structure((grid + rng.normal(0, 0.03, (12, 3))) @ cell, [78, 78, 8, 8, 8, 1, 1, 1, 1, 1, 1, 8], cell)
    tri = np.array([[3.0, 0, 0], [0.5, 3.1, 0], [0.2, 0.4, 3.2]])
    small = so.Structure(np.array([[0.1, 0.2, 0.1], [1.6, 1.4, 1.7]]), [29, 29], tri)
    return {"si8": (si, dict(nmax=3, lmax=3, rcut=3.5, alpha=2.0)),
            "ptho12": (ptho, dict(nmax=3, lmax=3, rcut=3.5, alpha=2.0)),
            "cu2_small_cell": (small, dict(nmax=4, lmax=4, rcut=4.0, alpha=1.5))}


def run_reference(struct, params):
    SO3 = load_reference_so3()
    d = SO3(derivative=True, stress=True, **params).calculate(struct)
    return {k: (np.asarray(v) if k != "elements" else v) for k, v in d.items()}


if __name__ == "__main__":
    for name, (st, prm) in test_structures().items():
        ref = run_reference(st, prm)
        mine = so.so3_calculate(st, stress=True, **prm)
        assert np.array_equal(ref["seq"], mine["seq"]), name
        for k in ("x", "dxdr", "rdxdr"):
            err = np.abs(ref[k] - mine[k]).max() / np.abs(ref[k]).max()
            print("%-16s %-6s shape %-16s max|.| %.4e  restatement err %.2e" % (name, k, ref[k].shape, np.abs(ref[k]).max(), err))
