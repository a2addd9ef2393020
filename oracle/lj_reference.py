"""oracle/lj_reference.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Plain-Python restatement of the reference's LJ base potential (cspbo/calculator.py:117-177, ase's pairwise LJ with a
half neighbour list): every unordered pair once, contributions mirrored onto the second atom.  The neighbour list is a
brute-force image loop (the reference uses ase.neighborlist.NeighborList, absent here; any complete list of the pairs
within rc gives the same sums).  Parity status: the formulas are the reference's line by line; the neighbour list is
ours -- pinned by the force = -dE/dr finite-difference check of tests/test_gp_gpu.py."""
import itertools

import numpy as np


def lj_calculate(positions, cell, pbc, sigma=1.0, epsilon=1.0, rc=5.0):
    positions, cell = np.asarray(positions, dtype=float), np.asarray(cell, dtype=float)
    natoms = len(positions)
    e0 = 4 * epsilon * ((sigma / rc) ** 12 - (sigma / rc) ** 6)                      # calculator.py:129
    energies, forces, stresses = np.zeros(natoms), np.zeros((natoms, 3)), np.zeros((natoms, 3, 3))
    inv = np.linalg.inv(cell)
    heights = 1.0 / np.linalg.norm(inv, axis=0)
    nimg = [int(np.ceil(rc / heights[k])) + 1 if pbc[k] else 0 for k in range(3)]
    shifts = list(itertools.product(*[range(-n, n + 1) for n in nimg]))
    for ii in range(natoms):
        for jj in range(ii, natoms):
            for sh in shifts:
                if jj == ii and sh <= (0, 0, 0):                                     # each unordered pair once
                    continue
                d = positions[jj] + np.asarray(sh, dtype=float) @ cell - positions[ii]   # pointing towards the neighbour (:141)
                r2 = float(d @ d)
                if r2 > rc ** 2:
                    continue
                c6 = (sigma ** 2 / r2) ** 3                                           # :144-146
                c12 = c6 ** 2
                pe = 4 * epsilon * (c12 - c6) - e0                                    # :148
                pf = (-24 * epsilon * (2 * c12 - c6) / r2) * d                        # :151-153
                energies[ii] += 0.5 * pe; energies[jj] += 0.5 * pe                     # :149,162
                forces[ii] += pf; forces[jj] -= pf                                     # :155,163
                stresses[ii] += 0.5 * np.outer(pf, d); stresses[jj] += 0.5 * np.outer(pf, d)   # :156-158,164-166
    s = stresses
    voigt = np.stack((s[:, 0, 0], s[:, 1, 1], s[:, 2, 2], (s[:, 1, 2] + s[:, 2, 1]) / 2, (s[:, 0, 2] + s[:, 2, 0]) / 2,
                      (s[:, 0, 1] + s[:, 1, 0]) / 2), axis=1)                         # full_3x3_to_voigt_6_stress
    return energies.sum(), forces, voigt / abs(np.linalg.det(cell))
