"""oracle/so3_reference.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (NumPy + SciPy) of the reference's SO(3) power-spectrum descriptor
`cspbo/SO3.py` -- the producer of the hot path's inputs {'x', 'dxdr', 'rdxdr', 'seq', 'elements'}:

    SO3.calculate            SO3.py:187-291   (power spectrum, its gradient per (centre, neighbour) pair,
                                               the stress tensor terms)
    build_neighbor_list      SO3.py:317-376   (ase.neighborlist.NeighborList(cutoffs=rcut/2, bothways=True,
                                               self_interaction=False, skin=0) -> here: brute force over
                                               periodic images, same pair set)
    Cosine cutoff            SO3.py:378-385
    W, phi, g                SO3.py:463-495   (orthonormalised radial basis)
    GaussChebyshevQuadrature SO3.py:497-505
    compute_cs / compute_dcs SO3.py:507-695   (expansion coefficients c_nlm of the Gaussian-smeared density
                                               and their Cartesian gradients)

ase is not installed here, so the reference class cannot be instantiated as is; tests/golden/make_golden.py
exec's the reference source with a stub `ase` whose NeighborList is `neighbor_list` below and pins this
restatement (and the CUDA product) against the arrays the reference's own code then produces.
"""
import numpy as np
from scipy.special import spherical_in, sph_harm_y


class Structure:
    """The few ase.Atoms members SO3.py touches (positions, numbers, cell, pbc, symbols, volume)."""
    _SYM = {1: "H", 6: "C", 8: "O", 14: "Si", 29: "Cu", 47: "Ag", 53: "I", 78: "Pt"}

    def __init__(self, positions, numbers, cell, pbc=True):
        self.positions = np.asarray(positions, dtype=np.float64)
        self.numbers = np.asarray(numbers, dtype=np.int64)
        self.cell = np.asarray(cell, dtype=np.float64)
        self.pbc = np.array([pbc] * 3 if isinstance(pbc, (bool, np.bool_)) else pbc, dtype=bool)
        self.symbols = [self._SYM.get(int(z), "X") for z in self.numbers]

    def __len__(self):
        return len(self.positions)

    def __getitem__(self, i):
        class _A:
            pass
        a = _A()
        a.number = int(self.numbers[i])
        return a

    def get_cell(self):
        return self.cell

    def get_volume(self):
        return abs(np.linalg.det(self.cell))

    def get_scaled_positions(self):
        return np.linalg.solve(self.cell.T, self.positions.T).T


def neighbor_list(positions, cell, pbc, rcut):
    """All (i, j, offset) with 0 < |r_j + offset.cell - r_i| < rcut (both directions, periodic images of
    i itself included, the zero-offset self pair excluded).  Sorted by (i, j, offset)."""
    pos = np.asarray(positions, dtype=np.float64)
    cell = np.asarray(cell, dtype=np.float64)
    n = len(pos)
    rec = np.linalg.inv(cell).T                       # rows: reciprocal vectors (without 2 pi)
    heights = 1.0 / np.linalg.norm(rec, axis=1)
    # atoms may sit outside the cell: bound the image search by the spread of the scaled positions too
    frac = pos @ np.linalg.inv(cell)
    span = np.ceil(frac.max(axis=0) - frac.min(axis=0)).astype(int) if n else np.zeros(3, int)
    nimg = [int(np.ceil(rcut / heights[k])) + int(span[k]) if pbc[k] else 0 for k in range(3)]
    offs = np.array([(a, b, c) for a in range(-nimg[0], nimg[0] + 1) for b in range(-nimg[1], nimg[1] + 1)
                     for c in range(-nimg[2], nimg[2] + 1)], dtype=np.int64)
    shift = offs @ cell                                # [nimg, 3]
    out = []
    for i in range(n):
        d = pos[None, :, :] + shift[:, None, :] - pos[i][None, None, :]   # [img, j, 3]
        r = np.linalg.norm(d, axis=2)
        ok = (r < rcut) & (r > 1e-12)
        img, j = np.nonzero(ok)
        order = np.lexsort((offs[img, 2], offs[img, 1], offs[img, 0], j))
        for k in order:
            out.append((i, int(j[k]), tuple(int(v) for v in offs[img[k]])))
    return out


class StubNeighborList:
    """Drop-in for ase.neighborlist.NeighborList as SO3.py uses it (SO3.py:331-345)."""

    def __init__(self, cutoffs, self_interaction=False, bothways=True, skin=0.0):
        self.rcut = 2.0 * cutoffs[0]

    def update(self, atoms):
        self._nl = {}
        for i, j, off in neighbor_list(atoms.positions, atoms.get_cell(), atoms.pbc, self.rcut):
            self._nl.setdefault(i, ([], []))
            self._nl[i][0].append(j)
            self._nl[i][1].append(off)

    def get_neighbors(self, i):
        js, offs = self._nl.get(i, ([], []))
        return np.array(js, dtype=np.int64), np.array(offs, dtype=np.int64).reshape(-1, 3)


# ----------------------------------------------------------------------------- radial part
def W(nmax):
    """SO3.py:463-479."""
    arr = np.zeros((nmax, nmax))
    for a in range(1, nmax + 1):
        t1 = (2 * a + 5) * (2 * a + 6) * (2 * a + 7)
        for b in range(1, a + 1):
            t2 = (2 * b + 5) * (2 * b + 6) * (2 * b + 7)
            arr[a - 1, b - 1] = arr[b - 1, a - 1] = np.sqrt(t1 * t2) / (5 + a + b) / (6 + a + b) / (7 + a + b)
    sinv = np.linalg.inv(arr)
    eigvals, V = np.linalg.eig(sinv)
    return np.dot(np.dot(V, np.diag(np.sqrt(eigvals))), np.linalg.inv(V))


def radial_tables(nmax, lmax, rcut, alpha):
    """Quadrature nodes r_q and the weights G[n,q] of the radial integral, derivative form
    (SO3.py:497-505, 600-626): I_nl(R) = sum_q G[n,q] i_l(2 alpha R r_q)."""
    nq = (nmax + lmax + 1) * 10
    xq = np.cos((2 * np.arange(1, nq + 1) - 1) * np.pi / 2 / nq)
    weight = np.pi / nq * rcut / 2
    rq = rcut / 2 * (xq + 1)
    w = W(nmax)
    G = np.zeros((nmax, nq))
    for n in range(1, nmax + 1):
        for a in range(1, nmax + 1):
            phi = (rcut - rq) ** (a + 2) / np.sqrt(2 * rcut ** (2 * a + 7) / (2 * a + 5) / (2 * a + 6) / (2 * a + 7))
            G[n - 1] += w[n - 1, a - 1] * phi
    G *= weight * rq ** 2 * np.exp(-alpha * rq ** 2) * np.sqrt(1 - xq ** 2)
    return rq, G


def cosine_cutoff(R, rc, derivative=False):
    """SO3.py:378-385."""
    if not derivative:
        return 0.5 * (np.cos(np.pi * R / rc) + 1.0)
    return -0.5 * np.pi / rc * np.sin(np.pi * R / rc)


def compute_dcs(pos, nmax, lmax, rcut, alpha):
    """c_nlm and dc_nlm/dr for every neighbour vector (SO3.py:581-695), cosine cutoff.
    Returns C[nb, nmax, lmax+1, 2lmax+1] and dC[..., 3] (complex)."""
    pos = np.asarray(pos, dtype=np.float64)
    R = np.linalg.norm(pos, axis=1)
    u = pos / R[:, None]
    rq, G = radial_tables(nmax, lmax, rcut, alpha)
    args = 2 * alpha * np.outer(R, rq)
    I = np.zeros((len(R), nmax, lmax + 1))
    dI = np.zeros((len(R), nmax, lmax + 1, 3))
    for l in range(lmax + 1):
        il = spherical_in(l, args)
        dil = spherical_in(l, args, derivative=True) * (2 * alpha * rq)[None, :]
        I[:, :, l] = il @ G.T
        dI[:, :, l, :] = (dil @ G.T)[:, :, None] * u[:, None, :]
    ex = 4 * np.pi * np.exp(-alpha * R ** 2)
    dex = (-2 * alpha * R * ex)[:, None] * u
    fc = cosine_cutoff(R, rcut)
    dfc = cosine_cutoff(R, rcut, True)[:, None] * u
    theta = np.arccos(pos[:, 2] / R)
    phi = np.arctan2(pos[:, 1], pos[:, 0])
    Y = np.zeros((len(R), lmax + 2, 2 * (lmax + 1) + 1), dtype=np.complex128)
    M0 = lmax + 1
    for l in range(lmax + 2):
        for m in range(-l, l + 1):
            Y[:, l, M0 + m] = sph_harm_y(l, m, theta, phi)
    dY = np.zeros((len(R), lmax + 1, 2 * lmax + 1, 3), dtype=np.complex128)
    for l in range(1, lmax + 1):
        for m in range(-l, l + 1):
            Mi, mi = M0 + m, lmax + m
            x0 = -np.sqrt(((l + 1) ** 2 - m ** 2) / (2 * l + 1) / (2 * l + 3)) * l * Y[:, l + 1, Mi] / R
            if abs(m) <= l - 1:
                x0 = x0 + np.sqrt((l ** 2 - m ** 2) / (2 * l - 1) / (2 * l + 1)) * (l + 1) * Y[:, l - 1, Mi] / R
            xp = -np.sqrt((l + m + 1) * (l + m + 2) / 2 / (2 * l + 1) / (2 * l + 3)) * l * Y[:, l + 1, Mi + 1] / R
            if abs(m + 1) <= l - 1:
                xp = xp - np.sqrt((l - m - 1) * (l - m) / 2 / (2 * l - 1) / (2 * l + 1)) * (l + 1) * Y[:, l - 1, Mi + 1] / R
            xm = -np.sqrt((l - m + 1) * (l - m + 2) / 2 / (2 * l + 1) / (2 * l + 3)) * l * Y[:, l + 1, Mi - 1] / R
            if abs(m - 1) <= l - 1:
                xm = xm - np.sqrt((l + m - 1) * (l + m) / 2 / (2 * l - 1) / (2 * l + 1)) * (l + 1) * Y[:, l - 1, Mi - 1] / R
            dY[:, l, mi, 0] = 1 / np.sqrt(2) * (xm - xp)
            dY[:, l, mi, 1] = 1j / np.sqrt(2) * (xm + xp)
            dY[:, l, mi, 2] = x0
    Yc = Y[:, :lmax + 1, 1:2 * lmax + 2]
    YI = Yc[:, None, :, :] * I[:, :, :, None]                                   # [nb, n, l, m]
    C = ex[:, None, None, None] * YI
    dC = dex[:, None, None, None, :] * YI[..., None]
    dC = dC + ex[:, None, None, None, None] * (dY[:, None, :, :, :] * I[:, :, :, None, None]
                                               + Yc[:, None, :, :, None] * dI[:, :, :, None, :])
    dC = dC * fc[:, None, None, None, None] + dfc[:, None, None, None, :] * C[..., None]
    C = C * fc[:, None, None, None]
    return C, dC


def so3_calculate(struct, nmax=3, lmax=3, rcut=3.5, alpha=2.0, stress=True):
    """Restates SO3.calculate with derivative=True, cosine cutoff, weight_on=False (SO3.py:187-277).
    Returns {'x','dxdr','rdxdr','seq','elements','numbers'}."""
    n_at = len(struct)
    nl = neighbor_list(struct.positions, struct.get_cell(), struct.pbc, rcut)
    cell = struct.get_cell()
    ci = np.array([p[0] for p in nl], dtype=np.int64)
    cj = np.array([p[1] for p in nl], dtype=np.int64)
    off = np.array([p[2] for p in nl], dtype=np.float64).reshape(-1, 3)
    Ri = struct.positions[ci]
    rel = struct.positions[cj] + off @ cell - Ri
    Rj = rel + Ri
    wts = struct.numbers[cj].astype(np.float64)
    seq = []
    for i in range(n_at):
        js = sorted(set(cj[ci == i].tolist()) | {i})
        seq.extend([i, j] for j in js)
    seq = np.array(seq, dtype=np.int64)
    ncoefs = nmax * (nmax + 1) // 2 * (lmax + 1)
    tril = np.tril_indices(nmax, k=0)
    norm = np.sqrt(2 * np.sqrt(2) * np.pi / np.sqrt(2 * np.arange(lmax + 1) + 1))
    C, dC = compute_dcs(rel, nmax, lmax, rcut, alpha)
    C = C * wts[:, None, None, None] * norm[None, None, :, None]
    dC = dC * wts[:, None, None, None, None] * norm[None, None, :, None, None]
    x = np.zeros((n_at, ncoefs))
    dxdr = np.zeros((len(seq), ncoefs, 3))
    pstress = np.zeros((len(seq), ncoefs, 3, 3))
    seq_index = {(int(a), int(b)): k for k, (a, b) in enumerate(seq)}
    for i in range(n_at):
        cen = np.nonzero(ci == i)[0]
        if len(cen) == 0:
            continue
        ctot = C[cen].sum(axis=0)
        P = np.einsum('ijk,ljk->ilj', ctot, np.conj(ctot)).real
        dP = np.einsum('wijkn,ljk->wiljn', dC[cen], np.conj(ctot))
        dP = (dP + np.conj(np.transpose(dP, axes=[0, 2, 1, 3, 4]))).real
        rdPi = np.einsum('wn,wijkm->wijknm', Ri[cen], dP)
        rdPj = np.einsum('wn,wijkm->wijknm', Rj[cen], dP)
        for w, j in enumerate(cj[cen]):
            k = seq_index[(i, int(j))]
            dxdr[k] += dP[w][tril].reshape(ncoefs, 3)
            pstress[k] -= rdPj[w][tril].reshape(ncoefs, 3, 3)
        x[i] = P[tril].flatten()
        kii = seq_index[(i, i)]
        own = np.nonzero(seq[:, 0] == i)[0]
        dxdr[kii] -= dxdr[own].sum(axis=0)
        pstress[kii] += rdPi.sum(axis=0)[tril].reshape(ncoefs, 3, 3)
    out = {'x': x, 'dxdr': dxdr, 'seq': seq, 'elements': list(struct.symbols), 'numbers': struct.numbers.copy()}
    out['rdxdr'] = -pstress / struct.get_volume() if stress else None
    return out
