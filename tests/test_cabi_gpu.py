"""The C ABI called directly (ctypes, host pointers) against the oracle's C port on seeded random
inputs that exercise the edge cases of the reference loops: several species, zero-norm rows
(skipped, dot_kernel.cpp:25,37), groups that own no row, 1-row groups, the [n2_start,n2_end) slice
(dot_kernel.cpp:245), non-zero initial pout (the functions accumulate), d not a multiple of 4,
d = 50 (AgI.json), zeta in {1, 2, 2.5, 3}, and empty inputs."""
import ctypes
import os

import numpy as np
import pytest

from tests._cases import max_norm_err

pytestmark = pytest.mark.gpu

c_int, c_dbl, c_vp = ctypes.c_int, ctypes.c_double, ctypes.c_void_p


def P(a):
    return c_vp(a.ctypes.data)


def make_side(rng, m, d, nc, species=(1, 8, 78), zero_rows=2, empty_groups=1, max_rows=9):
    """Random stacked set with ragged groups; returns x, dxdr, ele, inds, m."""
    counts = rng.integers(1, max_rows, size=m)
    if empty_groups and m > 2:
        counts[rng.integers(0, m, size=empty_groups)] = 0      # groups that own no row
    n = int(counts.sum())
    # power-spectrum descriptors are non-negative, so the cosine c stays positive (pow(c, zeta) real)
    x = (np.abs(rng.standard_normal((n, d))) + 0.05) * rng.uniform(0.5, 50.0, size=(n, 1))
    dx = rng.standard_normal((n, d, max(nc, 1))) * 3.0
    if zero_rows and n > 4:
        x[rng.integers(0, n, size=zero_rows)] = 0.0              # |x| = 0 rows are skipped
    ele = rng.choice(species, size=n).astype(np.int32)
    inds = np.repeat(np.arange(m, dtype=np.int32), counts)
    return np.ascontiguousarray(x), np.ascontiguousarray(dx), ele, inds, m


@pytest.fixture(scope="module")
def libs(oracle_built):
    from csp_bo_b200 import _lib
    from oracle import cpu_reference as cr
    lib = _lib.load()
    port = cr.get_backend("port")
    return lib, port, _lib


def both(libs, family, name, args, out_shapes, init=0.0):
    """Run cspbo_<family>_<name> and the oracle on identical zero/`init`-filled outputs."""
    lib, port, _lib = libs
    outs_g = [np.full(s, init) for s in out_shapes]
    outs_o = [np.full(s, init) for s in out_shapes]
    rc = getattr(lib, "cspbo_%s_%s" % (family, name))(*args, *[P(o) for o in outs_g])
    assert rc == 0, lib.cspbo_last_error()
    port.fn(family, name)(*args, *[P(o) for o in outs_o])
    return outs_g, outs_o


# d = 90 (nmax = lmax = 5, SO3.py:205) and 140 run the chunked kernels (k-chunks of 24 features, tiles.cu), d = 450 the
# unstaged ones (tiles longer than a shared-memory stage): the reference loops `i < d` for any d (dot_kernel.cpp:257-289)
@pytest.mark.parametrize("d,zeta,seed", [(24, 2.0, 0), (24, 3.0, 1), (10, 2.5, 2), (50, 2.0, 3), (24, 1.0, 4),
                                         (90, 2.0, 5), (140, 3.0, 6), (57, 2.5, 7), (450, 2.0, 8),
                                         # 24 < d <= 56: the KS = 14 kernels skip the k-steps beyond ceil(d / 4) (VK instantiations)
                                         (25, 2.0, 20), (30, 2.5, 21), (40, 2.0, 22), (48, 3.0, 23), (49, 2.0, 24), (53, 2.0, 25),
                                         (3, 2.0, 29), (9, 2.0, 30), (12, 3.0, 31), (18, 2.0, 32), (20, 2.5, 33), (21, 2.0, 34),
                                         (60, 2.0, 37), (75, 2.0, 38), (92, 3.0, 39), (97, 2.0, 40)])      # chunked: ragged last chunk
def test_dot_entry_points(libs, d, zeta, seed):
    rng = np.random.default_rng(seed)
    x1, _, e1, g1, m1 = make_side(rng, 11, d, 0)
    x2, dx2, e2, g2, m2 = make_side(rng, 13, d, 3)
    xa, dxa, ea, ga, ma = make_side(rng, 9, d, 9)
    n1, n2, na = len(x1), len(x2), len(xa)
    I = c_int
    g, o = both(libs, "dot", "kee_many", [I(n1), I(n1), I(d), I(m1), c_dbl(zeta), c_dbl(7.3), c_dbl(0.04),
                                          P(x1), P(e1), P(g1), P(x1), P(e1), P(g1)], [(m1, m1)], init=0.25)
    assert max_norm_err(g[0], o[0]) < 1e-12
    g, o = both(libs, "dot", "kef_many", [I(n1), I(n2), I(d), I(m2), c_dbl(zeta), P(x1), P(e1), P(g1),
                                          P(x2), P(dx2), P(e2), P(g2)], [(m1, m2, 3)])
    assert max_norm_err(g[0], o[0]) < 1e-12
    g, o = both(libs, "dot", "kef_many_stress", [I(n1), I(na), I(d), I(ma), c_dbl(zeta), P(x1), P(e1), P(g1),
                                                 P(xa), P(dxa), P(ea), P(ga)], [(m1, ma, 9)])
    assert max_norm_err(g[0], o[0]) < 1e-12
    for (s, e) in ((0, n2), (n2 // 3, 2 * n2 // 3), (5, 5)):
        g, o = both(libs, "dot", "kff_many", [I(n2), I(n2), I(s), I(e), I(d), I(m2), c_dbl(zeta),
                                              P(x2), P(dx2), P(e2), P(g2), P(x2), P(dx2), P(e2), P(g2)],
                    [(m2, 3, m2 * 3)], init=-1.5)
        assert max_norm_err(g[0], o[0]) < 1e-12, (s, e)
    g, o = both(libs, "dot", "kff_many_stress", [I(na), I(n2), I(0), I(n2), I(d), I(m2), c_dbl(zeta),
                                                 P(xa), P(dxa), P(ea), P(ga), P(x2), P(dx2), P(e2), P(g2)],
                [(ma, 9, m2 * 3)])
    assert max_norm_err(g[0], o[0]) < 1e-12


@pytest.mark.parametrize("d,zeta,l,seed", [(24, 2.0, 0.56, 10), (24, 3.0, 0.9, 11), (50, 2.0, 0.5, 12), (7, 2.5, 1.3, 13),
                                           (90, 2.0, 0.7, 14), (140, 2.5, 0.9, 15), (450, 2.0, 0.6, 16),
                                           (30, 2.0, 0.6, 26), (40, 2.5, 0.8, 27), (52, 2.0, 0.5, 28), (9, 2.0, 0.6, 35), (18, 2.0, 0.7, 36), (60, 2.0, 0.6, 41), (75, 2.5, 0.8, 42)])
def test_rbf_entry_points(libs, d, zeta, l, seed):
    rng = np.random.default_rng(seed)
    x1, _, e1, g1, m1 = make_side(rng, 10, d, 0)
    x2, dx2, e2, g2, m2 = make_side(rng, 12, d, 3)
    xa, dxa, ea, ga, ma = make_side(rng, 8, d, 9)
    n1, n2, na = len(x1), len(x2), len(xa)
    I, s2, l2 = c_int, 39.0, l * l
    g, o = both(libs, "rbf", "kee_many", [I(n1), I(n1), I(d), I(m1), c_dbl(zeta), c_dbl(s2), c_dbl(l2),
                                          P(x1), P(e1), P(g1), P(x1), P(e1), P(g1)], [(m1, m1)])
    assert max_norm_err(g[0], o[0]) < 1e-12
    g, o = both(libs, "rbf", "kee_many_with_grad", [I(n1), I(n1), I(d), I(m1), c_dbl(zeta), c_dbl(s2), c_dbl(l2),
                                                    P(x1), P(e1), P(g1), P(x1), P(e1), P(g1)], [(m1, m1), (m1, m1)], init=0.5)
    assert max_norm_err(g[0], o[0]) < 1e-12 and max_norm_err(g[1], o[1]) < 1e-12
    g, o = both(libs, "rbf", "kef_many", [I(n1), I(n2), I(d), I(m2), c_dbl(zeta), c_dbl(s2), c_dbl(l2),
                                          P(x1), P(e1), P(g1), P(x2), P(dx2), P(e2), P(g2)], [(m1, m2, 3)])
    assert max_norm_err(g[0], o[0]) < 1e-12
    g, o = both(libs, "rbf", "kef_many_with_grad", [I(n1), I(n2), I(d), I(m2), c_dbl(zeta), c_dbl(s2), c_dbl(l),
                                                    P(x1), P(e1), P(g1), P(x2), P(dx2), P(e2), P(g2)], [(m1, m2, 6)], init=2.0)
    assert max_norm_err(g[0], o[0]) < 1e-12
    g, o = both(libs, "rbf", "kef_many_stress", [I(n1), I(na), I(d), I(ma), c_dbl(zeta), c_dbl(s2), c_dbl(l2),
                                                 P(x1), P(e1), P(g1), P(xa), P(dxa), P(ea), P(ga)], [(m1, ma, 9)])
    assert max_norm_err(g[0], o[0]) < 1e-12
    for tol in (1e-12, 0.0, -1.0):
        g, o = both(libs, "rbf", "kff_many", [I(n2), I(n2), I(3), I(n2 - 2), I(d), I(m2), c_dbl(zeta), c_dbl(s2),
                                              c_dbl(l2), c_dbl(tol), P(x2), P(dx2), P(e2), P(g2),
                                              P(x2), P(dx2), P(e2), P(g2)], [(m2, 3, m2 * 3)], init=1.0)
        assert max_norm_err(g[0], o[0]) < 1e-12, tol
    g, o = both(libs, "rbf", "kff_many_with_grad", [I(n2), I(n2), I(0), I(n2), I(d), I(m2), c_dbl(zeta), c_dbl(s2),
                                                    c_dbl(l), P(x2), P(dx2), P(e2), P(g2),
                                                    P(x2), P(dx2), P(e2), P(g2)], [(m2, 3, m2 * 3), (m2, 3, m2 * 3)])
    assert max_norm_err(g[0], o[0]) < 1e-12 and max_norm_err(g[1], o[1]) < 1e-12
    g, o = both(libs, "rbf", "kff_many_stress", [I(na), I(n2), I(0), I(n2), I(d), I(m2), c_dbl(zeta), c_dbl(s2),
                                                 c_dbl(l2), c_dbl(1e-12), P(xa), P(dxa), P(ea), P(ga),
                                                 P(x2), P(dx2), P(e2), P(g2)], [(ma, 9, m2 * 3)])
    assert max_norm_err(g[0], o[0]) < 1e-12


@pytest.mark.parametrize("nspecies,d", [(20, 24), (20, 90), (40, 24)])
def test_many_species(libs, nspecies, d):
    """More species than any shipped model has: 20 go through the kernel-parameter species map, 40 through the
    device-memory one (the reference compares `ele1[a] == ele2[b]` for any labels, dot_kernel.cpp:255)."""
    rng = np.random.default_rng(100 + nspecies + d)
    species = tuple(int(v) for v in rng.choice(np.arange(1, 119), size=nspecies, replace=False))
    x1, _, e1, g1, m1 = make_side(rng, 14, d, 0, species=species, max_rows=14)
    x2, dx2, e2, g2, m2 = make_side(rng, 17, d, 3, species=species[: nspecies - 3] + (200, 201, 202), max_rows=14)
    n1, n2 = len(x1), len(x2)
    assert len(np.unique(e1)) > 16 and len(np.unique(e2)) > 16
    I = c_int
    g, o = both(libs, "dot", "kee_many", [I(n1), I(n1), I(d), I(m1), c_dbl(2.0), c_dbl(7.3), c_dbl(0.04),
                                          P(x1), P(e1), P(g1), P(x1), P(e1), P(g1)], [(m1, m1)])
    assert max_norm_err(g[0], o[0]) < 1e-12
    g, o = both(libs, "dot", "kef_many", [I(n1), I(n2), I(d), I(m2), c_dbl(2.0), P(x1), P(e1), P(g1),
                                          P(x2), P(dx2), P(e2), P(g2)], [(m1, m2, 3)])
    assert max_norm_err(g[0], o[0]) < 1e-12
    g, o = both(libs, "dot", "kff_many", [I(n2), I(n2), I(0), I(n2), I(d), I(m2), c_dbl(2.0),
                                          P(x2), P(dx2), P(e2), P(g2), P(x2), P(dx2), P(e2), P(g2)], [(m2, 3, m2 * 3)])
    assert max_norm_err(g[0], o[0]) < 1e-12
    g, o = both(libs, "rbf", "kff_many_with_grad", [I(n2), I(n2), I(0), I(n2), I(d), I(m2), c_dbl(2.0), c_dbl(39.0),
                                                    c_dbl(0.7), P(x2), P(dx2), P(e2), P(g2),
                                                    P(x2), P(dx2), P(e2), P(g2)], [(m2, 3, m2 * 3), (m2, 3, m2 * 3)])
    assert max_norm_err(g[0], o[0]) < 1e-12 and max_norm_err(g[1], o[1]) < 1e-12


def test_in_place_update_of_the_inputs_is_seen(oracle_built):
    """Packing is stateless like the reference wrappers (dot_kernel.py:207-258 copy their inputs on every call): an
    array updated in place between two calls -- MD loops and finite differences reuse buffers -- must give the new
    K.  (Round 1 cached packs on (address, shape, 8 probed values) and returned the stale matrix here.)"""
    from csp_bo_b200 import synthetic
    from csp_bo_b200.kernels.dot_kernel import kee_C, kff_C
    from oracle import cpu_reference as cr
    X = synthetic.force_set(40, seed=3)
    E = synthetic.energy_set(5, seed=4, rows_per_structure=16)
    K0 = kff_C(X, X, 18.5, 0.01, 2.0).copy()
    E0 = kee_C(E, E, 18.5, 0.01, 2.0).copy()
    X[0][5] *= 1.5            # one row that no sparse probe would have looked at, same buffers
    X[1][5] *= -0.5
    E[0][3] = E[0][3][::-1].copy()      # (a rescaled row would leave the normalised kernel unchanged)
    K1, E1 = kff_C(X, X, 18.5, 0.01, 2.0), kee_C(E, E, 18.5, 0.01, 2.0)
    assert np.abs(K1 - K0).max() > 1e-6 * np.abs(K0).max() and np.abs(E1 - E0).max() > 1e-9 * np.abs(E0).max()
    assert max_norm_err(K1, cr.dot_kff_C(X, X, 18.5, 0.01, 2.0, backend="port")) < 1e-12
    assert max_norm_err(E1, cr.dot_kee_C(E, E, 18.5, 0.01, 2.0, backend="port")) < 1e-12
    # explicit residency: a caller-held pack is a snapshot and is accepted wherever a stacked tuple is
    from csp_bo_b200.packing import resident
    held = resident({"force": X})
    assert np.array_equal(kff_C(held["force"], held["force"], 18.5, 0.01, 2.0), K1)


def test_reference_named_shims(libs, fixtures):
    """libdot_kernel_b200.so / librbf_kernel_b200.so: the exact names the reference's cffi binds."""
    lib, port, _lib = libs
    x, dx = fixtures["x2_x"], fixtures["x2_dxdr"]
    ele = np.ascontiguousarray(fixtures["x2_ele"], dtype=np.int32)
    inds = np.repeat(np.arange(27, dtype=np.int32), fixtures["x2_counts"])
    n, d, m, I = len(x), 24, 27, c_int
    for so, family, extra in (("libdot_kernel_b200.so", "dot", []),
                              ("librbf_kernel_b200.so", "rbf", [c_dbl(39.0), c_dbl(0.3), c_dbl(1e-12)])):
        shim = ctypes.CDLL(os.path.join(_lib.LIB_DIR, so))
        shim.kff_many.restype = None
        args = [I(n), I(n), I(0), I(n), I(d), I(m), c_dbl(2.0)] + extra + [P(x), P(dx), P(ele), P(inds), P(x), P(dx), P(ele), P(inds)]
        got, want = np.zeros((m, 3, m * 3)), np.zeros((m, 3, m * 3))
        shim.kff_many(*args, P(got))
        port.fn(family, "kff_many")(*args, P(want))
        assert max_norm_err(got, want) < 1e-12, so


def test_empty_and_bad_arguments(libs):
    lib, port, _lib = libs
    x = np.ones((4, 24)); e = np.ones(4, dtype=np.int32); g = np.zeros(4, dtype=np.int32); out = np.zeros(4)
    I = c_int
    # n1 = 0: nothing to do, pout untouched
    assert lib.cspbo_dot_kee_many(I(0), I(4), I(24), I(1), c_dbl(2.0), c_dbl(1.0), c_dbl(0.0),
                                  P(x), P(e), P(g), P(x), P(e), P(g), P(out)) == 0
    assert np.all(out == 0)
    # descriptor length out of range
    assert lib.cspbo_dot_kee_many(I(4), I(4), I(0), I(1), c_dbl(2.0), c_dbl(1.0), c_dbl(0.0),
                                  P(x), P(e), P(g), P(x), P(e), P(g), P(out)) == 2
    # group id outside [0, x2i)
    gbad = np.array([0, 0, 0, 7], dtype=np.int32)
    assert lib.cspbo_dot_kee_many(I(4), I(4), I(24), I(1), c_dbl(2.0), c_dbl(1.0), c_dbl(0.0),
                                  P(x), P(e), P(g), P(x), P(e), P(gbad), P(out)) == 2


def test_symmetric_and_general_builds_agree_at_scale():
    """Size-independent properties at 5 001 force rows: the mirrored (symmetric) build equals the general
    build, K_FF is symmetric, and K_FF + noise is positive definite (Cholesky succeeds)."""
    import torch
    from csp_bo_b200 import _lib, gp, synthetic
    from csp_bo_b200.packing import Pack, build_block
    X = synthetic.force_set(1667, seed=5)
    p = Pack(X[0], X[1], X[2], X[3])
    Ks, _ = build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=688.0, symmetric=True, device=True)
    Kg, _ = build_block(_lib.DOT, _lib.KFF, p, p, 2.0, scale=688.0, symmetric=False, device=True)
    Ks, Kg = Ks.reshape(5001, 5001), Kg.reshape(5001, 5001)
    den = float(Kg.abs().max())
    assert float((Ks - Kg).abs().max()) / den < 1e-13
    # mirrored off-diagonal blocks are bit-identical; the 8x8-group diagonal blocks compute (i,j) and
    # (j,i) in different threads (different summation order), so allow rounding there
    assert float((Ks - Ks.T).abs().max()) / den < 1e-13
    assert float((Kg - Kg.T).abs().max()) / den < 1e-13
    y = torch.ones(5001, dtype=torch.float64, device="cuda")
    alpha = gp.fit_alpha(Ks, y, n_energy=0, noise_e=0.005, f_coef=30.0)
    Kn = Ks + torch.eye(5001, dtype=torch.float64, device="cuda") * (30 * 0.005) ** 2
    assert float((Kn @ alpha - y).abs().max()) < 1e-6


def test_pipelined_host_copyback_equals_device_build():
    """Host-bound K_FF builds above 64 MB are computed in row slabs whose copy-back overlaps the next
    slab (api.cu); the result must be bit-identical to the one-shot device build, for the symmetric,
    the general and the stress (9-row) variants."""
    from csp_bo_b200 import _lib, synthetic
    from csp_bo_b200.packing import Pack, build_block
    X = synthetic.force_set(1100, seed=9)
    Y = synthetic.force_set(1000, seed=10)
    p, q = Pack(X[0], X[1], X[2], X[3]), Pack(Y[0], Y[1], Y[2], Y[3])
    x9 = np.concatenate((X[1], 0.5 * X[1][:, :, [1, 2, 0]], -0.25 * X[1][:, :, [2, 0, 1]]), axis=2)
    p9 = Pack(X[0], x9, X[2], X[3])
    pc = Pack(X[0], X[1], X[2], X[3], row_classes=12)      # class-strided copy-back (one 2-D copy per class)
    pc9 = Pack(X[0], x9, X[2], X[3], row_classes=5)
    for a, b, kw in ((p, p, dict(symmetric=True)), (p, q, {}), (p9, q, dict(stress=True)),
                     (pc, pc, dict(symmetric=True)), (pc, q, {}), (pc9, q, dict(stress=True))):
        dev, _ = build_block(_lib.RBF, _lib.KFF, a, b, 2.0, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True, device=True, **kw)
        host, _ = build_block(_lib.RBF, _lib.KFF, a, b, 2.0, sigma2=39.0, l=0.56, tol=1e-10, use_tol=True, **kw)
        assert host.nbytes >= 64 << 20
        assert np.array_equal(host, dev.cpu().numpy()), kw


def test_panel_and_cholesky_entry_points_reject_bad_arguments():
    """The multi-GPU layer of the ABI (cspbo_block_build_panel, cspbo_gp_chol_*) returns CSPBO_E_ARG with a message
    instead of touching memory when the layout contract is violated."""
    import torch
    from csp_bo_b200 import _lib, synthetic
    from csp_bo_b200._lib import BlockDesc
    from csp_bo_b200.packing import Pack
    lib = _lib.load()
    F = synthetic.force_set(20, seed=1)
    E = synthetic.energy_set(3, seed=2, rows_per_structure=16)
    pf, pe = Pack(F[0], F[1], F[2], F[3]), Pack(E[0], None, E[1], E[2])
    buf = torch.zeros((192, 96 + 60), dtype=torch.float64, device="cuda")     # 4 panels of 48 rows, ld = N = 156
    outs = (c_vp * 1)(c_vp(buf.data_ptr()))
    desc = BlockDesc(_lib.DOT, _lib.KFF, 2.0, 1.0, 0.0, 1.0, 0.0, 0, 0, 0, 1.0, 1.0, 0)

    def call(block, a, b, ld, nbw, r0a, r0b, rank=0, nranks=1):
        desc.block = block
        return lib.cspbo_block_build_panel(ctypes.byref(desc), a.handle, b.handle, rank, nranks, outs, ld, nbw, r0a, r0b)

    assert call(_lib.KFF, pf, pf, 156, 50, 96, 96) == 2 and b"multiple of 24" in lib.cspbo_last_error()
    assert call(_lib.KFF, pf, pf, 156, 48, 90, 90) == 2 and b"panel boundaries" in lib.cspbo_last_error()
    assert call(_lib.KFF, pf, pf, 100, 48, 96, 96) == 2 and b"does not fit" in lib.cspbo_last_error()
    assert call(_lib.KEF, pf, pe, 156, 48, 0, 96) == 2                      # K_EF needs (energy, force)
    assert call(_lib.KEE, pe, pf, 156, 48, 0, 96) == 2                      # K_EE needs the same energy pack twice
    assert call(_lib.KFF, pf, pf, 156, 48, 96, 96, rank=1, nranks=1) == 2
    # the valid call fills the force block and leaves the rest alone
    assert call(_lib.KFF, pf, pf, 156, 48, 96, 96) == 0
    torch.cuda.synchronize()
    assert float(buf[:, :96].abs().max()) == 0.0 and float(buf[:96].abs().max()) == 0.0
    blk = buf[96:156, 96:156]
    assert float(blk.abs().max()) > 0.0 and float((blk - blk.T).abs().max()) <= 1e-12 * float(blk.abs().max())
    # Cholesky building blocks
    D = torch.eye(8, dtype=torch.float64, device="cuda")
    inv = torch.zeros(64, dtype=torch.float64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    st = c_vp(torch.cuda.current_stream().cuda_stream)
    assert lib.cspbo_gp_chol_diag(c_vp(D.data_ptr()), 4, 8, c_vp(inv.data_ptr()), c_vp(info.data_ptr()), st) == 2   # ld < h
    assert lib.cspbo_gp_chol_diag(None, 8, 8, c_vp(inv.data_ptr()), c_vp(info.data_ptr()), st) == 2
    assert lib.cspbo_gp_chol_apply_inv(c_vp(inv.data_ptr()), 8, c_vp(D.data_ptr()), 4, 8, c_vp(D.data_ptr()), 8, st) == 2   # ldb < w
    assert lib.cspbo_gp_chol_update(c_vp(D.data_ptr()), 8, 8, 8, c_vp(D.data_ptr()), 4, c_vp(D.data_ptr()), 8, 8, st) == 2   # lda < nb
    # not positive definite: reported through the device-side info word, no exception, no host sync
    Dn = -torch.eye(8, dtype=torch.float64, device="cuda")
    assert lib.cspbo_gp_chol_diag(c_vp(Dn.data_ptr()), 8, 8, c_vp(inv.data_ptr()), c_vp(info.data_ptr()), st) == 0
    torch.cuda.synchronize()
    assert int(info.item()) > 0


@pytest.mark.parametrize("h", [33, 192, 384])
def test_fused_diagonal_block_factorisation_matches_cusolver(h):
    """csrc/chol.cu (opt-in CSPBO_CHOL_DIAG=fused): the single-kernel Cholesky of a diagonal block, run in a child
    process with the switch set, against cuSOLVER's potrf through the same entry point."""
    import subprocess, sys, textwrap
    code = textwrap.dedent("""
        import ctypes, sys, torch
        sys.path.insert(0, %r)
        from csp_bo_b200 import _lib
        lib = _lib.load(); c_vp = ctypes.c_void_p
        h = %d
        g = torch.Generator(device="cuda").manual_seed(1)
        A = torch.randn(h, h, dtype=torch.float64, device="cuda", generator=g)
        D0 = A @ A.T / h + torch.eye(h, dtype=torch.float64, device="cuda")
        rows = torch.zeros(h, h + 7, dtype=torch.float64, device="cuda"); rows[:, :h] = D0
        inv = torch.zeros(h, h, dtype=torch.float64, device="cuda"); info = torch.zeros(1, dtype=torch.int32, device="cuda")
        st = c_vp(torch.cuda.current_stream().cuda_stream)
        assert lib.cspbo_gp_chol_diag(c_vp(rows.data_ptr()), h + 7, h, c_vp(inv.data_ptr()), c_vp(info.data_ptr()), st) == 0
        torch.cuda.synchronize()
        U = torch.triu(rows[:, :h])
        print(float((U.T @ U - D0).abs().max()), float((inv @ U - torch.eye(h, dtype=torch.float64, device="cuda")).abs().max()), int(info))
        bad = -torch.eye(h, dtype=torch.float64, device="cuda")
        assert lib.cspbo_gp_chol_diag(c_vp(bad.data_ptr()), h, h, c_vp(inv.data_ptr()), c_vp(info.data_ptr()), st) == 0
        torch.cuda.synchronize(); print(int(info))
    """) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), h)
    env = dict(os.environ, CSPBO_CHOL_DIAG="fused")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    l1, l2 = out.stdout.strip().splitlines()[-2:]
    ferr, ierr, info = l1.split()
    assert float(ferr) < 1e-13 and float(ierr) < 1e-12 and int(info) == 0
    assert int(l2) == 1          # first non-positive pivot reported like cuSOLVER's devInfo


@pytest.mark.parametrize("single", [False, True])
def test_b_run_split_matches_the_unsplit_build(single, monkeypatch):
    """Small problems: gridDim.y CTAs share a block pair (each takes every n-th run of B tiles) and plane_reduce_kernel
    sums their partial blocks (tiles.cu, bsplit).  Same numbers as the one-CTA-per-pair build to rounding, for every block
    kind, symmetric and not, written into a strided sub-block of a larger matrix, and twice the same bits (deterministic)."""
    import torch
    from csp_bo_b200 import _lib, synthetic
    from csp_bo_b200.kernels import Dot_mb, RBF_mb
    from csp_bo_b200.kernels import device as kdev
    from csp_bo_b200.packing import resident
    E = synthetic.energy_set(40, seed=5, single_element=single, rows_per_structure=96)
    F = synthetic.force_set(160, seed=6, single_element=single)
    E2 = synthetic.energy_set(9, seed=7, single_element=single, rows_per_structure=96)
    F2 = synthetic.force_set(50, seed=8, single_element=single)
    d1, d2 = resident({"energy": E, "force": F}), resident({"energy": E2, "force": F2})
    launches = {}
    for kern in (Dot_mb(para=[18.5, 0.01], zeta=2), RBF_mb(para=[6.2, 0.56], zeta=2), Dot_mb(para=[18.5, 0.01], zeta=3)):
        res = {}
        for mode in ("1", "5", "auto", "5"):
            if mode == "auto":
                monkeypatch.delenv("CSPBO_BSPLIT", raising=False)
            else:
                monkeypatch.setenv("CSPBO_BSPLIT", mode)
            n0 = _lib.launch_count()
            K = kdev.k_total_device(kern, d1)                      # symmetric K_EE / K_FF + K_EF into one matrix
            Ks = kdev.k_total_device(kern, d2, d1)                 # K*: non-symmetric blocks
            launches.setdefault(mode, _lib.launch_count() - n0)
            res.setdefault(mode, []).append((K, Ks))
        (K1, Ks1), (K5, Ks5), (K5b, Ks5b) = res["1"][0], res["5"][0], res["5"][1]
        Ka, Ksa = res["auto"][0]
        for ref, got in ((K1, K5), (Ks1, Ks5), (K1, Ka), (Ks1, Ksa)):
            assert float((ref - got).abs().max()) <= 1e-13 * float(ref.abs().max())
        assert torch.equal(K5, K5b) and torch.equal(Ks5, Ks5b)
    assert launches["5"] > launches["1"]                           # the forced split really ran (one reduce launch per block)
    # reference `pout +=` semantics through the split: accumulate = 1 adds the summed planes onto what `out` held
    import ctypes
    from csp_bo_b200._lib import BlockDesc, c_vp
    pe = d1["energy"]
    desc = BlockDesc(_lib.DOT, _lib.KEE, 2.0, 1.0, 0.01, 1.0, 0.0, 0, 0, 0, 342.25, 1.0, 1)
    outs = {}
    for mode in ("1", "5"):
        monkeypatch.setenv("CSPBO_BSPLIT", mode)
        out = torch.full((pe.m, pe.m), 0.75, dtype=torch.float64, device="cuda")
        _lib.check(_lib.load().cspbo_block_build(ctypes.byref(desc), pe.handle, pe.handle, c_vp(out.data_ptr()), None, 1, 1))
        outs[mode] = out
    assert float((outs["1"] - outs["5"]).abs().max()) <= 1e-13 * float(outs["1"].abs().max())
    assert float(outs["5"].min()) > 0.75
