"""SO(3) descriptor: the NumPy restatement (oracle/so3_reference.py) is pinned on the CPU against
goldens produced by the reference's own cspbo/SO3.py source (tests/golden/so3_ref_harness.py), and the
CUDA implementation is checked against both on the GPU."""
import os

import numpy as np
import pytest

from tests._cases import max_norm_err

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "so3_golden.npz")
CASES = ["si8", "ptho12", "cu2_small_cell"]


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(GOLD))


def structure(gold, name):
    from oracle.so3_reference import Structure
    prm = gold[name + "_prm"]
    st = Structure(gold[name + "_pos"], gold[name + "_num"], gold[name + "_cell"])
    return st, dict(nmax=int(prm[0]), lmax=int(prm[1]), rcut=float(prm[2]), alpha=float(prm[3]))


@pytest.mark.parametrize("name", CASES)
def test_oracle_restatement_matches_reference_goldens(gold, name):
    from oracle import so3_reference as so
    st, prm = structure(gold, name)
    d = so.so3_calculate(st, stress=True, **prm)
    assert np.array_equal(d["seq"], gold[name + "_seq"])
    for k in ("x", "dxdr", "rdxdr"):
        assert max_norm_err(d[k], gold[name + "_" + k]) < 1e-12, (name, k)


def test_oracle_gradient_is_consistent_with_finite_differences(gold):
    """dxdr[(i,j)] is d x_i / d r_j: move atom j, watch x_i."""
    from oracle import so3_reference as so
    st, prm = structure(gold, "si8")
    d0 = so.so3_calculate(st, stress=False, **prm)
    h = 1e-5
    j, k = 3, 1
    pos = st.positions.copy()
    pos[j, k] += h
    dp = so.so3_calculate(so.Structure(pos, st.numbers, st.cell), stress=False, **prm)
    pos[j, k] -= 2 * h
    dm = so.so3_calculate(so.Structure(pos, st.numbers, st.cell), stress=False, **prm)
    fd = (dp["x"] - dm["x"]) / (2 * h)
    for s, (i, jj) in enumerate(d0["seq"]):
        if jj == j:
            assert np.abs(fd[i] - d0["dxdr"][s, :, k]).max() < 1e-6 * max(1.0, np.abs(d0["dxdr"]).max())


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_descriptor_matches_reference_goldens(gold, name):
    from csp_bo_b200.SO3 import SO3
    st, prm = structure(gold, name)
    d = SO3(derivative=True, stress=True, **prm).calculate(st)
    assert np.array_equal(d["seq"], gold[name + "_seq"])
    assert d["elements"] == list(st.symbols)
    for k in ("x", "dxdr", "rdxdr"):
        assert max_norm_err(d[k], gold[name + "_" + k]) < 1e-10, (name, k)
    # derivative=False / stress=False branches (SO3.py:282-291)
    d2 = SO3(derivative=True, stress=False, **prm).calculate(st)
    assert d2["rdxdr"] is None and np.array_equal(d2["dxdr"], d["dxdr"])
    d3 = SO3(derivative=False, **prm).calculate(st)
    assert d3["dxdr"] is None and np.array_equal(d3["x"], d["x"])


@pytest.mark.gpu
def test_gpu_descriptor_options(gold):
    from csp_bo_b200.SO3 import SO3
    st, prm = structure(gold, "ptho12")
    d = SO3(derivative=True, stress=True, weight_on=True, **prm).calculate(st)
    for k in ("x", "dxdr", "rdxdr"):
        assert max_norm_err(d[k], gold["ptho12_won_" + k]) < 1e-10, k
    for cut in ("tanh", "poly2"):
        d = SO3(derivative=True, stress=True, cutoff_function=cut, **prm).calculate(st)
        for k in ("x", "dxdr"):
            assert max_norm_err(d[k], gold["ptho12_%s_%s" % (cut, k)]) < 1e-10, (cut, k)


@pytest.mark.gpu
def test_gpu_descriptor_larger_cell_vs_oracle():
    """192-atom Pt/H/O-like cell (the README workload's size) against the NumPy oracle."""
    from csp_bo_b200.SO3 import SO3
    from oracle import so3_reference as so
    rng = np.random.default_rng(5)
    g = np.array([(i, j, k) for i in range(4) for j in range(6) for k in range(8)], dtype=float)
    cell = np.diag([4 * 2.6, 6 * 2.6, 8 * 2.6])
    pos = g * 2.6 + rng.normal(0, 0.25, g.shape)
    num = rng.choice([1, 8, 78], size=len(pos), p=[0.5, 0.3, 0.2])
    st = so.Structure(pos, num, cell)
    ref = so.so3_calculate(st, nmax=3, lmax=3, rcut=3.5, alpha=2.0, stress=True)
    d = SO3(nmax=3, lmax=3, rcut=3.5, alpha=2.0, derivative=True, stress=True).calculate(st)
    assert np.array_equal(d["seq"], ref["seq"])
    for k in ("x", "dxdr", "rdxdr"):
        assert max_norm_err(d[k], ref[k]) < 1e-10, k


@pytest.mark.gpu
@pytest.mark.parametrize("pbc", [(True, True, True), (True, False, True), (False, False, False)])
def test_gpu_neighbour_cells_unwrapped_positions_and_mixed_pbc(pbc):
    """The linked-cell neighbour search of csrc/so3.cu: atoms far outside the unit cell (wrap shifts), a triclinic cell with
    one axis thinner than rcut (several images of one cell) and non-periodic axes, against the oracle's plain image loop;
    `seq` must come out identical (same pair order), the descriptors within 1e-10."""
    from csp_bo_b200.SO3 import SO3
    from oracle import so3_reference as so
    rng = np.random.default_rng(11)
    cell = np.array([[7.9, 0.0, 0.0], [1.1, 3.1, 0.0], [-0.7, 0.9, 9.6]])       # heights ~7.9, 3.1 (< rcut), 9.6
    frac = rng.uniform(0, 1, (40, 3))
    frac += rng.integers(-2, 3, size=(40, 3))                                    # some atoms lie cells away from the origin
    pos = frac @ cell
    num = rng.choice([8, 78], size=40)
    st = so.Structure(pos, num, cell, pbc=list(pbc))
    ref = so.so3_calculate(st, nmax=3, lmax=2, rcut=3.6, alpha=2.0, stress=True)
    d = SO3(nmax=3, lmax=2, rcut=3.6, alpha=2.0, derivative=True, stress=True).calculate(st)
    assert np.array_equal(d["seq"], ref["seq"])
    for k in ("x", "dxdr", "rdxdr"):
        assert max_norm_err(d[k], ref[k]) < 1e-10, k
    dev = SO3(nmax=3, lmax=2, rcut=3.6, alpha=2.0, derivative=True, stress=True).calculate_device(st)
    assert np.array_equal(dev["seq"], ref["seq"]) and max_norm_err(dev["dxdr"].cpu().numpy(), ref["dxdr"]) < 1e-10


@pytest.mark.gpu
def test_full_pipeline_descriptor_to_forces(oracle_built):
    """Structures -> SO3 (GPU) -> training tuples -> GaussianProcess.fit -> predict_structure, against the
    same chain on the CPU oracle (NumPy SO3 + reference C++ kernels + SciPy)."""
    from csp_bo_b200.SO3 import SO3
    from csp_bo_b200.gaussianprocess import GaussianProcess
    from csp_bo_b200.kernels import RBF_mb
    from oracle import so3_reference as so
    from oracle.gp_reference import RefGaussianProcess
    rng = np.random.default_rng(11)
    prm = dict(nmax=3, lmax=3, rcut=3.5, alpha=2.0)
    sym = {"H": 1, "O": 8, "Pt": 78, "Si": 14, "Cu": 29}

    def make(seed):
        r = np.random.default_rng(seed)
        g = np.array([(i, j, k) for i in range(2) for j in range(2) for k in range(3)], dtype=float)
        return so.Structure(g * 2.5 + r.normal(0, 0.2, g.shape), r.choice([1, 8, 78], size=12), np.diag([5.0, 5.0, 7.5]))

    def tuples(d, ele, E, F):
        energy = [(d['x'], E / len(ele), ele)]
        force = []
        for i in range(len(ele)):
            ids = np.argwhere(d['seq'][:, 1] == i).flatten()
            _i = d['seq'][ids, 0]
            force.append((d['x'][_i, :], d['dxdr'][ids], F[i], ele[_i]))
        return energy, force

    desc = SO3(derivative=True, stress=True, **prm)
    data_g, data_o = {"energy": [], "force": []}, {"energy": [], "force": []}
    for s in range(3):
        st = make(s)
        E, F = float(rng.normal(-30, 1)), rng.normal(0, 0.5, (12, 3))
        for data, d in ((data_g, desc.calculate(st)), (data_o, so.so3_calculate(st, stress=True, **prm))):
            e, f = tuples(d, st.numbers, E, F)
            data["energy"] += e
            data["force"] += f[:6]
    para, bounds = [6.244955524643115, 0.5611029935713949], [[1e-2, 5e+1], [1e-1, 1e+1]]
    model = GaussianProcess(RBF_mb(para=list(para), bounds=bounds, zeta=2), descriptor=desc, f_coef=30, noise_e=[0.005, 0.002, 0.1])
    model.fit(data_g, show=False, opt=False)
    # oracle chain
    from csp_bo_b200.utilities import list_to_tuple
    y = np.concatenate(([e[1] for e in data_o["energy"]], np.concatenate([f[2] for f in data_o["force"]])))
    train_x = {"energy": list_to_tuple([(e[0], e[2]) for e in data_o["energy"]], mode="energy"),
               "force": list_to_tuple([(f[0], f[1], f[3]) for f in data_o["force"]])}
    ref = RefGaussianProcess("rbf", para, 2.0, bounds, f_coef=30, noise_e=(0.005, 0.002, 0.1))
    ref.set_train(train_x, y)
    ref.fit(opt=False)
    test = make(99)
    E, F, S = model.predict_structure(test, stress=True)
    Er, Fr, Sr = ref.predict_structure(so.so3_calculate(test, stress=True, **prm), 12, test.numbers, stress=True)
    assert abs(E - Er) <= 1e-8 * max(1.0, abs(Er))
    assert max_norm_err(F, Fr) < 1e-8 and max_norm_err(S, Sr) < 1e-8
