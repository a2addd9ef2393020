"""Real data, end to end: the reference's shipped Si model (examples/models/Si.json + Si.db) and its
database/Si_244_DFT.db (BASELINE.json config 2), stored as plain-array fixtures by
tests/golden/make_datasets.py.  No ase: structures -> SO3 (GPU) -> GP (GPU)."""
import os

import numpy as np
import pytest

from tests._cases import max_norm_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def datasets():
    return dict(np.load(os.path.join(G, "datasets.npz")))


def test_shipped_si_model_fits_and_generalises(datasets):
    from csp_bo_b200 import asedb
    from csp_bo_b200.gaussianprocess import GaussianProcess
    from csp_bo_b200.utilities import r2, mae
    model = GaussianProcess(kernel=None)
    model.load(os.path.join(G, "si_model.json"), db_rows=asedb.rows_from_npz_dict(datasets, "si_model"))
    assert (model.N_energy, model.N_forces) == (25, 82) and model.kernel.name == "RBF_mb"
    E, E1, F, F1 = model.validate_data()
    assert r2(E, E1) > 0.999 and mae(E, E1) < 0.01 and r2(F, F1) > 0.99
    test = asedb.rows_from_npz_dict(datasets, "si244")[::6]
    Es, Es1, Fs, Fs1 = [], [], [], []
    for r in test:
        e, f, s = model.predict_structure(r["atoms"], stress=True)
        Es.append(r["energy"] / len(r["atoms"])); Es1.append(e / len(r["atoms"]))
        Fs.append(r["force"].ravel()); Fs1.append(f.ravel())
        assert s.shape == (len(r["atoms"]), 6) and np.all(np.isfinite(s))
    assert r2(np.array(Es), np.array(Es1)) > 0.99 and mae(np.array(Es), np.array(Es1)) < 0.02
    assert r2(np.concatenate(Fs), np.concatenate(Fs1)) > 0.9


def test_real_si_structures_against_cpu_oracle_chain(datasets, oracle_built):
    """Same model built twice from the first structures of Si.db: GPU chain vs (NumPy SO3 + reference C++
    kernels + SciPy); predictions on unseen Si_244 structures must agree to 1e-8."""
    from csp_bo_b200 import asedb
    from csp_bo_b200.SO3 import SO3
    from csp_bo_b200.gaussianprocess import GaussianProcess
    from csp_bo_b200.kernels import RBF_mb
    from csp_bo_b200.utilities import list_to_tuple
    from oracle import so3_reference as so
    from oracle.gp_reference import RefGaussianProcess
    rows = [r for r in asedb.rows_from_npz_dict(datasets, "si_model") if len(r["atoms"]) <= 24][:8]
    prm = dict(nmax=3, lmax=3, rcut=4.5, alpha=2.0)
    para, bounds = [6.244955524643115, 0.5611029935713949], [[0.01, 20.0], [0.1, 10.0]]
    model = GaussianProcess(RBF_mb(para=list(para), bounds=bounds, zeta=2), descriptor=SO3(derivative=True, stress=True, **prm),
                            f_coef=20, noise_e=[0.008142338423350038, 0.003, 0.01])
    for r in rows:
        r["energy_in"], r["force_in"] = True, [0, 1]
    model.extract_db(rows)
    model.fit(show=False, opt=False)
    # oracle chain
    E, F, y_e, y_f = [], [], [], []
    for r in rows:
        at = r["atoms"]
        d = so.so3_calculate(so.Structure(at.positions, at.numbers, at.cell, at.pbc), stress=True, **prm)
        E.append((d['x'], at.numbers)); y_e.append(r["energy"] / len(at))
        for fid in r["force_in"]:
            ids = np.argwhere(d['seq'][:, 1] == fid).flatten()
            _i = d['seq'][ids, 0]
            F.append((d['x'][_i], d['dxdr'][ids], at.numbers[_i])); y_f.append(r["force"][fid])
    ref = RefGaussianProcess("rbf", para, 2.0, bounds, f_coef=20, noise_e=(0.008142338423350038, 0.003, 0.01))
    ref.set_train({"energy": list_to_tuple(E, mode="energy"), "force": list_to_tuple(F)}, np.concatenate((y_e, np.concatenate(y_f))))
    ref.fit(opt=False)
    assert max_norm_err(model.alpha_, ref.alpha_) < 1e-7
    for r in [t for t in asedb.rows_from_npz_dict(datasets, "si244") if len(t["atoms"]) <= 20][:2]:
        at = r["atoms"]
        e, f, s = model.predict_structure(at, stress=True)
        d = so.so3_calculate(so.Structure(at.positions, at.numbers, at.cell, at.pbc), stress=True, **prm)
        er, fr, sr = ref.predict_structure(d, len(at), at.numbers, stress=True)
        assert abs(e - er) <= 1e-8 * max(1.0, abs(er))
        assert max_norm_err(f, fr) < 1e-8 and max_norm_err(s, sr) < 1e-8
