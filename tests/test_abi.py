"""CPU-side checks of the boundary: the C-ABI library builds, loads and exports every symbol that
include/cspbo_b200.h declares; the reference-named shims export the cffi names; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from csp_bo_b200 import build
    build.build()
    return True


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "cspbo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cspbo_[a-z0-9_]+)\s*\(", hdr)))


def test_header_symbols_exported(built):
    from csp_bo_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), n
    # the ctypes table covers exactly the header
    assert set(_lib.SIGNATURES) == set(names)


def test_shims_export_reference_names(built):
    from csp_bo_b200 import _lib
    _lib.load()
    for so, names in _lib.SHIM_SYMBOLS.items():
        lib = ctypes.CDLL(os.path.join(_lib.LIB_DIR, so))
        for n in names:
            assert hasattr(lib, n), (so, n)


def test_no_device_fails_loudly(built):
    """Without a GPU a compute call must raise, never fall back to the CPU."""
    from csp_bo_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    from csp_bo_b200.kernels.dot_kernel import kee_C
    x = np.ones((4, 24))
    with pytest.raises(_lib.CspboError):
        kee_C((x, np.ones(4, dtype=int), [2, 2]), (x, np.ones(4, dtype=int), [4]))


def test_product_never_imports_oracle():
    """The product package must not reference oracle/ (SURVEY 8c / task rule 3)."""
    pkg = os.path.join(ROOT, "csp_bo_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".c", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                for banned in ("import oracle", "from oracle", "cpu_reference", "libcspbo_oracle", "oracle/_ref"):
                    assert banned not in txt, (f, banned)


def test_list_tuple_roundtrip():
    from csp_bo_b200.utilities import list_to_tuple, tuple_to_list
    rng = np.random.default_rng(1)
    data = [(rng.standard_normal((n, 5)), rng.standard_normal((n, 5, 3)), np.full(n, 14)) for n in (3, 1, 4)]
    X, dX, ELE, ind = list_to_tuple(data)
    assert X.shape == (8, 5) and dX.shape == (8, 5, 3) and ind == [3, 1, 4] and len(ELE) == 8
    back = tuple_to_list((X, dX, ELE, ind))
    for (a, b, c), (d, e, f) in zip(data, back):
        assert np.array_equal(a, d) and np.array_equal(b, e) and np.array_equal(c, f)
    e = [(rng.standard_normal((n, 5)), np.full(n, 8)) for n in (2, 2)]
    Xe, Ee, inde = list_to_tuple(e, mode='energy')
    assert Xe.shape == (4, 5) and inde == [2, 2]


def test_build_covariance_layouts():
    from csp_bo_b200.kernels.base import build_covariance, get_mask
    ee, ef, fe, ff = np.ones((2, 2)), 2 * np.ones((2, 6)), 3 * np.ones((6, 2)), 4 * np.ones((6, 6))
    K = build_covariance(ee, ef, fe, ff)
    assert K.shape == (8, 8) and K[0, 0] == 1 and K[0, 2] == 2 and K[2, 0] == 3 and K[7, 7] == 4
    assert build_covariance(None, None, fe, ff).shape == (6, 8)
    assert build_covariance(ee, ef, None, None).shape == (2, 8)
    assert build_covariance(None, None, None, None) is None
    assert get_mask(np.array([1, 1]), np.array([1, 1])) is None
    assert len(get_mask(np.array([1, 8]), np.array([1, 1]))[0]) == 2


def test_kernel_object_api():
    from csp_bo_b200.kernels import Dot_mb, RBF_mb
    k = Dot_mb()          # must be default-constructible (gaussianprocess.py:558-567)
    k.load_from_dict({"sigma": 2.0, "sigma0": 0.1, "zeta": 2, "bounds": [[1, 2], [3, 4]], "name": "Dot_mb"})
    assert k.parameters() == [2.0, 0.1] and k.save_dict()["zeta"] == 2
    assert str(k) == "2.000**2 *Dot(length=0.100)"
    r = RBF_mb(para=[3.0, 0.5])
    r.update([4.0, 0.25])
    assert r.parameters() == [4.0, 0.25] and r.save_dict()["l"] == 0.25
    assert str(r) == "4.000**2 *RBF(length=0.250)"


def test_gpr_calculator_glue():
    """csp_bo_b200.calculator.GPR (cspbo/calculator.py:9-64) with a stand-in model: results keys, stress summation over
    atoms, variances, getters, and the property getters of the ase-free base."""
    from csp_bo_b200.calculator import GPR

    class FF:
        def __init__(self):
            self.calls = []

        def predict_structure(self, struc, stress, return_std, f_tol=1e-8):
            self.calls.append((stress, return_std, f_tol))
            n = len(struc)
            E, F = -3.0 * n, np.arange(3 * n, dtype=float).reshape(n, 3)
            S = np.ones((n, 6)) if stress else None
            if return_std:
                return E, F, S, 0.25, np.full((n, 3), 0.5)
            return E, F, S

    atoms = [0] * 4
    ff = FF()
    calc = GPR(ff=ff, stress=True, return_std=True, f_tol=1e-10)
    calc.calculate(atoms)
    assert ff.calls[-1] == (True, True, 1e-10)
    r = calc.results
    assert r['energy'] == r['free_energy'] == -12.0 and r['forces'].shape == (4, 3)
    assert np.array_equal(r['stress'], np.full(6, 4.0))
    assert calc.get_var_e() == 0.25 and calc.get_var_e(total=True) == 1.0 and calc.get_var_f().shape == (4, 3)
    assert calc.get_e() == -3.0 and calc.get_e(peratom=False) == -12.0
    calc2 = GPR(ff=ff)                               # defaults: no stress, no std, f_tol 1e-12 (calculator.py:22-35)
    assert calc2.get_potential_energy(atoms) == -12.0 and ff.calls[-1] == (False, False, 1e-12)
    assert calc2.results['stress'] is None and 'var_e' not in calc2.results
    assert calc2.get_forces().shape == (4, 3)
