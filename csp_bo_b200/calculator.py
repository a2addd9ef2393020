"""`GPR`: the calculator that drives molecular dynamics / relaxations with the GPU Gaussian process
(reference cspbo/calculator.py:9-64).  Same constructor keywords (`ff`, `stress`, `return_std`, `f_tol`), same
`results` keys ('energy', 'free_energy', 'forces', 'stress', 'var_e', 'var_f') and helper getters.

With ase installed the class derives from `ase.calculators.calculator.Calculator` and plugs into ase's optimisers and
MD drivers unchanged; without ase (this image) a minimal base with the same calling convention is used, so the
on-the-fly MD loop of BASELINE.json config 4 runs either way:

    calc = GPR(ff=model, stress=False, return_std=True)      # model: csp_bo_b200.gaussianprocess.GaussianProcess
    calc.calculate(atoms)                                     # atoms: anything predict_structure accepts
    E, F = calc.results['energy'], calc.results['forces']
"""
from types import SimpleNamespace

try:                                                # pragma: no cover - ase is absent in the build image
    from ase.calculators.calculator import Calculator as _Base, all_changes
    _HAVE_ASE = True
except Exception:
    _HAVE_ASE = False
    all_changes = ['positions', 'numbers', 'cell', 'pbc', 'initial_charges', 'initial_magmoms']

    class _Base:
        """The part of ase's Calculator that GPR relies on: keyword parameters as attributes, a results dict, the atoms
        the results belong to, and property getters that (re)calculate on demand."""

        def __init__(self, **kwargs):
            self.parameters = SimpleNamespace(**kwargs)
            self.results = {}
            self.atoms = None

        def calculate(self, atoms=None, properties=('energy',), system_changes=all_changes):
            if atoms is not None:
                self.atoms = atoms

        def get_property(self, name, atoms=None):
            if atoms is not None or name not in self.results:
                self.calculate(atoms if atoms is not None else self.atoms, [name])
            return self.results[name]

        def get_potential_energy(self, atoms=None):
            return self.get_property('energy', atoms)

        def get_forces(self, atoms=None):
            return self.get_property('forces', atoms)

        def get_stress(self, atoms=None):
            return self.get_property('stress', atoms)

eV2GPa = 160.21766


class GPR(_Base):
    implemented_properties = ['energy', 'forces', 'stress', 'var_e', 'var_f']
    nolabel = True

    def __init__(self, **kwargs):
        _Base.__init__(self, **kwargs)
        self.results = {}

    def calculate(self, atoms=None, properties=['energy'], system_changes=all_changes):
        """calculator.py:17-49: one predict_structure call fills every result."""
        _Base.calculate(self, atoms, properties, system_changes)
        par = self.parameters
        stress = par.stress if hasattr(par, 'stress') else False
        f_tol = par.f_tol if hasattr(par, 'f_tol') else 1e-12
        return_std = par.return_std if hasattr(par, 'return_std') else False
        struc = atoms if atoms is not None else self.atoms
        if return_std:
            res = par.ff.predict_structure(struc, stress, True, f_tol=f_tol)
            self.results['var_e'] = res[3]
            self.results['var_f'] = res[4]
        else:
            res = par.ff.predict_structure(struc, stress, False, f_tol=f_tol)
        self.results['energy'] = res[0]
        self.results['free_energy'] = res[0]
        self.results['forces'] = res[1]
        self.results['stress'] = res[2].sum(axis=0) if stress else None

    def get_var_e(self, total=False):
        """calculator.py:51-55: eV per atom, or for the whole cell with total=True."""
        if total:
            return self.results["var_e"] * len(self.results["forces"])
        return self.results["var_e"]

    def get_var_f(self):
        return self.results["var_f"]

    def get_e(self, peratom=True):
        if peratom:
            return self.results["energy"] / len(self.results["forces"])
        return self.results["energy"]


class LJ():
    """Pairwise shifted Lennard-Jones base potential (reference cspbo/calculator.py:68-177, which copies ase's `lj.py`):
    `calculate(atoms)` -> (energy, forces [n,3], stress [n,6] | None) that the GP subtracts from its training targets and
    adds back to its predictions (gaussianprocess.py:457-462,686-692).  Same constructor dictionary, __str__,
    load_from_dict and save_dict; the pair sums run on the GPU (cspbo_lj_compute: linked-cell neighbour search + one
    kernel) and ase is not needed."""

    def __init__(self, parameters=None):
        _parameters = {'name': 'LJ', 'rc': 5.0, 'sigma': 1.0, 'epsilon': 1.0}
        if parameters is not None:
            _parameters.update(parameters)
        self.load_from_dict(_parameters)

    def __str__(self):
        return "LJ(eps: {:.3f}, sigma: {:.3f}, cutoff: {:.3f})".format(self.epsilon, self.sigma, self.rc)

    def load_from_dict(self, dict0):
        self._parameters = dict0
        self.name = self._parameters["name"]
        self.epsilon = self._parameters["epsilon"]
        self.sigma = self._parameters["sigma"]
        self.rc = self._parameters["rc"]

    def save_dict(self):
        return self._parameters

    def calculate(self, atoms):
        import ctypes
        import numpy as np
        from . import _lib
        c_vp = ctypes.c_void_p
        pos = np.ascontiguousarray(atoms.positions, dtype=np.float64)
        cell = np.ascontiguousarray(np.asarray(atoms.get_cell()), dtype=np.float64)
        pbc = np.broadcast_to(np.asarray(atoms.pbc), (3,)).astype(np.int32)
        periodic = np.linalg.matrix_rank(cell) == 3          # ase: atoms.number_of_lattice_vectors == 3
        if not periodic:                                     # a molecule: any non-singular frame, no images
            cell_s, pbc = np.eye(3), np.zeros(3, dtype=np.int32)
        else:
            cell_s = cell
        n = len(pos)
        e, f, s = np.zeros(n), np.zeros((n, 3)), np.zeros((n, 3, 3))
        _lib.check(_lib.load().cspbo_lj_compute(n, c_vp(pos.ctypes.data), c_vp(np.ascontiguousarray(cell_s).ctypes.data),
                                                c_vp(np.ascontiguousarray(pbc).ctypes.data), float(self.rc), float(self.sigma),
                                                float(self.epsilon), c_vp(e.ctypes.data), c_vp(f.ctypes.data), c_vp(s.ctypes.data)))
        stress = None
        if periodic:
            # ase.constraints.full_3x3_to_voigt_6_stress: (xx, yy, zz, yz, xz, xy), off-diagonals symmetrised
            voigt = np.stack((s[:, 0, 0], s[:, 1, 1], s[:, 2, 2], (s[:, 1, 2] + s[:, 2, 1]) / 2, (s[:, 0, 2] + s[:, 2, 0]) / 2,
                              (s[:, 0, 1] + s[:, 1, 0]) / 2), axis=1)
            stress = voigt / abs(np.linalg.det(cell))
        return float(e.sum()), f, stress
