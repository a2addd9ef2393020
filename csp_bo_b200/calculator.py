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
