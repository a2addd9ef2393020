"""SO(3) power-spectrum descriptor, GPU-backed drop-in for the reference's cspbo/SO3.py.

`SO3(nmax, lmax, rcut, alpha, derivative, stress, cutoff_function, weight_on, primitive).calculate(atoms)`
returns the dictionary the GP consumes (SO3.py:274-291):
    {'x': [n, nc], 'dxdr': [n_seq, nc, 3], 'rdxdr': [n_seq, nc, 3, 3] | None, 'seq': [n_seq, 2], 'elements': [...]}
`atoms` is anything with .positions, .numbers (or .get_atomic_numbers()), .get_cell(), .pbc and
.symbols / .get_chemical_symbols() -- e.g. an ase.Atoms object; ase itself is not needed, the
neighbour search runs inside libcspbo_b200 (csrc/so3.cu).
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_vp

_CUTOFFS = {'cosine': 0, 'tanh': 1, 'poly1': 2, 'poly2': 3, 'poly3': 4, 'poly4': 5, 'unity': 6}


class SO3:
    def __init__(self, nmax=3, lmax=3, rcut=3.5, alpha=2.0, derivative=True, stress=False,
                 cutoff_function='cosine', weight_on=False, primitive=False):
        self.nmax, self.lmax, self.rcut, self.alpha = nmax, lmax, rcut, alpha
        self.derivative, self.stress = derivative, stress
        self._type = "SO3"
        if cutoff_function not in _CUTOFFS:
            raise NotImplementedError("The requested cutoff function has not been implemented")
        self.cutoff_function = cutoff_function
        self.weight_on = weight_on
        self.primitive = primitive          # kept for API parity; the built-in neighbour search is always used
        self._h = None

    def __str__(self):
        s = "SO3 descriptor with Cutoff: {:6.3f}".format(self.rcut)
        s += " lmax: {:d}, nmax: {:d}, alpha: {:.3f}\n".format(self.lmax, self.nmax, self.alpha)
        return s

    def __repr__(self):
        return str(self)

    def load_from_dict(self, dict0):
        self.nmax, self.lmax, self.rcut, self.alpha = dict0["nmax"], dict0["lmax"], dict0["rcut"], dict0["alpha"]
        self.derivative, self.stress = dict0["derivative"], dict0["stress"]
        self._release()

    def save_dict(self):
        return {"nmax": self.nmax, "lmax": self.lmax, "rcut": self.rcut, "alpha": self.alpha,
                "derivative": self.derivative, "stress": self.stress, "_type": "SO3"}

    # ------------------------------------------------------------------
    def _release(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            _lib.load().cspbo_so3_destroy(h[0])

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _handle(self):
        key = (self.nmax, self.lmax, float(self.rcut), float(self.alpha), self.cutoff_function, bool(self.weight_on))
        if self._h is None or self._h[1] != key:
            self._release()
            h = c_vp()
            _lib.check(_lib.load().cspbo_so3_create(int(self.nmax), int(self.lmax), float(self.rcut), float(self.alpha),
                                                    _CUTOFFS[self.cutoff_function], int(bool(self.weight_on)),
                                                    ctypes.byref(h)))
            self._h = (h, key)
        return self._h[0]

    def _compute(self, atoms, want_stress):
        """Runs the descriptor kernels; returns (handle, n_atoms, n_seq, n_coefs, elements)."""
        lib = _lib.load()
        pos = np.ascontiguousarray(atoms.positions, dtype=np.float64)
        numbers = getattr(atoms, "numbers", None)
        if numbers is None:
            numbers = atoms.get_atomic_numbers()
        numbers = np.ascontiguousarray(numbers, dtype=np.int32)
        cell = np.ascontiguousarray(np.asarray(atoms.get_cell()), dtype=np.float64)
        pbc = np.ascontiguousarray(np.broadcast_to(np.asarray(atoms.pbc), (3,)).astype(np.int32))
        symbols = getattr(atoms, "symbols", None)
        elements = list(symbols) if symbols is not None else list(atoms.get_chemical_symbols())
        h = self._handle()
        nseq, nc = ctypes.c_int(0), ctypes.c_int(0)
        _lib.check(lib.cspbo_so3_compute(h, len(pos), c_vp(pos.ctypes.data), c_vp(numbers.ctypes.data), c_vp(cell.ctypes.data),
                                         c_vp(pbc.ctypes.data), int(want_stress), ctypes.byref(nseq), ctypes.byref(nc)))
        return h, len(pos), nseq.value, nc.value, elements

    def calculate_device(self, atoms, stress=None):
        """Like calculate(), but x / dxdr / rdxdr are CUDA tensors viewing the descriptor's own device buffers (valid
        until the next calculate* call on this object) -- nothing but `seq` crosses PCIe.  What
        GaussianProcess.predict_structure consumes on the MD path (gaussianprocess.py:335-351)."""
        import torch
        from .distributed import _RawCuda
        lib = _lib.load()
        want_stress = bool(self.stress if stress is None else stress) and bool(self.derivative)
        h, n, nseq, nc, elements = self._compute(atoms, want_stress)
        px, pd, pr = c_vp(), c_vp(), c_vp()
        _lib.check(lib.cspbo_so3_device_ptrs(h, ctypes.byref(px), ctypes.byref(pd), ctypes.byref(pr)))
        seq = np.empty((nseq, 2), dtype=np.int64)
        _lib.check(lib.cspbo_so3_fetch(h, None, None, None, c_vp(seq.ctypes.data)))
        x = torch.as_tensor(_RawCuda(px.value, (n, nc)), device="cuda")
        dxdr = torch.as_tensor(_RawCuda(pd.value, (nseq, nc, 3)), device="cuda")
        rdxdr = torch.as_tensor(_RawCuda(pr.value, (nseq, nc, 3, 3)), device="cuda") if want_stress else None
        return {'x': x, 'dxdr': dxdr, 'rdxdr': rdxdr, 'elements': elements, 'seq': seq}

    def calculate(self, atoms, atom_ids=None):
        if atom_ids is not None:
            raise NotImplementedError("atom_ids subsets are not supported by the GPU descriptor")
        lib = _lib.load()
        want_stress = bool(self.stress) and bool(self.derivative)
        h, n, nseq_v, nc_v, elements = self._compute(atoms, want_stress)
        nseq, nc = ctypes.c_int(nseq_v), ctypes.c_int(nc_v)
        x = np.empty((n, nc.value))
        if not self.derivative:
            _lib.check(lib.cspbo_so3_fetch(h, c_vp(x.ctypes.data), None, None, None))
            return {'x': x, 'dxdr': None, 'rdxdr': None, 'elements': elements}
        dxdr = np.empty((nseq.value, nc.value, 3))
        rdxdr = np.empty((nseq.value, nc.value, 3, 3)) if want_stress else None
        seq = np.empty((nseq.value, 2), dtype=np.int64)
        _lib.check(lib.cspbo_so3_fetch(h, c_vp(x.ctypes.data), c_vp(dxdr.ctypes.data),
                                       c_vp(rdxdr.ctypes.data) if want_stress else None, c_vp(seq.ctypes.data)))
        return {'x': x, 'dxdr': dxdr, 'rdxdr': rdxdr, 'elements': elements, 'seq': seq}
