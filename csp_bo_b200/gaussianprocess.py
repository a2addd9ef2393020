"""GPU-resident Gaussian-process regressor: the fit / prediction / log-marginal-likelihood slice of the
reference's `cspbo.gaussianprocess.GaussianProcess` (gaussianprocess.py:71-184, 325-390, 453-509) on top of
the CUDA covariance kernels, plus remove_train_pts / sparsify / CUR (:230-270, :754-790).  Same method
names, arguments and return values (load / load_from_dict / extract_db read the reference's json + ASE-sqlite
files through csp_bo_b200.asedb); what is NOT carried over is what writes ASE databases or needs pyxtal
(save, export_ase_db, add_structure): those stay in the reference.

Under torch.distributed (one process per GPU, torchrun) fit(opt=False) builds K across the ranks and factorises it
with the sharded Cholesky (csp_bo_b200/distributed.py: fit_sharded); every rank keeps alpha and the factor.

Device residency: K is assembled block by block straight into one CUDA matrix
(kernels/device.py), the noise is added and K is factorised in place by cuSOLVER
(csrc/gp.cu), `alpha_` and the Cholesky factor stay in HBM, predictions are K*.alpha GEMVs on the
device.  Only alpha-sized vectors and predictions cross PCIe.
"""
import json
import os
import warnings

import numpy as np
from scipy.optimize import minimize

from . import _lib, gp
from .kernels import device as kdev
from .utilities import list_to_tuple, tuple_to_list

# atomic numbers for descriptor.calculate() results that carry chemical symbols
# (gaussianprocess.py:335 uses pyxtal's Element(ele).z)
_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
            "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu "
            "Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()
_Z = {s: i for i, s in enumerate(_SYMBOLS)}


def _world_size():
    import torch
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        return torch.distributed.get_world_size()
    return 1


class GaussianProcess():
    """Gaussian-process regressor on energies and forces.

    Args (as in the reference, gaussianprocess.py:26-29):
        kernel: csp_bo_b200.kernels.Dot_mb | RBF_mb
        descriptor: object with .calculate(struc) -> {'x','dxdr','rdxdr','seq','elements'} (only needed by
                    predict_structure)
        base_potential: optional object with .calculate(atoms) -> (E, F, S) offsets
        f_coef: force noise = f_coef * energy noise
        noise_e: [value, lower bound, upper bound]
    """

    def __init__(self, kernel=None, descriptor=None, base_potential=None, f_coef=5, noise_e=[5e-3, 2e-3, 1e-1]):
        # kernel / descriptor default to None so that `gpr()` followed by load() works like the constructor the
        # reference's example scripts rely on (gaussianprocess_kang.py:28, examples/example_validate.py:20)
        self.noise_e = noise_e[0]
        self.f_coef = f_coef
        self.noise_f = self.f_coef * self.noise_e
        self.noise_bounds = noise_e[1:]
        self.descriptor = descriptor
        self.kernel = kernel
        self.base_potential = base_potential
        self.x = None
        self.train_x = None
        self.train_y = None
        self.train_db = None
        self.alpha_ = None
        self.L_ = None            # CUDA tensor: cuSOLVER factor of K + noise (its lower triangle, column-major)
        self._alpha_dev = None
        self._packs = None        # the training set as resident packs, snapshot taken by fit()
        self._lml_cache = None    # Dot_mb: (packs, zeta, K(sigma = 1, sigma0 = 0), P) of the hyper-optimisation (_dot_scaled_K)
        self._packs_pred = None   # the same for K* builds (small energy sets spread over the tile slots)
        self.N_energy = 0
        self.N_forces = 0
        self.N_energy_queue = 0
        self.N_forces_queue = 0

    def __str__(self):
        s = "------Gaussian Process Regression------\n"
        s += "Kernel: {:s}".format(str(self.kernel))
        if hasattr(self, "train_x"):
            s += " {:d} energy ({:.3f})".format(self.N_energy, self.noise_e)
            s += " {:d} forces ({:.3f})\n".format(self.N_forces, self.noise_f)
        return s

    def __repr__(self):
        return str(self)

    # ------------------------------------------------------------------ training data
    def set_train_pts(self, data, mode="w"):
        """data: {'energy': [(x, E, ele)], 'force': [(x, dxdr, F, ele)], 'db': [...] (optional)}
        (gaussianprocess.py:186-229)."""
        if mode == "w" or self.train_x is None:
            self.train_x = {'energy': [], 'force': []}
            self.train_y = {'energy': [], 'force': []}
            self.train_db = []
            self.N_energy = self.N_forces = 0
        N_E, N_F = 0, 0
        for d in data.get("db", []):
            (atoms, energy, force, energy_in, force_in) = d
            e_id = N_E + 1 if energy_in else None
            N_E += 1 if energy_in else 0
            f_ids = [N_F + i for i in range(len(force_in))]
            N_F += len(force_in)
            self.train_db.append((atoms, energy, force, energy_in, force_in, e_id, f_ids))
        if "db" not in data:
            N_E, N_F = len(data.get("energy", [])), len(data.get("force", []))
        if len(data.get("energy", [])) > 0:
            self.add_train_pts_energy(data["energy"])
        if len(data.get("force", [])) > 0:
            self.add_train_pts_force(data["force"])
        self.update_y_train()
        self.N_energy += N_E
        self.N_forces += N_F
        self.N_energy_queue += N_E
        self.N_forces_queue += N_F

    def add_train_pts_energy(self, energy_data):
        """energy_data: list of (X[n,d], E, ele)  (gaussianprocess.py:407-425)."""
        (X, ELE, indices, E) = list_to_tuple(energy_data, include_value=True, mode='energy')
        if len(self.train_x['energy']) == 3:
            (_X, _ELE, _indices) = self.train_x['energy']
            self.train_x['energy'] = (np.concatenate((_X, X), axis=0), np.concatenate((_ELE, ELE), axis=0),
                                      list(_indices) + list(indices))
            self.train_y['energy'] = list(self.train_y['energy']) + list(E)
        else:
            self.train_x['energy'] = (X, ELE, indices)
            self.train_y['energy'] = list(E)

    def add_train_pts_force(self, force_data):
        """force_data: list of (X[n,d], dXdR[n,d,3], F[3], ele)  (gaussianprocess.py:427-451)."""
        (X, dXdR, ELE, indices, F) = list_to_tuple(force_data, include_value=True)
        if len(self.train_x['force']) == 4:
            (_X, _dXdR, _ELE, _indices) = self.train_x['force']
            self.train_x['force'] = (np.concatenate((_X, X), axis=0), np.concatenate((_dXdR, dXdR), axis=0),
                                     np.concatenate((_ELE, ELE), axis=0), list(_indices) + list(indices))
            self.train_y['force'] = list(self.train_y['force']) + list(F)
        else:
            self.train_x['force'] = (X, dXdR, ELE, indices)
            self.train_y['force'] = list(F)

    def update_y_train(self):
        """Stack the targets: energies first, then forces atom-major (gaussianprocess.py:272-288)."""
        E = np.asarray(self.train_y["energy"], dtype=np.float64).reshape(-1)
        F = np.asarray(self.train_y["force"], dtype=np.float64).reshape(-1)
        self.y_train = np.concatenate((E, F)).reshape(-1, 1)

    def get_train_x(self):
        return self.train_x

    def _n_energy(self):
        return len(self.train_x['energy'][-1]) if len(self.train_x['energy']) > 0 else 0

    def _train_dict(self):
        return {k: v for k, v in self.train_x.items() if len(v) > 0}

    def _fitted(self):
        """The training set alpha_ belongs to: the packs fit() built (HBM-resident, one upload per fit).  Points
        queued by set_train_pts(mode='a+') after the last fit are not in it -- the reference's get_train_x()
        (gaussianprocess.py:392-405) drops them the same way -- so K* always has as many columns as alpha_ has rows."""
        if self._packs is None:
            raise RuntimeError("GaussianProcess: fit() has not been called")
        return self._packs_pred

    def remove_train_pts(self, e_ids, f_ids):
        """Drop the energy points `e_ids` and force points `f_ids` and refit (gaussianprocess.py:230-270;
        the ASE-database bookkeeping of the reference is kept only as far as train_db entries exist)."""
        data = {"energy": [], "force": []}
        energy_data = tuple_to_list(self.train_x['energy'], mode='energy') if len(self.train_x['energy']) else []
        force_data = tuple_to_list(self.train_x['force']) if len(self.train_x['force']) else []
        for i, (X, ele) in enumerate(energy_data):
            if i not in e_ids:
                data['energy'].append((X, self.train_y['energy'][i], ele))
        for i, (X, dxdr, ele) in enumerate(force_data):
            if i not in f_ids:
                data["force"].append((X, dxdr, self.train_y['force'][i], ele))
        if self.train_db:
            data["db"] = []
            for (atoms, energy, force, energy_in, force_in, e_id, _f_ids) in self.train_db:
                if e_id in e_ids:
                    energy_in = False
                _force_in = [force_in[i] for i, f_id in enumerate(_f_ids) if f_id not in f_ids]
                if energy_in or len(_force_in) > 0:
                    data['db'].append((atoms, energy, force, energy_in, _force_in))
        self.set_train_pts(data)
        self.fit(show=False, opt=False)

    def sparsify(self, e_tol=1e-10, f_tol=1e-10):
        """Remove training points that the CUR decomposition marks as redundant
        (gaussianprocess.py:754-772)."""
        K = kdev.k_total_device(self.kernel, self._train_dict())
        N_e = self._n_energy()
        N_f = len(self.train_x["force"][-1]) if len(self.train_x["force"]) else 0
        pts_e = CUR(K[:N_e, :N_e], e_tol) if N_e else np.array([], dtype=int)
        pts = CUR(K[N_e:, N_e:], f_tol) if N_f else np.array([], dtype=int)
        pts_f = []
        if N_f > 1:
            sel = set(int(v) for v in pts)
            pts_f = [i for i in range(N_f) if {3 * i, 3 * i + 1, 3 * i + 2} <= sel]
        print("{:d} energy and {:d} force points will be removed".format(len(pts_e), len(pts_f)))
        if len(pts_e) + len(pts_f) > 0:
            self.remove_train_pts(list(int(v) for v in pts_e), pts_f)
        return list(int(v) for v in pts_e), pts_f

    # ------------------------------------------------------------------ fit
    def fit(self, TrainData=None, show=True, opt=True):
        """gaussianprocess.py:71-131."""
        if TrainData is not None:
            self.set_train_pts(TrainData)
        else:
            self.update_y_train()
        if show:
            print(self)

        def obj_func(params, eval_gradient=True):
            if eval_gradient:
                lml, grad = self.log_marginal_likelihood(params, eval_gradient=True, clone_kernel=False)
                if show:
                    print("Loss: {:12.3f} ".format(-lml) + " ".join("{:6.3f}".format(p) for p in params))
                return (-lml, -grad)
            return -self.log_marginal_likelihood(params, clone_kernel=False)

        from .packing import SPREAD_MAX_GROUPS, resident, spread_energy
        self._packs = resident(self._train_dict())     # the one upload of the training descriptors of this fit
        # prediction side: a training set with few energy structures is packed spread over the tile slots
        self._packs_pred = dict(self._packs)
        if "energy" in self._packs and self._packs["energy"].m <= SPREAD_MAX_GROUPS:
            self._packs_pred["energy"] = spread_energy(self.train_x["energy"])
        hyper_params = self.kernel.parameters() + [self.noise_e]
        hyper_bounds = self.kernel.bounds + [self.noise_bounds]
        if opt:
            params, _ = self.optimize(obj_func, hyper_params, hyper_bounds)
            self._lml_cache = None            # the optimisation's cached unit-sigma K (Dot_mb) is not needed any longer
            self.kernel.update(params[:-1])
            self.noise_e = params[-1]
            self.noise_f = self.f_coef * params[-1]

        try:
            if _world_size() > 1:
                # one process per GPU (torchrun): K is built in the panel layout across the ranks and factorised by the
                # sharded Cholesky; every rank ends up with alpha and the replicated factor (csp_bo_b200/distributed.py)
                from .distributed import fit_sharded
                alpha, self._solver = fit_sharded(self.kernel, self._packs, self.y_train[:, 0], self.noise_e,
                                                  self.f_coef, return_solver=True)
                K, logdet = None, self._solver.logdet()
            else:
                self._solver = None
                K = kdev.k_total_device(self.kernel, self._packs)                 # CUDA [N,N]
                alpha, logdet = gp.fit_alpha(K, self.y_train[:, 0], n_energy=self._n_energy(), noise_e=self.noise_e,
                                             f_coef=self.f_coef, return_logdet=True, overwrite=True)
        except _lib.CspboError as exc:
            if exc.code == 5:    # CSPBO_E_NOTPD, the LinAlgError of gaussianprocess.py:118-125
                raise np.linalg.LinAlgError("The kernel, %s, is not returning a positive definite matrix. Try gradually "
                                            "increasing the noise." % self.kernel) from exc
            raise
        self.L_ = K                       # factorised in place
        self._alpha_dev = alpha
        self.alpha_ = alpha.cpu().numpy().reshape(-1, 1)
        self._logdet = logdet
        self.N_energy_queue = 0
        self.N_forces_queue = 0
        return self

    def optimize(self, fun, theta0, bounds):
        opt_res = minimize(fun, theta0, method="L-BFGS-B", bounds=bounds, jac=True,
                           options={'maxiter': 10, 'ftol': 1e-2})
        return opt_res.x, opt_res.fun

    def _dot_scaled_K(self, kernel, data):
        """Dot_mb only: the training covariance of k_total_with_grad is K(sigma, sigma0) = sigma^2 (K1 + sigma0^2 P) with
        K1 = K(1, 0) and P_ij = (valid same-species row pairs of structures i, j) / (N_i N_j) on the energy block
        (dot_kernel.cpp:45-47, dot_kernel.py:45,129-130,260): the descriptor-dependent part does not change while L-BFGS-B
        moves (sigma, sigma0, noise), so it is built ONCE per training set and every further evaluation is a 1 ms
        rescale instead of a 104 ms build (N = 20 001).  Costs one more resident N x N array; skipped (returns None)
        when that does not fit comfortably or with CSPBO_LML_CACHE=0."""
        import os
        import torch
        from .kernels import Dot_mb
        if os.environ.get("CSPBO_LML_CACHE", "1") == "0":
            return None
        c = getattr(self, "_lml_cache", None)
        if c is None or c[0] is not data or c[1] != float(kernel.zeta):
            self._lml_cache = None
            N = int(self.y_train.shape[0])
            free, _ = torch.cuda.mem_get_info()
            if free < 3 * 8 * N * N:
                return None
            unit = Dot_mb(para=[1.0, 0.0], zeta=kernel.zeta)
            K1 = kdev.k_total_device(unit, data, grad_semantics=True)
            P = None
            if "energy" in data:
                one = Dot_mb(para=[1.0, 1.0], zeta=kernel.zeta)
                NE = self._n_energy()
                P = kdev.k_total_device(one, {"energy": data["energy"]}, grad_semantics=True) - K1[:NE, :NE]
            c = self._lml_cache = (data, float(kernel.zeta), K1, P)
        s2 = float(kernel.sigma) ** 2
        K = torch.mul(c[2], s2)
        if c[3] is not None:
            NE = c[3].shape[0]
            K[:NE, :NE].add_(c[3], alpha=s2 * float(kernel.sigma0) ** 2)
        return K

    def log_marginal_likelihood(self, params, eval_gradient=False, clone_kernel=False):
        """gaussianprocess.py:453-503, evaluated on the device."""
        import torch
        kernel = self.kernel
        kernel.update(params[:-1])
        if self._packs is not None:
            # the packs are fit()'s snapshot: points added since (set_train_pts without a refit) make them stale for the
            # likelihood of the CURRENT training set, which is what the reference evaluates (gaussianprocess.py:456-464)
            rows = (self._packs["energy"].m_real if "energy" in self._packs else 0) + \
                   (3 * self._packs["force"].m if "force" in self._packs else 0)
            if rows != int(self.y_train.shape[0]):
                self._packs = None
        if self._packs is None:
            from .packing import resident
            self._packs = resident(self._train_dict())
        data = self._packs
        if _world_size() > 1:
            # one process per GPU: K distributed, sharded Cholesky, gradient traces per owned panel (distributed.lml_sharded)
            from .distributed import lml_sharded
            cache = None
            if kernel.name == 'Dot_mb' and os.environ.get("CSPBO_LML_CACHE", "1") != "0":
                # the sharded twin of _dot_scaled_K: every rank keeps its rows of K(1, 0) while L-BFGS-B runs
                c = self._lml_cache
                if not (isinstance(c, tuple) and c[0] is data and isinstance(c[1], dict)):
                    c = self._lml_cache = (data, {})
                cache = c[1]
            return lml_sharded(kernel, data, self.y_train[:, 0], params[-1], self.f_coef, eval_gradient=eval_gradient, cache=cache)
        K = self._dot_scaled_K(kernel, data) if (eval_gradient and kernel.name == 'Dot_mb') else None
        if K is None:
            K = kdev.k_total_device(kernel, data, grad_semantics=eval_gradient)   # the ONE N x N array of the evaluation
        N, NE = K.shape[0], self._n_energy()
        y = torch.from_numpy(np.ascontiguousarray(self.y_train[:, 0])).cuda()
        try:
            alpha, logdet = gp.fit_alpha(K, y, n_energy=NE, noise_e=params[-1], f_coef=self.f_coef,
                                         return_logdet=True, overwrite=True)
        except _lib.CspboError as exc:
            if exc.code == 5:
                return (-np.inf, np.zeros_like(params)) if eval_gradient else -np.inf
            raise
        MLL = float(-0.5 * torch.dot(y, alpha) - 0.5 * logdet - N / 2 * np.log(2 * np.pi))
        if not eval_gradient:
            return MLL
        # 0.5 tr((alpha alpha^T - Ky^-1) dKy/dtheta)  (GPML eq. 5.9, gaussianprocess.py:490-499) as FUSED TRACES: neither the
        # [N,N,3] K_gradient nor alpha alpha^T is formed and K^-1 only where a kernel really needs it; K's buffer (now the
        # factor) is the only N x N array of the whole evaluation.
        ne, nf = float(params[-1]), float(self.f_coef * params[-1])
        d = torch.full((N,), nf * nf, dtype=torch.float64, device="cuda")            # noise on the diagonal of Ky
        d[:NE] = ne * ne
        b = torch.full((N,), 2 * self.f_coef ** 2 * ne, dtype=torch.float64, device="cuda")   # dKy/dnoise (diagonal)
        b[:NE] = 2 * ne
        a2 = alpha * alpha
        if kernel.name == 'Dot_mb':
            # dK/dsigma = 2 K/sigma with K = Ky - D  =>  tr = (2/sigma) [alpha.(y - D alpha) - (N - sum_i d_i (Ky^-1)_ii)]
            # dK/dsigma0 = c0 1_E 1_E^T (dot_kernel.py:57, constant as is)  =>  tr = c0 [(sum_E alpha)^2 - 1_E^T Ky^-1 1_E]
            g_s0 = 0.0
            if NE > 0:
                ones = torch.zeros((N, 1), dtype=torch.float64, device="cuda")
                ones[:NE] = 1.0
                u = gp.cho_solve(K, ones)[:, 0]                  # before the factor is destroyed below
                c0 = 0.8 * 2 * kernel.sigma ** 2 * kernel.sigma0
                g_s0 = 0.5 * c0 * float(alpha[:NE].sum() ** 2 - u[:NE].sum())
            dinv = gp.inverse_diag(K)                            # diag(Ky^-1); K now holds L^-1
            g_s = (1.0 / kernel.sigma) * float(torch.dot(alpha, y) - torch.dot(d, a2) - N + torch.dot(d, dinv))
            llg = [g_s, g_s0]
        else:
            gp.cholesky_inverse_inplace(K)                       # K now holds Ky^-1 (upper triangle), in place
            dinv = torch.diagonal(K).clone()
            g_s = (1.0 / kernel.sigma) * float(torch.dot(alpha, y) - torch.dot(d, a2) - N + torch.dot(d, dinv))
            g_noise_now = 0.5 * float(torch.dot(b, a2 - dinv))   # before rbf_dl_trace rescales alpha / K^-1
            g_l = 0.5 * float(kdev.rbf_dl_trace(kernel, data, alpha.clone(), K))
            return MLL, np.array([g_s, g_l, g_noise_now])
        llg.append(0.5 * float(torch.dot(b, a2 - dinv)))
        return MLL, np.array(llg)

    # ------------------------------------------------------------------ model files
    def save_dict(self, db_filename):
        """gaussianprocess.py:539-551."""
        dict0 = {"noise": {"energy": self.noise_e, "f_coef": self.f_coef, "bounds": self.noise_bounds},
                 "kernel": self.kernel.save_dict(), "descriptor": self.descriptor.save_dict(),
                 "db_filename": db_filename}
        if self.base_potential is not None:
            dict0["base_potential"] = self.base_potential.save_dict()
        return dict0

    def load(self, filename, N_max=None, opt=False, device='cuda', db_rows=None):
        """Load a model JSON written by the reference (gaussianprocess.py:525-537), recompute the descriptors
        of its ASE database on the GPU and refit.  `db_rows` (list from asedb.read_rows / rows_from_npz_dict)
        overrides the db file named in the JSON."""
        with open(filename, "r") as fp:
            dict0 = json.load(fp)
        self.load_from_dict(dict0, N_max=N_max, device=device, db_rows=db_rows)
        self.fit(show=False, opt=opt)
        print("load the GP model from ", filename)

    def load_from_dict(self, dict0, N_max=None, device='cuda', db_rows=None):
        """gaussianprocess.py:553-592."""
        from .kernels import Dot_mb, RBF_mb
        from .SO3 import SO3
        name = dict0["kernel"]["name"]
        if name == "RBF_mb":
            self.kernel = RBF_mb()
        elif name == "Dot_mb":
            self.kernel = Dot_mb()
        else:
            raise NotImplementedError("unknow kernel {:s}".format(name))
        self.kernel.load_from_dict(dict0["kernel"])
        if dict0["descriptor"]["_type"] == "SO3":
            self.descriptor = SO3()
            self.descriptor.load_from_dict(dict0["descriptor"])
        else:
            raise NotImplementedError("unknow descriptors {:s}".format(str(dict0["descriptor"].get("_type"))))
        if "base_potential" in dict0:
            from .calculator import LJ
            if dict0["base_potential"]["name"] != "LJ":
                raise NotImplementedError("unknow base potential {:s}".format(str(dict0["base_potential"]["name"])))
            self.base_potential = LJ()
            self.base_potential.load_from_dict(dict0["base_potential"])
        self.kernel.device = device
        self.noise_e = dict0["noise"]["energy"]
        self.f_coef = dict0["noise"]["f_coef"]
        self.noise_bounds = dict0["noise"]["bounds"]
        self.noise_f = self.f_coef * self.noise_e
        self.extract_db(dict0["db_filename"] if db_rows is None else db_rows, N_max)

    def extract_db(self, db, N_max=None):
        """Descriptors of every structure of an ASE database (path) or of a list of rows, keeping the
        energies / forces flagged energy_in / force_in (gaussianprocess.py:627-671)."""
        from . import asedb
        rows = asedb.read_rows(db, limit=N_max) if isinstance(db, str) else (db if N_max is None else db[:N_max])
        pts_to_add = {"energy": [], "force": [], "db": []}
        for row in rows:
            atoms, energy, force = row["atoms"], row["energy"], row["force"].copy()
            d = self.descriptor.calculate(atoms)
            ele = np.array([e if isinstance(e, (int, np.integer)) else _Z[e] for e in d['elements']])
            if row["energy_in"]:
                pts_to_add["energy"].append((d['x'], energy / len(atoms), ele))
            for fid in row["force_in"]:
                ids = np.argwhere(d['seq'][:, 1] == fid).flatten()
                _i = d['seq'][ids, 0]
                pts_to_add["force"].append((d['x'][_i, :], d['dxdr'][ids], force[fid], ele[_i]))
            pts_to_add["db"].append((atoms, energy, force, row["energy_in"], row["force_in"]))
        self.set_train_pts(pts_to_add, "w")

    def add_structure(self, data, N_max=10, tol_e_var=1.2, tol_f_var=1.2, add_force=True):
        """Active-learning selection (gaussianprocess.py:673-744): predict (E, F) and their standard deviations for the
        structure `data = (atoms, energy, force)` with the current model -- K*, K*.alpha and diag(k) - diag(K* K^-1 K*^T)
        all on the device -- and propose its energy when E_std > tol_e_var * noise_e, and up to N_max of its forces whose
        std exceeds tol_f_var * noise_f and whose descriptor is new among those already picked (utilities.new_pt).
        Returns (pts_to_add, N_pts, errors) like the reference; the caller feeds pts_to_add to set_train_pts(.., 'a+')."""
        from .utilities import convert_train_data, new_pt
        tol_e_var *= self.noise_e
        tol_f_var *= self.noise_f
        pts_to_add = {"energy": [], "force": [], "db": []}
        (atoms, energy, force) = data
        if self.base_potential is not None:
            energy_off, force_off, _ = self.compute_base_potential(atoms)
        else:
            energy_off, force_off = 0, np.zeros((len(atoms), 3))
        energy = energy - energy_off
        force = np.asarray(force, dtype=np.float64) - force_off
        my_data = convert_train_data([(atoms, energy, force)], self.descriptor)
        if self.alpha_ is not None:
            E, E1, E_std, F, F1, F_std = self.validate_data(my_data, return_std=True)
            E_std = E_std[0]
            F_std = F_std.reshape((len(atoms), 3))
        else:
            E = E1 = [energy]
            F = F1 = force.flatten()
            E_std = 2 * tol_e_var
            F_std = 2 * tol_f_var * np.ones((len(atoms), 3))
        energy_in = bool(E_std > tol_e_var)
        if energy_in:
            pts_to_add["energy"] = my_data["energy"]
        force_in = []
        if add_force:
            xs_added = []
            for f_id in range(len(atoms)):
                include = False
                if np.max(F_std[f_id]) > tol_f_var:
                    X = my_data["energy"][0][0][f_id]
                    _ele = my_data["energy"][0][2][f_id]
                    include = len(xs_added) == 0 or new_pt((X, _ele), xs_added)
                if include:
                    force_in.append(f_id)
                    xs_added.append((X, _ele))
                    pts_to_add["force"].append(my_data["force"][f_id])
                if len(force_in) == N_max:
                    break
        N_pts = int(energy_in) + len(force_in)
        if N_pts > 0:
            pts_to_add["db"].append((atoms, energy, force, energy_in, force_in))
        errors = (E[0] + energy_off, E1[0] + energy_off, E_std, F + force_off.flatten(), F1 + force_off.flatten(), F_std)
        return pts_to_add, N_pts, errors

    def compute_base_potential(self, atoms):
        """(energy, forces, stress) offsets of the base potential (gaussianprocess.py:746-751)."""
        return self.base_potential.calculate(atoms)

    # ------------------------------------------------------------------ prediction
    def predict(self, X, stress=False, total_E=False, return_std=False, return_cov=False):
        """gaussianprocess.py:133-184."""
        import torch
        train = self._fitted()
        if stress:
            K_trans, K_trans1 = kdev.k_total_with_stress_device(self.kernel, X, train)
        else:
            K_trans = kdev.k_total_device(self.kernel, X, train)
        y_mean = gp.predict_mean(K_trans, self._alpha_dev).cpu().numpy()

        Npts, N_atoms = 0, None
        if 'energy' in X and len(X['energy']) > 0:
            if isinstance(X["energy"], tuple):
                N_atoms = np.array([x for x in X["energy"][-1]])
            else:
                N_atoms = np.array([len(x[0]) for x in X["energy"]])
            Npts += len(N_atoms)
        if 'force' in X and len(X['force']) > 0:
            Npts += 3 * (len(X["force"][-1]) if isinstance(X["force"], tuple) else len(X["force"]))
        factors = np.ones(Npts)
        if total_E and N_atoms is not None:
            factors[:len(N_atoms)] = N_atoms
        y_mean *= factors

        if return_cov:
            v = self._kinv(K_trans.T.contiguous())
            y_cov = kdev.k_total_device(self.kernel, X) - K_trans @ v
            return y_mean, y_cov.cpu().numpy()
        if return_std:
            y_var = self._variance(X, K_trans)
            if np.any(y_var < 0):
                warnings.warn("Predicted variances smaller than 0. Setting those variances to 0.")
            y_var[y_var < 0] = 0.0
            return y_mean, np.sqrt(y_var) * factors
        if stress:
            return y_mean, gp.predict_mean(K_trans1, self._alpha_dev).cpu().numpy()
        return y_mean

    def _kinv(self, B):
        """K^-1 B through the resident factor (cuSOLVER potrs; the sharded solver after a multi-GPU fit)."""
        if getattr(self, "_solver", None) is not None:
            return self._solver.solve(B)
        return gp.cho_solve(self.L_, B)

    def _variance(self, data, K_trans):
        """diag(k(X,X)) - diag(K* K^-1 K*^T)  (gaussianprocess.py:167-175,377-383) without forming K^-1."""
        v = self._kinv(K_trans.T.contiguous())                             # [N, n*]
        y_var = self.kernel.diag(data)
        y_var = y_var - (K_trans * v.T).sum(dim=1).cpu().numpy()
        return y_var

    def validate_data(self, test_data=None, total_E=False, return_std=False):
        """gaussianprocess.py:290-323."""
        if test_data is None:
            test_X_E = {"energy": self.train_x['energy']}
            test_X_F = {"force": self.train_x['force']}
            NE = self._n_energy()
            E = self.y_train[:NE].flatten().copy()
            F = self.y_train[NE:].flatten().copy()
            natoms = list(self.train_x['energy'][-1]) if NE else []
        else:
            test_X_E = {"energy": [(d[0], d[2]) for d in test_data['energy']]}
            test_X_F = {"force": [(d[0], d[1], d[3]) for d in test_data['force']]}
            E = np.array([d[1] for d in test_data['energy']], dtype=np.float64)
            F = np.array([d[2] for d in test_data['force']], dtype=np.float64).flatten()
            natoms = [len(d[0]) for d in test_data['energy']]
        if total_E:
            for i in range(len(E)):
                E[i] *= natoms[i]
        E_Pred = E_std = F_Pred = F_std = None
        if return_std:
            if len(test_X_E['energy']) > 0:
                E_Pred, E_std = self.predict(test_X_E, total_E=total_E, return_std=True)
            if len(test_X_F['force']) > 0:
                F_Pred, F_std = self.predict(test_X_F, return_std=True)
            return E, E_Pred, E_std, F, F_Pred, F_std
        if len(test_X_E['energy']) > 0:
            E_Pred = self.predict(test_X_E, total_E=total_E)
        if len(test_X_F['force']) > 0:
            F_Pred = self.predict(test_X_F)
        return E, E_Pred, F, F_Pred

    def predict_structure(self, struc, stress=True, return_std=False, f_tol=1e-8):
        """Energy, forces (and stress, std) of one structure from descriptor.calculate(struc)
        (gaussianprocess.py:325-390)."""
        n_atoms = len(struc)
        if hasattr(self.descriptor, "calculate_device"):
            # GPU descriptor: x / dxdr / rdxdr stay in HBM; only `seq` (which neighbour rows belong to which atom) comes
            # to the host for the stacked-tuple bookkeeping, and the gather by atom runs on the device
            import torch
            d = self.descriptor.calculate_device(struc, stress=stress)
            ele = np.array([e if isinstance(e, (int, np.integer)) else _Z[e] for e in d['elements']])
            seq = d['seq']
            order = np.argsort(seq[:, 1], kind='stable')
            centres = seq[order, 0]
            counts = np.bincount(seq[:, 1], minlength=n_atoms)
            order_d = torch.as_tensor(order, device="cuda")
            dX = d['dxdr'].index_select(0, order_d)
            if stress:
                rd = d['rdxdr'].index_select(0, order_d).reshape(len(order), dX.shape[1], 9)[:, :, [0, 4, 8, 1, 2, 5]]
                dX = torch.cat((dX, rd), dim=2)
            data = {"energy": (d['x'], ele, [len(ele)]),
                    "force": (d['x'].index_select(0, torch.as_tensor(centres, device="cuda")), dX, ele[centres], [int(c) for c in counts])}
        else:
            d = self.descriptor.calculate(struc)
            ele = np.array([e if isinstance(e, (int, np.integer)) else _Z[e] for e in d['elements']])
            _n, l = d['x'].shape
            # the reference gathers, per atom i, the seq rows with seq[:,1] == i in ascending order
            # (gaussianprocess.py:341-351); a stable sort on that column gives the same stacked tuple at once
            seq = np.asarray(d['seq'])
            order = np.argsort(seq[:, 1], kind='stable')
            centres = seq[order, 0]
            counts = np.bincount(seq[:, 1], minlength=n_atoms)
            dX = d['dxdr'][order]
            if stress:
                rd = d['rdxdr'][order].reshape([len(order), l, 9])[:, :, [0, 4, 8, 1, 2, 5]]
                dX = np.concatenate((dX, rd), axis=2)
            data = {"energy": (np.ascontiguousarray(d['x']), ele, [len(ele)]),
                    "force": (np.ascontiguousarray(d['x'][centres]), np.ascontiguousarray(dX), ele[centres], [int(c) for c in counts])}
        train = self._fitted()
        if stress:
            K_trans, K_trans1 = kdev.k_total_with_stress_device(self.kernel, data, train, f_tol)
        else:
            K_trans = kdev.k_total_device(self.kernel, data, train, f_tol)
        y_mean = gp.predict_mean(K_trans, self._alpha_dev).cpu().numpy()
        y_mean[0] *= n_atoms
        E = y_mean[0]
        F = y_mean[1:].reshape([n_atoms, 3])
        S = gp.predict_mean(K_trans1, self._alpha_dev).cpu().numpy().reshape([n_atoms, 6]) if stress else None
        if self.base_potential is not None:
            energy_off, force_off, stress_off = self.base_potential.calculate(struc)
            E += energy_off
            F += force_off
            if stress:
                S += stress_off
        if return_std:
            y_var = self._variance(data, K_trans)
            y_var[y_var < 0] = 0.0
            y_std = np.sqrt(y_var)
            return E, F, S, y_std[0], y_std[1:].reshape([n_atoms, 3])
        return E, F, S


def CUR(K, l_tol=1e-10):
    """CUR selection of Appendix D, Jinnouchi et al., PRB 014105 (2019), as in gaussianprocess.py:775-790:
    leverage scores omega_j = sum over eigenvectors with eigenvalue < l_tol of U[j,eta]^2; returns the
    indices of the N_low largest scores.  K: CUDA tensor or numpy array; eigh runs on the device."""
    import torch
    Kd = K if isinstance(K, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(K, dtype=np.float64)).cuda()
    L, U = torch.linalg.eigh(Kd)
    low = L < l_tol
    n_low = int(low.sum())
    omega = (U[:, low] ** 2).sum(dim=1)
    ids = torch.argsort(-omega, stable=True)
    return ids[:n_low].cpu().numpy()
