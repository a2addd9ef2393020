"""oracle/gp_reference.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (NumPy + SciPy, on top of oracle/cpu_reference.RefKernel) of the slice of the
reference's GaussianProcess that the product re-implements on the GPU:

    fit                      cspbo/gaussianprocess.py:71-131   (K + noise, cholesky, cho_solve, L-BFGS-B)
    predict                  :133-184  (K*.alpha, total_E factors, return_std / return_cov)
    log_marginal_likelihood  :453-503  (value and the 0.5 tr((aa^T - K^-1) dK) gradient)
    predict_structure        :325-390  (per-atom force tuples from descriptor output, stress rows)

The reference class itself cannot be imported here (ase / pyxtal / mpi4py are absent); its linear
algebra is SciPy's, which is what is used below.  Parity status: follows the cited lines one to one
on top of the pinned kernel oracle.
"""
import numpy as np
from scipy.linalg import cho_solve, cholesky, solve_triangular
from scipy.optimize import minimize

from . import cpu_reference as cr


class RefGaussianProcess:
    def __init__(self, family, para, zeta, bounds, f_coef=5, noise_e=(5e-3, 2e-3, 1e-1), backend="auto"):
        self.kernel = cr.RefKernel(family, para, zeta, backend=backend)
        self.bounds = bounds
        self.f_coef = f_coef
        self.noise_e = noise_e[0]
        self.noise_bounds = list(noise_e[1:])

    def set_train(self, train_x, y):
        self.train_x = train_x
        self.y_train = np.asarray(y, dtype=np.float64).reshape(-1, 1)
        self.NE = len(train_x["energy"][-1]) if "energy" in train_x and len(train_x["energy"]) > 0 else 0

    def _noise(self, N, noise_e):
        d = np.full(N, (self.f_coef * noise_e) ** 2)
        d[:self.NE] = noise_e ** 2
        return d

    def log_marginal_likelihood(self, params, eval_gradient=False):
        self.kernel.para = list(params[:-1])
        if eval_gradient:
            K, K_gradient = self.kernel.k_total_with_grad(self.train_x)
        else:
            K = self.kernel.k_total(self.train_x)
        K = K.copy()
        N = len(K)
        K[np.diag_indices(N)] += self._noise(N, params[-1])
        try:
            L = cholesky(K, lower=True)
        except np.linalg.LinAlgError:
            return (-np.inf, np.zeros_like(params)) if eval_gradient else -np.inf
        y = self.y_train
        alpha = cho_solve((L, True), y)
        ll = -0.5 * np.einsum("ik,ik->k", y, alpha)
        ll -= np.log(np.diag(L)).sum()
        ll -= N / 2 * np.log(2 * np.pi)
        MLL = ll.sum(-1)
        if not eval_gradient:
            return MLL
        base = np.zeros([N, N, 1])
        base[:self.NE, :self.NE, 0] += 2 * params[-1] * np.eye(self.NE)
        base[self.NE:, self.NE:, 0] += 2 * self.f_coef ** 2 * params[-1] * np.eye(N - self.NE)
        K_gradient = np.concatenate((K_gradient, base), axis=2)
        tmp = np.einsum("ik,jk->ijk", alpha, alpha)
        tmp -= cho_solve((L, True), np.eye(N))[:, :, np.newaxis]
        llg = 0.5 * np.einsum("ijl,jik->kl", tmp, K_gradient)
        return MLL, llg.sum(-1)

    def fit(self, opt=False):
        if opt:
            def obj(p):
                lml, g = self.log_marginal_likelihood(p, eval_gradient=True)
                return -lml, -g
            theta0 = list(self.kernel.para) + [self.noise_e]
            res = minimize(obj, theta0, method="L-BFGS-B", bounds=self.bounds + [self.noise_bounds], jac=True,
                           options={'maxiter': 10, 'ftol': 1e-2})
            self.kernel.para = list(res.x[:-1])
            self.noise_e = res.x[-1]
        K = self.kernel.k_total(self.train_x).copy()
        N = len(K)
        K[np.diag_indices(N)] += self._noise(N, self.noise_e)
        self.L_ = cholesky(K, lower=True)
        self.alpha_ = cho_solve((self.L_, True), self.y_train)
        return self

    def predict(self, X, total_E=False, return_std=False, return_cov=False):
        K_trans = self.kernel.k_total(X, self.train_x)
        y_mean = K_trans.dot(self.alpha_)[:, 0]
        factors = np.ones(len(y_mean))
        if total_E and "energy" in X:
            E = X["energy"]
            n_atoms = np.array(list(E[-1])) if isinstance(E, tuple) else np.array([len(e[0]) for e in E])
            factors[:len(n_atoms)] = n_atoms
        y_mean = y_mean * factors
        if return_cov:
            v = cho_solve((self.L_, True), K_trans.T)
            return y_mean, self.kernel.k_total(X) - K_trans.dot(v)
        if return_std:
            L_inv = solve_triangular(self.L_.T, np.eye(self.L_.shape[0]))
            K_inv = L_inv.dot(L_inv.T)
            y_var = self.kernel.diag(X)
            y_var = y_var - np.einsum("ij,ij->i", np.dot(K_trans, K_inv), K_trans)
            y_var[y_var < 0] = 0.0
            return y_mean, np.sqrt(y_var) * factors
        return y_mean

    def predict_structure(self, d, n_atoms, ele, stress=True, return_std=False, f_tol=1e-8):
        """`d` is the descriptor dictionary (gaussianprocess.py:335-351 builds the data from it)."""
        data = {"energy": [(d['x'], ele)], "force": []}
        l = d['x'].shape[1]
        for i in range(n_atoms):
            ids = np.argwhere(d['seq'][:, 1] == i).flatten()
            _i = d['seq'][ids, 0]
            _x, _dxdr, ele0 = d['x'][_i, :], d['dxdr'][ids], ele[_i]
            if stress:
                _r = d['rdxdr'][ids].reshape([len(ids), l, 9])[:, :, [0, 4, 8, 1, 2, 5]]
                data["force"].append((_x, np.concatenate((_dxdr, _r), axis=2), ele0))
            else:
                data["force"].append((_x, _dxdr, ele0))
        if stress:
            K_trans, K_trans1 = self.kernel.k_total_with_stress(data, self.train_x, f_tol)
        else:
            K_trans = self.kernel.k_total(data, self.train_x, f_tol)
        y = K_trans.dot(self.alpha_)[:, 0]
        y[0] *= n_atoms
        E, F = y[0], y[1:].reshape([n_atoms, 3])
        S = K_trans1.dot(self.alpha_)[:, 0].reshape([n_atoms, 6]) if stress else None
        if return_std:
            L_inv = solve_triangular(self.L_.T, np.eye(self.L_.shape[0]))
            K_inv = L_inv.dot(L_inv.T)
            y_var = self.kernel.diag(data)
            y_var = y_var - np.einsum("ij,ij->i", np.dot(K_trans, K_inv), K_trans)
            y_var[y_var < 0] = 0.0
            sd = np.sqrt(y_var)
            return E, F, S, sd[0], sd[1:].reshape([n_atoms, 3])
        return E, F, S


def CUR(K, l_tol=1e-10):
    """gaussianprocess.py:775-790 (the O(N^2) double loop written as a masked sum)."""
    L, U = np.linalg.eigh(K)
    low = L < l_tol
    omega = (U[:, low] ** 2).sum(axis=1)
    ids = np.argsort(-1 * omega, kind="stable")
    return ids[:int(low.sum())], omega
