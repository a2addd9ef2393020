"""RBF many-body kernel object, GPU-backed drop-in for the reference's cspbo/kernels/RBF_mb.py.

    k(x1, x2) = sigma^2 * exp(-(1 - (x1.x2 / |x1||x2|)^zeta) / (2 l^2)), summed over same-species pairs

API kept from the reference (RBF_mb.py:10-172).
"""
import numpy as np

from .base import build_covariance
from .rbf_kernel import kee_C, kef_C, kff_C


class RBF_mb():
    def __init__(self, para=[1., 1.], bounds=[[1e-2, 5e+1], [1e-1, 1e+1]], zeta=2, device='cuda'):
        self.name = 'RBF_mb'
        self.bounds = bounds
        self.update(para)
        self.zeta = zeta
        self.device = device

    def __str__(self):
        return "{:.3f}**2 *RBF(length={:.3f})".format(self.sigma, self.l)

    def load_from_dict(self, dict0):
        self.sigma = dict0["sigma"]
        self.l = dict0["l"]
        self.zeta = dict0["zeta"]
        self.bounds = dict0["bounds"]
        self.name = dict0["name"]

    def save_dict(self):
        return {"name": self.name, "sigma": self.sigma, "l": self.l, "zeta": self.zeta, "bounds": self.bounds}

    def parameters(self):
        return [self.sigma, self.l]

    def update(self, para):
        self.sigma, self.l = para[0], para[1]

    @staticmethod
    def _pairs(data1, data2):
        for key1, d1 in data1.items():
            for key2, d2 in data2.items():
                if len(d1) > 0 and len(d2) > 0:
                    yield key1, key2, d1, d2

    def k_total(self, data1, data2=None, tol=1e-10):
        """Covariance between data1 and data2 (RBF_mb.py:87-120); tol only reaches K_FF."""
        sigma, l, zeta = self.sigma, self.l, self.zeta
        same = data2 is None
        if same:
            data2 = data1
        C_ee, C_ef, C_fe, C_ff = None, None, None, None
        for key1, key2, d1, d2 in self._pairs(data1, data2):
            if key1 == 'energy' and key2 == 'energy':
                C_ee = kee_C(d1, d2, sigma, l, zeta)
            elif key1 == 'energy' and key2 == 'force':
                C_ef = kef_C(d1, d2, sigma, l, zeta)
            elif key1 == 'force' and key2 == 'energy':
                C_fe = C_ef.T if same else kef_C(d2, d1, sigma, l, zeta, transpose=True)
            elif key1 == 'force' and key2 == 'force':
                C_ff = kff_C(d1, d2, sigma, l, zeta, tol=tol)
        return build_covariance(C_ee, C_ef, C_fe, C_ff)

    def k_total_with_grad(self, data1):
        """(K, dK[N,N,2]) with dK/dsigma and dK/dl (RBF_mb.py:122-148)."""
        sigma, l, zeta = self.sigma, self.l, self.zeta
        blocks = [dict(ee=None, ef=None, fe=None, ff=None) for _ in range(3)]
        for key1, key2, d1, d2 in self._pairs(data1, data1):
            if key1 == 'energy' and key2 == 'energy':
                res = kee_C(d1, d2, sigma, l, zeta, grad=True)
                for blk, r in zip(blocks, res):
                    blk['ee'] = r
            elif key1 == 'energy' and key2 == 'force':
                res = kef_C(d1, d2, sigma, l, zeta, grad=True)
                for blk, r in zip(blocks, res):
                    blk['ef'], blk['fe'] = r, r.T
            elif key1 == 'force' and key2 == 'force':
                res = kff_C(d1, d2, sigma, l, zeta, grad=True)
                for blk, r in zip(blocks, res):
                    blk['ff'] = r
        C, C_s, C_l = [build_covariance(b['ee'], b['ef'], b['fe'], b['ff']) for b in blocks]
        return C, np.dstack((C_s, C_l))

    def k_total_with_stress(self, data1, data2, tol=1e-10):
        """(K*, K*_stress) (RBF_mb.py:150-172)."""
        sigma, l, zeta = self.sigma, self.l, self.zeta
        C_ee, C_ef, C_fe, C_ff, C_se, C_sf = None, None, None, None, None, None
        for key1, key2, d1, d2 in self._pairs(data1, data2):
            if key1 == 'energy' and key2 == 'energy':
                C_ee = kee_C(d1, d2, sigma, l, zeta)
            elif key1 == 'energy' and key2 == 'force':
                C_ef = kef_C(d1, d2, sigma, l, zeta)
            elif key1 == 'force' and key2 == 'energy':
                C_fe, C_se = kef_C(d2, d1, sigma, l, zeta, stress=True, transpose=True)
            elif key1 == 'force' and key2 == 'force':
                C_ff, C_sf = kff_C(d1, d2, sigma, l, zeta, stress=True, tol=tol)
        return build_covariance(C_ee, C_ef, C_fe, C_ff), build_covariance(None, None, C_se, C_sf)

    def diag(self, data):
        """Diagonal of k(X, X) (RBF_mb.py:45-85)."""
        from .diag import kernel_diag
        return kernel_diag(self, data)
