"""Dot many-body kernel object, GPU-backed drop-in for the reference's cspbo/kernels/Dot_mb.py.

    k(x1, x2) = sigma^2 * ((x1.x2 / |x1||x2|)^zeta + sigma0^2), summed over same-species atom pairs

API kept from the reference (Dot_mb.py:10-173): constructor arguments, `name`, `bounds`, `zeta`,
`device`, parameters()/update(), save_dict()/load_from_dict(), __str__, k_total,
k_total_with_grad, k_total_with_stress and diag.
"""
import numpy as np

from .base import build_covariance
from .dot_kernel import kee_C, kef_C, kff_C


class Dot_mb():
    def __init__(self, para=[1., 1.], bounds=[[1e-2, 5e+1], [1e-2, 1e+1]], zeta=3, device="cuda"):
        self.name = 'Dot_mb'
        self.bounds = bounds
        self.update(para)
        self.zeta = zeta
        self.device = device

    def __str__(self):
        return "{:.3f}**2 *Dot(length={:.3f})".format(self.sigma, self.sigma0)

    def load_from_dict(self, dict0):
        self.sigma = dict0["sigma"]
        self.sigma0 = dict0["sigma0"]
        self.zeta = dict0["zeta"]
        self.bounds = dict0["bounds"]
        self.name = dict0["name"]

    def save_dict(self):
        return {"name": self.name, "sigma": self.sigma, "sigma0": self.sigma0,
                "zeta": self.zeta, "bounds": self.bounds}

    def parameters(self):
        return [self.sigma, self.sigma0]

    def update(self, para):
        self.sigma, self.sigma0 = para[0], para[1]

    # ------------------------------------------------------------------ covariance builders
    @staticmethod
    def _pairs(data1, data2):
        for key1, d1 in data1.items():
            for key2, d2 in data2.items():
                if len(d1) > 0 and len(d2) > 0:
                    yield key1, key2, d1, d2

    def k_total(self, data1, data2=None, tol=1e-12):
        """Covariance between data1 and data2 (data1 itself when data2 is None); tol is unused.

        Parity note (Dot_mb.py:108-117): the reference calls kef_C(d1, d2, sigma, zeta) and
        kff_C(d1, d2, sigma, zeta) positionally, so `zeta` lands in the unused sigma0 slot and the
        K_EF / K_FF blocks are evaluated with the wrappers' default zeta=2.0 whatever self.zeta is.
        Reproduced here on purpose; K_EE and k_total_with_grad use self.zeta."""
        sigma, sigma0, zeta = self.sigma, self.sigma0, self.zeta
        same = data2 is None
        if same:
            data2 = data1
        C_ee, C_ef, C_fe, C_ff = None, None, None, None
        for key1, key2, d1, d2 in self._pairs(data1, data2):
            if key1 == 'energy' and key2 == 'energy':
                C_ee = kee_C(d1, d2, sigma, sigma0, zeta)
            elif key1 == 'energy' and key2 == 'force':
                C_ef = kef_C(d1, d2, sigma, zeta)
            elif key1 == 'force' and key2 == 'energy':
                C_fe = C_ef.T if same else kef_C(d2, d1, sigma, zeta, transpose=True)
            elif key1 == 'force' and key2 == 'force':
                C_ff = kff_C(d1, d2, sigma, zeta)
        return build_covariance(C_ee, C_ef, C_fe, C_ff)

    def k_total_with_grad(self, data1):
        """(K, dK[N,N,2]) with dK/dsigma and dK/dsigma0 (Dot_mb.py:121-148)."""
        sigma, sigma0, zeta = self.sigma, self.sigma0, self.zeta
        blocks = [dict(ee=None, ef=None, fe=None, ff=None) for _ in range(3)]
        for key1, key2, d1, d2 in self._pairs(data1, data1):
            if key1 == 'energy' and key2 == 'energy':
                res = kee_C(d1, d2, sigma, sigma0, zeta, grad=True)
                for blk, r in zip(blocks, res):
                    blk['ee'] = r
            elif key1 == 'energy' and key2 == 'force':
                res = kef_C(d1, d2, sigma, sigma0, zeta, grad=True)
                for blk, r in zip(blocks, res):
                    blk['ef'], blk['fe'] = r, r.T
            elif key1 == 'force' and key2 == 'force':
                res = kff_C(d1, d2, sigma, sigma0, zeta, grad=True)
                for blk, r in zip(blocks, res):
                    blk['ff'] = r
        C, C1, C2 = [build_covariance(b['ee'], b['ef'], b['fe'], b['ff']) for b in blocks]
        return C, np.dstack((C1, C2))

    def k_total_with_stress(self, data1, data2, tol=1e-10):
        """(K*, K*_stress) for energy/force/stress prediction (Dot_mb.py:150-173); same zeta quirk
        as k_total."""
        sigma, sigma0, zeta = self.sigma, self.sigma0, self.zeta
        C_ee, C_ef, C_fe, C_ff, C_se, C_sf = None, None, None, None, None, None
        for key1, key2, d1, d2 in self._pairs(data1, data2):
            if key1 == 'energy' and key2 == 'energy':
                C_ee = kee_C(d1, d2, sigma, sigma0, zeta)
            elif key1 == 'energy' and key2 == 'force':
                C_ef = kef_C(d1, d2, sigma, zeta)
            elif key1 == 'force' and key2 == 'energy':
                C_fe, C_se = kef_C(d2, d1, sigma, zeta, stress=True, transpose=True)
            elif key1 == 'force' and key2 == 'force':
                C_ff, C_sf = kff_C(d1, d2, sigma, zeta, stress=True)
        return build_covariance(C_ee, C_ef, C_fe, C_ff), build_covariance(None, None, C_se, C_sf)

    def diag(self, data):
        """Diagonal of k(X, X) (Dot_mb.py:45-85)."""
        from .diag import kernel_diag
        return kernel_diag(self, data)
