"""Shared parity cases: the same list is evaluated by the reference C++ (to make the goldens),
by the C port (tests/test_oracle.py) and by the CUDA product (tests/test_parity_gpu.py)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixtures():
    return dict(np.load(os.path.join(GOLDEN_DIR, "fixtures.npz")))


def load_golden():
    return dict(np.load(os.path.join(GOLDEN_DIR, "golden_outputs.npz")))


class OracleImpl:
    """The CPU oracle (oracle/cpu_reference.py) behind the uniform interface used by golden_cases."""

    def __init__(self, backend, threads=8):
        from oracle import cpu_reference as cr
        self.cr, self.backend, self.threads = cr, backend, threads

    def dot_kee_C(self, *a, **k): return self.cr.dot_kee_C(*a, backend=self.backend, **k)
    def dot_kef_C(self, *a, **k): return self.cr.dot_kef_C(*a, backend=self.backend, **k)
    def dot_kff_C(self, *a, **k): return self.cr.dot_kff_C(*a, backend=self.backend, threads=self.threads, **k)
    def rbf_kee_C(self, *a, **k): return self.cr.rbf_kee_C(*a, backend=self.backend, **k)
    def rbf_kef_C(self, *a, **k): return self.cr.rbf_kef_C(*a, backend=self.backend, **k)
    def rbf_kff_C(self, *a, **k): return self.cr.rbf_kff_C(*a, backend=self.backend, threads=self.threads, **k)
    def kernel(self, family, para, zeta): return self.cr.RefKernel(family, para, zeta, backend=self.backend)


class ProductImpl:
    """The CUDA product (csp_bo_b200.kernels) behind the same interface."""

    def __init__(self):
        from csp_bo_b200.kernels import dot_kernel, rbf_kernel, Dot_mb, RBF_mb
        self.d, self.r, self.Dot_mb, self.RBF_mb = dot_kernel, rbf_kernel, Dot_mb, RBF_mb

    def dot_kee_C(self, *a, **k): return self.d.kee_C(*a, **k)
    def dot_kef_C(self, *a, **k): return self.d.kef_C(*a, **k)
    def dot_kff_C(self, *a, **k): return self.d.kff_C(*a, **k)
    def rbf_kee_C(self, *a, **k): return self.r.kee_C(*a, **k)
    def rbf_kef_C(self, *a, **k): return self.r.kef_C(*a, **k)
    def rbf_kff_C(self, *a, **k): return self.r.kff_C(*a, **k)

    def kernel(self, family, para, zeta):
        return self.Dot_mb(para=para, zeta=zeta) if family == "dot" else self.RBF_mb(para=para, zeta=zeta)


def max_norm_err(a, b):
    """per-block max-normalised error."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = np.abs(b).max()
    return float(np.abs(a - b).max() / (den if den > 0 else 1.0))


HYPER = {
    "dot": dict(sigma=18.55544058601137, sigma0=0.01, zeta=2.0),
    "dot_kff_py": dict(sigma=28.835, sigma0=0.01, zeta=2.0),
    "rbf": dict(sigma=9.55544058601137, l=0.5, zeta=2.0),
    "rbf_si": dict(sigma=6.244955524643115, l=0.5611029935713949, zeta=2.0),
}


def stress_columns(dxdr):
    """No shipped fixture carries the 6 stress columns; derive a deterministic 9-column
    array from the 3 force columns so the *_stress entry points have pinned inputs."""
    return np.concatenate((dxdr, 0.5 * dxdr[:, :, [1, 2, 0]], -0.25 * dxdr[:, :, [2, 0, 1]]), axis=2)


def take_groups(force, g0, g1):
    x, dx, ele, counts = force
    off = np.concatenate(([0], np.cumsum(counts)))
    s, e = off[g0], off[g1]
    return (x[s:e], dx[s:e], ele[s:e], list(counts[g0:g1]))


def take_groups_e(energy, g0, g1):
    x, ele, counts = energy
    off = np.concatenate(([0], np.cumsum(counts)))
    s, e = off[g0], off[g1]
    return (x[s:e], ele[s:e], list(counts[g0:g1]))


def force_list(F):
    x, dx, ele, counts = F
    off = np.concatenate(([0], np.cumsum(counts)))
    return [(x[off[i]:off[i + 1]], dx[off[i]:off[i + 1]], ele[off[i]:off[i + 1]]) for i in range(len(counts))]


def energy_list(E):
    x, ele, counts = E
    off = np.concatenate(([0], np.cumsum(counts)))
    return [(x[off[i]:off[i + 1]], ele[off[i]:off[i + 1]]) for i in range(len(counts))]


def force_tuple(fx, name):
    return (fx[name + "_x"], fx[name + "_dxdr"], fx[name + "_ele"], list(fx[name + "_counts"]))


def energy_tuple(fx, name):
    return (fx[name + "_x"], fx[name + "_ele"], list(fx[name + "_counts"]))


def golden_cases(fx, cr):
    """name -> array for every parity case; `cr` is an OracleImpl or a ProductImpl."""
    big, x2 = force_tuple(fx, "big"), force_tuple(fx, "x2")
    ee1, ee2 = energy_tuple(fx, "ee1"), energy_tuple(fx, "ee2")
    big64 = take_groups(big, 0, 64)
    big9 = (big64[0], stress_columns(big64[1]), big64[2], big64[3])
    x29 = (x2[0], stress_columns(x2[1]), x2[2], x2[3])
    ee2s = take_groups_e(ee2, 0, 4)
    out = {}
    h = HYPER["dot"]
    s, s0, z = h["sigma"], h["sigma0"], h["zeta"]
    r = cr.dot_kee_C(ee1, ee2, s, s0, z, grad=True)
    out["dot_kee"], out["dot_kee_ds"], out["dot_kee_ds0"] = r
    out["dot_kee_z3"] = cr.dot_kee_C(ee2s, ee2s, s, s0, 3.0)
    out["dot_kef"] = cr.dot_kef_C(ee1, big64, s, s0, z)
    out["dot_kef_x2"] = cr.dot_kef_C(ee2s, x2, s, s0, z)
    out["dot_kfe_T"] = cr.dot_kef_C(ee2s, x2, s, s0, z, transpose=True)
    c, cs = cr.dot_kef_C(ee2s, x29, s, s0, z, stress=True, transpose=True)
    out["dot_kfe_stress_C"], out["dot_kfe_stress_Cs"] = c, cs
    out["dot_kef_z3"] = cr.dot_kef_C(ee2s, x2, s, s0, 3.0)
    h = HYPER["dot_kff_py"]
    out["dot_kff_x2x2"] = cr.dot_kff_C(x2, x2, h["sigma"], h["sigma0"], h["zeta"])
    out["dot_kff_x2x2_z3"] = cr.dot_kff_C(x2, x2, h["sigma"], h["sigma0"], 3.0)
    out["dot_kff_x2x2_z25"] = cr.dot_kff_C(x2, x2, h["sigma"], h["sigma0"], 2.5)
    out["dot_kff_x2big64"] = cr.dot_kff_C(x2, big64, s, s0, z)
    c, cs = cr.dot_kff_C(x29, big64, s, s0, z, stress=True)
    out["dot_kff_stress_C"], out["dot_kff_stress_Cs"] = c, cs
    out["dot_kff_bigbig"] = cr.dot_kff_C(big, big, s, s0, z)

    for tag in ("rbf", "rbf_si"):
        h = HYPER[tag]
        s, l, z = h["sigma"], h["l"], h["zeta"]
        r = cr.rbf_kee_C(ee1, ee2, s, l, z, grad=True)
        out[tag + "_kee"], out[tag + "_kee_ds"], out[tag + "_kee_dl"] = r
        out[tag + "_kee_nograd"] = cr.rbf_kee_C(ee1, ee2, s, l, z)
        out[tag + "_kef"] = cr.rbf_kef_C(ee2s, x2, s, l, z)
        r = cr.rbf_kef_C(ee2s, x2, s, l, z, grad=True)
        out[tag + "_kef_g"], out[tag + "_kef_ds"], out[tag + "_kef_dl"] = r
        c, cs = cr.rbf_kef_C(ee2s, x29, s, l, z, stress=True, transpose=True)
        out[tag + "_kfe_stress_C"], out[tag + "_kfe_stress_Cs"] = c, cs
        out[tag + "_kff_x2x2"] = cr.rbf_kff_C(x2, x2, s, l, z, tol=1e-12)
        out[tag + "_kff_x2x2_tol"] = cr.rbf_kff_C(x2, x2, s, l, z, tol=s * s / (2 * l * l) * np.exp(-0.05 / (2 * l * l)))
        r = cr.rbf_kff_C(x2, x2, s, l, z, grad=True)
        out[tag + "_kff_g"], out[tag + "_kff_ds"], out[tag + "_kff_dl"] = r
        c, cs = cr.rbf_kff_C(x29, big64, s, l, z, stress=True, tol=1e-12)
        out[tag + "_kff_stress_C"], out[tag + "_kff_stress_Cs"] = c, cs
    out["rbf_kff_x2x2_z3"] = cr.rbf_kff_C(x2, x2, 9.55544058601137, 0.5, 3.0, tol=1e-12)
    h = HYPER["rbf_si"]
    out["rbf_si_kff_big64"] = cr.rbf_kff_C(big64, big64, h["sigma"], h["l"], h["zeta"], tol=1e-10)

    # kernel-object level (k_total family) on a mixed energy+force training set
    train = {"energy": ee2s, "force": x2}
    test = {"energy": take_groups_e(ee2, 4, 6), "force": take_groups(big, 0, 12)}
    test9 = {"energy": take_groups_e(ee2, 4, 5),
             "force": (test["force"][0], stress_columns(test["force"][1]), test["force"][2], test["force"][3])}
    for fam, para, zeta in (("dot", [18.55544058601137, 0.01], 2.0), ("rbf", [6.244955524643115, 0.5611029935713949], 2.0)):
        k = cr.kernel(fam, para, zeta)
        out[fam + "_ktotal_train"] = k.k_total(train)
        out[fam + "_ktotal_cross"] = k.k_total(test, train)
        K, dK = k.k_total_with_grad(train)
        out[fam + "_ktotal_grad_K"], out[fam + "_ktotal_grad_dK"] = K, dK
        K, Ks = k.k_total_with_stress(test9, train)
        out[fam + "_ktotal_stress_K"], out[fam + "_ktotal_stress_Ks"] = K, Ks
        out[fam + "_diag"] = k.diag(test)           # tuple form
        out[fam + "_diag_list"] = k.diag({"energy": energy_list(test["energy"]), "force": force_list(test9["force"])})
    return out


