"""oracle/cpu_reference.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

CPU restatement of the reference's block wrappers and kernel objects for the GP
kernel-matrix hot path, driven through ctypes on one of two native back ends:

  backend "ref"   the UNMODIFIED reference C++ (cspbo/kernels/dot_kernel.cpp, rbf_kernel.cpp)
                  compiled by oracle/Makefile into oracle/_ref/lib{dot,rbf}_ref.so
  backend "port"  our plain-C restatement oracle/kernels_port.c -> oracle/_build/libcspbo_oracle.so

What is restated here, and from where:
  kee/kef/kff wrappers   cspbo/kernels/dot_kernel.py:8-274, cspbo/kernels/rbf_kernel.py:8-333
                         (group-id expansion, output reshapes, the /N_i, *(-sigma^2),
                         *(sigma^2 zeta), /l^3 post-scalings, hyper-gradient companions)
  block assembly         cspbo/kernels/base.py:3-26
  k_total*               cspbo/kernels/Dot_mb.py:87-173, cspbo/kernels/RBF_mb.py:87-172
                         (including the Dot quirk that K_EF/K_FF in k_total are evaluated with
                         zeta = 2.0 because `zeta` is passed in the sigma0 slot, Dot_mb.py:108-117)
  MPI split of K_FF      dot_kernel.py:188-203 -- emulated with threads, one [n2_start,n2_end)
                         slice each (ctypes releases the GIL), summed afterwards like Allreduce

Parity status: PINNED (see the header of oracle/kernels_port.c and tests/test_oracle.py).
Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_c_int = ctypes.c_int
_c_dbl = ctypes.c_double
_P = ctypes.c_void_p


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


class Backend:
    """Loads a CPU implementation of the 13 native entry points."""

    def __init__(self, kind="auto"):
        ref_dot = os.path.join(_HERE, "_ref", "libdot_ref.so")
        ref_rbf = os.path.join(_HERE, "_ref", "librbf_ref.so")
        port = os.path.join(_HERE, "_build", "libcspbo_oracle.so")
        if kind == "auto":
            kind = "ref" if os.path.exists(ref_dot) and os.path.exists(ref_rbf) else "port"
        self.kind = kind
        if kind == "ref":
            self._dot = ctypes.CDLL(ref_dot)
            self._rbf = ctypes.CDLL(ref_rbf)
            self._prefix = {"dot": "", "rbf": ""}
        elif kind in ("port", "port_native"):
            if kind == "port_native":       # -O3 -march=native build of the same restatement (make -C oracle port_native)
                port = os.path.join(_HERE, "_build", "libcspbo_oracle_native.so")
            if not os.path.exists(port):
                raise FileNotFoundError(port + " missing: run `make -C oracle %s`" % kind)
            self._dot = self._rbf = ctypes.CDLL(port)
            self._prefix = {"dot": "oracle_dot_", "rbf": "oracle_rbf_"}
        else:
            raise ValueError(kind)

    def fn(self, family, name):
        lib = self._dot if family == "dot" else self._rbf
        f = getattr(lib, self._prefix[family] + name)
        f.restype = None
        return f


_BACKENDS = {}


def get_backend(kind="auto"):
    if kind not in _BACKENDS:
        _BACKENDS[kind] = Backend(kind)
    return _BACKENDS[kind]


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libdot_ref.so"))


# ----------------------------------------------------------------------------- packing
def list_to_tuple(data, stress=False, mode="force"):
    """cspbo/utilities.py:350-400 (without the include_value branch)."""
    if mode == "force":
        X = np.concatenate([np.asarray(f[0], dtype=np.float64) for f in data], axis=0)
        nc = 9 if stress else 3
        dX = np.concatenate([np.asarray(f[1], dtype=np.float64)[:, :, :nc] for f in data], axis=0)
        ELE = np.concatenate([np.ravel(f[2]) for f in data])
        return (X, dX, ELE, [len(f[0]) for f in data])
    X = np.concatenate([np.asarray(f[0], dtype=np.float64) for f in data], axis=0)
    ELE = np.concatenate([np.ravel(f[1]) for f in data])
    return (X, ELE, [len(f[0]) for f in data])


def _prep_energy(X):
    if isinstance(X, list):
        X = list_to_tuple(X, mode="energy")
    x, ele, indices = X
    x = np.ascontiguousarray(x, dtype=np.float64)
    ele = np.ascontiguousarray(ele, dtype=np.int32)
    counts = np.asarray(indices, dtype=np.int64)
    gid = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
    return x, ele, gid, counts


def _prep_force(X, stress=False):
    if isinstance(X, list):
        X = list_to_tuple(X, stress=stress)
    x, dxdr, ele, indices = X
    x = np.ascontiguousarray(x, dtype=np.float64)
    dxdr = np.ascontiguousarray(dxdr, dtype=np.float64)
    ele = np.ascontiguousarray(ele, dtype=np.int32)
    counts = np.asarray(indices, dtype=np.int64)
    gid = np.repeat(np.arange(len(counts), dtype=np.int32), counts)
    return x, dxdr, ele, gid, counts


def _split(n, parts):
    """dot_kernel.py:188-203: contiguous slices, the first n%parts get one more row."""
    base, rest = divmod(n, parts)
    out, start = [], 0
    for r in range(parts):
        size = base + (1 if r < rest else 0)
        out.append((start, start + size))
        start += size
    return out


def _run_kff_sliced(call, n_out, m2p, threads):
    """Run `call(n2_start, n2_end, out)` over `threads` slices and sum (MPI emulation)."""
    if threads <= 1:
        out = np.zeros(n_out)
        call(0, m2p, out)
        return out
    outs = [np.zeros(n_out) for _ in range(threads)]
    ts = []
    for (s, e), o in zip(_split(m2p, threads), outs):
        t = threading.Thread(target=call, args=(s, e, o))
        t.start()
        ts.append(t)
    for t in ts:
        t.join()
    total = outs[0]
    for o in outs[1:]:
        total += o
    return total


# ----------------------------------------------------------------------------- Dot blocks
def dot_kee_C(X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False, backend="auto"):
    """dot_kernel.py:8-62."""
    be = get_backend(backend)
    x1, e1, g1, c1 = _prep_energy(X1)
    x2, e2, g2, c2 = _prep_energy(X2)
    m1, m2, d = len(c1), len(c2), x1.shape[1]
    out = np.zeros(m1 * m2)
    be.fn("dot", "kee_many")(_c_int(len(x1)), _c_int(len(x2)), _c_int(d), _c_int(m2),
                             _c_dbl(zeta), _c_dbl(sigma ** 2), _c_dbl(sigma0 ** 2),
                             _ptr(x1), _ptr(e1), _ptr(g1), _ptr(x2), _ptr(e2), _ptr(g2), _ptr(out))
    C = out.reshape(m1, m2)
    C /= (c1[:, None] * c2[None, :])
    if grad:
        return C, 2 * C / sigma, 0.8 * 2 * sigma ** 2 * sigma0 * np.ones([m1, m2])
    return C


def dot_kef_C(X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False, stress=False, transpose=False,
              backend="auto"):
    """dot_kernel.py:65-159."""
    be = get_backend(backend)
    x1, e1, g1, c1 = _prep_energy(X1)
    x2, dx2, e2, g2, c2 = _prep_force(X2, stress)
    m1, m2, d = len(c1), len(c2), x1.shape[1]
    nc = 9 if stress else 3
    assert dx2.shape[2] == nc
    out = np.zeros(m1 * m2 * nc)
    be.fn("dot", "kef_many_stress" if stress else "kef_many")(
        _c_int(len(x1)), _c_int(len(x2)), _c_int(d), _c_int(m2), _c_dbl(zeta),
        _ptr(x1), _ptr(e1), _ptr(g1), _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
    out = out.reshape(m1, m2, nc)
    out /= c1[:, None, None]
    out *= -sigma * sigma
    C = out[:, :, :3].reshape(m1, m2 * 3)
    Cs = out[:, :, 3:].reshape(m1, m2 * 6) if stress else np.zeros([m1, m2 * 6])
    if transpose:
        C, Cs = C.T, Cs.T
    if grad:
        return C, 2 * C / sigma, np.zeros([m1, m2 * 3])
    if stress:
        return C, Cs
    return C


def dot_kff_C(X1, X2, sigma=1.0, sigma0=1.0, zeta=2.0, grad=False, stress=False, backend="auto",
              threads=1):
    """dot_kernel.py:161-274."""
    be = get_backend(backend)
    x1, dx1, e1, g1, c1 = _prep_force(X1, stress)
    x2, dx2, e2, g2, c2 = _prep_force(X2, False)
    m1, m2, d = len(c1), len(c2), x1.shape[1]
    nc1 = 9 if stress else 3
    assert dx1.shape[2] == nc1 and dx2.shape[2] == 3
    fn = be.fn("dot", "kff_many_stress" if stress else "kff_many")

    def call(s, e, out):
        fn(_c_int(len(x1)), _c_int(len(x2)), _c_int(s), _c_int(e), _c_int(d), _c_int(m2),
           _c_dbl(zeta), _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1),
           _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))

    out = _run_kff_sliced(call, m1 * nc1 * m2 * 3, len(x2), threads).reshape(m1, nc1, m2 * 3)
    out *= sigma * sigma * zeta
    C = out[:, :3, :].reshape(m1 * 3, m2 * 3)
    if grad:
        return C, 2 * C / sigma, np.zeros([m1 * 3, m2 * 3])
    if stress:
        return C, out[:, 3:, :].reshape(m1 * 6, m2 * 3)
    return C


# ----------------------------------------------------------------------------- RBF blocks
def rbf_kee_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False, backend="auto"):
    """rbf_kernel.py:8-84."""
    be = get_backend(backend)
    x1, e1, g1, c1 = _prep_energy(X1)
    x2, e2, g2, c2 = _prep_energy(X2)
    m1, m2, d = len(c1), len(c2), x1.shape[1]
    out = np.zeros(m1 * m2)
    norm = (c1[:, None] * c2[None, :])
    if grad:
        dl = np.zeros(m1 * m2)
        be.fn("rbf", "kee_many_with_grad")(
            _c_int(len(x1)), _c_int(len(x2)), _c_int(d), _c_int(m2), _c_dbl(zeta),
            _c_dbl(sigma * sigma), _c_dbl(l * l),
            _ptr(x1), _ptr(e1), _ptr(g1), _ptr(x2), _ptr(e2), _ptr(g2), _ptr(out), _ptr(dl))
        C = out.reshape(m1, m2) / norm
        C_l = dl.reshape(m1, m2) / norm
        C_l *= 1 / (l * (l * l))
        return C, (2 / sigma) * C, C_l
    be.fn("rbf", "kee_many")(
        _c_int(len(x1)), _c_int(len(x2)), _c_int(d), _c_int(m2), _c_dbl(zeta),
        _c_dbl(sigma * sigma), _c_dbl(l * l),
        _ptr(x1), _ptr(e1), _ptr(g1), _ptr(x2), _ptr(e2), _ptr(g2), _ptr(out))
    return out.reshape(m1, m2) / norm


def rbf_kef_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False, stress=False, transpose=False,
              backend="auto"):
    """rbf_kernel.py:86-185."""
    be = get_backend(backend)
    x1, e1, g1, c1 = _prep_energy(X1)
    x2, dx2, e2, g2, c2 = _prep_force(X2, stress)
    m1, m2, d = len(c1), len(c2), x1.shape[1]
    if stress:
        name, nco, hyper = "kef_many_stress", 9, l * l
    elif grad:
        name, nco, hyper = "kef_many_with_grad", 6, l
    else:
        name, nco, hyper = "kef_many", 3, l * l
    out = np.zeros(m1 * m2 * nco)
    be.fn("rbf", name)(_c_int(len(x1)), _c_int(len(x2)), _c_int(d), _c_int(m2), _c_dbl(zeta),
                       _c_dbl(sigma * sigma), _c_dbl(hyper),
                       _ptr(x1), _ptr(e1), _ptr(g1), _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))
    out = out.reshape(m1, m2, nco)
    out /= c1[:, None, None]
    C = out[:, :, :3].reshape(m1, m2 * 3)
    Cs = np.zeros([m1, m2 * 6])
    C_l = C_s = None
    if stress:
        Cs = out[:, :, 3:].reshape(m1, m2 * 6)
    elif grad:
        C_l = out[:, :, 3:].reshape(m1, m2 * 3)
        C_s = (2 / sigma) * C
    if transpose:
        C, Cs = C.T, Cs.T
    if grad:
        return C, C_s, C_l
    if stress:
        return C, Cs
    return C


def rbf_kff_C(X1, X2, sigma=1.0, l=1.0, zeta=2.0, grad=False, stress=False, tol=1e-12,
              backend="auto", threads=1):
    """rbf_kernel.py:187-333."""
    be = get_backend(backend)
    x1, dx1, e1, g1, c1 = _prep_force(X1, stress)
    x2, dx2, e2, g2, c2 = _prep_force(X2, False)
    m1, m2, d = len(c1), len(c2), x1.shape[1]
    s2 = sigma * sigma
    if stress:
        fn = be.fn("rbf", "kff_many_stress")

        def call(s, e, out):
            fn(_c_int(len(x1)), _c_int(len(x2)), _c_int(s), _c_int(e), _c_int(d), _c_int(m2),
               _c_dbl(zeta), _c_dbl(s2), _c_dbl(l * l), _c_dbl(tol),
               _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1), _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))

        out = _run_kff_sliced(call, m1 * 9 * m2 * 3, len(x2), threads).reshape(m1, 9, m2 * 3)
        return out[:, :3, :].reshape(m1 * 3, m2 * 3), out[:, 3:, :].reshape(m1 * 6, m2 * 3)
    if grad:
        out = np.zeros(m1 * 3 * m2 * 3)
        dl = np.zeros(m1 * 3 * m2 * 3)
        be.fn("rbf", "kff_many_with_grad")(
            _c_int(len(x1)), _c_int(len(x2)), _c_int(0), _c_int(len(x2)), _c_int(d), _c_int(m2),
            _c_dbl(zeta), _c_dbl(s2), _c_dbl(l),
            _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1), _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2),
            _ptr(out), _ptr(dl))
        C = out.reshape(m1 * 3, m2 * 3)
        return C, (2 / sigma) * C, dl.reshape(m1 * 3, m2 * 3)
    fn = be.fn("rbf", "kff_many")

    def call(s, e, out):
        fn(_c_int(len(x1)), _c_int(len(x2)), _c_int(s), _c_int(e), _c_int(d), _c_int(m2),
           _c_dbl(zeta), _c_dbl(s2), _c_dbl(l * l), _c_dbl(tol),
           _ptr(x1), _ptr(dx1), _ptr(e1), _ptr(g1), _ptr(x2), _ptr(dx2), _ptr(e2), _ptr(g2), _ptr(out))

    return _run_kff_sliced(call, m1 * 3 * m2 * 3, len(x2), threads).reshape(m1 * 3, m2 * 3)


# ----------------------------------------------------------------------------- assembly
def build_covariance(c_ee, c_ef, c_fe, c_ff):
    """cspbo/kernels/base.py:3-26."""
    have = [c is not None for c in (c_ee, c_ef, c_fe, c_ff)]
    if all(have):
        return np.block([[c_ee, c_ef], [c_fe, c_ff]])
    table = {
        (False, False, True, True): lambda: np.hstack((c_fe, c_ff)),
        (True, True, False, False): lambda: np.hstack((c_ee, c_ef)),
        (False, True, False, False): lambda: c_ef,
        (True, False, False, False): lambda: c_ee,
        (False, False, False, True): lambda: c_ff,
        (False, False, True, False): lambda: c_fe,
    }
    fn = table.get(tuple(have))
    return fn() if fn else None


def _nonempty(v):
    return len(v) > 0


class RefKernel:
    """Restates Dot_mb / RBF_mb k_total, k_total_with_grad, k_total_with_stress on the CPU."""

    def __init__(self, family, para, zeta, backend="auto"):
        assert family in ("dot", "rbf")
        self.family, self.para, self.zeta, self.backend = family, list(para), zeta, backend

    # -- block dispatch with the reference's exact positional-argument behaviour
    def _blocks(self, quirk):
        p0, p1, z, be = self.para[0], self.para[1], self.zeta, self.backend
        if self.family == "rbf":
            kee = lambda a, b, **k: rbf_kee_C(a, b, p0, p1, z, backend=be, **k)
            kef = lambda a, b, **k: rbf_kef_C(a, b, p0, p1, z, backend=be, **k)
            kff = lambda a, b, **k: rbf_kff_C(a, b, p0, p1, z, backend=be, **k)
            return kee, kef, kff
        kee = lambda a, b, **k: dot_kee_C(a, b, p0, p1, z, backend=be, **k)
        if quirk:
            # Dot_mb.py:108-117: kef_C(d1, d2, sigma, zeta) -> sigma0 := zeta, zeta := default 2.0
            kef = lambda a, b, **k: dot_kef_C(a, b, p0, z, backend=be, **k)
            kff = lambda a, b, **k: dot_kff_C(a, b, p0, z, backend=be, **k)
        else:
            kef = lambda a, b, **k: dot_kef_C(a, b, p0, p1, z, backend=be, **k)
            kff = lambda a, b, **k: dot_kff_C(a, b, p0, p1, z, backend=be, **k)
        return kee, kef, kff

    def k_total(self, data1, data2=None, tol=None):
        kee, kef, kff = self._blocks(quirk=True)
        same = data2 is None
        if same:
            data2 = data1
        ffkw = {}
        if self.family == "rbf":
            ffkw["tol"] = 1e-10 if tol is None else tol
        C = {"ee": None, "ef": None, "fe": None, "ff": None}
        for k1, d1 in data1.items():
            for k2, d2 in data2.items():
                if not (_nonempty(d1) and _nonempty(d2)):
                    continue
                if k1 == "energy" and k2 == "energy":
                    C["ee"] = kee(d1, d2)
                elif k1 == "energy" and k2 == "force":
                    C["ef"] = kef(d1, d2)
                elif k1 == "force" and k2 == "energy":
                    C["fe"] = C["ef"].T if same else kef(d2, d1, transpose=True)
                elif k1 == "force" and k2 == "force":
                    C["ff"] = kff(d1, d2, **ffkw)
        return build_covariance(C["ee"], C["ef"], C["fe"], C["ff"])

    def k_total_with_grad(self, data1):
        kee, kef, kff = self._blocks(quirk=False)
        B = [dict(ee=None, ef=None, fe=None, ff=None) for _ in range(3)]
        for k1, d1 in data1.items():
            for k2, d2 in data1.items():
                if not (_nonempty(d1) and _nonempty(d2)):
                    continue
                if k1 == "energy" and k2 == "energy":
                    r = kee(d1, d2, grad=True)
                    for t in range(3):
                        B[t]["ee"] = r[t]
                elif k1 == "energy" and k2 == "force":
                    r = kef(d1, d2, grad=True)
                    for t in range(3):
                        B[t]["ef"], B[t]["fe"] = r[t], r[t].T
                elif k1 == "force" and k2 == "force":
                    r = kff(d1, d2, grad=True)
                    for t in range(3):
                        B[t]["ff"] = r[t]
        C, C1, C2 = [build_covariance(b["ee"], b["ef"], b["fe"], b["ff"]) for b in B]
        return C, np.dstack((C1, C2))

    def k_total_with_stress(self, data1, data2, tol=1e-10):
        kee, kef, kff = self._blocks(quirk=True)
        ffkw = {"tol": tol} if self.family == "rbf" else {}
        C = dict(ee=None, ef=None, fe=None, ff=None)
        C_se = C_sf = None
        for k1, d1 in data1.items():
            for k2, d2 in data2.items():
                if not (_nonempty(d1) and _nonempty(d2)):
                    continue
                if k1 == "energy" and k2 == "energy":
                    C["ee"] = kee(d1, d2)
                elif k1 == "energy" and k2 == "force":
                    C["ef"] = kef(d1, d2)
                elif k1 == "force" and k2 == "energy":
                    C["fe"], C_se = kef(d2, d1, stress=True, transpose=True)
                elif k1 == "force" and k2 == "force":
                    C["ff"], C_sf = kff(d1, d2, stress=True, **ffkw)
        return (build_covariance(C["ee"], C["ef"], C["fe"], C["ff"]),
                build_covariance(None, None, C_se, C_sf))


def _refkernel_diag(self, data):
    """Dot_mb.diag / RBF_mb.diag (eps-shifted NumPy formulas), see kernel_diag below."""
    return kernel_diag(self.family, self.para, self.zeta, data)


RefKernel.diag = _refkernel_diag


# ----------------------------------------------------------------------------- GP algebra
def gp_fit(K, y, n_energy, noise_e, f_coef):
    """gaussianprocess.py:105-127: K + noise -> Cholesky -> alpha.  Returns (L, alpha)."""
    from scipy.linalg import cholesky, cho_solve
    K = np.array(K, dtype=np.float64, copy=True)
    n = len(K)
    diag = np.full(n, (f_coef * noise_e) ** 2)
    diag[:n_energy] = noise_e ** 2
    K[np.diag_indices(n)] += diag
    L = cholesky(K, lower=True)
    alpha = cho_solve((L, True), np.asarray(y, dtype=np.float64).reshape(n, -1))
    return L, alpha


# ----------------------------------------------------------------------------- diag (eps variants)
def _single_pair_blocks(family, x1, x2, dx1, dx2, ele1, ele2, sigma2, p1, zeta, eps=1e-8):
    """NumPy restatement of the reference's single-pair functions K_ee / K_ff used by diag()
    (Dot_mb.py:177-258, RBF_mb.py:176-271): every norm carries +eps, nothing is skipped.
    p1 = sigma0^2 (Dot) or l^2 (RBF).  Returns (kee scalar, kff[3,3] or None)."""
    n1 = np.linalg.norm(x1, axis=1) + eps
    n2 = np.linalg.norm(x2, axis=1) + eps
    s = x1 @ x2.T
    same = (np.asarray(ele1)[:, None] == np.asarray(ele2)[None, :])
    d_ee = s / (eps + n1[:, None] * n2[None, :])
    if family == "dot":
        kee = (sigma2 * same * (d_ee ** zeta + p1)).sum() / (len(x1) * len(x2))
    else:
        kee = (sigma2 * np.exp(-(0.5 / p1) * (1 - d_ee ** zeta)) * same).sum() / (len(x1) * len(x2))
    if dx1 is None:
        return kee, None
    n12 = n1[:, None] * n2[None, :]
    d = s / n12 if family == "dot" else s / (eps + n12)
    D2 = d ** (zeta - 2)
    D1 = d * D2
    dd1 = (x2[None, :, :] * n1[:, None, None] - s[:, :, None] * (x1 / n1[:, None])[:, None, :]) / (n1 ** 2)[:, None, None] / n2[None, :, None]
    dd2 = (x1[:, None, :] * n2[None, :, None] - s[:, :, None] * (x2 / n2[:, None])[None, :, :]) / n1[:, None, None] / (n2 ** 2)[None, :, None]
    eye = np.eye(x1.shape[1])
    t33 = eye[None, :, :] - x2[:, :, None] * (x2 / (n2 ** 2)[:, None])[:, None, :]
    x1n3, x2n3 = x1 / (n1 ** 3)[:, None], x2 / (n2 ** 3)[:, None]
    out1 = (x1n3[:, None, :, None] * (x1[:, None, None, :] / n2[None, :, None, None])
            - x1n3[:, None, :, None] * x2n3[None, :, None, :] * s[:, :, None, None])
    d2d = t33[None, :, :, :] / n12[:, :, None, None] - out1
    pq = dd1[:, :, :, None] * dd2[:, :, None, :]
    d2D = zeta * (pq * (D2 * (zeta - 1))[:, :, None, None] + D1[:, :, None, None] * d2d)
    if family == "dot":
        M = d2D * (sigma2 * same)[:, :, None, None]
    else:
        k = sigma2 * np.exp(-(0.5 / p1) * (1 - d * D1)) * same
        zd2 = -0.5 / p1 * zeta * zeta * D1 ** 2
        M = (-d2D + zd2[:, :, None, None] * pq) * ((-0.5 / p1) * k)[:, :, None, None]
    kff = np.einsum("apk,abpq,bql->kl", dx1, M, dx2)
    return kee, kff


def kernel_diag(family, para, zeta, data):
    """Restates Dot_mb.diag / RBF_mb.diag (Dot_mb.py:45-85, RBF_mb.py:45-85)."""
    sigma2 = para[0] ** 2
    p1 = para[1] ** 2
    out = []
    if "energy" in data:
        E = data["energy"]
        if isinstance(E, tuple):
            x, ele, counts = E
            off = np.concatenate(([0], np.cumsum(counts)))
            E = [(x[off[i]:off[i + 1]], ele[off[i]:off[i + 1]]) for i in range(len(counts))]
        out.append(np.array([_single_pair_blocks(family, x, x, None, None, e, e, sigma2, p1, zeta)[0] for x, e in E]))
    if "force" in data:
        F = data["force"]
        if isinstance(F, tuple):
            x, dx, ele, counts = F
            off = np.concatenate(([0], np.cumsum(counts)))
            F = [(x[off[i]:off[i + 1]], dx[off[i]:off[i + 1]], ele[off[i]:off[i + 1]]) for i in range(len(counts))]
        vals = []
        for x, dx, e in F:
            dx3 = np.asarray(dx)[:, :, :3]
            vals.extend(np.diag(_single_pair_blocks(family, x, x, dx3, dx3, e, e, sigma2, p1, zeta)[1]))
        out.append(np.array(vals))
    return np.hstack(out)
