"""GP algebra on the device (reference cspbo/gaussianprocess.py:105-127 fit, :138-141 predict):
K + noise -> cuSOLVER Cholesky -> alpha, and the K*.alpha GEMV, through the C ABI."""
import ctypes

import numpy as np

from . import _lib
from ._lib import c_vp


def _cuda(a):
    import torch
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def fit_alpha(K, y, n_energy, noise_e, f_coef, return_logdet=False, overwrite=False):
    """alpha = (K + noise)^-1 y.  K: [N,N] numpy array or CUDA tensor (symmetric), y: [N].
    Returns a numpy vector (or a CUDA tensor when K is one).  Raises CspboError(CSPBO_E_NOTPD) when
    K + noise is not positive definite, like the LinAlgError of gaussianprocess.py:115-125."""
    import torch
    lib = _lib.load()
    on_dev = isinstance(K, torch.Tensor) and K.is_cuda
    Kd = _cuda(K)
    if on_dev and not overwrite and Kd.data_ptr() == K.data_ptr():
        Kd = Kd.clone()
    N = Kd.shape[0]
    yd = _cuda(y).reshape(-1).clone()
    assert Kd.shape == (N, N) and yd.numel() == N
    _lib.check(lib.cspbo_gp_add_noise(c_vp(Kd.data_ptr()), N, int(n_energy), float(noise_e), float(f_coef)))
    logdet = ctypes.c_double(0.0)
    _lib.check(lib.cspbo_gp_cholesky_solve(c_vp(Kd.data_ptr()), N, c_vp(yd.data_ptr()), 1,
                                           ctypes.byref(logdet) if return_logdet else None))
    alpha = yd if on_dev else yd.cpu().numpy()
    return (alpha, logdet.value) if return_logdet else alpha


def predict_mean(Kstar, alpha):
    """K* . alpha (gaussianprocess.py:140,359)."""
    import torch
    lib = _lib.load()
    on_dev = isinstance(Kstar, torch.Tensor) and Kstar.is_cuda
    Kd, ad = _cuda(Kstar), _cuda(alpha).reshape(-1)
    n, N = Kd.shape
    assert ad.numel() == N
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(lib.cspbo_gp_predict(c_vp(Kd.data_ptr()), n, N, c_vp(ad.data_ptr()), c_vp(out.data_ptr())))
    return out if on_dev else out.cpu().numpy()


def cho_solve(L, Bt):
    """(L L^T)^-1 B for the factor left in K by fit_alpha(..., overwrite=True).
    Bt: CUDA tensor [N, nrhs] (any layout); returns a CUDA tensor [N, nrhs]."""
    import torch
    lib = _lib.load()
    N = L.shape[0]
    assert Bt.shape[0] == N
    X = Bt.T.contiguous().clone()          # row-major [nrhs, N] == column-major [N, nrhs]
    _lib.check(lib.cspbo_gp_cho_solve(c_vp(L.data_ptr()), N, c_vp(X.data_ptr()), int(X.shape[0])))
    return X.T


def cholesky_inverse(L):
    """K^-1 from the in-place factor (csrc/dense.cu from 1024 rows on, cuSOLVER potri below), CUDA tensor [N,N]."""
    import torch
    lib = _lib.load()
    N = L.shape[0]
    out = torch.empty_like(L)
    _lib.check(lib.cspbo_gp_cholesky_inverse(c_vp(L.data_ptr()), N, c_vp(out.data_ptr())))
    return out


def cholesky_inverse_inplace(L, mirror=False):
    """The factor in L becomes K^-1 in place (row-major upper triangle; both with mirror=True)."""
    _lib.check(_lib.load().cspbo_gp_cholesky_inverse_inplace(c_vp(L.data_ptr()), int(L.shape[0]), int(bool(mirror))))
    return L


def inverse_diag(L):
    """diag(K^-1) from the in-place factor, which is destroyed (triangular inverse in place, csrc/dense.cu, + row sums of
    squares): no second N x N array."""
    import torch
    out = torch.empty(L.shape[0], dtype=torch.float64, device=L.device)
    _lib.check(_lib.load().cspbo_gp_inverse_diag(c_vp(L.data_ptr()), int(L.shape[0]), c_vp(out.data_ptr())))
    return out


def tri_inverse_inplace(U):
    """U (upper triangular CUDA tensor [n,n], contiguous; its lower triangle is ignored) -> U^-1 in place, lower triangle zeroed."""
    _lib.check(_lib.load().cspbo_gp_tri_inverse_inplace(c_vp(U.data_ptr()), int(U.shape[0])))
    return U.triu_()


def inverse_workspace_bytes(N):
    """Device workspace (library block cache, not torch's allocator) that inverse_diag / cholesky_inverse* use for an
    N x N factor: leaf results + the chunk buffer of the out-of-place triangular products."""
    return int(_lib.load().cspbo_gp_inverse_workspace_bytes(int(N)))
