"""Per-group self covariance (`diag`) of the kernel objects (reference Dot_mb.py:45-85,
RBF_mb.py:45-85), used for the predictive variance (gaussianprocess.py:173,381).

The reference evaluates its NumPy single-pair functions K_ee / K_ff (Dot_mb.py:177-258,
RBF_mb.py:176-271), which add eps = 1e-8 to every row norm and never skip a row.  Here the same
quantity comes from the GPU tile kernels run in "diagonal" mode on a pack built with
norm_shift = 1e-8 (cspbo_pack_create_eps + cspbo_block_diag): only the (i,i) group pairs are
evaluated, one warp per 8 groups.
"""
import numpy as np

from .. import _lib
from ..packing import Pack, block_diag
from ..utilities import list_to_tuple

_EPS = 1e-8


def _family_args(kernel):
    if kernel.name == 'Dot_mb':
        return dict(family=_lib.DOT, sigma2=kernel.sigma ** 2, sigma02=kernel.sigma0 ** 2, l=1.0)
    return dict(family=_lib.RBF, sigma2=kernel.sigma ** 2, sigma02=0.0, l=kernel.l)


def kernel_diag(kernel, data):
    fa, zeta = _family_args(kernel), kernel.zeta
    C_ee, C_ff = None, None
    if "energy" in data:
        E = data["energy"]
        if isinstance(E, list):
            E = list_to_tuple(E, mode='energy')
        x, ele, indices = E
        p = Pack(x, None, ele, indices, norm_shift=_EPS)
        scale = fa["sigma2"] if fa["family"] == _lib.DOT else 1.0
        C_ee = block_diag(fa["family"], _lib.KEE, p, zeta, fa["sigma2"], fa["sigma02"], fa["l"], scale=scale)
        C_ee /= p.counts.astype(np.float64) ** 2
    if "force" in data:
        F = data["force"]
        if isinstance(F, list):
            F = [(f[0], np.asarray(f[1])[:, :, :3], f[2]) for f in F]
            F = list_to_tuple(F)
        x, dxdr, ele, indices = F
        d3 = dxdr[:, :, :3].contiguous() if type(dxdr).__module__.startswith("torch") else np.ascontiguousarray(np.asarray(dxdr)[:, :, :3])
        p = Pack(x, d3, ele, indices, norm_shift=_EPS)
        scale = fa["sigma2"] * zeta if fa["family"] == _lib.DOT else 1.0
        blocks = block_diag(fa["family"], _lib.KFF, p, zeta, fa["sigma2"], fa["sigma02"], fa["l"], scale=scale)
        C_ff = np.ascontiguousarray(np.diagonal(blocks, axis1=1, axis2=2)).reshape(-1)
    if C_ff is None:
        return C_ee
    if C_ee is None:
        return C_ff
    return np.hstack((C_ee, C_ff))
