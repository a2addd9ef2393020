"""ctypes binding of libcspbo_b200.so (include/cspbo_b200.h).

There is no CPU fallback: if the library is missing, or a call fails (no GPU, CUDA error),
an exception is raised.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_PKG, "lib")
LIB_PATH = os.environ.get("CSPBO_LIB") or os.path.join(LIB_DIR, "libcspbo_b200.so")     # CSPBO_LIB: A/B builds of the library

c_int, c_dbl, c_ll, c_vp = ctypes.c_int, ctypes.c_double, ctypes.c_longlong, ctypes.c_void_p

MEM_HOST, MEM_DEVICE = 0, 1
DOT, RBF = 0, 1
KEE, KEF, KFF = 0, 1, 2


class CspboError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cspbo_b200 error %d: %s" % (code, msg))
        self.code = code


class BlockDesc(ctypes.Structure):
    """cspbo_block_desc"""
    _fields_ = [("family", c_int), ("block", c_int), ("zeta", c_dbl), ("sigma2", c_dbl), ("sigma02", c_dbl),
                ("l", c_dbl), ("tol", c_dbl), ("use_tol", c_int), ("with_grad", c_int), ("stress", c_int),
                ("scale", c_dbl), ("scale_grad", c_dbl), ("symmetric", c_int)]


_E = [c_vp, c_vp, c_vp]          # x, ele, inds
_F = [c_vp, c_vp, c_vp, c_vp]    # x, dxdr, ele, inds

# name -> argtypes, exactly the declarations of include/cspbo_b200.h
SIGNATURES = {
    "cspbo_last_error": [],
    "cspbo_version": [],
    "cspbo_launch_count": [],
    "cspbo_set_device": [c_int],
    "cspbo_device_count": [ctypes.POINTER(c_int)],
    "cspbo_synchronize": [],
    "cspbo_release_scratch": [],
    "cspbo_dot_kee_many": [c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl] + _E + _E + [c_vp],
    "cspbo_dot_kef_many": [c_int, c_int, c_int, c_int, c_dbl] + _E + _F + [c_vp],
    "cspbo_dot_kef_many_stress": [c_int, c_int, c_int, c_int, c_dbl] + _E + _F + [c_vp],
    "cspbo_dot_kff_many": [c_int, c_int, c_int, c_int, c_int, c_int, c_dbl] + _F + _F + [c_vp],
    "cspbo_dot_kff_many_stress": [c_int, c_int, c_int, c_int, c_int, c_int, c_dbl] + _F + _F + [c_vp],
    "cspbo_rbf_kee_many": [c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl] + _E + _E + [c_vp],
    "cspbo_rbf_kee_many_with_grad": [c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl] + _E + _E + [c_vp, c_vp],
    "cspbo_rbf_kef_many": [c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl] + _E + _F + [c_vp],
    "cspbo_rbf_kef_many_with_grad": [c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl] + _E + _F + [c_vp],
    "cspbo_rbf_kef_many_stress": [c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl] + _E + _F + [c_vp],
    "cspbo_rbf_kff_many": [c_int, c_int, c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_dbl] + _F + _F + [c_vp],
    "cspbo_rbf_kff_many_with_grad": [c_int, c_int, c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl] + _F + _F + [c_vp, c_vp],
    "cspbo_rbf_kff_many_stress": [c_int, c_int, c_int, c_int, c_int, c_int, c_dbl, c_dbl, c_dbl, c_dbl] + _F + _F + [c_vp],
    "cspbo_pack_create": [c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int,
                          ctypes.POINTER(c_vp)],
    "cspbo_pack_create_eps": [c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_dbl,
                              ctypes.POINTER(c_vp)],
    "cspbo_pack_create_ex": [c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_int, c_dbl, c_int,
                             ctypes.POINTER(c_vp)],
    "cspbo_pack_destroy": [c_vp],
    "cspbo_pack_groups": [c_vp],
    "cspbo_pack_bytes": [c_vp],
    "cspbo_pack_pair_count": [c_vp, c_vp],
    "cspbo_block_build": [ctypes.POINTER(BlockDesc), c_vp, c_vp, c_vp, c_vp, c_int, c_int],
    "cspbo_block_build_into": [ctypes.POINTER(BlockDesc), c_vp, c_vp, c_vp, c_ll, c_ll, c_ll],
    "cspbo_block_diag": [ctypes.POINTER(BlockDesc), c_vp, c_vp, c_int],
    "cspbo_dev_alloc": [c_ll, ctypes.POINTER(c_vp)],
    "cspbo_dev_free": [c_vp],
    "cspbo_ipc_export": [c_vp, c_vp],
    "cspbo_ipc_open": [c_vp, ctypes.POINTER(c_vp)],
    "cspbo_ipc_close": [c_vp],
    "cspbo_copy_rows_to_host": [c_vp, c_ll, c_vp, c_ll, c_ll, c_ll, c_vp],
    "cspbo_sharded_rows": [c_vp, c_int, c_int],
    "cspbo_pack_order": [c_vp, c_vp, c_ll],
    "cspbo_block_build_sharded": [ctypes.POINTER(BlockDesc), c_vp, c_int, c_int, ctypes.POINTER(c_vp)],
    "cspbo_sharded_rows_ex": [c_vp, c_int, c_int, c_int, c_int],
    "cspbo_block_build_sharded_ex": [ctypes.POINTER(BlockDesc), c_vp, c_int, c_int, ctypes.POINTER(c_vp), c_int, c_int, c_int],
    "cspbo_block_build_sharded_ld": [ctypes.POINTER(BlockDesc), c_vp, c_int, c_int, ctypes.POINTER(c_vp), c_int, c_int, c_int, c_ll],
    "cspbo_block_build_sharded_slab": [ctypes.POINTER(BlockDesc), c_vp, c_int, c_int, ctypes.POINTER(c_vp), c_int, c_int, c_int, c_ll,
                                       c_int, c_int, c_vp, c_vp, c_vp],
    "cspbo_pack_block_work": [c_vp, c_vp, c_ll],
    "cspbo_block_build_panel": [ctypes.POINTER(BlockDesc), c_vp, c_vp, c_int, c_int, ctypes.POINTER(c_vp), c_ll, c_int, c_ll, c_ll],
    "cspbo_so3_create": [c_int, c_int, c_dbl, c_dbl, c_int, c_int, ctypes.POINTER(c_vp)],
    "cspbo_so3_destroy": [c_vp],
    "cspbo_so3_compute": [c_vp, c_int, c_vp, c_vp, c_vp, c_vp, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)],
    "cspbo_so3_device_ptrs": [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_vp)],
    "cspbo_so3_fetch": [c_vp, c_vp, c_vp, c_vp, c_vp],
    "cspbo_lj_compute": [c_int, c_vp, c_vp, c_vp, c_dbl, c_dbl, c_dbl, c_vp, c_vp, c_vp],
    "cspbo_gp_add_noise": [c_vp, c_int, c_int, c_dbl, c_dbl],
    "cspbo_gp_cholesky_solve": [c_vp, c_int, c_vp, c_int, ctypes.POINTER(c_dbl)],
    "cspbo_gp_cho_solve": [c_vp, c_int, c_vp, c_int],
    "cspbo_gp_cholesky_inverse": [c_vp, c_int, c_vp],
    "cspbo_gp_cholesky_inverse_inplace": [c_vp, c_int, c_int],
    "cspbo_gp_inverse_diag": [c_vp, c_int, c_vp],
    "cspbo_gp_inverse_workspace_bytes": [c_int],
    "cspbo_gp_tri_inverse_inplace": [c_vp, c_int],
    "cspbo_block_trace_slab": [ctypes.POINTER(BlockDesc), c_vp, c_vp, c_vp, c_vp, c_ll, c_ll, c_ll, c_ll, c_int, c_int, c_vp, c_int],
    "cspbo_block_trace": [ctypes.POINTER(BlockDesc), c_vp, c_vp, c_vp, c_vp, c_ll, c_ll, c_ll, c_vp, c_int],
    "cspbo_gp_predict": [c_vp, c_int, c_int, c_vp, c_vp],
    "cspbo_gp_chol_diag": [c_vp, c_ll, c_int, c_vp, c_vp, c_vp],
    "cspbo_gp_chol_apply_inv": [c_vp, c_int, c_vp, c_ll, c_ll, c_vp, c_ll, c_vp],
    "cspbo_gp_chol_update": [c_vp, c_ll, c_int, c_ll, c_vp, c_ll, c_vp, c_ll, c_int, c_vp],
    "cspbo_gp_stream_sm_target": [c_vp, c_int],
    "cspbo_gp_chol_streams": [c_int, ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_vp), ctypes.POINTER(c_vp),
                              ctypes.POINTER(c_int), ctypes.POINTER(c_int)],
}
_RESTYPE = {"cspbo_last_error": ctypes.c_char_p, "cspbo_launch_count": c_ll, "cspbo_pack_bytes": c_ll,
            "cspbo_pack_pair_count": c_ll, "cspbo_sharded_rows": c_ll, "cspbo_sharded_rows_ex": c_ll,
            "cspbo_gp_inverse_workspace_bytes": c_ll}

# reference-named shims (libdot_builder.py:5-9, librbf_builder.py:5-12)
SHIM_SYMBOLS = {
    "libdot_kernel_b200.so": ["kee_many", "kef_many", "kef_many_stress", "kff_many", "kff_many_stress"],
    "librbf_kernel_b200.so": ["kee_many", "kee_many_with_grad", "kef_many", "kef_many_with_grad", "kef_many_stress",
                              "kff_many", "kff_many_with_grad", "kff_many_stress"],
}

_lib = None


def load():
    """Load libcspbo_b200.so (once).  Raises if it has not been built: no fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("%s not found: build it with `python -m csp_bo_b200.build` "
                           "(csp_bo_b200 has no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _RESTYPE.get(name, c_int)
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise CspboError(rc, load().cspbo_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(load().cspbo_launch_count())


def device_count():
    n = c_int(0)
    rc = load().cspbo_device_count(ctypes.byref(n))
    return n.value if rc == 0 else 0
