#!/usr/bin/env python
"""Regenerates tests/golden/fixtures.npz and tests/golden/golden_outputs.npz.

Run in the build container only (needs /root/reference and `make -C oracle ref`):

    python tests/golden/make_golden.py

fixtures.npz      the reference's own test inputs, converted from pickled object arrays
                  (Cpp_code/X1_big.npy, X2.npy, Kff_QZ/X1_EE.npy, Kff_QZ/X2_EE.npy) to plain
                  float64/int32 arrays so that nothing on the GPU box needs the reference tree
                  or allow_pickle.  X1.npy holds the same 27 atoms as X2.npy in list form; the
                  list form is rebuilt from X2 by the tests.
golden_outputs.npz  outputs of the UNMODIFIED reference C++ (oracle/_ref, g++ -O2) on those
                  inputs with the hyper-parameters the reference uses with them
                  (Cpp_code/K_FF.py:180, cspbo/kernels/dot_kernel.py:283-284,
                  Cpp_code/RBF/rbf_kernel.py:287-288, examples/models/Si.json).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("CSPBO_REFERENCE", "/root/reference")

from oracle import cpu_reference as cr  # noqa: E402

from tests._cases import OracleImpl, golden_cases  # noqa: E402


def load_fixtures_from_reference():
    ld = lambda p: np.load(os.path.join(REF, "Cpp_code", p), allow_pickle=True)
    big, x2 = ld("X1_big.npy"), ld("X2.npy")
    e1, e2 = ld("Kff_QZ/X1_EE.npy"), ld("Kff_QZ/X2_EE.npy")
    fx = {}
    for name, f in (("big", big), ("x2", x2)):
        fx[name + "_x"] = np.asarray(f[0], np.float64)
        fx[name + "_dxdr"] = np.asarray(f[1], np.float64)
        fx[name + "_ele"] = np.asarray(f[2], np.int32)
        fx[name + "_counts"] = np.asarray(f[3], np.int32)
    for name, f in (("ee1", e1), ("ee2", e2)):
        fx[name + "_x"] = np.asarray(f[0], np.float64)
        fx[name + "_ele"] = np.asarray(f[1], np.int32)
        fx[name + "_counts"] = np.asarray(f[2], np.int32)
    # sanity: X1.npy is the list form of X2.npy
    x1 = ld("X1.npy")
    assert np.array_equal(np.concatenate([r[0] for r in x1]), fx["x2_x"])
    return fx


def main():
    assert cr.have_ref(), "run `make -C oracle ref` first"
    fx = load_fixtures_from_reference()
    np.savez(os.path.join(HERE, "fixtures.npz"), **fx)
    out = golden_cases(fx, OracleImpl("ref"))
    # diag(): the goldens come from the reference's OWN NumPy single-pair functions, exec'd from source
    # (importing the package needs mpi4py/ase); the oracle's restatement must agree with them.
    from tests._cases import energy_list, energy_tuple, force_list, force_tuple, stress_columns, take_groups, take_groups_e
    big, ee2 = force_tuple(fx, "big"), energy_tuple(fx, "ee2")
    tf = take_groups(big, 0, 12)
    test = {"energy": take_groups_e(ee2, 4, 6), "force": tf}
    test_l = {"energy": energy_list(test["energy"]),
              "force": force_list((tf[0], stress_columns(tf[1]), tf[2], tf[3]))}
    for fam, fname, cls, para in (("dot", "Dot_mb.py", "Dot_mb", [18.55544058601137, 0.01]),
                                  ("rbf", "RBF_mb.py", "RBF_mb", [6.244955524643115, 0.5611029935713949])):
        src = [l for l in open(os.path.join(REF, "cspbo", "kernels", fname)).read().split("\n") if not l.startswith("from .")]
        ns = {}
        exec("def build_covariance(*a, **k): pass\n"
             "def get_mask(ele1, ele2):\n    import numpy as np\n    ids = np.where((ele1[:,None] - ele2[None,:]) != 0)\n"
             "    return None if len(ids[0]) == 0 else ids\nkee_C = kef_C = kff_C = None\n" + "\n".join(src), ns)
        k = ns[cls](para=para, zeta=2.0)
        ref_list = k.diag({"energy": test_l["energy"], "force": [(a, b[:, :, :3], c) for a, b, c in test_l["force"]]})
        for key in (fam + "_diag", fam + "_diag_list"):
            assert np.abs(out[key] - ref_list).max() / np.abs(ref_list).max() < 1e-12, key
            out[key] = ref_list
    np.savez_compressed(os.path.join(HERE, "golden_outputs.npz"), **out)
    for k, v in out.items():
        print("%-28s %-14s max|.|=%.6e" % (k, v.shape, np.abs(v).max()))


if __name__ == "__main__":
    main()
