"""Pins the oracle: the C port (oracle/kernels_port.c) and, when present, the compiled unmodified
reference (oracle/_ref) must reproduce tests/golden/golden_outputs.npz, which the unmodified
reference C++ produced from the reference's own fixtures (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from tests._cases import OracleImpl, golden_cases, max_norm_err


@pytest.fixture(scope="module")
def port_outputs(oracle_built, fixtures):
    return golden_cases(fixtures, OracleImpl("port"))


def test_port_matches_reference_goldens(port_outputs, golden):
    assert set(port_outputs) == set(golden)
    worst = {k: max_norm_err(port_outputs[k], golden[k]) for k in golden}
    bad = {k: v for k, v in worst.items() if not v <= 1e-13}
    assert not bad, bad


def test_compiled_reference_matches_goldens(oracle_built, fixtures, golden):
    from oracle import cpu_reference as cr
    if not cr.have_ref():
        pytest.skip("oracle/_ref not built (reference tree not mounted)")
    out = golden_cases(fixtures, OracleImpl("ref"))
    for k in golden:
        # diag goldens come from the reference's NumPy functions, the oracle restates them (1e-12);
        # everything else is the same compiled C++ and must reproduce to rounding of the thread sum
        assert max_norm_err(out[k], golden[k]) <= (1e-12 if "diag" in k else 1e-14), k


def test_kff_symmetry_and_psd(golden):
    K = golden["dot_kff_bigbig"]
    assert max_norm_err(K, K.T) < 1e-13
    w = np.linalg.eigvalsh(0.5 * (K + K.T))
    assert w.min() > -1e-8 * w.max()


def test_dot_hyper_gradient_is_scaled_copy(golden):
    # dot_kernel.py:55-57: dK_EE/dsigma = 2K/sigma
    assert np.allclose(golden["dot_kee_ds"], 2 * golden["dot_kee"] / 18.55544058601137, rtol=1e-15)


def test_rbf_dl_matches_finite_difference(oracle_built, fixtures):
    from tests._cases import force_tuple
    from oracle import cpu_reference as cr
    x2 = force_tuple(fixtures, "x2")
    s, l, h = 6.244955524643115, 0.5611029935713949, 1e-6
    _, _, dl = cr.rbf_kff_C(x2, x2, s, l, 2.0, grad=True, backend="port")
    kp = cr.rbf_kff_C(x2, x2, s, l + h, 2.0, tol=-1.0, backend="port")
    km = cr.rbf_kff_C(x2, x2, s, l - h, 2.0, tol=-1.0, backend="port")
    assert max_norm_err((kp - km) / (2 * h), dl) < 1e-6


def test_gp_fit_oracle(oracle_built, golden):
    from oracle import cpu_reference as cr
    K = golden["dot_ktotal_train"]
    rng = np.random.default_rng(0)
    y = rng.standard_normal(len(K))
    L, alpha = cr.gp_fit(K, y, n_energy=4, noise_e=0.005, f_coef=30)
    Kn = K.copy()
    Kn[np.diag_indices(len(K))] += np.r_[np.full(4, 0.005 ** 2), np.full(len(K) - 4, (30 * 0.005) ** 2)]
    assert np.allclose(Kn @ alpha[:, 0], y, atol=1e-8 * np.abs(y).max())


def test_bench_reference_arm_prints_one_contract_line(oracle_built):
    """`bench.py --impl reference` (the reference's own CPU code on a bounded slab) prints exactly one JSON line with the
    keys the driver reads; runs without a GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-500:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "K_FF entries/s (fp64)" and d["unit"] == "entries/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


def _random_sets(rng, d, mE=5, mF=9, species=(1, 8, 78)):
    """ragged random energy / force sets with zero-norm rows and 1-row groups (positive descriptors: pow(c, zeta) real)"""
    def rows(m, lo, hi):
        counts = rng.integers(lo, hi, size=m)
        counts[rng.integers(0, m)] = 1
        n = int(counts.sum())
        x = (np.abs(rng.standard_normal((n, d))) + 0.05) * rng.uniform(0.5, 50.0, size=(n, 1))
        x[rng.integers(0, n, size=2)] = 0.0
        return counts, x, rng.choice(species, size=n).astype(np.int64)
    ce, xe, ee = rows(mE, 3, 12)
    cf, xf, ef = rows(mF, 2, 8)
    dxf = rng.standard_normal((len(xf), d, 9)) * 3.0
    return (xe, ee, [int(c) for c in ce]), (xf, dxf, ef, [int(c) for c in cf])


@pytest.mark.parametrize("d,zeta", [(24, 2.0), (10, 1.0), (50, 2.5), (24, 3.0)])
def test_port_equals_compiled_reference_on_random_inputs(oracle_built, d, zeta):
    """The C restatement against the compiled UNMODIFIED reference beyond the fixtures: all wrappers (plain, grad,
    stress), ragged groups, zero-norm rows, d not a multiple of 4, non-integer zeta."""
    from oracle import cpu_reference as cr
    if not cr.have_ref():
        pytest.skip("oracle/_ref not built (reference tree not mounted)")
    rng = np.random.default_rng(int(d * 100 + zeta * 10))
    E1, F1 = _random_sets(rng, d)
    E2, F2 = _random_sets(rng, d, mE=4, mF=7)
    F1_3 = (F1[0], F1[1][:, :, :3], F1[2], F1[3])
    F2_3 = (F2[0], F2[1][:, :, :3], F2[2], F2[3])
    s, s0, l = 7.3, 0.02, 0.61

    def both(fn, *a, **k):
        return fn(*a, backend="port", **k), fn(*a, backend="ref", **k)

    cases = {
        "dot_kee": both(cr.dot_kee_C, E1, E2, s, s0, zeta, grad=True),
        "dot_kef": both(cr.dot_kef_C, E1, F2_3, s, s0, zeta),
        "dot_kef_stress": both(cr.dot_kef_C, E1, F2, s, s0, zeta, stress=True),
        "dot_kff": both(cr.dot_kff_C, F1_3, F2_3, s, s0, zeta),
        "dot_kff_stress": both(cr.dot_kff_C, F1, F2_3, s, s0, zeta, stress=True),
        "rbf_kee": both(cr.rbf_kee_C, E1, E2, s, l, zeta, grad=True),
        "rbf_kef": both(cr.rbf_kef_C, E1, F2_3, s, l, zeta),
        "rbf_kef_grad": both(cr.rbf_kef_C, E1, F2_3, s, l, zeta, grad=True),
        "rbf_kef_stress": both(cr.rbf_kef_C, E1, F2, s, l, zeta, stress=True),
        "rbf_kff": both(cr.rbf_kff_C, F1_3, F2_3, s, l, zeta, tol=1e-10),
        "rbf_kff_grad": both(cr.rbf_kff_C, F1_3, F2_3, s, l, zeta, grad=True),
        "rbf_kff_stress": both(cr.rbf_kff_C, F1, F2_3, s, l, zeta, stress=True, tol=1e-10),
    }
    for name, (p, r) in cases.items():
        p, r = (p, r) if isinstance(p, tuple) else ((p,), (r,))
        assert len(p) == len(r), name
        for a, b in zip(p, r):
            assert a.shape == b.shape, name
            assert max_norm_err(a, b) <= 1e-13 or np.abs(b).max() == 0.0, (name, max_norm_err(a, b))
