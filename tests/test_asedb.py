"""The ase-free ASE-sqlite reader against the committed plain-array fixtures (container only: the
sqlite files live in the reference tree)."""
import os

import numpy as np
import pytest

REF = os.environ.get("CSPBO_REFERENCE", "/root/reference")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reader_matches_fixtures():
    from csp_bo_b200 import asedb
    path = os.path.join(REF, "examples", "models", "Si.db")
    if not os.path.exists(path):
        pytest.skip("reference tree not mounted")
    rows = asedb.read_rows(path)
    d = dict(np.load(os.path.join(G, "datasets.npz")))
    back = asedb.rows_from_npz_dict(d, "si_model")
    assert len(rows) == len(back) == 74
    for a, b in zip(rows, back):
        assert np.array_equal(a["atoms"].positions, b["atoms"].positions) and np.array_equal(a["atoms"].numbers, b["atoms"].numbers)
        assert np.array_equal(a["atoms"].cell, b["atoms"].cell) and a["energy"] == b["energy"]
        assert np.array_equal(a["force"], b["force"]) and a["force_in"] == b["force_in"] and bool(a["energy_in"]) == b["energy_in"]
    assert rows[0]["atoms"].symbols[0] == "Si" and abs(rows[0]["atoms"].get_volume() - 1170.291223762801) < 1e-6


def test_structure_roundtrip_and_metrics():
    from csp_bo_b200 import asedb
    from csp_bo_b200.utilities import mae, r2, rmse, metric_single
    st = asedb.Structure([[0, 0, 0], [1, 1, 1]], [14, 8], np.eye(3) * 3)
    rows = [{"atoms": st, "energy": -1.0, "force": np.zeros((2, 3)), "energy_in": True, "force_in": [1], "stress": None, "group": None}]
    back = asedb.rows_from_npz_dict(asedb.rows_to_npz_dict(rows, "t"), "t")
    assert back[0]["force_in"] == [1] and back[0]["atoms"].symbols == ["Si", "O"] and len(back[0]["atoms"]) == 2
    y, p = np.array([1.0, 2.0, 3.0]), np.array([1.1, 1.9, 3.2])
    assert abs(rmse(y, p) - np.sqrt(0.06 / 3)) < 1e-12 and abs(mae(y, p) - 0.4 / 3) < 1e-12 and abs(r2(y, p) - (1 - 0.06 / 2)) < 1e-12
    assert metric_single(y, p, "Train Energy").startswith("Train Energy [   3]: R2 0.9700 MAE  0.133 RMSE  0.141")


def test_text_json_data_column(tmp_path):
    """Older ASE files (db version < 9) keep `data` as text JSON with inlined arrays
    ({'__ndarray__': [shape, dtype, flat_list]}); the reader turns them into numpy arrays."""
    import json
    import sqlite3
    from csp_bo_b200 import asedb
    path = str(tmp_path / "old.db")
    con = sqlite3.connect(path)
    con.execute("create table systems (id integer primary key, numbers blob, positions blob, cell blob, pbc integer, data text)")
    force = [[0.1, -0.2, 0.3], [0.0, 0.5, -0.5]]
    data = {"energy": -3.5, "force": {"__ndarray__": [[2, 3], "float64", sum(force, [])]}, "energy_in": True,
            "force_in": [0, 1]}
    con.execute("insert into systems values (1, ?, ?, ?, 7, ?)",
                (np.array([14, 14], dtype=np.int32).tobytes(), np.array([[0, 0, 0], [1.3, 1.3, 1.3]], dtype=np.float64).tobytes(),
                 (np.eye(3) * 5.4).tobytes(), json.dumps(data)))
    con.commit(); con.close()
    rows = asedb.read_rows(path)
    assert len(rows) == 1 and rows[0]["energy"] == -3.5 and rows[0]["force_in"] == [0, 1]
    assert isinstance(rows[0]["force"], np.ndarray) and np.allclose(rows[0]["force"], force)
    # the same inlined form inside the decode helper's fallback path
    d = asedb._decode_data(json.dumps(data).encode())
    assert d["force"].shape == (2, 3) and d["force"].dtype == np.float64
