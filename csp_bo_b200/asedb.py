"""Read-only access to ASE sqlite databases WITHOUT ase (the reference stores its training sets and
models that way: database/*.db, examples/models/*.db; written by gaussianprocess.py:594-625).

Row layout (ase.db.sqlite): table `systems`; `numbers` int32 blob, `positions` / `cell` float64 blobs, `pbc`
bit mask, `data` = either JSON text or a binary blob whose first 8 bytes are the little-endian offset of a
JSON tail in which arrays appear as {"__ndarray__": [shape, dtype, byte_offset]}.  The reference keeps
energy / force / stress / energy_in / force_in / group inside `data`.
"""
import json
import sqlite3
import struct

import numpy as np

_SYMBOLS = ("X H He Li Be B C N O F Ne Na Mg Al Si P S Cl Ar K Ca Sc Ti V Cr Mn Fe Co Ni Cu Zn Ga Ge As Se Br Kr "
            "Rb Sr Y Zr Nb Mo Tc Ru Rh Pd Ag Cd In Sn Sb Te I Xe Cs Ba La Ce Pr Nd Pm Sm Eu Gd Tb Dy Ho Er Tm Yb Lu "
            "Hf Ta W Re Os Ir Pt Au Hg Tl Pb Bi Po At Rn").split()


class Structure:
    """The subset of ase.Atoms the descriptor and the GP read: positions, numbers, cell, pbc, symbols."""

    def __init__(self, positions, numbers, cell, pbc=(True, True, True)):
        self.positions = np.ascontiguousarray(positions, dtype=np.float64).reshape(-1, 3)
        self.numbers = np.ascontiguousarray(numbers, dtype=np.int64).reshape(-1)
        self.cell = np.ascontiguousarray(cell, dtype=np.float64).reshape(3, 3)
        self.pbc = np.array(np.broadcast_to(np.asarray(pbc, dtype=bool), (3,)))

    @property
    def symbols(self):
        return [_SYMBOLS[int(z)] for z in self.numbers]

    def get_chemical_symbols(self):
        return self.symbols

    def get_atomic_numbers(self):
        return self.numbers

    def get_cell(self):
        return self.cell

    def get_volume(self):
        return float(abs(np.linalg.det(self.cell)))

    def __len__(self):
        return len(self.positions)


def _inline_arrays(obj):
    """Older ASE files (db version < 9) keep `data` as text JSON with arrays inlined as
    {'__ndarray__': [shape, dtype, flat_list]}: turn those into numpy arrays (complex data is out of scope)."""
    if isinstance(obj, dict):
        if '__ndarray__' in obj and len(obj) == 1:
            shape, dtype, flat = obj['__ndarray__']
            if isinstance(flat, (list, tuple)):
                return np.array(flat, dtype=dtype).reshape(shape)
            return obj
        return {k: _inline_arrays(v) for k, v in obj.items()}
    return obj


def _decode_data(blob):
    if blob is None:
        return {}
    if isinstance(blob, str):
        return _inline_arrays(json.loads(blob))
    b = bytes(blob)
    try:
        off = struct.unpack('<q', b[:8])[0]
        js = json.loads(b[off:].decode())
    except Exception:
        return _inline_arrays(json.loads(b.decode()))
    out = {}
    for k, v in js.items():
        if isinstance(v, dict) and '__ndarray__' in v:
            shape, dtype, o = v['__ndarray__']
            if isinstance(o, (list, tuple)):                 # inlined array inside a binary record
                out[k] = np.array(o, dtype=dtype).reshape(shape)
                continue
            n = int(np.prod(shape)) if len(shape) else 1
            out[k] = np.frombuffer(b, dtype=dtype, count=n, offset=o).reshape(shape).copy()
        else:
            out[k] = _inline_arrays(v)
    return out


def read_rows(path, limit=None):
    """List of dicts: atoms (Structure), energy, force [n,3], stress, energy_in, force_in, group, row id."""
    con = sqlite3.connect('file:' + path + '?mode=ro', uri=True)
    try:
        cur = con.cursor()
        q = "select id, numbers, positions, cell, pbc, data from systems order by id"
        if limit is not None:
            q += " limit %d" % int(limit)
        rows = []
        for rid, numbers, positions, cell, pbc, data in cur.execute(q):
            num = np.frombuffer(numbers, dtype=np.int32)
            pos = np.frombuffer(positions, dtype=np.float64).reshape(-1, 3)
            cl = np.frombuffer(cell, dtype=np.float64).reshape(3, 3) if cell is not None else np.zeros((3, 3))
            mask = [(int(pbc) >> k) & 1 == 1 for k in range(3)]
            d = _decode_data(data)
            rows.append({"id": rid, "atoms": Structure(pos, num, cl, mask),
                         "energy": d.get("energy"), "force": np.asarray(d["force"], dtype=np.float64) if "force" in d else None,
                         "stress": np.asarray(d["stress"], dtype=np.float64) if d.get("stress") is not None else None,
                         "energy_in": d.get("energy_in"), "force_in": list(d.get("force_in", []) or []),
                         "group": d.get("group")})
        return rows
    finally:
        con.close()


def rows_to_npz_dict(rows, prefix):
    """Flatten rows into plain arrays (for fixtures that must travel without the sqlite files)."""
    nat = np.array([len(r["atoms"]) for r in rows], dtype=np.int64)
    out = {prefix + "_natoms": nat,
           prefix + "_pos": np.concatenate([r["atoms"].positions for r in rows]),
           prefix + "_num": np.concatenate([r["atoms"].numbers for r in rows]).astype(np.int32),
           prefix + "_cell": np.stack([r["atoms"].cell for r in rows]),
           prefix + "_pbc": np.stack([r["atoms"].pbc for r in rows]),
           prefix + "_energy": np.array([np.nan if r["energy"] is None else r["energy"] for r in rows], dtype=np.float64),
           prefix + "_force": np.concatenate([r["force"] if r["force"] is not None else np.full((len(r["atoms"]), 3), np.nan) for r in rows]),
           prefix + "_energy_in": np.array([bool(r["energy_in"]) for r in rows]),
           prefix + "_force_in_n": np.array([len(r["force_in"]) for r in rows], dtype=np.int64),
           prefix + "_force_in": np.array([i for r in rows for i in r["force_in"]], dtype=np.int64)}
    return out


def rows_from_npz_dict(d, prefix):
    nat = d[prefix + "_natoms"]
    off = np.concatenate(([0], np.cumsum(nat)))
    foff = np.concatenate(([0], np.cumsum(d[prefix + "_force_in_n"])))
    rows = []
    for i in range(len(nat)):
        s, e = off[i], off[i + 1]
        rows.append({"id": i + 1, "atoms": Structure(d[prefix + "_pos"][s:e], d[prefix + "_num"][s:e], d[prefix + "_cell"][i], d[prefix + "_pbc"][i]),
                     "energy": float(d[prefix + "_energy"][i]), "force": d[prefix + "_force"][s:e].copy(),
                     "energy_in": bool(d[prefix + "_energy_in"][i]),
                     "force_in": [int(v) for v in d[prefix + "_force_in"][foff[i]:foff[i + 1]]], "stress": None, "group": None})
    return rows
