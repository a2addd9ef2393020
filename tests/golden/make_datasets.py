#!/usr/bin/env python
"""Converts the reference's ASE databases to plain-array fixtures (container only):
    examples/models/Si.db      -> si_model_*   (74 structures with energy_in / force_in flags; model Si.json)
    database/Si_244_DFT.db     -> si244_*      (244 structures, BASELINE.json config 2)
    database/Cu-EMT-300K.db    -> cu300_*      (100 x 32 atoms, config 4)
so that the GPU box can run real-data configurations without the reference tree or ase."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = os.environ.get("CSPBO_REFERENCE", "/root/reference")
from csp_bo_b200 import asedb  # noqa: E402

out = {}
for prefix, rel in (("si_model", "examples/models/Si.db"), ("si244", "database/Si_244_DFT.db"), ("cu300", "database/Cu-EMT-300K.db")):
    rows = asedb.read_rows(os.path.join(REF, rel))
    out.update(asedb.rows_to_npz_dict(rows, prefix))
    print(prefix, len(rows), "structures", int(out[prefix + "_natoms"].sum()), "atoms")
np.savez_compressed(os.path.join(HERE, "datasets.npz"), **out)
with open(os.path.join(HERE, "si_model.json"), "w") as fp:
    json.dump(json.load(open(os.path.join(REF, "examples/models/Si.json"))), fp, indent=1)
print(os.path.getsize(os.path.join(HERE, "datasets.npz")), "bytes")
