"""Real-data configuration: the reference's shipped Si model (examples/models/Si.json + Si.db, here as
tests/golden/si_model.json + datasets.npz) loaded, refitted and validated on the GPU, then used to
predict the 244 structures of database/Si_244_DFT.db (BASELINE.json config 2).  Mirrors what
examples/example_validate.py prints."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from csp_bo_b200 import asedb
from csp_bo_b200.gaussianprocess import GaussianProcess
from csp_bo_b200.utilities import metric_single

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
d = dict(np.load(os.path.join(G, "datasets.npz")))
rows = asedb.rows_from_npz_dict(d, "si_model")
t0 = time.perf_counter()
model = GaussianProcess(kernel=None)
model.load(os.path.join(G, "si_model.json"), db_rows=rows, opt="--opt" in sys.argv)
torch.cuda.synchronize(); t_load = time.perf_counter() - t0
print(model)
E, E1, F, F1 = model.validate_data()
metric_single(E, E1, "Train Energy"); metric_single(F, F1, "Train Forces")
test = asedb.rows_from_npz_dict(d, "si244")
t0 = time.perf_counter()
Es, Es1, Fs, Fs1 = [], [], [], []
for r in test:
    e, f, s = model.predict_structure(r["atoms"], stress=True)
    Es.append(r["energy"] / len(r["atoms"])); Es1.append(e / len(r["atoms"])); Fs.append(r["force"].ravel()); Fs1.append(f.ravel())
torch.cuda.synchronize(); t_test = time.perf_counter() - t0
metric_single(np.array(Es), np.array(Es1), "Test Energy"); metric_single(np.concatenate(Fs), np.concatenate(Fs1), "Test Forces")
print("load+fit %.2f s (%d E + %d F points); %d test structures (%d atoms) in %.2f s = %.1f ms/structure"
      % (t_load, model.N_energy, model.N_forces, len(test), sum(len(r["atoms"]) for r in test), t_test, 1e3 * t_test / len(test)))
