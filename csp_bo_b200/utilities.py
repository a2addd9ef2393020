"""Data packing helpers with the reference's formats (cspbo/utilities.py:350-416).

"energy" data:  list of (x[n_i,d], ele[n_i])              <-> stacked (X, ELE, indices)
"force"  data:  list of (x[n_i,d], dxdr[n_i,d,3|9], ele)  <-> stacked (X, dXdR, ELE, indices)
`indices` is the list of row counts per group (structure / force atom).
"""
import numpy as np


def list_to_tuple(data, stress=False, include_value=False, mode='force'):
    """Stack a list of per-group tuples.  Mirrors cspbo/utilities.py:350-400, including the
    optional target values carried as the middle element when include_value=True."""
    data = list(data)
    if len(data) == 0:
        raise ValueError("list_to_tuple needs at least one group")
    indices = [int(np.shape(fd[0])[0]) for fd in data]
    X = np.concatenate([np.asarray(fd[0], dtype=np.float64) for fd in data], axis=0)
    ELE = np.ravel(np.concatenate([np.ravel(fd[-1]) for fd in data]))
    values = [fd[-2] for fd in data] if include_value else None
    if mode == 'force':
        length = 9 if stress else 3
        dXdR = np.concatenate([np.asarray(fd[1], dtype=np.float64) for fd in data], axis=0)
        if dXdR.shape[1:] != (X.shape[1], length):
            raise ValueError("dxdr blocks must have shape [n, %d, %d] (stress=%s), got %s"
                             % (X.shape[1], length, stress, dXdR.shape))
        return (X, dXdR, ELE, indices, values) if include_value else (X, dXdR, ELE, indices)
    return (X, ELE, indices, values) if include_value else (X, ELE, indices)


def tuple_to_list(data, mode='force'):
    """Inverse of list_to_tuple (cspbo/utilities.py:403-416)."""
    out, c = [], 0
    if mode == 'force':
        X, dXdR, ELE, indices = data
        for ind in indices:
            out.append((X[c:c + ind], dXdR[c:c + ind], ELE[c:c + ind]))
            c += ind
    else:
        X, ELE, indices = data
        for ind in indices:
            out.append((X[c:c + ind], ELE[c:c + ind]))
            c += ind
    return out


def new_pt(data, Refs, d_tol=1e-1, eps=1e-8):
    """Is the descriptor (X, ele) farther than d_tol (in 1 - cos^2) from every same-species reference?
    (cspbo/utilities.py:32-42, used by the active-learning selection of add_structure; the reference's
    `norm(X1 + eps)` is kept as written)."""
    (X, ele) = data
    X = X / (np.linalg.norm(X) + eps)
    for (X1, ele1) in Refs:
        if ele1 == ele:
            X1 = X1 / np.linalg.norm(X1 + eps)
            d = X @ X1.T
            if 1 - d ** 2 < d_tol:
                return False
    return True


# ---------------------------------------------------------------------------------------------
# metrics printed by the reference's example scripts (cspbo/utilities.py:44-90)
def rmse(true, predicted):
    true, predicted = np.asarray(true), np.asarray(predicted)
    return float(np.sqrt(np.sum((true - predicted) ** 2) / len(true)))


def mae(true, predicted):
    true, predicted = np.asarray(true), np.asarray(predicted)
    return float(np.sum(np.abs(true - predicted)) / len(true))


def r2(true, predicted):
    true, predicted = np.asarray(true), np.asarray(predicted)
    t_bar = np.sum(true) / len(true)
    square_error = np.sum((true - predicted) ** 2)
    true_variance = np.sum((true - t_bar) ** 2)
    return float(1 - square_error / true_variance)


def metric_single(y_train, y_train_pred, header, show_max=False):
    """Prints and returns the label of cspbo/utilities.py:80-90, e.g. 'Train Energy [ 16]: R2 0.9480 MAE ...'."""
    r2_train, mae_train, rmse_train = r2(y_train, y_train_pred), mae(y_train, y_train_pred), rmse(y_train, y_train_pred)
    str1 = "{:s} [{:4d}]: ".format(header, len(y_train))
    str1 += "R2 {:6.4f} MAE {:6.3f} RMSE {:6.3f}".format(r2_train, mae_train, rmse_train)
    if show_max:
        str1 += " Max {:6.3f}".format(float(np.max(np.abs(np.asarray(y_train) - np.asarray(y_train_pred)))))
    print(str1)
    return str1


def convert_train_data(data, des, N_force=100000):
    """[(structure, energy, forces)] -> {'energy', 'force', 'db'} training dictionary through the descriptor
    `des` (cspbo/utilities.py:104-136; every atom contributes a force point until N_force is reached)."""
    from .gaussianprocess import _Z
    energy_data, force_data, db_data = [], [], []
    for (struc, energy, forces) in data:
        d = des.calculate(struc)
        ele = np.array([e if isinstance(e, (int, np.integer)) else _Z[e] for e in d['elements']])
        f_ids = []
        for i in range(len(struc)):
            if len(force_data) < N_force:
                ids = np.argwhere(d['seq'][:, 1] == i).flatten()
                _i = d['seq'][ids, 0]
                force_data.append((d['x'][_i, :], d['dxdr'][ids], forces[i], ele[_i]))
                f_ids.append(i)
        energy_data.append((d['x'], energy / len(struc), ele))
        db_data.append((struc, energy, forces, True, f_ids))
    return {"energy": energy_data, "force": force_data, "db": db_data}
