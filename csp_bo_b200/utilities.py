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
