"""Covariance assembly helpers (reference cspbo/kernels/base.py:3-34)."""
import numpy as np

_LAYOUTS = {
    # which of (ee, ef, fe, ff) exist -> how they are glued   (base.py:13-26)
    (True, True, True, True): lambda ee, ef, fe, ff: np.block([[ee, ef], [fe, ff]]),
    (False, False, True, True): lambda ee, ef, fe, ff: np.hstack((fe, ff)),   # E/F in train, F in predict
    (True, True, False, False): lambda ee, ef, fe, ff: np.hstack((ee, ef)),   # E/F in train, E in predict
    (False, True, False, False): lambda ee, ef, fe, ff: ef,
    (True, False, False, False): lambda ee, ef, fe, ff: ee,
    (False, False, False, True): lambda ee, ef, fe, ff: ff,
    (False, False, True, False): lambda ee, ef, fe, ff: fe,
}


def build_covariance(c_ee, c_ef, c_fe, c_ff, c_se=None, c_sf=None):
    """[[K_EE, K_EF], [K_FE, K_FF]] from whichever blocks exist; None for any other combination,
    exactly like the reference's if/elif chain."""
    key = tuple(x is not None for x in (c_ee, c_ef, c_fe, c_ff))
    fn = _LAYOUTS.get(key)
    return None if fn is None else fn(c_ee, c_ef, c_fe, c_ff)


def get_mask(ele1, ele2):
    """Index arrays of the species-mismatched (row1,row2) pairs, or None (base.py:28-34)."""
    ids = np.nonzero(np.asarray(ele1)[:, None] != np.asarray(ele2)[None, :])
    return None if len(ids[0]) == 0 else ids
