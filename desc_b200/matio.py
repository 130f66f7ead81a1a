"""Reading / writing the reference's MATLAB data (host-side plumbing only; needs scipy).

``model_out`` structs saved from MATLAB (``save('graph.mat', '-struct', 'model_out')`` or ``save('graph.mat',
'model_out')``) come back as the dict the Python entry points take: ``Ind`` m x 2, ``RijMat`` 3 x 3 x m,
``R_orig`` 3 x 3 x n, ``ErrVec`` (m,), plus whatever else the file holds.  ``save_mat`` writes results so that
MATLAB's ``load`` sees the reference's shapes (``S_vec`` 1 x m, rotations 3 x 3 x n).
"""
from __future__ import annotations

import numpy as np


def _unwrap(v):
    """scipy returns MATLAB structs as 1x1 object arrays of void records"""
    while isinstance(v, np.ndarray) and v.dtype == object and v.size == 1:
        v = v.ravel()[0]
    return v


def load_mat(path, struct="model_out"):
    """-> dict with at least ``Ind`` and ``RijMat`` in the layout ``desc_b200.DESC*`` expects."""
    import scipy.io
    raw = scipy.io.loadmat(path)
    d = {k: v for k, v in raw.items() if not k.startswith("__")}
    if struct in d:                                     # saved as one struct variable
        rec = _unwrap(d[struct])
        if getattr(rec, "dtype", None) is not None and rec.dtype.names:
            d = {name: _unwrap(rec[name]) for name in rec.dtype.names}
    if "Ind" not in d or "RijMat" not in d:
        raise ValueError("%s holds neither Ind/RijMat nor a struct %r with them" % (path, struct))
    out = dict(d)
    out["Ind"] = np.asfortranarray(np.asarray(d["Ind"], dtype=np.float64))
    R = np.asarray(d["RijMat"], dtype=np.float64)
    if R.ndim == 2:                                      # a single edge: MATLAB drops the trailing dimension
        R = R[:, :, None]
    out["RijMat"] = np.asfortranarray(R)
    if out["Ind"].ndim != 2 or out["Ind"].shape[1] != 2 or out["RijMat"].shape != (3, 3, out["Ind"].shape[0]):
        raise ValueError("Ind must be m x 2 and RijMat 3 x 3 x m")
    for key in ("R_orig", "Rij_orig"):
        if key in d:
            out[key] = np.asfortranarray(np.asarray(d[key], dtype=np.float64))
    if "ErrVec" in d:
        out["ErrVec"] = np.asarray(d["ErrVec"], dtype=np.float64).ravel()
    return out


def save_mat(path, **arrays):
    """Write results for MATLAB: vectors named ``S_vec`` / ``SVec`` / ``ErrVec`` become 1 x m rows (DESC.m:148)."""
    import scipy.io
    out = {}
    for k, v in arrays.items():
        a = np.asarray(v)
        if k in ("S_vec", "SVec", "ErrVec"):
            a = a.reshape(1, -1)
        out[k] = a
    scipy.io.savemat(path, out, do_compression=True)
