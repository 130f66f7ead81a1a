"""Host-side mirror of the reference's data generators, on top of the C ABI.

* ``Uniform_Topology(n, p, q, sigma, model='uniform')``                          Models/Uniform_Topology.m:24
* ``Nonuniform_Topology(n, p, p_node_crpt, p_edge_crpt, sigma_in, sigma_out, crpt_type='uniform')``
                                                                                 Models/Nonuniform_Topology.m:26
* ``Ring_Topology(n, deg, window, q, sigma)`` -- the SfM-shaped graph of BASELINE.json configs[4] (no reference
  counterpart): ``Uniform_Topology`` restricted to pairs within circular distance ``window``, p = deg/(2 window).

Each returns ``model_out`` with the reference's field names (``Ind`` m x 2, ``RijMat`` 3x3xm, ``Rij_orig``,
``R_orig`` 3x3xn, ``ErrVec`` 1 x m; ``AdjMat`` only on request -- it is n x n dense).  Extra keyword arguments the
reference cannot express: ``seed`` (MATLAB's global RNG stream is replaced by counter-based draws),
``device``, and ``on_device=True`` to get a :class:`Model` whose buffers stay in HBM and can be passed straight to
``desc_b200.Solver`` / ``DESC*`` (``model.Ind``, ``model.RijMat``).  All arithmetic runs in ``csrc/gen.cu``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


class DeviceArray:
    """A float64 buffer in HBM owned by a :class:`Model` (just enough surface for ``Solver`` to take it)."""
    is_cuda = True
    dtype = "torch.float64"   # Solver's device-input check compares the dtype by name

    def __init__(self, ptr, count, owner):
        self._ptr, self._count, self._owner = int(ptr), int(count), owner

    def data_ptr(self):
        return self._ptr

    def numel(self):
        return self._count

    def is_contiguous(self):
        return True


class Model:
    """Generated inputs resident on the device (``desc_b200_model``)."""

    def __init__(self, opts):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.desc_b200_generate(C.byref(opts), C.byref(self._h)))
        info = (C.c_int64 * 4)()
        ms = C.c_double(0.0)
        _lib.check(self._lib.desc_b200_model_info(self._h, info, C.byref(ms)))
        self.n, self.m, self.launches, self.device = int(info[0]), int(info[1]), int(info[2]), int(info[3])
        self.gen_ms = float(ms.value)
        p = [C.c_void_p() for _ in range(4)]
        _lib.check(self._lib.desc_b200_model_device(self._h, *[C.byref(x) for x in p]))
        self.Ind = DeviceArray(p[0].value, 2 * self.m, self)
        self.RijMat = DeviceArray(p[1].value, 9 * self.m, self)
        self.R_orig_dev = DeviceArray(p[2].value, 9 * self.n, self)
        self.ErrVec_dev = DeviceArray(p[3].value, self.m, self)

    def to_host(self, want_adj=False):
        """``model_out`` of the reference as numpy arrays."""
        m, n = self.m, self.n
        Ind = np.empty((m, 2), dtype=np.float64, order="F")
        RijMat = np.empty((3, 3, m), dtype=np.float64, order="F")
        Rij_orig = np.empty((3, 3, m), dtype=np.float64, order="F")
        R_orig = np.empty((3, 3, n), dtype=np.float64, order="F")
        ErrVec = np.empty(m, dtype=np.float64)
        corr = np.empty(m, dtype=np.uint8)
        vp = lambda a: C.c_void_p(a.ctypes.data)   # noqa: E731
        _lib.check(self._lib.desc_b200_model_fetch(self._h, vp(Ind), vp(RijMat), vp(R_orig), vp(ErrVec), vp(Rij_orig),
                                                  vp(corr)))
        out = dict(Ind=Ind, RijMat=RijMat, Rij_orig=Rij_orig, R_orig=R_orig, ErrVec=ErrVec.reshape(1, -1),
                   corrupted=corr.astype(bool))
        if want_adj:                                    # Uniform_Topology.m:32 (n x n dense: on request only)
            A = np.zeros((n, n))
            i, j = Ind[:, 0].astype(np.int64) - 1, Ind[:, 1].astype(np.int64) - 1
            A[i, j] = 1.0
            A[j, i] = 1.0
            out["AdjMat"] = A
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.desc_b200_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _make(on_device, want_adj, **kw):
    opts = _lib.GenOpts(device=int(kw.pop("device", -1)), topology=int(kw.pop("topology", 0)), n=int(kw.pop("n")),
                        window=int(kw.pop("window", 0)), kind=int(kw.pop("kind")), reserved=0, p=float(kw.pop("p")),
                        q=float(kw.pop("q", 0.0)), sigma=float(kw.pop("sigma")), sigma_out=float(kw.pop("sigma_out", 0.0)),
                        p_node_crpt=float(kw.pop("p_node_crpt", 0.0)), p_edge_crpt=float(kw.pop("p_edge_crpt", 0.0)),
                        seed=int(kw.pop("seed", 0)) & (2 ** 64 - 1))
    assert not kw, kw
    mo = Model(opts)
    if on_device:
        return mo
    try:
        return mo.to_host(want_adj=want_adj)
    finally:
        mo.close()


def Uniform_Topology(n, p, q, sigma, model="uniform", seed=0, device=-1, on_device=False, want_adj=False):
    """``model_out = Uniform_Topology(n,p,q,sigma,model)`` (Models/Uniform_Topology.m:24).  As in the reference any
    ``model`` other than 'uniform' means self-consistent corruption (:76,83)."""
    return _make(on_device, want_adj, n=n, p=p, q=q, sigma=sigma, kind=0 if model == "uniform" else 1, seed=seed,
                 device=device)


def Nonuniform_Topology(n, p, p_node_crpt, p_edge_crpt, sigma_in, sigma_out, crpt_type="uniform", seed=0, device=-1,
                        on_device=False, want_adj=False):
    """``model_out = Nonuniform_Topology(n,p,p_node_crpt,p_edge_crpt,sigma_in,sigma_out,crpt_type)``
    (Models/Nonuniform_Topology.m:26)."""
    kinds = {"uniform": 2, "self-consistent": 3, "adv": 4}
    if crpt_type not in kinds:
        raise ValueError("crpt_type must be 'uniform', 'self-consistent' or 'adv' (Nonuniform_Topology.m:95-118)")
    return _make(on_device, want_adj, n=n, p=p, sigma=sigma_in, sigma_out=sigma_out, p_node_crpt=p_node_crpt,
                 p_edge_crpt=p_edge_crpt, kind=kinds[crpt_type], seed=seed, device=device)


def Ring_Topology(n, deg, window, q, sigma, model="uniform", seed=0, device=-1, on_device=False, want_adj=False):
    """SfM-shaped graph (BASELINE.json configs[4]): cameras on a closed track, node i may be connected to the nodes
    within circular distance ``window``, each with probability deg/(2 window) (mean degree ``deg``)."""
    return _make(on_device, want_adj, n=n, p=deg / (2.0 * window), q=q, sigma=sigma, topology=1, window=window,
                 kind=0 if model == "uniform" else 1, seed=seed, device=device)
