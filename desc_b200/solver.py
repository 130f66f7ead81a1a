"""Host-side mirror of the reference's solver interface, on top of the C ABI.

Function names, argument meaning and outputs follow the reference's MATLAB functions:

* ``DESC_PGD(Ind, RijMat, params)  -> S_vec``             Algorithms/DESC_PGD.m:14
* ``DESC_init(Ind, RijMat, params) -> (R_est, S_vec)``    Algorithms/DESC_init.m:14
* ``DESC(Ind, RijMat, params) -> (R_est, R_init, S_vec)`` Algorithms/DESC.m:14
* ``GCW(Ind, AdjMat, RijMat, S_vec) -> R_est``            Utils/GCW.m:1
* step rules ``ConstantStepSize``, ``PiecewiseStepSize``, ``HybridGradient``  (Utils/*.m)

* ``CEMP(Ind, RijMat, CEMP_parameters) -> SVec``          Algorithms/CEMP.m:25      
* ``CEMP_GCW(Ind, RijMat, CEMP_parameters) -> R_est``     Algorithms/CEMP_GCW.m:25
* ``MPLS(Ind, RijMat, CEMP_parameters, MPLS_parameters) -> (R_est, R_init)``   Algorithms/MPLS.m:28
* ``Spectral(Ind, RijMat) -> R_est``                      Algorithms/Spectral.m:15
* ``Rotation_Alignment(R_est, R_gt) -> (R_out, R_align, mean_error, median_error)``  Utils/Rotation_Alignment.m:13

``Ind`` is the reference's m x 2, 1-based, i<j, (i,j)-sorted edge list; ``RijMat`` is
3 x 3 x m; ``params`` is a dict with the reference's field names (``iters``, ``Gradient``,
``make_plots``, ``ErrVec``, ``R_orig``; ``learning_rate`` is accepted and ignored exactly as
in DESC.m:169).  Extra, optional keys that the reference cannot express: ``n_sample``
(0 = reference rule), ``seed`` (sampler seed), ``cycles`` (explicit ``(ptr, apex)`` lists),
``verbose`` (print the reference's per-iteration progress line, DESC.m:241).  ``make_plots=True``
(DESC.m:235-239) needs ``ErrVec`` and ``R_orig``; the four convergence curves are computed on the device and
left in ``desc_b200.solver.last_diagnostics`` (and drawn like DESC.m:315-344 when matplotlib is importable).

All arithmetic happens on the GPU inside ``libdesc_b200.so``.  Nothing here computes.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import DescError, StepRule, Timings, Opts


# ---------------------------------------------------------------------------------------
# Step rules: parameter holders with the reference's property names.  GetStep is evaluated
# on the device inside the fused PGD kernel; the objects only carry state (``t``).
# ---------------------------------------------------------------------------------------
class ConstantStepSize:
    """Utils/ConstantStepSize.m: ``step = -learning_rate * grad``."""

    def __init__(self, learning_rate):
        self.learning_rate = float(learning_rate)

    def _to_c(self):
        return StepRule(kind=0, strategy=0, lr=self.learning_rate, decay_interval=1.0, beta_1=0.0, beta_2=0.0, t=0)

    def _from_c(self, r):
        pass


class PiecewiseStepSize:
    """Utils/PiecewiseStepSize.m: ``t++; step = -lr/(fix(t/decay_interval)+1) * grad``."""

    def __init__(self, learning_rate, decay_interval):
        self.learning_rate = float(learning_rate)
        self.decay_interval = float(decay_interval)
        self.t = 0

    def _to_c(self):
        return StepRule(kind=1, strategy=0, lr=self.learning_rate, decay_interval=self.decay_interval,
                        beta_1=0.0, beta_2=0.0, t=int(self.t))

    def _from_c(self, r):
        self.t = int(r.t)


class HybridGradient:
    """Utils/HybridGradient.m: Adam (strategy 0) or 100x decayed SGD (strategy 1, after ``stopAdam``).

    The Adam moments ``m_t``/``v_t`` live on the device inside the solver handle (they are
    m_cycle long); they are zeroed when ``t == 0`` as HybridGradient.m:24-27 does.
    """

    def __init__(self, lr, beta_1, beta_2, decay_interval):
        self.lr, self.beta_1, self.beta_2 = float(lr), float(beta_1), float(beta_2)
        self.decay_interval = float(decay_interval)
        self.t = 0
        self.strategy = 0

    def stopAdam(self):
        self.strategy = 1
        return self

    def _to_c(self):
        return StepRule(kind=2, strategy=int(self.strategy), lr=self.lr, decay_interval=self.decay_interval,
                        beta_1=self.beta_1, beta_2=self.beta_2, t=int(self.t))

    def _from_c(self, r):
        self.t = int(r.t)


# ---------------------------------------------------------------------------------------
# buffers
# ---------------------------------------------------------------------------------------
def _is_device(x):
    return hasattr(x, "data_ptr") and getattr(x, "is_cuda", False)


def _host_f64(x, shape_check=None):
    a = np.asfortranarray(np.asarray(x, dtype=np.float64))
    if shape_check is not None:
        shape_check(a)
    return a


def _ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


class Solver:
    """One graph on one GPU (or one rank of a multi-GPU solve): owns a ``desc_b200_handle``.

    ``Ind``/``RijMat`` may be numpy arrays (copied to the device by the library) or CUDA torch
    tensors already laid out like MATLAB memory (``Ind``: 2m doubles, all i then all j;
    ``RijMat``: 9m doubles, element (r,c) of edge e at 9e+r+3c), in which case they are used
    in place (``RijMat`` must outlive the solver).
    """

    def __init__(self, Ind, RijMat, n=0, device=-1, stream=None, rank=0, world=1, nccl_id=None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self._keep = []
        flags = 0
        if _is_device(Ind) != _is_device(RijMat):
            raise ValueError("Ind and RijMat must both be host arrays or both be CUDA tensors")
        if _is_device(Ind):
            flags |= _lib.INPUTS_ON_DEVICE
            m = int(Ind.numel() // 2)
            if Ind.numel() != 2 * m or RijMat.numel() != 9 * m or str(Ind.dtype) != "torch.float64" \
                    or str(RijMat.dtype) != "torch.float64" or not Ind.is_contiguous() or not RijMat.is_contiguous():
                raise ValueError("device inputs must be contiguous float64 with 2m and 9m elements")
            ind_buf, r_buf = Ind, RijMat
        else:
            Ind = np.asarray(Ind)
            if Ind.ndim != 2 or Ind.shape[1] != 2:
                raise ValueError("Ind must be m x 2")
            RijMat = np.asarray(RijMat)
            if RijMat.ndim != 3 or RijMat.shape[:2] != (3, 3) or RijMat.shape[2] != Ind.shape[0]:
                raise ValueError("RijMat must be 3 x 3 x m")
            m = Ind.shape[0]
            ind_buf, r_buf = _host_f64(Ind), _host_f64(RijMat)
        self._keep += [ind_buf, r_buf]
        idbuf = None
        if world > 1:
            idbuf = C.create_string_buffer(bytes(nccl_id), 128)
            self._keep.append(idbuf)
        opts = Opts(device=int(device), flags=flags, stream=C.c_void_p(stream) if stream else None,
                    rank=int(rank), world=int(world), nccl_id=C.cast(idbuf, C.c_void_p) if idbuf else None)
        _lib.check(self._lib.desc_b200_create(C.byref(self._h), int(n), int(m), _ptr(ind_buf), _ptr(r_buf),
                                             C.byref(opts)))
        self.m = m
        self.rank, self.world = int(rank), int(world)

    # -- lifetime ------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.desc_b200_destroy(self._h)
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

    # -- stages (DESC.m:19-263) ------------------------------------------------------------
    def build_incidence(self, n_sample=0, seed=0, cycles=None):
        ptr = apex = None
        if cycles is not None:
            ptr = np.ascontiguousarray(cycles[0], dtype=np.int64)
            apex = np.ascontiguousarray(cycles[1], dtype=np.int32)
            if ptr.size != self.m + 1 or apex.size != int(ptr[-1]):
                raise ValueError("cycles=(ptr, apex): ptr needs m+1 entries and apex ptr[-1] entries")
        _lib.check(self._lib.desc_b200_build_incidence(self._h, int(n_sample), int(seed) & (2 ** 64 - 1),
                                                      _ptr(ptr), _ptr(apex)))
        return self.info()

    def cycle_inconsistency(self):
        _lib.check(self._lib.desc_b200_cycle_inconsistency(self._h))

    def pgd(self, iters, rule, want_S=True, want_hist=True):
        iters = int(iters)
        S = np.empty(self.m, dtype=np.float64) if want_S else None
        hist = np.zeros((max(iters, 1), 2), dtype=np.float64) if want_hist else None
        r = rule._to_c()
        run = C.c_int32(0)
        _lib.check(self._lib.desc_b200_pgd(self._h, iters, C.byref(r), _ptr(S), _ptr(hist), C.byref(run)))
        rule._from_c(r)
        k = int(run.value)
        return S, (hist[:k] if hist is not None else None), k

    def gcw(self, S_vec=None, want_R=True):
        info = self.info()
        S = None if S_vec is None else np.ascontiguousarray(np.asarray(S_vec, dtype=np.float64).ravel())
        if S is not None and S.size != self.m:
            raise ValueError("S_vec must have m entries")
        R = np.empty((3, 3, info["n"]), dtype=np.float64, order="F") if want_R else None
        _lib.check(self._lib.desc_b200_gcw(self._h, _ptr(S), _ptr(R)))
        return R

    def refine(self, S_vec=None, R_init=None):
        """DESC step 5 (DESC.m:265-312): weighted Lie-algebraic averaging started from R_init (default: the
        rotations of the last gcw) with S_vec (default: the last pgd).  Returns (R_est 3x3xn, scores)."""
        info = self.info()
        S = None if S_vec is None else np.ascontiguousarray(np.asarray(S_vec, dtype=np.float64).ravel())
        if S is not None and S.size != self.m:
            raise ValueError("S_vec must have m entries")
        R0 = None
        if R_init is not None:
            R0 = np.asfortranarray(np.asarray(R_init, dtype=np.float64))
            if R0.shape != (3, 3, info["n"]):
                raise ValueError("R_init must be 3 x 3 x n")
        R = np.empty((3, 3, info["n"]), dtype=np.float64, order="F")
        scores = np.zeros(100, dtype=np.float64)
        run = C.c_int32(0)
        _lib.check(self._lib.desc_b200_refine(self._h, _ptr(S), _ptr(R0), _ptr(R), C.byref(run), _ptr(scores)))
        return R, scores[:run.value].copy()

    # -- CEMP on the same incidence, evaluation, diagnostics -----------
    def cemp(self, max_iter, reweighting):
        """CEMP.m:98-129 on the handle's incidence (after build_incidence + cycle_inconsistency) -> SVec (m,)."""
        beta = np.ascontiguousarray(np.asarray(reweighting, dtype=np.float64).ravel())
        if int(max_iter) > 0 and beta.size == 0:
            raise ValueError("reweighting must not be empty")
        S = np.empty(self.m, dtype=np.float64)
        _lib.check(self._lib.desc_b200_cemp(self._h, int(max_iter), _ptr(beta), int(beta.size), _ptr(S)))
        return S

    def cemp_gcw(self, SVec=None):
        """CEMP_GCW.m:127-159 -> R_est 3x3xn (SVec default: the last cemp on this handle)."""
        S = None if SVec is None else np.ascontiguousarray(np.asarray(SVec, dtype=np.float64).ravel())
        if S is not None and S.size != self.m:
            raise ValueError("SVec must have m entries")
        R = np.empty((3, 3, self.info()["n"]), dtype=np.float64, order="F")
        _lib.check(self._lib.desc_b200_cemp_gcw(self._h, _ptr(S), _ptr(R)))
        return R

    def spectral(self):
        """Algorithms/Spectral.m:15-47 -> R_est 3x3xn."""
        R = np.empty((3, 3, self.info()["n"]), dtype=np.float64, order="F")
        _lib.check(self._lib.desc_b200_spectral(self._h, _ptr(R)))
        return R

    def mst_init(self, SVec=None):
        """MPLS.m:152-195: rotations multiplied along the minimum spanning tree of SVec+1 -> R 3x3xn (CEMP+MST)."""
        S = None if SVec is None else np.ascontiguousarray(np.asarray(SVec, dtype=np.float64).ravel())
        if S is not None and S.size != self.m:
            raise ValueError("SVec must have m entries")
        R = np.empty((3, 3, self.info()["n"]), dtype=np.float64, order="F")
        _lib.check(self._lib.desc_b200_mst_init(self._h, _ptr(S), _ptr(R)))
        return R

    def mpls_refine(self, MPLS_parameters, SVec=None, R_init=None):
        """MPLS.m:198-256 -> (R_est 3x3xn, scores).  Defaults: SVec of the last cemp, R_init of the last mst_init."""
        n = self.info()["n"]
        S = None if SVec is None else np.ascontiguousarray(np.asarray(SVec, dtype=np.float64).ravel())
        if S is not None and S.size != self.m:
            raise ValueError("SVec must have m entries")
        R0 = None
        if R_init is not None:
            R0 = np.asfortranarray(np.asarray(R_init, dtype=np.float64))
            if R0.shape != (3, 3, n):
                raise ValueError("R_init must be 3 x 3 x n")
        vecs = [np.ascontiguousarray(np.asarray(_param(MPLS_parameters, k), dtype=np.float64).ravel())
                for k in ("reweighting", "thresholding", "cycle_info_ratio")]
        if any(v.size == 0 for v in vecs):
            raise ValueError("MPLS_parameters.reweighting / thresholding / cycle_info_ratio must not be empty")
        max_iter = int(_param(MPLS_parameters, "max_iter"))
        p = _lib.MplsParams(stop_threshold=float(_param(MPLS_parameters, "stop_threshold")), max_iter=max_iter,
                            n_reweighting=vecs[0].size, n_thresholding=vecs[1].size, n_cycle_info_ratio=vecs[2].size,
                            reweighting=vecs[0].ctypes.data, thresholding=vecs[1].ctypes.data,
                            cycle_info_ratio=vecs[2].ctypes.data)
        R = np.empty((3, 3, n), dtype=np.float64, order="F")
        scores = np.zeros(max(max_iter, 1), dtype=np.float64)
        run = C.c_int32(0)
        _lib.check(self._lib.desc_b200_mpls_refine(self._h, _ptr(S), _ptr(R0), C.byref(p), _ptr(R), C.byref(run),
                                                  _ptr(scores)))
        return R, scores[:run.value].copy()

    def cycle_reweight(self, x, beta, empty_value=1.0):
        """One cycle reweighting of an edge vector (CEMP.m:109-125; HVec of MPLS.m:219-233)."""
        x = np.ascontiguousarray(np.asarray(x, dtype=np.float64).ravel())
        if x.size != self.m:
            raise ValueError("x must have m entries")
        out = np.empty(self.m, dtype=np.float64)
        _lib.check(self._lib.desc_b200_cycle_reweight(self._h, _ptr(x), float(beta), float(empty_value), _ptr(out)))
        return out

    def rotation_alignment(self, R_est, R_gt):
        """Utils/Rotation_Alignment.m:13-38 -> (R_out, R_align, mean_error, median_error)."""
        n = self.info()["n"]
        Re = np.asfortranarray(np.asarray(R_est, dtype=np.float64))
        Rg = np.asfortranarray(np.asarray(R_gt, dtype=np.float64))
        if Re.shape != (3, 3, n) or Rg.shape != (3, 3, n):
            raise ValueError("R_est and R_gt must be 3 x 3 x n")
        R_out = np.empty((3, 3, n), dtype=np.float64, order="F")
        R_align = np.empty((3, 3), dtype=np.float64, order="F")
        mean, med = C.c_double(0.0), C.c_double(0.0)
        _lib.check(self._lib.desc_b200_rotation_alignment(self._h, _ptr(Re), _ptr(Rg), _ptr(R_out), _ptr(R_align),
                                                         C.byref(mean), C.byref(med)))
        return R_out, R_align, float(mean.value), float(med.value)

    def pgd_diag(self, iters, rule, ErrVec, R_orig):
        """pgd with the make_plots branch (DESC.m:235-239) -> (S_vec, hist, iters_run, diag) where diag[t] =
        [mean(abs(ErrVec-S_vec)), MSE_mean, MSE_median] after iteration t+1."""
        iters = int(iters)
        n = self.info()["n"]
        Err = np.ascontiguousarray(np.asarray(ErrVec, dtype=np.float64).ravel())
        Ro = np.asfortranarray(np.asarray(R_orig, dtype=np.float64))
        if Err.size != self.m or Ro.shape != (3, 3, n):
            raise ValueError("ErrVec must have m entries and R_orig must be 3 x 3 x n")
        S = np.empty(self.m, dtype=np.float64)
        hist = np.zeros((max(iters, 1), 2), dtype=np.float64)
        diag = np.zeros((max(iters, 1), 3), dtype=np.float64)
        r = rule._to_c()
        run = C.c_int32(0)
        _lib.check(self._lib.desc_b200_pgd_diag(self._h, iters, C.byref(r), _ptr(Err), _ptr(Ro), _ptr(S), _ptr(hist),
                                               _ptr(diag), C.byref(run)))
        rule._from_c(r)
        k = int(run.value)
        return S, hist[:k], k, diag[:k]

    # -- getters -------------------------------------------------------------------------
    def info(self):
        a = (C.c_int64 * 10)()
        _lib.check(self._lib.desc_b200_get_info(self._h, a))
        keys = ["n", "m", "m_pos", "m_cycle", "n_sample", "max_slots_per_edge", "edge_begin", "edge_end",
                "local_slots", "max_codeg"]
        return dict(zip(keys, [int(v) for v in a]))

    def codeg(self):
        out = np.empty(self.m, dtype=np.int32)
        _lib.check(self._lib.desc_b200_get_codeg(self._h, _ptr(out)))
        return out

    def incidence(self):
        info = self.info()
        rowptr = np.empty(self.m + 1, dtype=np.int64)
        apex = np.empty(info["m_cycle"], dtype=np.int32)
        _lib.check(self._lib.desc_b200_get_incidence(self._h, _ptr(rowptr), _ptr(apex)))
        return rowptr, apex

    def slots(self):
        ns = self.info()["local_slots"]
        e_jk = np.empty(ns, dtype=np.int32)
        e_ki = np.empty(ns, dtype=np.int32)
        a = np.empty(ns, dtype=np.uint8)
        b = np.empty(ns, dtype=np.uint8)
        _lib.check(self._lib.desc_b200_get_slots(self._h, _ptr(e_jk), _ptr(e_ki), _ptr(a), _ptr(b)))
        return e_jk, e_ki, a.astype(bool), b.astype(bool)

    def S0(self):
        out = np.empty(self.info()["local_slots"], dtype=np.float64)
        _lib.check(self._lib.desc_b200_get_s0(self._h, _ptr(out)))
        return out

    def w(self):
        out = np.empty(self.info()["local_slots"], dtype=np.float64)
        _lib.check(self._lib.desc_b200_get_w(self._h, _ptr(out)))
        return out

    def gcw_info(self):
        a = (C.c_double * 8)()
        _lib.check(self._lib.desc_b200_get_gcw_info(self._h, a))
        return {"iters": int(a[0]), "residual": float(a[1]), "theta": [float(a[2]), float(a[3]), float(a[4])]}

    def timings(self):
        t = Timings()
        _lib.check(self._lib.desc_b200_get_timings(self._h, C.byref(t)))
        return t.as_dict()

    def sync(self):
        _lib.check(self._lib.desc_b200_sync(self._h))


def device_count():
    n = _lib.load().desc_b200_device_count()
    if n < 0:
        _lib.check(n)
    return n


def nccl_unique_id():
    buf = C.create_string_buffer(128)
    _lib.check(_lib.load().desc_b200_nccl_unique_id(buf))
    return buf.raw


# ---------------------------------------------------------------------------------------
# Replaying a MATLAB draw: the reference's cycle lists -> the (ptr, apex) lists ``cycles=`` takes
# ---------------------------------------------------------------------------------------
def cycles_from_desc(cum_ind, CoDeg_pos_ind, IJK, m):
    """DESC.m's ``cum_ind`` (m_pos+1), ``CoDeg_pos_ind`` (1-based ids of the edges with a 3-cycle, :36) and ``IJK``
    (1-based apex per slot, :93) -> ``(ptr over ALL m edges, 0-based apex)``."""
    cum_ind = np.asarray(cum_ind, dtype=np.int64).ravel()
    pos = np.asarray(CoDeg_pos_ind, dtype=np.int64).ravel() - 1
    cnt = np.zeros(int(m), dtype=np.int64)
    cnt[pos] = np.diff(cum_ind)
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    apex = (np.asarray(IJK, dtype=np.int64).ravel() - 1).astype(np.int32)
    if apex.size != int(ptr[-1]):
        raise ValueError("IJK must have cum_ind(end) entries")
    return ptr, apex


def cycles_from_cemp(CoIndMat):
    """CEMP.m's / MPLS.m's ``CoIndMat`` (nsample x m, 1-based apices; all-zero columns for edges without a 3-cycle,
    :63) -> ``(ptr over ALL m edges, 0-based apex)``.  Repeated apices (the draw is WITH replacement) are kept."""
    Co = np.asarray(CoIndMat, dtype=np.int64)
    if Co.ndim != 2:
        raise ValueError("CoIndMat must be nsample x m")
    has = (Co > 0).all(axis=0)
    if ((Co > 0).any(axis=0) != has).any():
        raise ValueError("a CoIndMat column must be all zero (no 3-cycle) or all positive")
    cnt = np.where(has, Co.shape[0], 0).astype(np.int64)
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    apex = (Co.T[has].ravel() - 1).astype(np.int32)
    return ptr, apex


# ---------------------------------------------------------------------------------------
# The reference's entry points
# ---------------------------------------------------------------------------------------
def _param(params, key, default=None):
    if isinstance(params, dict):
        return params.get(key, default)
    return getattr(params, key, default)


#: convergence curves of the last solve run with ``make_plots=True`` (DESC.m:174-178,235-239)
last_diagnostics = None


def _check_params(params):
    rule = _param(params, "Gradient")
    if rule is None or not hasattr(rule, "_to_c"):
        raise ValueError("params.Gradient must be a ConstantStepSize / PiecewiseStepSize / HybridGradient object")
    if _param(params, "make_plots", False) and (_param(params, "ErrVec") is None or _param(params, "R_orig") is None):
        raise ValueError("params.make_plots=true needs params.ErrVec and params.R_orig (DESC.m:236-238)")
    return rule


def _pgd_stage(s, params, rule):
    """DESC.m:164-261 on solver ``s``; with make_plots the diagnostics branch runs on the device as well."""
    global last_diagnostics
    iters = int(_param(params, "iters"))
    if not _param(params, "make_plots", False):
        return s.pgd(iters, rule)
    ErrVec, R_orig = _param(params, "ErrVec"), _param(params, "R_orig")
    S_vec, hist, iters_run, diag = s.pgd_diag(iters, rule, ErrVec, R_orig)
    last_diagnostics = dict(svec_errors=diag[:, 0].copy(), obj_vals=hist[:, 1].copy(), MSE_means=diag[:, 1].copy(),
                            MSE_medians=diag[:, 2].copy())
    _draw_plots(last_diagnostics)
    return S_vec, hist, iters_run


def _draw_plots(d):
    """DESC.m:315-344: the 2x2 figure.  Only if matplotlib is installed; the curves are in last_diagnostics."""
    try:
        import matplotlib.pyplot as plt
    except Exception:
        return
    fig, ax = plt.subplots(2, 2)
    for a, key, title, yl in ((ax[0, 0], "svec_errors", "Convergence of Corruption Estimate Vector (S_vec, sampled)",
                               "Average distance to true corruption"),
                              (ax[0, 1], "obj_vals", "Convergence of Objective Function (sampled)",
                               "Value of Objective Function"),
                              (ax[1, 0], "MSE_means", "Convergence of Rotation Estimate, Mean (sampled)",
                               "Mean Error in R estimate (degrees)"),
                              (ax[1, 1], "MSE_medians", "Convergence of Rotation Estimate, Median (sampled)",
                               "Median Error in R estimate (degrees)")):
        a.plot(d[key])
        a.set_title(title)
        a.set_xlabel("Iteration number")
        a.set_ylabel(yl)
    d["figure"] = fig


def _run_pgd(Ind, RijMat, params, want_gcw, **solver_kw):
    rule = _check_params(params)
    iters = int(_param(params, "iters"))
    s = Solver(Ind, RijMat, **solver_kw)
    try:
        s.build_incidence(n_sample=int(_param(params, "n_sample", 0) or 0), seed=int(_param(params, "seed", 0) or 0),
                          cycles=_param(params, "cycles"))
        s.cycle_inconsistency()
        if _param(params, "verbose", False):
            print("Initialization completed!")              # DESC.m:160
            print("Reweighting Procedure Started ...")      # DESC.m:162
        S_vec, hist, iters_run = _pgd_stage(s, params, rule)
        if _param(params, "verbose", False):
            for t in range(iters_run):                       # DESC.m:241
                print("iter %d: average change in S_vec %f, objective value: %f" % (t + 1, hist[t, 0], hist[t, 1]))
        R = s.gcw() if want_gcw else None
        extras = dict(hist=hist, iters_run=iters_run, info=s.info(), timings=s.timings())
    finally:
        s.close()
    return S_vec.reshape(1, -1), R, extras


def DESC_PGD(Ind, RijMat, params, **solver_kw):
    """``[S_vec] = DESC_PGD(Ind, RijMat, params)`` (Algorithms/DESC_PGD.m:14).  S_vec is 1 x m."""
    S_vec, _, _ = _run_pgd(Ind, RijMat, params, False, **solver_kw)
    return S_vec


def _dlmwrite_append(path, row):
    """``dlmwrite(path, row, 'delimiter', ',', '-append')``: one comma-separated line, 5 significant digits."""
    with open(path, "a") as f:
        f.write(",".join("%.5g" % float(v) for v in row) + "\n")


def DESC_init(Ind, RijMat, params, **solver_kw):
    """``[R_est, S_vec] = DESC_init(Ind, RijMat, params)`` (Algorithms/DESC_init.m:14).  With ``make_plots`` the
    reference also appends the convergence curves to two CSV files in the working directory
    (DESC_init.m:261-262); so does this."""
    S_vec, R, _ = _run_pgd(Ind, RijMat, params, True, **solver_kw)
    if _param(params, "make_plots", False) and last_diagnostics is not None:
        _dlmwrite_append("linear_convergence_rotation_error.csv", last_diagnostics["MSE_means"])
        _dlmwrite_append("linear_convergence_svec_error.csv", last_diagnostics["svec_errors"])
    return R, S_vec


def DESC(Ind, RijMat, params, **solver_kw):
    """``[R_est, R_init, S_vec] = DESC(Ind, RijMat, params)`` (Algorithms/DESC.m:14): the hot path
    (DESC.m:14-263) followed by the weighted Lie-algebraic refinement (DESC.m:265-312), all on the device."""
    rule = _check_params(params)
    verbose = _param(params, "verbose", False)
    s = Solver(Ind, RijMat, **solver_kw)
    try:
        s.build_incidence(n_sample=int(_param(params, "n_sample", 0) or 0), seed=int(_param(params, "seed", 0) or 0),
                          cycles=_param(params, "cycles"))
        s.cycle_inconsistency()
        S_vec, hist, iters_run = _pgd_stage(s, params, rule)
        R_init = s.gcw()
        if verbose:
            print("Rotation Initialized!")                   # DESC.m:283
            print("Start DESC refinement ...")
        R_est, scores = s.refine()
        if verbose:
            for t, sc in enumerate(scores):                  # DESC.m:305
                print("Iter %d: ||\u0394R||= %f" % (t + 1, sc))
            print("DONE!")                                   # DESC.m:313
    finally:
        s.close()
    return R_est, R_init, S_vec.reshape(1, -1)


def GCW(Ind, AdjMat, RijMat, S_vec, **solver_kw):
    """``R_est = GCW(Ind, AdjMat, RijMat, S_vec)`` (Utils/GCW.m:1).  ``AdjMat`` is only a mask in the
    reference (GCW.m:20) and is implied by ``Ind``; it is accepted and ignored."""
    s = Solver(Ind, RijMat, **solver_kw)
    try:
        return s.gcw(S_vec)
    finally:
        s.close()


# ---------------------------------------------------------------------------------------
# the comparators on the same incidence, and the evaluation metric
# ---------------------------------------------------------------------------------------
def _run_cemp(Ind, RijMat, CEMP_parameters, want_gcw, **solver_kw):
    nsample = int(_param(CEMP_parameters, "nsample"))
    s = Solver(Ind, RijMat, **solver_kw)
    try:
        s.build_incidence(n_sample=nsample, seed=int(_param(CEMP_parameters, "seed", 0) or 0),
                          cycles=_param(CEMP_parameters, "cycles"))
        s.cycle_inconsistency()
        SVec = s.cemp(int(_param(CEMP_parameters, "max_iter")), _param(CEMP_parameters, "reweighting"))
        R = s.cemp_gcw() if want_gcw else None
    finally:
        s.close()
    return SVec.reshape(1, -1), R


def CEMP(Ind, RijMat, CEMP_parameters, **solver_kw):
    """``SVec = CEMP(Ind, RijMat, CEMP_parameters)`` (Algorithms/CEMP.m:25), SVec is 1 x m.

    ``CEMP_parameters``: ``max_iter``, ``reweighting``, ``nsample`` as in the reference (``gcw_beta`` is unused there
    too); optional ``seed`` / ``cycles`` as for DESC.  The reference samples WITH replacement (CEMP.m:63); the
    device sampler keeps ``nsample`` distinct common neighbours (all of them when there are fewer) -- pass
    ``cycles=(ptr, apex)`` with repeated apices to replay a MATLAB draw exactly."""
    return _run_cemp(Ind, RijMat, CEMP_parameters, False, **solver_kw)[0]


def CEMP_GCW(Ind, RijMat, CEMP_parameters, **solver_kw):
    """``R_est = CEMP_GCW(Ind, RijMat, CEMP_parameters)`` (Algorithms/CEMP_GCW.m:25)."""
    return _run_cemp(Ind, RijMat, CEMP_parameters, True, **solver_kw)[1]


def MPLS(Ind, RijMat, CEMP_parameters, MPLS_parameters, **solver_kw):
    """``[R_est, R_init] = MPLS(Ind, RijMat, CEMP_parameters, MPLS_parameters)`` (Algorithms/MPLS.m:28): CEMP
    (:66-150), CEMP+MST initialisation (:152-195) and the MPLS reweighting loop (:198-256), all on the device."""
    verbose = _param(MPLS_parameters, "verbose", False)
    s = Solver(Ind, RijMat, **solver_kw)
    try:
        s.build_incidence(n_sample=int(_param(CEMP_parameters, "nsample")), seed=int(_param(CEMP_parameters, "seed", 0) or 0),
                          cycles=_param(CEMP_parameters, "cycles"))
        s.cycle_inconsistency()
        s.cemp(int(_param(CEMP_parameters, "max_iter")), _param(CEMP_parameters, "reweighting"))
        if verbose:
            print("Building minimum spanning tree ...")          # MPLS.m:153
        R_init = s.mst_init()
        if verbose:
            print("Rotation Initialized!")                       # MPLS.m:216
            print("Start MPLS reweighting ...")
        R_est, scores = s.mpls_refine(MPLS_parameters)
        if verbose:
            for t, sc in enumerate(scores):                      # MPLS.m:248
                print("Iter %d: ||\u0394R||= %f" % (t + 1, sc))
            print("DONE!")
    finally:
        s.close()
    return R_est, R_init


def Spectral(Ind, RijMat, **solver_kw):
    """``R_est = Spectral(Ind, RijMat)`` (Algorithms/Spectral.m:15)."""
    s = Solver(Ind, RijMat, **solver_kw)
    try:
        return s.spectral()
    finally:
        s.close()


def Rotation_Alignment(R_est, R_gt, **solver_kw):
    """``[R_out, R_align, mean_error, median_error] = Rotation_Alignment(R_est, R_gt)``
    (Utils/Rotation_Alignment.m:13).  Needs no graph: a two-node dummy handle carries the device context."""
    R_est = np.asarray(R_est, dtype=np.float64)
    n = R_est.shape[2]
    if n < 2:
        raise ValueError("need at least two rotations")
    Ind = np.stack([np.arange(1, n), np.arange(2, n + 1)], axis=1).astype(np.float64)   # a path: every node has an edge
    Rij = np.tile(np.eye(3)[:, :, None], (1, 1, n - 1))
    s = Solver(Ind, Rij, **solver_kw)
    try:
        return s.rotation_alignment(R_est, R_gt)
    finally:
        s.close()
