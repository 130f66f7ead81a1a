"""ctypes binding of the C/OpenMP restatement under oracle/ (TEST INFRASTRUCTURE):

* ``oracle/desc_full.c``  graph, co-degree, sampler, slot lists, reciprocal slots, d_ijk, GCW operator
  (Algorithms/DESC.m:19-147, Utils/GCW.m:13-27)
* ``oracle/desc_pgd.c``   the projected-gradient loop (Algorithms/DESC.m:148-261)

Same contracts as ``oracle/desc_oracle.py`` (it returns that module's ``Incidence``), all stages threaded, int32 indices:
this is what the full-size parity tests compare the CUDA path with and what ``bench.py`` times as the CPU arm.
Pinned against the numpy oracle in ``tests/test_oracle_c.py``.  Only tests/, ``__graft_entry__.smoke()`` and
bench.py's CPU legs may import this.
"""
import ctypes as C
import math
import os
import subprocess
import time

import numpy as np

from . import desc_oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libdesc_oracle_c.so")
_lib = None

p64, p32, pd, pu64 = C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_uint64)


def load():
    global _lib
    if _lib is None:
        srcs = [os.path.join(_HERE, f) for f in ("desc_pgd.c", "desc_full.c", "Makefile")]
        if not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs if os.path.exists(s)):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_LIB)
        _lib.desc_c_pgd.restype = C.c_int
        _lib.desc_c_pgd.argtypes = [C.c_int64, C.c_int64, p64, p64, p32, p32, p64, p64, pd, C.c_int, C.c_int, C.c_double,
                                    C.c_double, p64, C.c_int, C.c_double, C.c_int, pd, pd, pd]
        _lib.desc_c_max_threads.restype = C.c_int
        _lib.desc_c_graph.restype = C.c_int
        _lib.desc_c_graph.argtypes = [C.c_int64, C.c_int64, p32, p32, p64, p32, p32, pu64, C.c_int64]
        _lib.desc_c_codeg.restype = None
        _lib.desc_c_codeg.argtypes = [C.c_int64, p32, p32, pu64, C.c_int64, p32, C.c_int]
        _lib.desc_c_fill.restype = None
        _lib.desc_c_fill.argtypes = [C.c_int64, C.c_int64, p32, p32, pu64, C.c_int64, p64, p32, p32, p32, p64, C.c_int,
                                     C.c_uint64, p32, p32, p32, C.c_int]
        _lib.desc_c_recip.restype = None
        _lib.desc_c_recip.argtypes = [C.c_int64, p32, p32, p64, p32, p32, p32, C.c_int, p64, p64, C.c_int]
        _lib.desc_c_cycle.restype = None
        _lib.desc_c_cycle.argtypes = [C.c_int64, p32, p32, p64, p32, p32, p32, pd, pd, C.c_int]
        _lib.desc_c_gcw_weights.restype = None
        _lib.desc_c_gcw_weights.argtypes = [C.c_int64, C.c_int64, p32, p32, pd, C.c_int, pd, pd]
        _lib.desc_c_gcw_matvec.restype = None
        _lib.desc_c_gcw_matvec.argtypes = [C.c_int64, p64, p32, p32, pd, pd, pd, pd, C.c_int, C.c_int]
    return _lib


def host_threads():
    """threads the CPU arm uses: the cores this process may run on (torchrun exports OMP_NUM_THREADS=1 to its
    workers; the CPU baseline must not inherit that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def max_threads():
    return host_threads()


def _P(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


class Graph:
    """A1 (DESC.m:19-24): int32 endpoints, symmetric CSR adjacency with edge ids, adjacency bitmap."""

    def __init__(self, Ind, n=None):
        lib = load()
        Ind = np.asarray(Ind)
        self.m = int(Ind.shape[0])
        self.ei = np.ascontiguousarray(Ind[:, 0], dtype=np.float64).astype(np.int32) - 1
        self.ej = np.ascontiguousarray(Ind[:, 1], dtype=np.float64).astype(np.int32) - 1
        self.n = int(self.ej.max()) + 1 if n is None else int(n)
        key = self.ei.astype(np.int64) * self.n + self.ej
        if (self.ei < 0).any() or (self.ei >= self.ej).any() or (np.diff(key) <= 0).any():
            raise ValueError("Ind rows must satisfy 1 <= i < j and be strictly sorted by (i, j)")
        self.nw64 = (self.n + 63) // 64
        self.rowstart = np.zeros(self.n + 1, dtype=np.int64)
        self.nbr = np.zeros(2 * self.m, dtype=np.int32)
        self.eid = np.zeros(2 * self.m, dtype=np.int32)
        self.bm = np.zeros(self.n * self.nw64, dtype=np.uint64)
        rc = lib.desc_c_graph(self.n, self.m, _P(self.ei, C.c_int32), _P(self.ej, C.c_int32), _P(self.rowstart, C.c_int64),
                              _P(self.nbr, C.c_int32), _P(self.eid, C.c_int32), _P(self.bm, C.c_uint64), self.nw64)
        if rc != 0:
            raise ValueError("bad edge list")


def build_incidence(Ind, n_sample=None, seed=0, cycles=None, threads=0, graph=None, timings=None):
    """DESC.m:19-127; same contract and result type as desc_oracle.build_incidence (index arrays are int32 / int64)."""
    lib = load()
    threads = threads or host_threads()
    t0 = time.perf_counter()
    g = graph or Graph(Ind)
    m, n = g.m, g.n
    codeg = np.zeros(m, dtype=np.int32)
    lib.desc_c_codeg(m, _P(g.ei, C.c_int32), _P(g.ej, C.c_int32), _P(g.bm, C.c_uint64), g.nw64, _P(codeg, C.c_int32), threads)
    pos_mask = codeg > 0
    pos_edges = np.nonzero(pos_mask)[0].astype(np.int64)
    if cycles is None:
        if n_sample is None:
            n_sample = max(int(math.ceil(O.matlab_median(codeg[pos_mask]) / 4.0)), 30) if pos_edges.size else 30
        elif n_sample < 0:
            n_sample = int(codeg.max()) + 1 if m else 1
        n_sample = int(n_sample)
        ns_all = np.minimum(codeg, n_sample).astype(np.int64)
    else:
        ptr, apex_in = cycles
        ns_all = np.diff(np.asarray(ptr, dtype=np.int64))
        n_sample = int(n_sample) if n_sample else int(ns_all.max())
    rowptr_all = np.concatenate([[0], np.cumsum(ns_all)]).astype(np.int64)
    m_cycle = int(rowptr_all[-1])
    apex = np.zeros(max(m_cycle, 1), dtype=np.int32)
    e_jk = np.zeros(max(m_cycle, 1), dtype=np.int32)
    e_ki = np.zeros(max(m_cycle, 1), dtype=np.int32)
    if cycles is None:
        lib.desc_c_fill(n, m, _P(g.ei, C.c_int32), _P(g.ej, C.c_int32), _P(g.bm, C.c_uint64), g.nw64,
                        _P(g.rowstart, C.c_int64), _P(g.nbr, C.c_int32), _P(g.eid, C.c_int32), _P(codeg, C.c_int32),
                        _P(rowptr_all, C.c_int64), n_sample, C.c_uint64(int(seed) & (2 ** 64 - 1)), _P(apex, C.c_int32),
                        _P(e_jk, C.c_int32), _P(e_ki, C.c_int32), threads)
        srt = 1
    else:
        raise NotImplementedError("explicit cycle lists: use the numpy oracle (small cases only)")
    IKJ = np.zeros(max(m_cycle, 1), dtype=np.int64)
    JKI = np.zeros(max(m_cycle, 1), dtype=np.int64)
    lib.desc_c_recip(m, _P(g.ei, C.c_int32), _P(g.ej, C.c_int32), _P(rowptr_all, C.c_int64), _P(apex, C.c_int32),
                     _P(e_jk, C.c_int32), _P(e_ki, C.c_int32), srt, _P(IKJ, C.c_int64), _P(JKI, C.c_int64), threads)
    rowptr = np.concatenate([[0], np.cumsum(ns_all[pos_edges])]).astype(np.int64)
    e_ij = np.repeat(np.arange(m, dtype=np.int32), ns_all)
    inc = O.Incidence(n=n, m=m, ei=g.ei, ej=g.ej, codeg=codeg, n_sample=n_sample, pos_edges=pos_edges, rowptr=rowptr,
                      e_ij=e_ij, e_jk=e_jk[:m_cycle], e_ki=e_ki[:m_cycle], k=apex[:m_cycle], IKJ=IKJ[:m_cycle],
                      JKI=JKI[:m_cycle], extras=dict(rowptr_all=rowptr_all, graph=g))
    if timings is not None:
        timings["build_s"] = time.perf_counter() - t0
    return inc


def cycle_inconsistency(inc, RijMat, threads=0, timings=None):
    """DESC.m:129-147 (unfused, reference order).  RijMat: MATLAB 3x3xm."""
    lib = load()
    threads = threads or host_threads()
    t0 = time.perf_counter()
    R = np.asarray(RijMat, dtype=np.float64)
    Rf = R.reshape(-1, order="F") if R.ndim == 3 else R.ravel()
    Rf = np.ascontiguousarray(Rf)
    S0 = np.zeros(max(inc.m_cycle, 1), dtype=np.float64)
    rp = np.ascontiguousarray(inc.extras["rowptr_all"], dtype=np.int64)
    lib.desc_c_cycle(inc.m, _P(inc.ei, C.c_int32), _P(inc.ej, C.c_int32), _P(rp, C.c_int64),
                     _P(np.ascontiguousarray(inc.k, dtype=np.int32), C.c_int32),
                     _P(np.ascontiguousarray(inc.e_jk, dtype=np.int32), C.c_int32),
                     _P(np.ascontiguousarray(inc.e_ki, dtype=np.int32), C.c_int32), _P(Rf, C.c_double), _P(S0, C.c_double),
                     threads)
    if timings is not None:
        timings["cycle_s"] = time.perf_counter() - t0
    return S0[:inc.m_cycle]


def pgd(inc, S0, iters, rule, patience=30, tol=1e-5, threads=0, return_w=False, timings=None):
    """Same contract as desc_oracle.pgd for ConstantStepSize / PiecewiseStepSize rules (DESC.m:148-261)."""
    lib = load()
    threads = threads or host_threads()
    t0 = time.perf_counter()
    kind = type(rule).__name__
    if kind == "ConstantStepSize":
        rk, lr, dec, t = 0, float(rule.learning_rate), 1.0, 0
    elif kind == "PiecewiseStepSize":
        rk, lr, dec, t = 1, float(rule.learning_rate), float(rule.decay_interval), int(rule.t)
    else:
        raise ValueError("desc_pgd.c implements the constant and piecewise step rules")
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)   # noqa: E731
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)   # noqa: E731
    pos, ptr, ejk, eki, ikj, jki = (i64(inc.pos_edges), i64(inc.rowptr), i32(inc.e_jk), i32(inc.e_ki), i64(inc.IKJ),
                                    i64(inc.JKI))
    S0 = np.ascontiguousarray(S0, dtype=np.float64)
    S_vec = np.empty(inc.m, dtype=np.float64)
    w = np.empty(max(inc.m_cycle, 1), dtype=np.float64)
    hist = np.zeros(2 * max(iters, 1), dtype=np.float64)
    t_io = C.c_int64(t)
    run = lib.desc_c_pgd(inc.m, inc.m_pos, _P(pos, C.c_int64), _P(ptr, C.c_int64), _P(ejk, C.c_int32), _P(eki, C.c_int32),
                         _P(ikj, C.c_int64), _P(jki, C.c_int64), _P(S0, C.c_double), int(iters), rk, lr, dec, C.byref(t_io),
                         int(patience), float(tol), int(threads), _P(S_vec, C.c_double), _P(w, C.c_double),
                         _P(hist, C.c_double))
    if rk == 1:
        rule.t = int(t_io.value)
    if timings is not None:
        timings["pgd_s"] = time.perf_counter() - t0
    out = (S_vec, hist[:2 * run].reshape(-1, 2).copy(), int(run))
    return out + (w[:inc.m_cycle],) if return_w else out


def gcw(Ind, RijMat, S_vec, power=1.5, threads=0, graph=None, timings=None):
    """Utils/GCW.m:1-38 with the block operator applied by desc_c_gcw_matvec (all host threads) inside ARPACK
    (scipy ``eigsh`` -- the reference calls ``eigs``, GCW.m:27), then GCW.m:28-36 as in the numpy oracle."""
    from scipy.sparse.linalg import LinearOperator, eigsh
    lib = load()
    threads = threads or host_threads()
    t0 = time.perf_counter()
    g = graph or Graph(Ind)
    n, m = g.n, g.m
    R = np.asarray(RijMat, dtype=np.float64)
    Rf = np.ascontiguousarray(R.reshape(-1, order="F") if R.ndim == 3 else R.ravel())
    S = np.ascontiguousarray(np.asarray(S_vec, dtype=np.float64).ravel())
    rule = {1.5: 0, 1: 1, None: 2}[power]
    coef = np.zeros(m)
    isd = np.zeros(n)
    lib.desc_c_gcw_weights(n, m, _P(g.ei, C.c_int32), _P(g.ej, C.c_int32), _P(S, C.c_double), rule, _P(coef, C.c_double),
                           _P(isd, C.c_double))

    def matmat(X):
        X = np.asarray(X, dtype=np.float64)
        one = X.ndim == 1
        Xf = np.asfortranarray(X.reshape(3 * n, -1))
        Y = np.zeros_like(Xf, order="F")
        lib.desc_c_gcw_matvec(n, _P(g.rowstart, C.c_int64), _P(g.nbr, C.c_int32), _P(g.eid, C.c_int32), _P(Rf, C.c_double),
                              _P(coef, C.c_double), _P(Xf, C.c_double), _P(Y, C.c_double), Xf.shape[1], threads)
        return Y[:, 0] if one else Y

    if 3 * n <= 1200:
        N = matmat(np.eye(3 * n))
        lam, U = np.linalg.eigh((N + N.T) / 2.0)
        U = U[:, ::-1][:, :3]
    else:
        op = LinearOperator((3 * n, 3 * n), matvec=matmat, matmat=matmat, dtype=np.float64)
        lam, U = eigsh(op, k=3, which="LA", tol=1e-14, v0=np.random.default_rng(0).standard_normal(3 * n))
        U = U[:, np.argsort(-lam)]
    V = U * np.repeat(isd, 3)[:, None]
    V = V / np.linalg.norm(V, axis=0, keepdims=True)
    if np.linalg.det(V[:3, :]) < 0:                              # GCW.m:28
        V[:, 0] = -V[:, 0]
    out = O.to_matlab(O.proj_so3(V.reshape(n, 3, 3)))            # GCW.m:30-36
    if timings is not None:
        timings["gcw_s"] = time.perf_counter() - t0
    return out


def DESC_init(Ind, RijMat, params, n_sample=None, seed=0, threads=0, full=False, want_R=True):
    """Algorithms/DESC_init.m:14 on the C port (``want_R=False``: DESC_PGD.m:14, no GCW); ``full`` returns every
    intermediate and the per-stage seconds."""
    tm = {}
    t0 = time.perf_counter()
    g = Graph(Ind)
    tm["graph_s"] = time.perf_counter() - t0
    inc = build_incidence(Ind, n_sample=n_sample, seed=seed, threads=threads, graph=g, timings=tm)
    S0 = cycle_inconsistency(inc, RijMat, threads=threads, timings=tm)
    S_vec, hist, iters_run, w = pgd(inc, S0, int(params["iters"]), params["Gradient"], threads=threads, return_w=True,
                                    timings=tm)
    R = None
    tm["gcw_s"] = 0.0
    if want_R:
        R = gcw(Ind, RijMat, S_vec, threads=threads, graph=g, timings=tm)
    if full:
        return dict(R=R, S_vec=S_vec, hist=hist, iters_run=iters_run, w=w, S0=S0, inc=inc, timings=tm)
    return R, S_vec
