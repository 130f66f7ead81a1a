"""ctypes binding of oracle/desc_pgd.c (TEST INFRASTRUCTURE: the C/OpenMP restatement of DESC.m:148-261).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libdesc_oracle_c.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
        _lib = C.CDLL(_LIB)
        p64, pd = C.POINTER(C.c_int64), C.POINTER(C.c_double)
        _lib.desc_c_pgd.restype = C.c_int
        _lib.desc_c_pgd.argtypes = [C.c_int64, C.c_int64, p64, p64, p64, p64, p64, p64, pd, C.c_int, C.c_int, C.c_double,
                                    C.c_double, p64, C.c_int, C.c_double, C.c_int, pd, pd, pd]
        _lib.desc_c_max_threads.restype = C.c_int
    return _lib


def max_threads():
    return int(load().desc_c_max_threads())


def pgd(inc, S0, iters, rule, patience=30, tol=1e-5, threads=0, return_w=False):
    """Same contract as desc_oracle.pgd for ConstantStepSize / PiecewiseStepSize rules."""
    lib = load()
    kind = type(rule).__name__
    if kind == "ConstantStepSize":
        rk, lr, dec, t = 0, float(rule.learning_rate), 1.0, 0
    elif kind == "PiecewiseStepSize":
        rk, lr, dec, t = 1, float(rule.learning_rate), float(rule.decay_interval), int(rule.t)
    else:
        raise ValueError("desc_pgd.c implements the constant and piecewise step rules")
    i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)   # noqa: E731
    pos, ptr, ejk, eki, ikj, jki = (i64(inc.pos_edges), i64(inc.rowptr), i64(inc.e_jk), i64(inc.e_ki), i64(inc.IKJ),
                                    i64(inc.JKI))
    S0 = np.ascontiguousarray(S0, dtype=np.float64)
    S_vec = np.empty(inc.m, dtype=np.float64)
    w = np.empty(max(inc.m_cycle, 1), dtype=np.float64)
    hist = np.zeros(2 * max(iters, 1), dtype=np.float64)
    t_io = C.c_int64(t)
    P = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))          # noqa: E731
    run = lib.desc_c_pgd(inc.m, inc.m_pos, P(pos, C.c_int64), P(ptr, C.c_int64), P(ejk, C.c_int64), P(eki, C.c_int64),
                         P(ikj, C.c_int64), P(jki, C.c_int64), P(S0, C.c_double), int(iters), rk, lr, dec, C.byref(t_io),
                         int(patience), float(tol), int(threads), P(S_vec, C.c_double), P(w, C.c_double),
                         P(hist, C.c_double))
    if rk == 1:
        rule.t = int(t_io.value)
    out = (S_vec, hist[:2 * run].reshape(-1, 2).copy(), int(run))
    return out + (w[:inc.m_cycle],) if return_w else out
