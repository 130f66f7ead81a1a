"""ctypes binding of ``libdesc_b200.so`` (C ABI declared in ``include/desc_b200.h``).

There is no CPU fallback: if the shared library is missing this module raises, and every
entry point fails with ``DESC_B200_ERR_CUDA`` when no CUDA device is visible.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DESC_B200_LIB: load another build of the library (kernel-tuning experiments)
LIB_PATH = os.environ.get("DESC_B200_LIB") or os.path.join(_HERE, "libdesc_b200.so")

OK, ERR_ARG, ERR_CUDA, ERR_STATE, ERR_LIMIT, ERR_NCCL, ERR_NOCONV = 0, -1, -2, -3, -4, -5, -6
INPUTS_ON_DEVICE = 1

ERROR_NAMES = {ERR_ARG: "DESC_B200_ERR_ARG", ERR_CUDA: "DESC_B200_ERR_CUDA", ERR_STATE: "DESC_B200_ERR_STATE",
               ERR_LIMIT: "DESC_B200_ERR_LIMIT", ERR_NCCL: "DESC_B200_ERR_NCCL", ERR_NOCONV: "DESC_B200_ERR_NOCONV"}

# every symbol include/desc_b200.h declares (tests check the library exports all of them)
SYMBOLS = [
    "desc_b200_last_error", "desc_b200_version", "desc_b200_device_count", "desc_b200_nccl_unique_id", "desc_b200_comm_finalize", "desc_b200_trim",
    "desc_b200_create", "desc_b200_destroy", "desc_b200_build_incidence", "desc_b200_cycle_inconsistency",
    "desc_b200_pgd", "desc_b200_gcw", "desc_b200_refine", "desc_b200_solve", "desc_b200_get_info", "desc_b200_get_codeg",
    "desc_b200_get_incidence", "desc_b200_get_slots", "desc_b200_get_s0", "desc_b200_get_w",
    "desc_b200_get_gcw_info", "desc_b200_get_timings", "desc_b200_sync",
    "desc_b200_cemp", "desc_b200_cemp_gcw", "desc_b200_cycle_reweight", "desc_b200_rotation_alignment",
    "desc_b200_pgd_diag", "desc_b200_mst_init", "desc_b200_mpls_refine", "desc_b200_spectral",
    "desc_b200_generate", "desc_b200_model_destroy", "desc_b200_model_info", "desc_b200_model_fetch",
    "desc_b200_model_device", "desc_b200_plan_shards",
]


class DescError(RuntimeError):
    """A C-ABI call returned a negative code; ``code`` holds it."""

    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERROR_NAMES.get(code, "DESC_B200_ERR"), code, msg))
        self.code = code


class Opts(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("stream", C.c_void_p), ("rank", C.c_int32),
                ("world", C.c_int32), ("nccl_id", C.c_void_p)]


class GenOpts(C.Structure):
    _fields_ = [("device", C.c_int32), ("topology", C.c_int32), ("n", C.c_int32), ("window", C.c_int32),
                ("kind", C.c_int32), ("reserved", C.c_int32), ("p", C.c_double), ("q", C.c_double),
                ("sigma", C.c_double), ("sigma_out", C.c_double), ("p_node_crpt", C.c_double),
                ("p_edge_crpt", C.c_double), ("seed", C.c_uint64)]


class MplsParams(C.Structure):
    _fields_ = [("stop_threshold", C.c_double), ("max_iter", C.c_int32), ("n_reweighting", C.c_int32),
                ("n_thresholding", C.c_int32), ("n_cycle_info_ratio", C.c_int32), ("reweighting", C.c_void_p),
                ("thresholding", C.c_void_p), ("cycle_info_ratio", C.c_void_p)]


class StepRule(C.Structure):
    _fields_ = [("kind", C.c_int32), ("strategy", C.c_int32), ("lr", C.c_double), ("decay_interval", C.c_double),
                ("beta_1", C.c_double), ("beta_2", C.c_double), ("t", C.c_int64)]


class Timings(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("graph_ms", C.c_double), ("build_ms", C.c_double), ("cycle_ms", C.c_double),
                ("pgd_ms", C.c_double), ("gcw_ms", C.c_double), ("d2h_ms", C.c_double), ("pgd_iter_ms", C.c_double),
                ("pgd_launches", C.c_int32), ("gcw_iters", C.c_int32), ("total_launches", C.c_int32),
                ("reserved", C.c_int32), ("pgd_pass1_ms", C.c_double), ("pgd_pass2_ms", C.c_double),
                ("pgd_comm_ms", C.c_double), ("laa_ms", C.c_double), ("laa_iters", C.c_int32), ("laa_cg_iters", C.c_int32),
                ("cemp_ms", C.c_double), ("cemp_iters", C.c_int32), ("reserved2", C.c_int32),
                ("mst_ms", C.c_double), ("gcw_spmv_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if not k.startswith("reserved")}


_lib = None


def load():
    """Load the library once.  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "desc_b200: %s not found -- build it with `make -C desc_b200/csrc` (or __graft_entry__.build()); "
            "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dp = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_void_p
    lib.desc_b200_last_error.restype = C.c_char_p
    lib.desc_b200_last_error.argtypes = []
    lib.desc_b200_version.restype = C.c_int
    lib.desc_b200_device_count.restype = C.c_int
    lib.desc_b200_nccl_unique_id.argtypes = [vp]
    lib.desc_b200_create.argtypes = [C.POINTER(vp), i32, i64, dp, dp, C.POINTER(Opts)]
    lib.desc_b200_destroy.argtypes = [vp]
    lib.desc_b200_destroy.restype = None
    lib.desc_b200_build_incidence.argtypes = [vp, i32, u64, vp, vp]
    lib.desc_b200_cycle_inconsistency.argtypes = [vp]
    lib.desc_b200_pgd.argtypes = [vp, i32, C.POINTER(StepRule), dp, dp, C.POINTER(i32)]
    lib.desc_b200_gcw.argtypes = [vp, dp, dp]
    lib.desc_b200_refine.argtypes = [vp, dp, dp, dp, C.POINTER(i32), dp]
    lib.desc_b200_solve.argtypes = [vp, i32, u64, i32, C.POINTER(StepRule), dp, dp, dp, C.POINTER(i32)]
    lib.desc_b200_get_info.argtypes = [vp, C.POINTER(i64)]
    lib.desc_b200_get_codeg.argtypes = [vp, vp]
    lib.desc_b200_get_incidence.argtypes = [vp, vp, vp]
    lib.desc_b200_get_slots.argtypes = [vp, vp, vp, vp, vp]
    lib.desc_b200_get_s0.argtypes = [vp, vp]
    lib.desc_b200_get_w.argtypes = [vp, vp]
    lib.desc_b200_get_gcw_info.argtypes = [vp, C.POINTER(C.c_double)]
    lib.desc_b200_get_timings.argtypes = [vp, C.POINTER(Timings)]
    lib.desc_b200_sync.argtypes = [vp]
    lib.desc_b200_generate.argtypes = [C.POINTER(GenOpts), C.POINTER(vp)]
    lib.desc_b200_model_destroy.argtypes = [vp]
    lib.desc_b200_model_destroy.restype = None
    lib.desc_b200_model_info.argtypes = [vp, C.POINTER(i64), C.POINTER(C.c_double)]
    lib.desc_b200_model_fetch.argtypes = [vp, dp, dp, dp, dp, dp, vp]
    lib.desc_b200_model_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    lib.desc_b200_spectral.argtypes = [vp, dp]
    lib.desc_b200_mst_init.argtypes = [vp, dp, dp]
    lib.desc_b200_mpls_refine.argtypes = [vp, dp, dp, C.POINTER(MplsParams), dp, C.POINTER(i32), dp]
    lib.desc_b200_cemp.argtypes = [vp, i32, dp, i32, dp]
    lib.desc_b200_cemp_gcw.argtypes = [vp, dp, dp]
    lib.desc_b200_cycle_reweight.argtypes = [vp, dp, C.c_double, C.c_double, dp]
    lib.desc_b200_rotation_alignment.argtypes = [vp, dp, dp, dp, dp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.desc_b200_pgd_diag.argtypes = [vp, i32, C.POINTER(StepRule), dp, dp, dp, dp, dp, C.POINTER(i32)]
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise DescError(rc, load().desc_b200_last_error().decode("utf-8", "replace"))
