"""desc_b200 -- B200-native DESC solver hot path (ColeWyeth/DESC, Algorithms/DESC.m:14-263 + Utils/GCW.m).

Hand-written sm_100a CUDA behind a C ABI (``include/desc_b200.h``, ``libdesc_b200.so``); this
package is the thin host-side mirror of the reference's MATLAB interface.  No CPU fallback.
"""
from ._lib import DescError, LIB_PATH  # noqa: F401
from .solver import (ConstantStepSize, PiecewiseStepSize, HybridGradient, Solver, DESC, DESC_PGD,  # noqa: F401
                     DESC_init, GCW, CEMP, CEMP_GCW, MPLS, Spectral, Rotation_Alignment, cycles_from_desc, cycles_from_cemp, device_count, nccl_unique_id)

from .matio import load_mat, save_mat  # noqa: F401
from .models import Uniform_Topology, Nonuniform_Topology, Ring_Topology, Model  # noqa: F401

__all__ = ["load_mat", "save_mat", "Uniform_Topology", "Nonuniform_Topology", "Ring_Topology", "Model", "ConstantStepSize", "PiecewiseStepSize", "HybridGradient", "Solver", "DESC", "DESC_PGD", "DESC_init",
           "GCW", "CEMP", "CEMP_GCW", "MPLS", "Spectral", "Rotation_Alignment", "cycles_from_desc", "cycles_from_cemp", "device_count", "nccl_unique_id", "DescError", "LIB_PATH"]
