import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _all_golden():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def golden_names():
    """the DESC fixtures (DESC.m:14-312)"""
    return [n for n in _all_golden() if not n.startswith(("cemp_", "gen_"))]


def gen_golden_names():
    """outputs of the counter-based generators for fixed seeds"""
    return [n for n in _all_golden() if n.startswith("gen_")]


def cemp_golden_names():
    """the CEMP / CEMP_GCW fixtures (CEMP.m, CEMP_GCW.m)"""
    return [n for n in _all_golden() if n.startswith("cemp_")]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_rule(g, mod):
    """Step-rule object of fixture ``g`` built from module ``mod`` (oracle or desc_b200)."""
    kind = str(g["rule_kind"])
    r = g["rule"]
    if kind == "const":
        return mod.ConstantStepSize(r[0])
    if kind == "piecewise":
        return mod.PiecewiseStepSize(r[0], r[1])
    return mod.HybridGradient(r[0], r[1], r[2], r[3])


def golden_csr(g):
    """(ptr over ALL m edges, apex 0-based) of the fixture's cycle lists."""
    m = g["Ind"].shape[0]
    cnt = np.zeros(m, dtype=np.int64)
    cnt[g["CoDeg_pos_ind"] - 1] = np.diff(g["cum_ind"])
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return ptr, (g["IJK"] - 1).astype(np.int32)


@pytest.fixture(scope="session")
def have_gpu():
    try:
        import desc_b200
        return desc_b200.device_count() > 0
    except Exception:
        return False
