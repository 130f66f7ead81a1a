"""The multi-GPU shard plan (csrc/build.cu: slot-balanced start + cost-model descent) through its host-only entry
point: no GPU needed.  Pins the properties the sharded solve relies on and the direction of the cost model."""
import ctypes as C

import numpy as np
import pytest

from desc_b200 import _lib


def plan(cs, ca, world, slots_only=False):
    lib = _lib.load()
    n = cs.size - 1
    out = np.zeros(world + 1, dtype=np.int32)
    cs = np.ascontiguousarray(cs, dtype=np.int64)
    ca = np.ascontiguousarray(ca, dtype=np.int32)
    rc = lib.desc_b200_plan_shards(C.c_int32(n), C.c_int32(world), cs.ctypes.data_as(C.POINTER(C.c_int64)),
                                   ca.ctypes.data_as(C.POINTER(C.c_int32)), C.c_int32(1 if slots_only else 0),
                                   out.ctypes.data_as(C.POINTER(C.c_int32)))
    assert rc == 0
    return out


def er_profile(n, p, ns):
    """cumulative slots / adjacency entries of an Erdos-Renyi graph's vertex blocks: block v has (n-1-v) p edges"""
    own = (n - 1 - np.arange(n)) * p
    cs = np.concatenate([[0], np.cumsum(np.round(own * ns))]).astype(np.int64)
    ca = np.concatenate([[0], np.cumsum(np.full(n, int(round((n - 1) * p))))]).astype(np.int32)
    return cs, ca


def cost(cs, ca, vb):
    n = cs.size - 1
    m1 = max((cs[vb[r + 1]] - cs[vb[r]]) + 4.5 * (ca[vb[r + 1]] - ca[vb[r]]) for r in range(len(vb) - 1))
    m2 = max(1.2 * (cs[vb[r + 1]] - cs[vb[r]]) + 3.9 * (ca[n] - ca[vb[r]]) for r in range(len(vb) - 1))
    return m1 + m2


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_boundaries_partition_the_vertices(world):
    cs, ca = er_profile(2000, 0.1, 30)
    for so in (True, False):
        vb = plan(cs, ca, world, slots_only=so)
        assert vb[0] == 0 and vb[-1] == 2000 and np.all(np.diff(vb) >= 0)


def test_slot_balanced_plan_balances_slots():
    cs, ca = er_profile(10000, 0.1, 30)
    vb = plan(cs, ca, 8, slots_only=True)
    share = np.diff(cs[vb]) / cs[-1]
    assert np.max(np.abs(share - 1 / 8)) < 0.01
    # slots per vertex block fall linearly, so the ranges grow: n (1 - sqrt(1 - r/8))
    np.testing.assert_allclose(vb[1:-1], 10000 * (1 - np.sqrt(1 - np.arange(1, 8) / 8)), atol=12)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cost_model_never_worse_and_shifts_slots_to_the_middle(world):
    cs, ca = er_profile(10000, 0.1, 30)
    a, b = plan(cs, ca, world, slots_only=True), plan(cs, ca, world)
    assert cost(cs, ca, b) <= cost(cs, ca, a) * (1 + 1e-12)
    # the last rank owns the most vertex blocks (pass-1 tables): the model gives it fewer slots than 1/world
    assert (cs[b[-1]] - cs[b[-2]]) < (cs[a[-1]] - cs[a[-2]])


def test_degenerate_inputs():
    cs = np.zeros(11, dtype=np.int64)          # no triangles at all
    ca = np.arange(11, dtype=np.int32) * 3
    vb = plan(cs, ca, 4)
    assert vb[0] == 0 and vb[-1] == 10 and np.all(np.diff(vb) >= 0)
    lib = _lib.load()
    assert lib.desc_b200_plan_shards(C.c_int32(0), C.c_int32(2), None, None, C.c_int32(0), None) == _lib.ERR_ARG
