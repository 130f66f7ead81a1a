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


def test_sharded_pgd_model_on_cost_model_shards_equals_single_rank():
    """the N-rank algorithm (oracle/desc_sharded.py, the CPU model of what libdesc_b200 runs on N GPUs) on the UNEQUAL,
    vertex-aligned shards the cost model produces: same S_vec / history / stop iteration as one rank.  Three ranks as
    threads of this process; collectives = barriers over shared buffers."""
    import threading
    from oracle import desc_oracle as O
    from oracle.desc_sharded import pgd_sharded

    W = 3
    mo = O.uniform_topology(90, 0.45, 0.25, 0.05, rng=5)
    inc = O.build_incidence(mo["Ind"], n_sample=10, seed=2)
    S0 = O.cycle_inconsistency(inc, mo["RijMat"])
    n, m = inc.n, inc.m
    ns_all = np.zeros(m, dtype=np.int64)
    ns_all[inc.pos_edges] = np.diff(inc.rowptr)
    rowptr_all = np.concatenate([[0], np.cumsum(ns_all)])
    estart = np.searchsorted(inc.ei, np.arange(n + 1))                  # first edge of every vertex block
    deg = np.bincount(inc.ei, minlength=n) + np.bincount(inc.ej, minlength=n)
    vb = plan(rowptr_all[estart], np.concatenate([[0], np.cumsum(deg)]), W)
    bounds = estart[vb]                                                 # vertex-aligned edge boundaries
    assert bounds[0] == 0 and bounds[-1] == m and len(set(np.diff(rowptr_all[bounds]))) > 1   # really unequal

    bar = threading.Barrier(W)
    box = {"red": [None] * W, "S": None}

    def make(rank):
        def allreduce(x):
            box["red"][rank] = np.array(x, dtype=np.float64, copy=True)
            bar.wait()
            out = sum(box["red"][r] for r in range(W))                  # same order on every rank
            bar.wait()
            return out

        def allgather(S, b):
            if rank == 0:
                box["S"] = np.array(S, copy=True)
            bar.wait()
            box["S"][b[rank]:b[rank + 1]] = S[b[rank]:b[rank + 1]]
            bar.wait()
            out = box["S"].copy()
            bar.wait()
            return out
        return allreduce, allgather

    res = [None] * W

    def run(rank):
        ar, ag = make(rank)
        res[rank] = pgd_sharded(inc, S0, 30, 0.02, rank, W, bounds, ar, ag)

    th = [threading.Thread(target=run, args=(r,)) for r in range(W)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    S_ref, hist_ref, k_ref = O.pgd(inc, S0, 30, O.ConstantStepSize(0.02))
    for r in range(W):
        S, hist, k = res[r]
        assert k == k_ref
        assert np.max(np.abs(S - S_ref)) <= 1e-12
        assert np.max(np.abs(hist[:, 1] - hist_ref[:, 1]) / np.abs(hist_ref[:, 1])) <= 1e-11
