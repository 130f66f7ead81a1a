"""World-size-2 tests of the multi-GPU path on CPU (gloo): rendezvous helpers, slot-balanced
sharding, the sharded scatter-form PGD and the sharded CEMP reweighting (the algorithms libdesc_b200 runs over
NCCL) against the single-process oracle."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from desc_b200 import dist as ddist
        from oracle import desc_oracle as O
        from oracle.desc_sharded import pgd_sharded, cemp_sharded

        ident = ddist.exchange_nccl_id(lambda: bytes(range(128)))
        assert ident == bytes(range(128))

        mo = O.uniform_topology(70, 0.5, 0.25, 0.05, rng=3)      # same seed on every rank
        inc = O.build_incidence(mo["Ind"], n_sample=12, seed=4)
        S0 = O.cycle_inconsistency(inc, mo["RijMat"])
        ns_all = np.zeros(inc.m, dtype=np.int64)
        ns_all[inc.pos_edges] = np.diff(inc.rowptr)
        rowptr_all = np.concatenate([[0], np.cumsum(ns_all)])
        bounds = ddist.shard_bounds(rowptr_all, world)

        def allreduce(x):
            t = torch.from_numpy(np.ascontiguousarray(x))
            dist.all_reduce(t)
            return t.numpy()

        def allgather(S, b):
            out = S.copy()
            for r in range(world):
                t = torch.from_numpy(np.ascontiguousarray(out[b[r]:b[r + 1]]))
                dist.broadcast(t, r)
                out[b[r]:b[r + 1]] = t.numpy()
            return out

        S, hist, k = pgd_sharded(inc, S0, 40, 0.02, rank, world, bounds, allreduce, allgather)
        S_ref, hist_ref, k_ref = O.pgd(inc, S0, 40, O.ConstantStepSize(0.02))
        ok = (k == k_ref and np.max(np.abs(S - S_ref)) <= 1e-12 and
              np.max(np.abs(hist[:, 1] - hist_ref[:, 1]) / np.abs(hist_ref[:, 1])) <= 1e-11)
        # CEMP on the same shards (csrc/cemp.cu: local reweighting + all-gather): bit-identical to the single-rank oracle
        beta = [1.0, 4.0, 16.0]
        ok = ok and np.array_equal(cemp_sharded(inc, S0, 5, beta, rank, world, bounds, allgather), O.cemp(inc, S0, 5, beta))
        slots = rowptr_all[bounds[rank + 1]] - rowptr_all[bounds[rank]]
        q.put((rank, bool(ok), int(slots), int(inc.m_cycle)))
    finally:
        dist.destroy_process_group()


def test_world2_sharded_pgd_matches_single_rank_oracle():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res
    tot = res[0][3]
    assert sum(r[2] for r in res) == tot
    assert max(r[2] for r in res) - min(r[2] for r in res) <= 2 * 12    # slot-balanced within one edge's list


def test_shard_bounds_properties():
    sys.path.insert(0, ROOT)
    from desc_b200.dist import shard_bounds
    rng = np.random.default_rng(0)
    cnt = rng.integers(0, 31, size=1000)
    cnt[rng.random(1000) < 0.1] = 0
    rowptr = np.concatenate([[0], np.cumsum(cnt)])
    for world in (1, 2, 3, 8):
        b = shard_bounds(rowptr, world)
        assert b[0] == 0 and b[-1] == 1000 and (np.diff(b) >= 0).all()
        per = np.diff(rowptr[b])
        assert per.sum() == rowptr[-1] and per.max() - per.min() <= 2 * 30
