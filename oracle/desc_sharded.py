"""Sharded restatement of the PGD loop -- TEST INFRASTRUCTURE (CPU model of the multi-GPU path).

Same arithmetic as ``desc_oracle.pgd`` (DESC.m:148-261) but organised the way libdesc_b200 runs on
N GPUs: contiguous slot-balanced edge shards; the partner sums A_l, B_l (DESC.m:189-190)
in *scatter* form -- slot (ij;k) adds its weight to the "via shared vertex" accumulator of each
partner edge whose reciprocal slot exists -- into a local length-2m partial followed by an
all-reduce; S_vec shards all-gathered; objective / change scalars all-reduced.

``allreduce`` / ``allgather`` are injected so the same code runs in one process (world of
callables) or under torch.distributed with the gloo backend (tests/test_dist_cpu.py).
"""
from __future__ import annotations

import numpy as np


def shard_state(inc, S0, rank, bounds):
    """slots / edges owned by ``rank``; ``bounds`` are edge boundaries over ALL m edges"""
    ns_all = np.zeros(inc.m, dtype=np.int64)
    ns_all[inc.pos_edges] = np.diff(inc.rowptr)
    rowptr_all = np.concatenate([[0], np.cumsum(ns_all)])
    e0, e1 = int(bounds[rank]), int(bounds[rank + 1])
    s0, s1 = int(rowptr_all[e0]), int(rowptr_all[e1])
    sl = slice(s0, s1)
    i = inc.ei[inc.e_ij[sl]]
    j = inc.ej[inc.e_ij[sl]]
    k = inc.k[sl]
    return dict(e0=e0, e1=e1, sl=sl, own=inc.e_ij[sl], e_jk=inc.e_jk[sl], e_ki=inc.e_ki[sl],
                fa=inc.IKJ[sl] >= 0, fb=inc.JKI[sl] >= 0,
                # accumulator index: 2*edge + (0 if the shared vertex is that edge's smaller endpoint else 1)
                tgt_a=2 * inc.e_ki[sl] + np.where(i < k, 0, 1), tgt_b=2 * inc.e_jk[sl] + np.where(j < k, 0, 1),
                S0=S0[sl], ns=ns_all, rowptr_all=rowptr_all)


def _project(w, starts, counts):
    """Euclidean projection of each segment onto the simplex (Michelot active-set iteration, the
    algorithm of csrc/pgd.cu; same fixed point as the sort-and-scan of DESC.m:215-224)."""
    seg = np.repeat(np.arange(counts.size), counts)
    T = (np.add.reduceat(w, starts) - 1.0) / counts
    cnt = counts.copy()
    for _ in range(int(counts.max()) + 2):
        act = w > T[seg]
        s2 = np.add.reduceat(np.where(act, w, 0.0), starts)
        c2 = np.add.reduceat(act.astype(np.int64), starts)
        ch = (c2 != cnt) & (c2 > 0)
        if not ch.any():
            break
        T = np.where(ch, (s2 - 1.0) / np.maximum(c2, 1), T)
        cnt = np.where(ch, c2, cnt)
    return np.maximum(w - T[seg], 0.0)


def pgd_sharded(inc, S0, iters, lr, rank, world, bounds, allreduce, allgather, patience=30, tol=1e-5):
    """Returns (S_vec (m,), hist, iters_run) -- identical on every rank."""
    m = inc.m
    st = shard_state(inc, S0, rank, bounds)
    own = st["own"]
    local_edges = np.arange(st["e0"], st["e1"])
    has = st["ns"][local_edges] > 0
    pos_local = local_edges[has]
    counts = st["ns"][pos_local]
    starts = (st["rowptr_all"][pos_local] - st["rowptr_all"][st["e0"]]).astype(np.int64)
    seg = np.repeat(np.arange(counts.size), counts)

    def scatter(w):
        acc = np.zeros(2 * m)
        np.add.at(acc, st["tgt_a"][st["fa"]], w[st["fa"]])
        np.add.at(acc, st["tgt_b"][st["fb"]], w[st["fb"]])
        return allreduce(acc)

    def gather_S(S_local_pos, S_prev):
        S = S_prev.copy()
        S[pos_local] = S_local_pos
        # every rank contributes its own edge range
        return allgather(S, bounds)

    w = np.repeat(1.0 / np.maximum(counts, 1), counts) if counts.size else np.zeros(0)
    S = gather_S(np.add.reduceat(w * st["S0"], starts) if counts.size else np.zeros(0), np.ones(m))
    acc = scatter(w)
    hist, misses, iters_run = [], 0, 0
    for it in range(1, iters + 1):
        if counts.size:
            A = acc[2 * pos_local][seg]
            B = acc[2 * pos_local + 1][seg]
            grad = S[st["e_jk"]] + S[st["e_ki"]] + (np.where(st["fa"], A, 0.0) + np.where(st["fb"], B, 0.0)) * st["S0"]
            grad = grad - (np.add.reduceat(grad, starts) / counts)[seg]
            w = _project(w - lr * grad, starts, counts)
            S_new_local = np.add.reduceat(w * st["S0"], starts)
        else:
            S_new_local = np.zeros(0)
        S_new = gather_S(S_new_local, S)
        acc = scatter(w)
        red = allreduce(np.array([float(np.dot(w, S_new[st["e_jk"]] + S_new[st["e_ki"]])),
                                  float(np.abs(S_new[local_edges] - S[local_edges]).sum())]))
        hist.append((red[1] / m, red[0]))
        iters_run = it
        S = S_new
        if it > 1 and hist[-2][1] - hist[-1][1] < tol:
            misses += 1
            if misses >= patience:
                break
        else:
            misses = 0
    return S, np.array(hist).reshape(-1, 2), iters_run


def cemp_sharded(inc, S0, max_iter, reweighting, rank, world, bounds, allgather):
    """CEMP.m:98-129 the way libdesc_b200 runs it on N GPUs (csrc/cemp.cu): every rank reweights its own contiguous
    edge range from the full vector of the previous iteration, then the ranges are all-gathered.  Per-edge
    arithmetic does not depend on the shard, so the result is bit-identical to ``desc_oracle.cemp``."""
    from .desc_oracle import cemp_betas
    m = inc.m
    st = shard_state_light(inc, S0, rank, bounds)
    x = np.ones(m)

    def reweight(x_cur, beta, first):
        out = x_cur.copy()
        loc = np.ones(st["e1"] - st["e0"])
        if st["counts"].size:
            W = np.ones(st["S0"].size) if first else np.exp(-beta * (x_cur[st["e_ki"]] + x_cur[st["e_jk"]]))
            wsum = np.add.reduceat(W, st["starts"])
            loc[st["has"]] = np.add.reduceat(W / np.repeat(wsum, st["counts"]) * st["S0"], st["starts"]) if not first \
                else np.add.reduceat(st["S0"], st["starts"]) / st["counts"]
        out[st["e0"]:st["e1"]] = loc
        return allgather(out, bounds)

    x = reweight(x, 0.0, True)
    for beta in cemp_betas(max_iter, reweighting):
        x = reweight(x, beta, False)
    return x


def shard_state_light(inc, S0, rank, bounds):
    """slot range of ``rank`` for incidences without reciprocal maps (CEMP)"""
    ns_all = np.zeros(inc.m, dtype=np.int64)
    ns_all[inc.pos_edges] = np.diff(inc.rowptr)
    rowptr_all = np.concatenate([[0], np.cumsum(ns_all)])
    e0, e1 = int(bounds[rank]), int(bounds[rank + 1])
    sl = slice(int(rowptr_all[e0]), int(rowptr_all[e1]))
    local = np.arange(e0, e1)
    has = ns_all[local] > 0
    counts = ns_all[local][has]
    starts = (rowptr_all[local][has] - rowptr_all[e0]).astype(np.int64)
    return dict(e0=e0, e1=e1, e_jk=inc.e_jk[sl], e_ki=inc.e_ki[sl], S0=S0[sl], has=has, counts=counts, starts=starts)
