"""Counter-based restatement of the reference's data generators  --  TEST INFRASTRUCTURE ONLY.

SURVEY 8(f) #2.  ``Models/Uniform_Topology.m:24-111`` and ``Models/Nonuniform_Topology.m:26-157`` draw from
MATLAB's global RNG stream (``rand``, ``randn``, ``randperm``), which cannot be restated; PARITY UNPINNED in
that sense (see ``oracle/desc_oracle.py``).  What is restated here is every deterministic statement of the two
files, with each random draw replaced by a *counter-based* draw -- a pure function of (seed, stream, item,
index) built on the sampler's 64-bit mix (``desc_oracle.sampler_keys`` == ``desc_key`` in
``csrc/internal.cuh``) -- so that the CUDA generators (``csrc/gen.cu``) can be checked value for value:

    u64(seed, stream, a, b)  = desc_key(seed + stream * 0xA0761D6478BD642F, a, b)
    uniform(...)             = (u64 >> 11) * 2^-53                       in [0, 1)
    normal(seed, s, a, q)    = Box-Muller on the pair (u64(.., a, 2t), u64(.., a, 2t+1)), t = q >> 1:
                               r = sqrt(-2 log(((z1 >> 11) + 1) 2^-53)), angle = 2 pi (z2 >> 11) 2^-53,
                               q even -> r cos(angle), q odd -> r sin(angle)
    randperm(k)              = order of the k items by (u64 key, index)

Also the SfM-shaped ring generator that BASELINE.json's configs[4] needs and the reference lacks
(``ring=window``: an edge {i,j} is possible only if the circular distance of i and j is <= window).
"""
from __future__ import annotations

import math

import numpy as np

from .desc_oracle import _err_vec, proj_so3, sampler_keys, to_matlab

S_ADJ, S_RORIG, S_MASK, S_NOISE, S_RCORR, S_CORR, S_NODEPERM, S_NBRPERM, S_R0, S_NOISE_OUT = range(1, 11)
_STREAM_MUL = 0xA0761D6478BD642F
_MASK64 = (1 << 64) - 1


def u64(seed, stream, a, b):
    return sampler_keys((int(seed) + stream * _STREAM_MUL) & _MASK64, a, b)


def uniform(seed, stream, a, b):
    return (u64(seed, stream, a, b) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def normal(seed, stream, a, q):
    a, q = np.broadcast_arrays(np.asarray(a, dtype=np.int64), np.asarray(q, dtype=np.int64))
    t = q >> 1
    z1 = u64(seed, stream, a, 2 * t)
    z2 = u64(seed, stream, a, 2 * t + 1)
    u1 = ((z1 >> np.uint64(11)).astype(np.float64) + 1.0) * 2.0 ** -53
    u2 = (z2 >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    r = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    return np.where((q & 1) == 0, r * np.cos(ang), r * np.sin(ang))


def randn3(seed, stream, items):
    """``randn(3)`` per item as (k,3,3) with [item, r, c] = the (r+3c)-th normal (MATLAB fills column-major)."""
    items = np.asarray(items, dtype=np.int64)
    q = (np.arange(3)[:, None] + 3 * np.arange(3)[None, :])[None, :, :]
    return normal(seed, stream, items[:, None, None], q)


def rand_rot(seed, stream, items):
    """``Q=randn(3); [U,~,V]=svd(Q); U*diag([1,1,det(U*V')])*V'`` (Uniform_Topology.m:40-45)."""
    return proj_so3(randn3(seed, stream, items))


def graph(n, p, seed, ring=0):
    """``G = tril(rand(n,n)<p,-1); [Ind_j,Ind_i]=find(G)`` (Uniform_Topology.m:29-34): pairs lo<hi sorted by
    (lo, hi); the draw of pair {lo,hi} is uniform(S_ADJ, lo, hi)."""
    lo, hi = np.triu_indices(n, 1)
    if ring:
        d = hi - lo
        keep = np.minimum(d, n - d) <= ring
        lo, hi = lo[keep], hi[keep]
    sel = uniform(seed, S_ADJ, lo, hi) < p
    return lo[sel].astype(np.int64), hi[sel].astype(np.int64)


def uniform_topology(n, p, q, sigma, model="uniform", seed=0, ring=0):
    """Models/Uniform_Topology.m:24-111 on counter-based draws."""
    ei, ej = graph(n, p, seed, ring)
    m = ei.size
    e = np.arange(m, dtype=np.int64)
    R_orig = rand_rot(seed, S_RORIG, np.arange(n))                                   # :37-45
    Rij_orig = R_orig[ei] @ R_orig[ej].transpose(0, 2, 1)                            # :47-51
    RijMat = Rij_orig.copy()
    noise = uniform(seed, S_MASK, e, 0) >= q                                         # :53
    corr = ~noise
    RijMat[noise] = proj_so3(RijMat[noise] + sigma * randn3(seed, S_NOISE, e[noise]))   # :58-65
    R_corr = rand_rot(seed, S_RCORR, np.arange(n))                                   # :67-74
    if model == "uniform":
        RijMat[corr] = rand_rot(seed, S_CORR, e[corr])                               # :77-82
    else:
        Q = R_corr[ei[corr]] @ R_corr[ej[corr]].transpose(0, 2, 1) + sigma * randn3(seed, S_CORR, e[corr])   # :84-90
        RijMat[corr] = proj_so3(Q)
    ErrVec = _err_vec(Rij_orig, RijMat)                                              # :94-101
    Ind = np.stack([ei + 1, ej + 1], axis=1).astype(np.float64)
    return dict(Ind=Ind, RijMat=to_matlab(RijMat), Rij_orig=to_matlab(Rij_orig), R_orig=to_matlab(R_orig),
                ErrVec=ErrVec, corrupted=corr)


def nonuniform_topology(n, p, p_node_crpt, p_edge_crpt, sigma_in, sigma_out, crpt_type="uniform", seed=0):
    """Models/Nonuniform_Topology.m:26-157 on counter-based draws.

    ``randperm`` orders are key orders, so the sequential loops :80-124 become per-edge rules: node i is
    corrupted iff it is among the floor(n p_node_crpt) nodes of smallest key; it corrupts the
    floor(p_edge_crpt deg_i) neighbours of smallest key (i, j); an edge picked from both ends is written last by
    the endpoint that comes later in the node order (the loop overwrites), which is what :93-121 leaves behind."""
    ei, ej = graph(n, p, seed)
    m = ei.size
    e = np.arange(m, dtype=np.int64)
    R_orig = rand_rot(seed, S_RORIG, np.arange(n))
    Rij_orig = R_orig[ei] @ R_orig[ej].transpose(0, 2, 1)
    RijMat = Rij_orig.copy()
    nodes = np.arange(n, dtype=np.int64)
    nkey = u64(seed, S_NODEPERM, nodes, 0)
    order = np.lexsort((nodes, nkey))                                                # randperm(n), :62
    n_node = int(math.floor(n * p_node_crpt))                                        # :63
    pos = np.full(n, -1, dtype=np.int64)                                             # position in node_crpt, -1 = clean
    pos[order[:n_node]] = np.arange(n_node)
    R_crpt = rand_rot(seed, S_RCORR, nodes)                                          # :68-73
    # per node: neighbours by (key, j); the first floor(p_edge_crpt*deg) are corrupted (:81-86)
    full_a = np.concatenate([ei, ej])
    full_b = np.concatenate([ej, ei])
    full_e = np.concatenate([e, e])
    bkey = u64(seed, S_NBRPERM, full_a, full_b)
    o = np.lexsort((full_b, bkey, full_a))
    fa, fe = full_a[o], full_e[o]
    deg = np.bincount(full_a, minlength=n)
    start = np.concatenate([[0], np.cumsum(deg)])[fa]
    rank = np.arange(fa.size) - start
    nn = np.floor(p_edge_crpt * deg.astype(np.float64)).astype(np.int64)
    picked = (pos[fa] >= 0) & (rank < nn[fa])
    # winner per edge: the picking endpoint that is later in node_crpt
    win = np.full(m, -1, dtype=np.int64)
    wpos = np.full(m, -1, dtype=np.int64)
    for a_, e_ in zip(fa[picked], fe[picked]):
        if pos[a_] > wpos[e_]:
            wpos[e_] = pos[a_]
            win[e_] = a_
    crpt = win >= 0
    ce = e[crpt]
    w = win[crpt]
    other = np.where(w == ei[crpt], ej[crpt], ei[crpt])
    if crpt_type == "uniform":
        M = rand_rot(seed, S_R0, 2 * ce + (w == ej[crpt]))                           # :91-94 (a draw per visit)
    elif crpt_type == "self-consistent":
        M = R_crpt[w] @ R_crpt[other].transpose(0, 2, 1)                             # :104-109
    elif crpt_type == "adv":
        M = R_crpt[w] @ R_orig[other].transpose(0, 2, 1)                             # :112-118
    else:
        raise ValueError(crpt_type)
    RijMat[crpt] = np.where((w == ei[crpt])[:, None, None], M, M.transpose(0, 2, 1))  # k>0: M, else M' (:97-101)
    noise = ~crpt
    RijMat[noise] = RijMat[noise] + sigma_in * randn3(seed, S_NOISE, e[noise])       # :128-130
    RijMat[crpt] = RijMat[crpt] + sigma_out * randn3(seed, S_NOISE_OUT, ce)          # :132-133
    RijMat = proj_so3(RijMat)                                                        # :139-143
    ErrVec = _err_vec(Rij_orig, RijMat)
    Ind = np.stack([ei + 1, ej + 1], axis=1).astype(np.float64)
    return dict(Ind=Ind, RijMat=to_matlab(RijMat), Rij_orig=to_matlab(Rij_orig), R_orig=to_matlab(R_orig),
                ErrVec=ErrVec, corrupted=crpt)


def nonuniform_topology_sequential(n, p, p_node_crpt, p_edge_crpt, sigma_in, sigma_out, crpt_type="uniform", seed=0):
    """The same model with the reference's SEQUENTIAL loops kept literally (Nonuniform_Topology.m:80-124: for every
    corrupted node in ``randperm`` order, for every picked neighbour in ``randperm`` order, overwrite the edge), the
    ``randperm`` orders being the key orders of the counter-based draws.  Exists to check the per-edge rules of
    ``nonuniform_topology`` (and of ``csrc/gen.cu``) against the loop they replace."""
    ei, ej = graph(n, p, seed)
    m = ei.size
    e = np.arange(m, dtype=np.int64)
    R_orig = rand_rot(seed, S_RORIG, np.arange(n))
    Rij_orig = R_orig[ei] @ R_orig[ej].transpose(0, 2, 1)
    RijMat = Rij_orig.copy()
    IndMat = {}
    for k in range(m):                                                               # :57-58
        IndMat[(int(ei[k]), int(ej[k]))] = k + 1
        IndMat[(int(ej[k]), int(ei[k]))] = -(k + 1)
    nodes = np.arange(n, dtype=np.int64)
    node_crpt = np.lexsort((nodes, u64(seed, S_NODEPERM, nodes, 0)))                 # randperm(n), :62
    node_crpt = node_crpt[: int(math.floor(n * p_node_crpt))]                        # :63-64
    crptInd = np.zeros(m, dtype=bool)
    R_crpt = rand_rot(seed, S_RCORR, nodes)
    full_a = np.concatenate([ej, ei])                                                # Ind_full, :42
    full_b = np.concatenate([ei, ej])
    for i in node_crpt:                                                              # :80
        cand = full_b[full_a == i]                                                   # :81
        perm = np.lexsort((cand, u64(seed, S_NBRPERM, np.full(cand.size, i), cand)))  # randperm(length(cand)), :83
        nn = int(math.floor(p_edge_crpt * cand.size))                                # :84
        for j in cand[perm[:nn]]:                                                    # :88
            k = IndMat[(int(i), int(j))]
            ke = abs(k) - 1
            crptInd[ke] = True
            if crpt_type == "uniform":
                M = rand_rot(seed, S_R0, np.array([2 * ke + (0 if k > 0 else 1)]))[0]   # a fresh draw per visit, :91-94
            elif crpt_type == "self-consistent":
                M = R_crpt[i] @ R_crpt[j].T
            else:
                M = R_crpt[i] @ R_orig[j].T
            RijMat[ke] = M if k > 0 else M.T                                         # :97-101
    noise = ~crptInd
    RijMat[noise] = RijMat[noise] + sigma_in * randn3(seed, S_NOISE, e[noise])
    RijMat[crptInd] = RijMat[crptInd] + sigma_out * randn3(seed, S_NOISE_OUT, e[crptInd])
    RijMat = proj_so3(RijMat)
    return dict(Ind=np.stack([ei + 1, ej + 1], axis=1).astype(np.float64), RijMat=to_matlab(RijMat),
                ErrVec=_err_vec(Rij_orig, RijMat), corrupted=crptInd)
