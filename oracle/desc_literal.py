"""Literal, loop-for-loop restatement of the reference  --  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see ``oracle/desc_oracle.py`` header): no MATLAB/Octave in this image and
no golden vectors in the reference.  This second restatement keeps the reference's *dense*
data structures (``AdjMat`` n x n, ``IndMat``, ``RijMat4d``, ``IJK_Mat`` n x m_pos,
``IKJ_appears``/``JKI_appears`` n x m_pos) and its per-edge interpreter loops, so that the
CSR oracle (the one the CUDA build is tested against) has an independent cross-check.
Only usable for n <= ~200.

Follows ``Algorithms/DESC.m:14-263`` and ``Utils/GCW.m:1-38`` statement by statement;
line numbers in comments refer to those files.  Indices are kept **1-based** inside this
file (arrays get a dummy row/column 0 where that keeps the statements literal).
"""
from __future__ import annotations

import math

import numpy as np

from .desc_oracle import abs_acos, matlab_median, sampler_keys


def _datasample(CoInd_ij, n_sample, edge0, seed, permute_rng=None):
    """Stand-in for ``datasample(CoInd_ij, n_sample, 'Replace', false)`` (DESC.m:84): the
    n_sample entries with the smallest shared-sampler key.  MATLAB returns them in random
    order; ``permute_rng`` reproduces that (order only changes FP summation order)."""
    keys = sampler_keys(seed, np.full(CoInd_ij.size, edge0), CoInd_ij - 1)
    order = np.lexsort((CoInd_ij, keys))[:n_sample]
    out = np.sort(CoInd_ij[order])
    if permute_rng is not None:
        out = permute_rng.permutation(out)
    return out


def desc_literal(Ind, RijMat, params, seed=0, n_sample=None, permute_rng=None, run_gcw=True):
    """[R_est(GCW), S_vec, extras] following DESC.m:14-263 (DESC_init.m == same lines)."""
    Ind = np.asarray(Ind).astype(np.int64)
    RijMat = np.asarray(RijMat, dtype=np.float64)            # 3 x 3 x m (MATLAB layout)
    Ind_i = Ind[:, 0]
    Ind_j = Ind[:, 1]
    n = int(Ind.max())                                         # :21
    m = Ind_i.size                                             # :22
    AdjMat = np.zeros((n + 1, n + 1))                          # :23-24 (row/col 0 unused)
    AdjMat[Ind_i, Ind_j] = 1
    AdjMat = AdjMat + AdjMat.T

    CoDeg = (AdjMat @ AdjMat) * AdjMat                         # :29
    CoDeg[(CoDeg == 0) & (AdjMat > 0)] = -1                    # :30
    CoDeg_low = np.tril(CoDeg[1:, 1:], -1)                     # :31
    CoDeg_vec = CoDeg_low.flatten(order="F")                   # :32 column-major (:)
    CoDeg_vec = CoDeg_vec[CoDeg_vec != 0]                      # :34
    assert CoDeg_vec.size == m, "Ind must be in tril(:) order (SURVEY H9)"
    CoDeg_pos_ind = np.nonzero(CoDeg_vec > 0)[0] + 1           # :36 (1-based edge ids)
    CoDeg_vec_pos = CoDeg_vec[CoDeg_pos_ind - 1]               # :37
    if n_sample is None:
        n_sample = max(int(math.ceil(matlab_median(CoDeg_vec_pos) / 4.0)), 30)   # :43
    CoDeg_vec_pos_sampled = np.minimum(CoDeg_vec_pos, n_sample).astype(np.int64)   # :45
    cum_ind = np.concatenate([[0], np.cumsum(CoDeg_vec_pos_sampled)]).astype(np.int64)  # :49
    m_pos = CoDeg_pos_ind.size                                 # :50
    m_cycle = int(cum_ind[-1])                                 # :51
    CoDeg_pos_ind_long = np.zeros(m + 1, dtype=np.int64)       # :53-54
    CoDeg_pos_ind_long[CoDeg_pos_ind] = np.arange(1, m_pos + 1)

    Ind_ij = np.zeros(m_cycle + 1, dtype=np.int64)             # :56-58 (slot 0 unused)
    Ind_jk = np.zeros(m_cycle + 1, dtype=np.int64)
    Ind_ki = np.zeros(m_cycle + 1, dtype=np.int64)
    RijMat4d = np.zeros((3, 3, n + 1, n + 1))                  # :60
    IndMat = np.zeros((n + 1, n + 1), dtype=np.int64)
    for l in range(1, m + 1):                                  # :63-69
        i, j = Ind_i[l - 1], Ind_j[l - 1]
        RijMat4d[:, :, i, j] = RijMat[:, :, l - 1]
        RijMat4d[:, :, j, i] = RijMat[:, :, l - 1].T
        IndMat[i, j] = l
        IndMat[j, i] = l
    Rjk0Mat = np.zeros((3, 3, m_cycle + 1))                    # :71-75
    Rki0Mat = np.zeros((3, 3, m_cycle + 1))
    IJK = np.zeros(m_cycle + 1, dtype=np.int64)
    IKJ = np.zeros(m_cycle + 1, dtype=np.int64)
    JKI = np.zeros(m_cycle + 1, dtype=np.int64)
    IJK_Mat = np.zeros((n + 1, m_pos + 1), dtype=np.int64)     # :77

    for l in range(1, m_pos + 1):                              # :79-96
        IJ = CoDeg_pos_ind[l - 1]
        i, j = Ind_i[IJ - 1], Ind_j[IJ - 1]
        CoInd_ij = np.nonzero(AdjMat[1:, i] * AdjMat[1:, j])[0] + 1     # :82
        if CoInd_ij.size >= n_sample:                          # :83-85
            CoInd_ij = _datasample(CoInd_ij, n_sample, IJ - 1, seed, permute_rng)
        rng_l = slice(cum_ind[l - 1] + 1, cum_ind[l] + 1)
        Ind_ij[rng_l] = IJ                                     # :86
        Ind_jk[rng_l] = IndMat[j, CoInd_ij]                    # :87
        Ind_ki[rng_l] = IndMat[CoInd_ij, i]                    # :88
        Rjk0Mat[:, :, rng_l] = RijMat4d[:, :, j, CoInd_ij]     # :89
        Rki0Mat[:, :, rng_l] = RijMat4d[:, :, CoInd_ij, i]     # :91
        IJK[rng_l] = CoInd_ij                                  # :93
        IJK_Mat[1:CoDeg_vec_pos_sampled[l - 1] + 1, l] = CoInd_ij      # :94

    IKJ_appears = np.zeros((n + 1, m_pos + 1), dtype=bool)     # :100-102
    JKI_appears = np.zeros((n + 1, m_pos + 1), dtype=bool)
    for l in range(1, m_pos + 1):                              # :103-127
        IJ = CoDeg_pos_ind[l - 1]
        i, j = Ind_i[IJ - 1], Ind_j[IJ - 1]
        nsl = CoDeg_vec_pos_sampled[l - 1]
        range_l = np.arange(cum_ind[l - 1] + 1, cum_ind[l] + 1)
        IK = CoDeg_pos_ind_long[IndMat[i, IJK[range_l]]]       # :106
        IK_cum = cum_ind[IK - 1]                               # :110  cum_ind(IK), 1-based
        eq = IJK_Mat[1:, IK] == j                              # :111  n x ns_l
        # [J_ind,~] = find(eq): row indices in column-major order
        cc, rr = np.nonzero(eq.T)
        J_ind = rr + 1
        IKJ_appears[1:nsl + 1, l] = eq.any(axis=0)             # :113
        app = IKJ_appears[1:nsl + 1, l]
        IKJ[range_l[app]] = IK_cum[app] + J_ind                # :116
        JK = CoDeg_pos_ind_long[IndMat[j, IJK[range_l]]]       # :119
        JK_cum = cum_ind[JK - 1]                               # :122
        eq = IJK_Mat[1:, JK] == i                              # :123
        cc, rr = np.nonzero(eq.T)
        I_ind = rr + 1
        JKI_appears[1:nsl + 1, l] = eq.any(axis=0)             # :124
        app = JKI_appears[1:nsl + 1, l]
        JKI[range_l[app]] = JK_cum[app] + I_ind                # :125

    Rij0Mat = np.zeros((3, 3, m_cycle + 1))
    Rij0Mat[:, :, 1:] = RijMat[:, :, Ind_ij[1:] - 1]           # :129
    R_cycle0 = np.zeros((3, 3, m_cycle + 1))                   # :133-143
    R_cycle = np.zeros((3, 3, m_cycle + 1))
    for jj in range(3):
        R_cycle0 = R_cycle0 + Rij0Mat[:, jj:jj + 1, :] * Rjk0Mat[jj:jj + 1, :, :]
    for jj in range(3):
        R_cycle = R_cycle + R_cycle0[:, jj:jj + 1, :] * Rki0Mat[jj:jj + 1, :, :]
    R_trace = (R_cycle[0, 0, :] + R_cycle[1, 1, :]) + R_cycle[2, 2, :]   # :146
    S0_long = abs_acos((R_trace - 1.0) / 2.0) / np.pi          # :147
    S0_long[0] = 0.0
    S_vec = np.ones(m + 1)                                     # :148

    wijk = np.ones(m_cycle + 1)                                # :151-157
    for l in range(1, m_pos + 1):
        IJ = CoDeg_pos_ind[l - 1]
        rng_l = slice(cum_ind[l - 1] + 1, cum_ind[l] + 1)
        weight = wijk[rng_l]
        wijk[rng_l] = weight / weight.sum()
        S_vec[IJ] = wijk[rng_l] @ S0_long[rng_l]

    sum_ikj = np.zeros(m_cycle + 1)                            # :164-165
    sum_jki = np.zeros(m_cycle + 1)
    S_vec_last = S_vec.copy()
    learning_iters = int(params["iters"])                      # :170
    rule = params["Gradient"]
    obj_vals = []
    changes = []
    patience = 30                                              # :180
    misses = 0
    iters_run = 0
    for it in range(1, learning_iters + 1):                    # :182
        for l in range(1, m_pos + 1):                          # :185-191
            nsl = CoDeg_vec_pos_sampled[l - 1]
            range_l = np.arange(cum_ind[l - 1] + 1, cum_ind[l] + 1)
            a = IKJ_appears[1:nsl + 1, l]
            b = JKI_appears[1:nsl + 1, l]
            sum_ikj[range_l[a]] = wijk[IKJ[range_l[a]]].sum()
            sum_jki[range_l[b]] = wijk[JKI[range_l[b]]].sum()
        grad_long = np.zeros(m_cycle + 1)
        grad_long[1:] = S_vec[Ind_jk[1:]] + S_vec[Ind_ki[1:]] + (sum_ikj[1:] + sum_jki[1:]) * S0_long[1:]  # :193
        for l in range(1, m_pos + 1):                          # :195-204
            nsample = CoDeg_vec_pos_sampled[l - 1]
            rng_l = slice(cum_ind[l - 1] + 1, cum_ind[l] + 1)
            grad = grad_long[rng_l]
            nv = np.ones(nsample) / (nsample ** 0.5)
            grad = grad - (grad @ nv) * nv
            grad_long[rng_l] = grad
        wijk[1:] = wijk[1:] + rule.GetStep(grad_long[1:])      # :207
        for l in range(1, m_pos + 1):                          # :208-230
            IJ = CoDeg_pos_ind[l - 1]
            nsample = CoDeg_vec_pos_sampled[l - 1]
            rng_l = slice(cum_ind[l - 1] + 1, cum_ind[l] + 1)
            w_new = wijk[rng_l]
            w = np.sort(w_new)                                 # :215
            Ti = 0
            for i in range(1, nsample + 1):                    # :217-222
                if (w[i - 1:] - w[i - 1]).sum() < 1:
                    Ti = i
                    break
            T = w[Ti - 1] - (1 - (w[Ti - 1:] - w[Ti - 1]).sum()) / w[Ti - 1:].size   # :223
            wijk[rng_l] = np.maximum(w_new - T, 0)             # :224
            S_vec[IJ] = wijk[rng_l] @ S0_long[rng_l]           # :229
        average_change = float(np.mean(np.abs(S_vec[1:] - S_vec_last[1:])))          # :232
        obj_vals.append(float(wijk[1:] @ (S_vec[Ind_jk[1:]] + S_vec[Ind_ki[1:]])))    # :233
        changes.append(average_change)
        iters_run = it
        if it > 1 and obj_vals[-2] - obj_vals[-1] < 10 ** (-5):   # :243
            misses += 1
            if misses >= patience:
                break
        else:
            misses = 0
        S_vec_last = S_vec.copy()                              # :257

    extras = dict(n_sample=n_sample, cum_ind=cum_ind, CoDeg_pos_ind=CoDeg_pos_ind, Ind_ij=Ind_ij[1:],
                  Ind_jk=Ind_jk[1:], Ind_ki=Ind_ki[1:], IJK=IJK[1:], IKJ=IKJ[1:], JKI=JKI[1:],
                  S0_long=S0_long[1:], wijk=wijk[1:], hist=np.stack([changes, obj_vals], axis=1),
                  iters_run=iters_run)
    R_est = gcw_literal(Ind, RijMat, S_vec[1:]) if run_gcw else None   # :263
    return R_est, S_vec[1:].copy(), extras


def gcw_literal(Ind, RijMat, S_vec):
    """Utils/GCW.m:1-38 with the dense non-symmetric matrix and a dense eigen-solver
    (``eigs(RijW,3,'la')`` read as: the 3 eigenvalues of largest real part, unit-norm vectors)."""
    Ind = np.asarray(Ind).astype(np.int64)
    RijMat = np.asarray(RijMat, dtype=np.float64)
    S_vec = np.asarray(S_vec, dtype=np.float64).ravel()
    n = int(Ind.max())
    m = Ind.shape[0]
    d = 3
    Rij_blk = np.zeros((n * d, n * d))                         # :9-13
    for k in range(m):
        i, j = Ind[k, 0] - 1, Ind[k, 1] - 1
        Rij_blk[3 * i:3 * i + 3, 3 * j:3 * j + 3] = RijMat[:, :, k]
    Rij_blk = Rij_blk + Rij_blk.T                              # :15
    AdjMat = np.zeros((n, n))
    AdjMat[Ind[:, 0] - 1, Ind[:, 1] - 1] = 1
    AdjMat = AdjMat + AdjMat.T
    SMat_sq = np.zeros((n, n))                                 # :17-18
    SMat_sq[Ind[:, 0] - 1, Ind[:, 1] - 1] = S_vec
    SMat_sq = SMat_sq + SMat_sq.T
    Weights = (1.0 / (SMat_sq ** 1.5 + 1e-8)) * AdjMat         # :20
    Weights = np.diag(1.0 / Weights.sum(axis=1)) @ Weights     # :21
    Weights = np.kron(Weights, np.ones((d, d)))                # :22
    RijW = Rij_blk * Weights                                   # :23
    lam, V = np.linalg.eig(RijW)                               # :27
    top = np.argsort(-lam.real)[:d]
    V = np.real(V[:, top])
    V = V / np.linalg.norm(V, axis=0, keepdims=True)
    s = np.sign(np.linalg.det(V[:d, :]))                       # :28
    V[:, 0] = V[:, 0] * s
    R_est = np.zeros((d, d, n))                                # :29-36
    for i in range(n):
        Ri = V[3 * i:3 * i + 3, :]
        Ur, _, Vrt = np.linalg.svd(Ri)
        S0 = np.diag([1.0, 1.0, np.linalg.det(Ur @ Vrt)])
        R_est[:, :, i] = Ur @ S0 @ Vrt
    return R_est


def cemp_literal(Ind, RijMat, CEMP_parameters, CoIndMat, return_gcw=False):
    """Algorithms/CEMP.m:25-131 statement by statement (== CEMP_GCW.m:25-125), dense structures.

    ``CoIndMat`` (nsample x m, 1-based apices, column l used only for edges with a triangle) is what
    ``datasample(find(AdjMat(:,i).*AdjMat(:,j)), nsample)`` (:63, WITH replacement) returned: MATLAB's
    RNG cannot be restated, so the draw is an input.  ``return_gcw``: also CEMP_GCW.m:127-159."""
    Ind = np.asarray(Ind).astype(np.int64)
    RijMat = np.asarray(RijMat, dtype=np.float64)
    T = int(CEMP_parameters["max_iter"])                        # :27
    beta_cemp = [float(b) for b in np.asarray(CEMP_parameters["reweighting"], dtype=np.float64).ravel()]   # :28
    nsample = int(CEMP_parameters["nsample"])                   # :29
    T_beta = len(beta_cemp)                                     # :30
    if T_beta < T:                                              # :31-35
        beta_cemp = beta_cemp + [beta_cemp[-1]] * (T - T_beta)
    Ind_i = Ind[:, 0]                                           # :38
    Ind_j = Ind[:, 1]
    n = int(Ind.max())                                          # :40
    m = Ind_i.size                                              # :41
    AdjMat = np.zeros((n + 1, n + 1))                           # :42-43 (row/col 0 unused)
    AdjMat[Ind_i, Ind_j] = 1
    AdjMat = AdjMat + AdjMat.T
    CoDeg = (AdjMat @ AdjMat) * AdjMat                          # :50
    AdjPos = AdjMat.copy()                                      # :51
    AdjPos[CoDeg > 0] = -1                                      # :53
    AdjPosLow = np.tril(AdjPos[1:, 1:]).flatten(order="F")      # :54
    AdjPosLow = AdjPosLow[AdjPosLow != 0]                       # :55
    IndPos = np.nonzero(AdjPosLow < 0)[0] + 1                   # :57 (1-based edge ids)
    IndPosbin = np.zeros(m + 1, dtype=bool)                     # :58-59
    IndPosbin[IndPos] = True
    CoIndMat = np.asarray(CoIndMat).astype(np.int64)
    assert CoIndMat.shape == (nsample, m)
    for l in IndPos:                                            # :63 (the draw itself is the input)
        i, j = Ind_i[l - 1], Ind_j[l - 1]
        assert (AdjMat[CoIndMat[:, l - 1], i] * AdjMat[CoIndMat[:, l - 1], j] == 1).all()
    RijMat4d = np.zeros((3, 3, n + 1, n + 1))                   # :69-76
    IndMat = np.zeros((n + 1, n + 1), dtype=np.int64)
    for l in range(1, m + 1):
        i, j = Ind_i[l - 1], Ind_j[l - 1]
        RijMat4d[:, :, i, j] = RijMat[:, :, l - 1]
        RijMat4d[:, :, j, i] = RijMat[:, :, l - 1].T
        IndMat[i, j] = l
        IndMat[j, i] = -l
    Rki0 = np.zeros((3, 3, m, nsample))                         # :79-84
    Rjk0 = np.zeros((3, 3, m, nsample))
    for l in IndPos:
        Rki0[:, :, l - 1, :] = RijMat4d[:, :, CoIndMat[:, l - 1], Ind_i[l - 1]]
        Rjk0[:, :, l - 1, :] = RijMat4d[:, :, Ind_j[l - 1], CoIndMat[:, l - 1]]
    # :87-89 reshape with the edge index fastest: slot s*m + l
    Rki0Mat = Rki0.reshape(3, 3, m * nsample, order="F")
    Rjk0Mat = Rjk0.reshape(3, 3, m * nsample, order="F")
    Rij0Mat = np.tile(RijMat, (1, 1, nsample))
    R_cycle0 = np.zeros((3, 3, m * nsample))                    # :91-95
    R_cycle = np.zeros((3, 3, m * nsample))
    for j in range(3):
        R_cycle0 = R_cycle0 + Rij0Mat[:, j, None, :] * Rjk0Mat[j, None, :, :]
    for j in range(3):                                          # :96-98
        R_cycle = R_cycle + R_cycle0[:, j, None, :] * Rki0Mat[j, None, :, :]
    R_trace = ((R_cycle[0, 0, :] + R_cycle[1, 1, :]) + R_cycle[2, 2, :]).reshape(m, nsample, order="F").T   # :99
    S0Mat = abs_acos((R_trace - 1.0) / 2.0) / np.pi             # :100  (nsample x m)
    SVec = np.zeros(m + 1)
    SVec[1:] = S0Mat.mean(axis=0)                               # :101
    SVec[~IndPosbin] = 1                                        # :102
    SVec[0] = 0
    hist = [SVec[1:].copy()]
    for it in range(T):                                         # :106
        beta = beta_cemp[it]                                    # :108
        Ski = np.zeros((nsample, m))                            # :109-110
        Sjk = np.zeros((nsample, m))
        for l in IndPos:                                        # :111-115
            i, j = Ind_i[l - 1], Ind_j[l - 1]
            Ski[:, l - 1] = SVec[np.abs(IndMat[i, CoIndMat[:, l - 1]])]
            Sjk[:, l - 1] = SVec[np.abs(IndMat[j, CoIndMat[:, l - 1]])]
        Smax = Ski + Sjk                                        # :116
        WeightMat = np.exp(-beta * Smax)                        # :118
        weightsum = WeightMat.sum(axis=0)                       # :119
        WeightMat = WeightMat / weightsum[None, :]              # :121
        SMat = WeightMat * S0Mat                                # :122
        SVec[1:] = SMat.sum(axis=0)                             # :124
        SVec[~IndPosbin] = 1                                    # :125
        SVec[0] = 0
        hist.append(SVec[1:].copy())
    out = SVec[1:].copy()
    if not return_gcw:
        return out, dict(S0Mat=S0Mat, hist=hist, IndPos=IndPos)
    # CEMP_GCW.m:127-159: GCW.m with Weights = 1./(SMat_sq+1e-8) (:141)
    d = 3
    Rij_blk = np.zeros((n * d, n * d))
    for k in range(m):
        i, j = Ind[k, 0] - 1, Ind[k, 1] - 1
        Rij_blk[3 * i:3 * i + 3, 3 * j:3 * j + 3] = RijMat[:, :, k]
    Rij_blk = Rij_blk + Rij_blk.T
    SMat_sq = np.zeros((n, n))
    SMat_sq[Ind[:, 0] - 1, Ind[:, 1] - 1] = out
    SMat_sq = SMat_sq + SMat_sq.T
    Weights = (1.0 / (SMat_sq + 1e-8)) * AdjMat[1:, 1:]         # :141
    Weights = np.diag(1.0 / Weights.sum(axis=1)) @ Weights      # :142
    Weights = np.kron(Weights, np.ones((d, d)))
    RijW = Rij_blk * Weights
    lam, V = np.linalg.eig(RijW)                                # :148
    top = np.argsort(-lam.real)[:d]
    V = np.real(V[:, top])
    V = V / np.linalg.norm(V, axis=0, keepdims=True)
    V[:, 0] = V[:, 0] * np.sign(np.linalg.det(V[:d, :]))        # :149
    R_est = np.zeros((d, d, n))
    for i in range(n):                                          # :151-157
        Ur, _, Vrt = np.linalg.svd(V[3 * i:3 * i + 3, :])
        R_est[:, :, i] = Ur @ np.diag([1.0, 1.0, np.linalg.det(Ur @ Vrt)]) @ Vrt
    return out, dict(S0Mat=S0Mat, hist=hist, IndPos=IndPos, R_est=R_est)


def cemp_draw(Ind, nsample, rng):
    """A with-replacement draw of CEMP.m:63 (numpy RNG standing in for datasample): returns
    (CoIndMat nsample x m 1-based, zeros for edges without a triangle; ptr over all edges; apex 0-based)."""
    Ind = np.asarray(Ind).astype(np.int64)
    n = int(Ind.max())
    m = Ind.shape[0]
    A = np.zeros((n + 1, n + 1), dtype=bool)
    A[Ind[:, 0], Ind[:, 1]] = True
    A = A | A.T
    rng = np.random.default_rng(rng)
    CoIndMat = np.zeros((nsample, m), dtype=np.int64)
    cnt = np.zeros(m, dtype=np.int64)
    for l in range(m):
        c = np.nonzero(A[:, Ind[l, 0]] & A[:, Ind[l, 1]])[0]
        if c.size:
            CoIndMat[:, l] = c[rng.integers(0, c.size, nsample)]
            cnt[l] = nsample
    ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    apex = (CoIndMat.T[cnt > 0].ravel() - 1).astype(np.int32)
    return CoIndMat, ptr, apex
