"""Synthetic inputs with the distribution and memory layout of the reference's generators,
drawn on the device with torch (input generation only -- not part of the solver hot path).

``uniform_topology`` follows Models/Uniform_Topology.m:24-111: Erdos-Renyi G(n,p) edge list
sorted by (i,j); Haar rotations R_i; R_ij = R_i R_j'; every edge independently corrupted with
probability q by a fresh Haar rotation ('uniform') or R^c_i R^c_j' + noise ('self-consistent');
inliers get sigma*randn(3) added and are projected back to SO(3); ErrVec = geodesic/pi distance to
the clean relative rotation.  The projection uses the Newton polar iteration instead of an SVD
(identical result for det>0 inputs); Haar rotations come from normalised Gaussian quaternions
instead of the SVD of a Gaussian matrix (same distribution).

Returns tensors already in MATLAB memory order: ``Ind`` (2, m) float64 = all i then all j
(1-based), ``RijMat`` (m, 3, 3) float64 with [e, c, r] = R_e(r, c), i.e. byte-identical to MATLAB's
3x3xm array.  ``to_host`` converts to the numpy arrays the reference-style API takes.
"""
from __future__ import annotations

import numpy as np
import torch


def _haar(k, gen, device):
    q = torch.randn(k, 4, generator=gen, device=device, dtype=torch.float64)
    q = q / q.norm(dim=1, keepdim=True)
    w, x, y, z = q.unbind(1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], dim=1)
    return R.view(k, 3, 3)


def _inv_t(X):
    """inverse-transpose of a batch of 3x3 via cofactors"""
    a, b, c = X[:, 0, 0], X[:, 0, 1], X[:, 0, 2]
    d, e, f = X[:, 1, 0], X[:, 1, 1], X[:, 1, 2]
    g, h, i = X[:, 2, 0], X[:, 2, 1], X[:, 2, 2]
    C = torch.stack([e * i - f * h, f * g - d * i, d * h - e * g,
                     c * h - b * i, a * i - c * g, b * g - a * h,
                     b * f - c * e, c * d - a * f, a * e - b * d], dim=1).view(-1, 3, 3)
    det = a * C[:, 0, 0] + b * C[:, 0, 1] + c * C[:, 0, 2]
    return C / det[:, None, None]


def _polar(X, iters=14):
    for _ in range(iters):
        X = 0.5 * (X + _inv_t(X))
    return X


def uniform_topology(n, p, q, sigma, model="uniform", seed=0, device="cuda"):
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    mask = torch.rand(n, n, generator=gen, device=device) < p
    mask = torch.triu(mask, 1)
    ij = mask.nonzero()                        # sorted by i then j, i < j
    del mask
    ei, ej = ij[:, 0], ij[:, 1]
    m = ei.numel()
    R_orig = _haar(n, gen, device)
    Rij_orig = R_orig[ei] @ R_orig[ej].transpose(1, 2)
    corr = torch.rand(m, generator=gen, device=device, dtype=torch.float64) < q
    Rij = _polar(Rij_orig + sigma * torch.randn(m, 3, 3, generator=gen, device=device, dtype=torch.float64)) \
        if sigma > 0 else Rij_orig.clone()
    nc = int(corr.sum())
    if model == "uniform":
        Rij[corr] = _haar(nc, gen, device)
    else:
        Rc = _haar(n, gen, device)
        Q = Rc[ei[corr]] @ Rc[ej[corr]].transpose(1, 2)
        if sigma > 0:
            Q = _polar(Q + sigma * torch.randn(nc, 3, 3, generator=gen, device=device, dtype=torch.float64))
        Rij[corr] = Q
    tr = (Rij_orig * Rij).sum(dim=(1, 2))
    ErrVec = torch.acos(((tr - 1) / 2).clamp(-1, 1)) / np.pi
    Ind = torch.stack([ei + 1, ej + 1], dim=0).to(torch.float64).contiguous()      # (2, m): all i, then all j
    RijMat = Rij.transpose(1, 2).contiguous()                                        # [e, c, r]
    return dict(n=n, m=m, Ind=Ind, RijMat=RijMat, R_orig=R_orig.transpose(1, 2).contiguous(), ErrVec=ErrVec,
                corrupted=corr)


def ring_topology(n, deg, window, q, sigma, seed=0, device="cuda"):
    """SfM-shaped graph for configs[4] of BASELINE.json (the reference has no such generator):
    cameras on a closed track, node i is connected to each of the `window` following nodes (mod n)
    independently with probability deg / (2 window), so the mean degree is `deg` and neighbours share
    many neighbours (co-degree ~ (2 window - |i-j|) (deg / 2 window)^2), unlike an Erdos-Renyi graph of
    the same density.  Rotations, corruption and noise as in ``uniform_topology`` ('uniform' model)."""
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    p = deg / (2.0 * window)
    off = torch.arange(1, window + 1, device=device)
    keep = torch.rand(n, window, generator=gen, device=device) < p
    a = torch.arange(n, device=device)[:, None].expand(n, window)[keep]
    b = (a + off[None, :].expand(n, window)[keep]) % n
    lo, hi = torch.minimum(a, b), torch.maximum(a, b)
    key = torch.unique(lo * n + hi)                       # sorted by (i, j), duplicates removed
    ei, ej = key // n, key % n
    m = ei.numel()
    R_orig = _haar(n, gen, device)
    Rij_orig = R_orig[ei] @ R_orig[ej].transpose(1, 2)
    corr = torch.rand(m, generator=gen, device=device, dtype=torch.float64) < q
    Rij = _polar(Rij_orig + sigma * torch.randn(m, 3, 3, generator=gen, device=device, dtype=torch.float64)) \
        if sigma > 0 else Rij_orig.clone()
    Rij[corr] = _haar(int(corr.sum()), gen, device)
    tr = (Rij_orig * Rij).sum(dim=(1, 2))
    ErrVec = torch.acos(((tr - 1) / 2).clamp(-1, 1)) / np.pi
    Ind = torch.stack([ei + 1, ej + 1], dim=0).to(torch.float64).contiguous()
    return dict(n=n, m=m, Ind=Ind, RijMat=Rij.transpose(1, 2).contiguous(),
                R_orig=R_orig.transpose(1, 2).contiguous(), ErrVec=ErrVec, corrupted=corr)


def to_host(model_out):
    """numpy views in the reference's shapes: Ind (m,2), RijMat (3,3,m), R_orig (3,3,n), ErrVec (m,)."""
    Ind = model_out["Ind"].cpu().numpy().T                       # (m, 2), Fortran-contiguous
    R = model_out["RijMat"].cpu().numpy().transpose(2, 1, 0)     # (3, 3, m), Fortran-contiguous
    Ro = model_out["R_orig"].cpu().numpy().transpose(2, 1, 0)
    return dict(Ind=Ind, RijMat=R, R_orig=Ro, ErrVec=model_out["ErrVec"].cpu().numpy())
