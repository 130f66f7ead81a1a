#!/usr/bin/env python
"""Python twin of the reference's Demo/compare_algorithms.m on the B200 library.

Same parameters and the same order of calls as the MATLAB script (file:line in the comments); every algorithm and
the data generator run on the GPU through the C ABI.  The two IRLS comparators (IRLS_GM.m, IRLS_L12.m: third-party
L1/IRLS kit) are outside this repo's scope and are not in the table.

    python demo/compare_algorithms.py [--n 100] [--p 0.5] [--q 0.2] [--sigma 0.1] [--model uniform] [--seed 0]
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import desc_b200  # noqa: E402


def run(n=100, p=0.5, q=0.2, sigma=0.1, model="uniform", seed=0, verbose=False):
    # parameters with uniform topology; generate data                                   compare_algorithms.m:9-13
    model_out = desc_b200.Uniform_Topology(n, p, q, sigma, model, seed=seed)
    Ind, RijMat = model_out["Ind"], model_out["RijMat"]                                 # :20-23
    ErrVec, R_orig = model_out["ErrVec"], model_out["R_orig"]
    # CEMP default parameters                                                            :26-30
    CEMP_parameters = dict(max_iter=6, reweighting=2.0 ** (np.arange(1, 7) - 1), nsample=50, gcw_beta=5, seed=seed)
    # MPLS default parameters                                                            :33-37
    MPLS_parameters = dict(stop_threshold=1e-3, max_iter=100, reweighting=CEMP_parameters["reweighting"][-1:],
                           thresholding=[0.95, 0.9, 0.85, 0.8], cycle_info_ratio=1.0 / (np.arange(1, 101) + 1))
    # DESC default parameters                                                            :40-46
    lr = 0.01
    DESC_parameters = dict(iters=100, learning_rate=lr, make_plots=False, Gradient=desc_b200.ConstantStepSize(lr),
                           R_orig=R_orig, ErrVec=ErrVec, seed=seed, verbose=verbose)
    R_MPLS, R_CEMP_MST = desc_b200.MPLS(Ind, RijMat, CEMP_parameters, MPLS_parameters)   # :58
    R_SP = desc_b200.Spectral(Ind, RijMat)                                               # :61
    R_CEMP_GCW = desc_b200.CEMP_GCW(Ind, RijMat, CEMP_parameters)                        # :64
    R_DESC, R_DESC_init, S_vec = desc_b200.DESC(Ind, RijMat, DESC_parameters)            # :71
    rows = []
    for name, R in (("Spectral", R_SP), ("CEMP+MST", R_CEMP_MST), ("CEMP+GCW", R_CEMP_GCW), ("MPLS", R_MPLS),
                    ("DESC_init", R_DESC_init), ("DESC", R_DESC)):                       # :74-95
        _, _, mean_error, median_error = desc_b200.Rotation_Alignment(R, R_orig)
        rows.append((name, mean_error, median_error))
    return rows, dict(S_vec=S_vec, ErrVec=ErrVec, m=Ind.shape[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100)
    ap.add_argument("--p", type=float, default=0.5)
    ap.add_argument("--q", type=float, default=0.2)
    ap.add_argument("--sigma", type=float, default=0.1)
    ap.add_argument("--model", default="uniform")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    rows, extra = run(a.n, a.p, a.q, a.sigma, a.model, a.seed, a.verbose)
    print("%-12s %12s %12s" % ("Algorithms", "MeanError", "MedianError"))                # the `Results` table, :84-98
    for name, mean_error, median_error in rows:
        print("%-12s %12.4f %12.4f" % (name, mean_error, median_error))
    print("edges %d, mean |S_vec - ErrVec| = %.4f" % (extra["m"], float(np.mean(np.abs(extra["S_vec"].ravel() - extra["ErrVec"].ravel())))))


if __name__ == "__main__":
    main()
