"""cfg-4 run of the whole DESC() pipeline incl. the LAA refinement (DESC.m:265-312), for timing / ncu.
usage: python profiles/profile_laa.py [q] [sigma]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import desc_b200  # noqa: E402
from desc_b200 import synth  # noqa: E402

q = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
mo = synth.uniform_topology(10000, 0.1, q, sigma, "uniform", seed=0, device="cuda")
Ind_d, R_d = mo["Ind"].reshape(-1).contiguous(), mo["RijMat"].reshape(-1).contiguous()
with desc_b200.Solver(Ind_d, R_d, n=10000) as s:
    s.build_incidence(n_sample=0, seed=1)
    s.cycle_inconsistency()
    s.pgd(100, desc_b200.ConstantStepSize(0.01), want_S=False)
    s.gcw(want_R=False)
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        R, sc = s.refine()
        t1 = time.perf_counter()
        tm = s.timings()
        print("refine call %d: %d IRLS iterations, scores %s, wall %.1f ms, laa_ms %.1f, CG iterations %d" % (
            rep, len(sc), sc[:4], 1e3 * (t1 - t0), tm["laa_ms"], tm["laa_cg_iters"]))
