"""Short cfg-4 run for ncu: build + d_ijk + a few PGD iterations (+ optional GCW).
usage: python profiles/profile_pgd.py [iters] [workload n] [p] [gcw|-] [n_sample]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import desc_b200  # noqa: E402
from desc_b200 import synth  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.1
do_gcw = len(sys.argv) > 4 and sys.argv[4] == "gcw"
n_sample = int(sys.argv[5]) if len(sys.argv) > 5 else 0
mo = synth.uniform_topology(n, p, 0.2, 0.1, "uniform", seed=0, device="cuda")
Ind_d, R_d = mo["Ind"].reshape(-1).contiguous(), mo["RijMat"].reshape(-1).contiguous()
torch.cuda.synchronize()
t0 = time.perf_counter()
with desc_b200.Solver(Ind_d, R_d, n=n) as s:
    info = s.build_incidence(n_sample=n_sample, seed=1)
    s.cycle_inconsistency()
    _, hist, k = s.pgd(iters, desc_b200.ConstantStepSize(0.01), want_S=False)
    if do_gcw:
        s.gcw(want_R=False)
    tm = s.timings()
print("wall %.3f s" % (time.perf_counter() - t0), info, k, hist[-1] if k else None)
print({k2: round(v, 3) if isinstance(v, float) else v for k2, v in tm.items()})
