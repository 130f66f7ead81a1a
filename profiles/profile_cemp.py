"""Profiling driver for the CEMP reweighting kernel at the bench workload (cfg 4 graph,
CEMP_parameters.nsample = 50): generates the graph on the device (csrc/gen.cu), builds the incidence, d_ijk, then
runs CEMP.  Used as the command under ncu (see profiles/README.md):

    ncu --set full --clock-control none --import-source on -k regex:k_cemp_reweight -s 1 -c 1 \
        -o gpurun_out/r01_cemp python profiles/profile_cemp.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import desc_b200  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
with desc_b200.Uniform_Topology(n, 0.1, 0.2, 0.1, seed=0, on_device=True) as mo:
    with desc_b200.Solver(mo.Ind, mo.RijMat) as s:
        info = s.build_incidence(n_sample=50, seed=1)
        s.cycle_inconsistency()
        s.cemp(3, 2.0 ** np.arange(3))
        s.cemp(3, 2.0 ** np.arange(3))
        t = s.timings()
        print("n=%d m=%d m_cycle=%d generator %.2f ms, cemp %.3f ms (%.3f ms per reweighting)" % (
            n, info["m"], info["m_cycle"], mo.gen_ms, t["cemp_ms"], t["cemp_ms"] / 4.0))
