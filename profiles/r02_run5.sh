#!/bin/bash
# final single-GPU pass of round 2: full GPU suite, the default bench line, and two A/Bs (cfg-3 fill kernel, cfg-5 path)
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_run5_tests.log 2>&1
tail -4 gpurun_out/r02_run5_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_cfg4_n1.json 2> gpurun_out/r02_bench_cfg4_n1.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02_bench_cfg4_n1.json").read().strip().splitlines()[-1])
print("cfg4", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, "frac", round(j["roofline"]["frac"],4), "traffic", j["roofline"]["traffic"], j["roofline"]["traffic_source"], "cpu", "%.3g"%j["cpu_baseline"]["value"], j["cpu_baseline"]["same_graph"])
PY
L=gpurun_out/r02_run5_ab.log
cat > /tmp/ab.py <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import desc_b200
kind, iters = sys.argv[1], int(sys.argv[2])
if kind == "cfg3":
    mo = desc_b200.Nonuniform_Topology(2000, 0.5, 0.3, 0.5, 0.1, 0.1, "adv", seed=0, on_device=True); ns, lr = 0, 0.01
else:
    mo = desc_b200.Ring_Topology(50000, 100, 75, 0.2, 0.05, seed=2, on_device=True); ns, lr = 50, 1.0
for rep in range(2):
    with desc_b200.Solver(mo.Ind, mo.RijMat, n=mo.n) as s:
        s.build_incidence(n_sample=ns, seed=1); s.cycle_inconsistency()
        s.pgd(iters, desc_b200.ConstantStepSize(lr), want_S=False, want_hist=False)
        t = s.timings()
print({k: round(t[k], 3) for k in ("build_ms", "cycle_ms", "pgd_ms", "pgd_pass1_ms", "pgd_pass2_ms", "pgd_iter_ms")})
PY
echo "== cfg3 fill reg (T=32)" >> $L; python /tmp/ab.py cfg3 3 >> $L 2>&1
echo "== cfg3 fill generic" >> $L; DESC_B200_FILL=generic python /tmp/ab.py cfg3 3 >> $L 2>&1
for st in 8,2,2,4 8,4,4,2 8,2,2,6; do echo "== cfg5 stream $st" >> $L; DESC_B200_ST=$st python /tmp/ab.py cfg5 30 >> $L 2>&1; done
echo "== cfg5 blocked" >> $L; DESC_B200_PGD_PATH=blocked python /tmp/ab.py cfg5 30 >> $L 2>&1
echo "== cfg5 generic" >> $L; DESC_B200_PGD_PATH=generic python /tmp/ab.py cfg5 30 >> $L 2>&1
cat $L
