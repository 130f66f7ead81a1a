#!/bin/bash
# single GPU: full GPU test suite on the final kernels, bench lines of every configuration, launch list + ncu of cfg 4
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_run3_tests.log 2>&1
tail -4 gpurun_out/r02_run3_tests.log
for wl in cfg4 cfg2 cfg3 cfg3sc cfg5; do
  extra="--no-side"; [ $wl = cfg4 ] && extra=""
  python bench.py --workload $wl --steps 5 --warmup 3 $extra > gpurun_out/r02_bench_${wl}_n1.json 2> gpurun_out/r02_bench_${wl}_n1.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_${wl}_n1.json").read().strip().splitlines()[-1])
    c=j["cpu_baseline"]
    print("$wl", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, "frac", round(j["roofline"]["frac"],4), "traffic", j["roofline"]["traffic"], "cpu", "%.3g"%c["value"], c["cores"], c["same_graph"], "iters", j["config"]["iters_run"], "m_cycle", j["config"]["m_cycle"], j["quality"])
except Exception as e:
    print("$wl failed", e); print(open("gpurun_out/r02_bench_${wl}_n1.err").read()[-1500:])
PY
done
python bench.py --steps 1 --warmup 0 --no-side --no-cpu > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 0 --no-side --no-cpu > gpurun_out/r02_run3_ncu0.log 2>&1
python profiles/profile_pgd.py 6 10000 0.1 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pgd_stream|k_pgd_passb" -s 6 -c 2 -o gpurun_out/r02_pgd_v2 python profiles/profile_pgd.py 6 10000 0.1 > gpurun_out/r02_run3_ncu1.log 2>&1
tail -n 2 gpurun_out/r02_run3_ncu1.log
