#!/bin/bash
mkdir -p gpurun_out
for wl in cfg2 cfg5; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-side > gpurun_out/r02_bench_${wl}_n1.json 2> gpurun_out/r02_bench_${wl}_n1.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_${wl}_n1.json").read().strip().splitlines()[-1])
    c=j["cpu_baseline"]
    print("$wl", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, "frac", round(j["roofline"]["frac"],4), "cpu", "%.3g"%c["value"], c["cores"], c["same_graph"], "iters", j["config"]["iters_run"], "m_cycle", j["config"]["m_cycle"], j["quality"])
except Exception as e:
    print("$wl failed", e); print(open("gpurun_out/r02_bench_${wl}_n1.err").read()[-1500:])
PY
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_cfg4_reference.json 2> gpurun_out/r02_bench_cfg4_reference.err; tail -c 600 gpurun_out/r02_bench_cfg4_reference.json
