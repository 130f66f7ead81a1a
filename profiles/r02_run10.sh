#!/bin/bash
mkdir -p gpurun_out
for wl in cfg5 cfg4; do
  extra="--no-side"; [ $wl = cfg4 ] && extra=""
  python bench.py --workload $wl --steps 5 --warmup 3 $extra > gpurun_out/r02_bench_${wl}_n1.json 2> gpurun_out/r02_bench_${wl}_n1.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/r02_bench_${wl}_n1.json").read().strip().splitlines()[-1])
print("$wl", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, "frac", round(j["roofline"]["frac"],4), j["roofline"]["traffic"], j["clocks"], j["roofline"]["kernel"][:50], "cpu %.3g"%j["cpu_baseline"]["value"])
PY
done
