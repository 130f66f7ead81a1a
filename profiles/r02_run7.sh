#!/bin/bash
mkdir -p gpurun_out
for wl in cfg5 cfg2; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-side > gpurun_out/r02_bench_${wl}_n1.json 2> gpurun_out/r02_bench_${wl}_n1.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_${wl}_n1.json").read().strip().splitlines()[-1])
    print("$wl", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, "frac", round(j["roofline"]["frac"],4), j["clocks"], j["roofline"]["kernel"][:60])
except Exception as e:
    print("$wl failed", e); print(open("gpurun_out/r02_bench_${wl}_n1.err").read()[-1500:])
PY
done
python bench.py --steps 3 --warmup 3 --no-side --no-cpu | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['ms_per_step'],2), j['roofline']['frac'], j['roofline']['traffic'], j['clocks'])"
python -c "import __graft_entry__ as g; g.smoke()"
