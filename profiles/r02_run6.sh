#!/bin/bash
# last job of round 2: parity after the two path switches (fill kernel by co-degree, blocked PGD for sparse graphs),
# bench lines of the affected workloads, the PGD capture that tags roofline.traffic to the final sources
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
for wl in cfg3 cfg3sc cfg5; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-side > gpurun_out/r02_bench_${wl}_n1.json 2> gpurun_out/r02_bench_${wl}_n1.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_${wl}_n1.json").read().strip().splitlines()[-1])
    print("$wl", round(j["ms_per_step"],2), "e2e", round(j["e2e"]["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, "frac", round(j["roofline"]["frac"],4), "cpu %.3g"%j["cpu_baseline"]["value"])
except Exception as e:
    print("$wl failed", e); print(open("gpurun_out/r02_bench_${wl}_n1.err").read()[-1500:])
PY
done
python profiles/profile_pgd.py 6 10000 0.1 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pgd_stream|k_pgd_passb" -s 6 -c 2 -o gpurun_out/r02_pgd_v3 python profiles/profile_pgd.py 6 10000 0.1 > gpurun_out/r02_run6_ncu.log 2>&1
tail -n 2 gpurun_out/r02_run6_ncu.log
