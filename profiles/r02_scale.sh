#!/bin/bash
# strong scaling of the bench line on N GPUs of one box: cfg 4 (and cfg 5 when asked), peer-memory exchanges (default)
mkdir -p gpurun_out
N=$1; shift
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
for wl in "$@"; do
  timeout 600 $TR bench.py --gpus $N --workload $wl --steps 5 --warmup 3 > gpurun_out/r02_bench_${wl}_n${N}.json 2> gpurun_out/r02_bench_${wl}_n${N}.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_${wl}_n${N}.json").read().strip().splitlines()[-1])
    print("$wl", j["n_gpus"], round(j["ms_per_step"],2), "e2e", round(j["e2e"]["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, "frac", round(j["roofline"]["frac"],4), j["per_rank"], j["parity"])
except Exception as e:
    print("$wl failed", e); print(open("gpurun_out/r02_bench_${wl}_n${N}.err").read()[-1500:])
PY
done
