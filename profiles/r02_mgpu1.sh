#!/bin/bash
# 2 GPUs: N-rank == 1-rank check (small cases + cfg-4 size) with the NCCL-collective exchanges and with the
# point-to-point ones, then the cfg-4 bench line for both
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
$TR tests/multi_gpu_check.py --full > gpurun_out/r02_mgpu_check_n${N}_coll.log 2>&1
tail -6 gpurun_out/r02_mgpu_check_n${N}_coll.log
$TR bench.py --gpus $N --steps 3 --warmup 2 --no-side > gpurun_out/r02_bench_n${N}_coll.json 2> gpurun_out/r02_bench_n${N}_coll.err
DESC_B200_COMM=p2p $TR bench.py --gpus $N --steps 3 --warmup 2 --no-side > gpurun_out/r02_bench_n${N}_p2p.json 2> gpurun_out/r02_bench_n${N}_p2p.err
for f in coll p2p; do python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_n${N}_$f.json").read().strip().splitlines()[-1])
    print("$f", j["n_gpus"], round(j["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, j["per_rank"], j["parity"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r02_bench_n${N}_$f.err").read()[-2000:])
PY
done
