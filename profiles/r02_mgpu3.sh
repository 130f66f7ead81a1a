#!/bin/bash
# N GPUs: peer-memory exchanges (default) vs NCCL point-to-point: parity check + cfg-4 bench line
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551"
timeout 600 $TR tests/multi_gpu_check.py --full > gpurun_out/r02_mgpu_check_n${N}_peer.log 2>&1
tail -4 gpurun_out/r02_mgpu_check_n${N}_peer.log
run() { # name, env...
  name=$1; shift
  timeout 300 env "$@" $TR bench.py --gpus $N --steps 3 --warmup 2 --no-side --no-cpu > gpurun_out/r02_bench_n${N}_$name.json 2> gpurun_out/r02_bench_n${N}_$name.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", j["n_gpus"], round(j["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, j["per_rank"], j["parity"]["ok"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_bench_n${N}_$name.err").read()[-1500:])
PY
}
run peer DESC_B200_COMM=peer
run p2pb DESC_B200_COMM=p2p
