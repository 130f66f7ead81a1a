#!/bin/bash
# N GPUs, cfg 4: collective vs point-to-point exchanges, default vs more NCCL p2p channels
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
run() { # name, env...
  name=$1; shift
  env "$@" $TR bench.py --gpus $N --steps 2 --warmup 1 --no-side --no-cpu > gpurun_out/r02_bench_n${N}_$name.json 2> gpurun_out/r02_bench_n${N}_$name.err
  python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/r02_bench_n${N}_$name.json").read().strip().splitlines()[-1])
    print("$name", j["n_gpus"], round(j["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, j["per_rank"], j["parity"]["ok"])
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r02_bench_n${N}_$name.err").read()[-1500:])
PY
}
run coll DESC_B200_COMM=coll
run p2p DESC_B200_COMM=p2p
run p2p32 DESC_B200_COMM=p2p NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run coll32 DESC_B200_COMM=coll NCCL_MIN_NCHANNELS=32
