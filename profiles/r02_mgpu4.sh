#!/bin/bash
# debugging: N-rank == 1-rank at cfg-4 size for (shard model) x (exchange implementation)
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571"
for sh in cost slots; do for cm in peer p2p; do
  echo "== SHARD=$sh COMM=$cm"
  DESC_B200_SHARD=$sh DESC_B200_COMM=$cm timeout 300 $TR bench.py --gpus $N --steps 1 --warmup 1 --no-side --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print(round(j['ms_per_step'],2), j['per_rank']['pass1_ms'], j['parity'])
"
done; done
