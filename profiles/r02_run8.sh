#!/bin/bash
# diagnostic: does the clock sampler perturb the resident timing of the cfg-5 steps?
for e in "DESC_BENCH_NO_CLOCKS=1" "DESC_BENCH_CLOCK_PERIOD=0.1" "DESC_BENCH_CLOCK_PERIOD=0.5"; do
  env $e python bench.py --workload cfg5 --steps 5 --warmup 3 --no-side --no-cpu | python -c "
import sys,json
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$e', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['ms_per_step'],2), j['clocks'])"
done
