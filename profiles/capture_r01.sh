#!/bin/bash
# Round-1 profile capture (run under gpurun, one GPU).  Writes into gpurun_out/; the summaries are
# then copied into profiles/ by profiles/summarise.py (run on the build box).
set -x
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(void )?k_" -c 1500 --csv \
    --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_bench.log 2>&1
python profiles/profile_pgd.py 6 10000 0.1 gcw > gpurun_out/plain_pgd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pgd_passb|k_pgd_stream" -s 4 -c 2 \
    -o gpurun_out/r01_pgd_v7 -f python profiles/profile_pgd.py 6 10000 0.1 gcw > gpurun_out/ncu_pgd.log 2>&1
tail -2 gpurun_out/ncu_pgd.log
