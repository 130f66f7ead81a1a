#!/bin/bash
# Round-2 profile capture (run under gpurun, one GPU; ~3 GPU-minutes).  Same passes as capture_r01.sh plus the
# SURVEY 8(f) stages added at the end of round 1 (CEMP, CEMP+MST, MPLS, Spectral, generator).  Every ncu pass
# comes after the same command has exited 0 without ncu.  Outputs go to gpurun_out/; copy the summaries into
# profiles/ with profiles/summarise.py on the build box.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/plain_bench_r02.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^(void )?k_" -c 3000 --csv \
    --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-cpu > gpurun_out/ncu_bench_r02.log 2>&1
python profiles/profile_pgd.py 6 10000 0.1 gcw > gpurun_out/plain_pgd_r02.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pgd_passb|k_pgd_stream" -s 4 -c 2 \
    -o gpurun_out/r02_pgd -f python profiles/profile_pgd.py 6 10000 0.1 gcw > gpurun_out/ncu_pgd_r02.log 2>&1
python profiles/profile_cemp.py > gpurun_out/plain_cemp_r02.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_cemp_reweight|k_gen_edges|k_gen_rows" -s 1 -c 4 \
    -o gpurun_out/r02_cemp_gen -f python profiles/profile_cemp.py > gpurun_out/ncu_cemp_r02.log 2>&1
tail -2 gpurun_out/ncu_pgd_r02.log gpurun_out/ncu_cemp_r02.log
