#!/bin/bash
# full GPU test suite (with the full-size element-wise parity tests), default bench line, and the ncu capture of the
# two PGD kernels that roofline.traffic is read from
mkdir -p gpurun_out
( time python -m pytest tests -x -q -m gpu --durations=12 ) > gpurun_out/r02_run1_tests.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_a.json 2> gpurun_out/r02_bench_n1_a.err
python profiles/profile_pgd.py 6 10000 0.1 > gpurun_out/r02_run1_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_pgd_stream|k_pgd_passb" -s 6 -c 2 -o gpurun_out/r02_pgd_v1 python profiles/profile_pgd.py 6 10000 0.1 > gpurun_out/r02_run1_ncu.log 2>&1
tail -25 gpurun_out/r02_run1_tests.log; tail -c 1500 gpurun_out/r02_bench_n1_a.json; tail -5 gpurun_out/r02_bench_n1_a.err; tail -3 gpurun_out/r02_run1_ncu.log
