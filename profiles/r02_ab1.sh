#!/bin/bash
# first A/B of round 2: lane-per-edge path (default) vs the round-1 streamed path at cfg 4, plus build variants
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r02_ab1_tests.log
for v in "" keepd tb256 tb64; do
  lib=desc_b200/libdesc_b200${v:+_$v}.so
  echo "== ell ${v:-default}" >> gpurun_out/r02_ab1.log
  DESC_B200_LIB=$PWD/$lib python profiles/profile_pgd.py 30 10000 0.1 gcw >> gpurun_out/r02_ab1.log 2>&1
done
echo "== stream (round 1)" >> gpurun_out/r02_ab1.log
DESC_B200_PGD_PATH=stream python profiles/profile_pgd.py 30 10000 0.1 gcw >> gpurun_out/r02_ab1.log 2>&1
echo "== ell cfg2" >> gpurun_out/r02_ab1.log
python profiles/profile_pgd.py 30 1000 0.5 gcw >> gpurun_out/r02_ab1.log 2>&1
echo "== stream cfg2" >> gpurun_out/r02_ab1.log
DESC_B200_PGD_PATH=stream python profiles/profile_pgd.py 30 1000 0.5 gcw >> gpurun_out/r02_ab1.log 2>&1
cat gpurun_out/r02_ab1_tests.log gpurun_out/r02_ab1.log
