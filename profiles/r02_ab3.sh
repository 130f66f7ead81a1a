#!/bin/bash
# k_cycle with cooperative record gathers: parity tests + timing (cycle_ms) for 3 and 2 resident CTAs per SM
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not lane_per_edge and not streamed_path and not slot_lists" 2>&1 | tail -4 > gpurun_out/r02_ab3.log
for v in "" cyc2; do
  echo "== cycle ${v:-minb3}" >> gpurun_out/r02_ab3.log
  DESC_B200_LIB=$PWD/desc_b200/libdesc_b200${v:+_$v}.so python profiles/profile_pgd.py 3 10000 0.1 2>&1 | grep -o "'cycle_ms': [0-9.]*" >> gpurun_out/r02_ab3.log
  DESC_B200_LIB=$PWD/desc_b200/libdesc_b200${v:+_$v}.so python profiles/profile_pgd.py 3 1000 0.5 2>&1 | grep -o "'cycle_ms': [0-9.]*" >> gpurun_out/r02_ab3.log
done
python profiles/profile_pgd.py 2 10000 0.1 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_cycle -c 1 -o gpurun_out/r02_cycle_v1 python profiles/profile_pgd.py 2 10000 0.1 > gpurun_out/r02_ab3_ncu.log 2>&1
cat gpurun_out/r02_ab3.log
