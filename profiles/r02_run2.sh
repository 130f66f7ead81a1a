#!/bin/bash
# single GPU: parity tests after the fill / SpMV / cycle rewrites, fresh bench line, launch-shape experiments,
# one rank's share of an 8-GPU solve profiled on one GPU (DESC_B200_FAKE_SHARD), ncu of the rewritten kernels
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize_parity.py tests/test_gpu_cemp_diag.py -x -q -m gpu ) > gpurun_out/r02_run2_tests.log 2>&1
tail -5 gpurun_out/r02_run2_tests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02_bench_n1_b.json").read().strip().splitlines()[-1])
print("bench", round(j["ms_per_step"],2), {k:round(v,3) for k,v in j["stages_ms"].items()}, j["roofline"]["frac"], j["roofline"]["stages"], j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"], j["cpu_baseline"]["same_graph"])
PY
L=gpurun_out/r02_run2_shapes.log
for st in 8,4,4,2 8,2,2,3 8,2,2,4 8,4,2,2; do
  echo "== DESC_B200_ST=$st" >> $L
  DESC_B200_ST=$st python profiles/profile_pgd.py 20 10000 0.1 2>&1 | grep -o "'pgd_pass1_ms': [0-9.]*, 'pgd_pass2_ms': [0-9.]*" >> $L
done
for fs in 0/8 7/8 0/2; do
  for pw in "" 2 4; do
    echo "== FAKE_SHARD=$fs PB_WARPS=${pw:-auto}" >> $L
    DESC_B200_FAKE_SHARD=$fs DESC_B200_PB_WARPS=$pw python profiles/profile_pgd.py 20 10000 0.1 2>&1 | grep -o "'pgd_pass1_ms': [0-9.]*, 'pgd_pass2_ms': [0-9.]*" >> $L
  done
done
cat $L
DESC_B200_FAKE_SHARD=0/8 python profiles/profile_pgd.py 5 10000 0.1 > /dev/null 2>&1 &&
DESC_B200_FAKE_SHARD=0/8 ncu --set full --clock-control none --import-source on -k regex:"k_pgd_stream|k_pgd_passb" -s 6 -c 2 -o gpurun_out/r02_pgd_shard0of8 python profiles/profile_pgd.py 5 10000 0.1 > gpurun_out/r02_run2_ncu1.log 2>&1
python profiles/profile_pgd.py 2 10000 0.1 gcw > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_fill_slots_reg|k_gcw_spmv|k_recip_flags_thr" -c 3 -o gpurun_out/r02_fill_spmv python profiles/profile_pgd.py 2 10000 0.1 gcw > gpurun_out/r02_run2_ncu2.log 2>&1
tail -2 gpurun_out/r02_run2_ncu1.log gpurun_out/r02_run2_ncu2.log
