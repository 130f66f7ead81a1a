#!/bin/bash
# ablations of k_pgd_ell at cfg 4 (timing only) + one ncu --set full capture of the default build
mkdir -p gpurun_out
for v in nored nogather noproj noredgather; do
  echo "== ell $v" >> gpurun_out/r02_ab2.log
  DESC_B200_LIB=$PWD/desc_b200/libdesc_b200_$v.so python profiles/profile_pgd.py 12 10000 0.1 >> gpurun_out/r02_ab2.log 2>&1
done
echo "== ell default" >> gpurun_out/r02_ab2.log
python profiles/profile_pgd.py 6 10000 0.1 >> gpurun_out/r02_ab2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_pgd_ell -s 3 -c 1 -o gpurun_out/r02_ell_v1 python profiles/profile_pgd.py 6 10000 0.1 > gpurun_out/r02_ab2_ncu.log 2>&1
grep -E "==|pgd_iter_ms" gpurun_out/r02_ab2.log | sed -E "s/.*'pgd_iter_ms': ([0-9.]+).*/\1/"
tail -3 gpurun_out/r02_ab2_ncu.log
