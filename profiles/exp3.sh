DESC_B200_DBG=32 python profiles/profile_pgd.py 6 > gpurun_out/plain_s.log 2>&1 && DESC_B200_DBG=32 ncu --set full --clock-control none --import-source on -k regex:k_pgd_stream -s 2 -c 1 -o gpurun_out/stream_v3 -f python profiles/profile_pgd.py 6 > gpurun_out/ncu_s.log 2>&1
tail -3 gpurun_out/ncu_s.log
