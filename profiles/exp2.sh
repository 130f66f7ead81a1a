python profiles/profile_pgd.py 6 > gpurun_out/plain_s.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pgd -c 40 --csv --log-file gpurun_out/launches_stream.csv python profiles/profile_pgd.py 6 > gpurun_out/ncu_s.log 2>&1
DESC_B200_DBG=31 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pgd -c 40 --csv --log-file gpurun_out/launches_stream31.csv python profiles/profile_pgd.py 6 > gpurun_out/ncu_s31.log 2>&1
tail -3 gpurun_out/ncu_s.log
