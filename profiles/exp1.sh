timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
python profiles/profile_pgd.py 20 > gpurun_out/plain_s.log 2>&1 && tail -1 gpurun_out/plain_s.log && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_pgd -c 40 --csv --log-file gpurun_out/launches_stream.csv python profiles/profile_pgd.py 6 > gpurun_out/ncu_s.log 2>&1
