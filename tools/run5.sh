mkdir -p gpurun_out
python tools/profile_step.py 128 2 > gpurun_out/plain_step128.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python tools/profile_step.py 128 1 > gpurun_out/ncu_step128.log 2>&1
tail -n 3 gpurun_out/plain_step128.log
