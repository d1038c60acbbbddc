mkdir -p gpurun_out
python tools/profile_kd.py 256 3 > gpurun_out/kd256.log 2>&1; tail -n 3 gpurun_out/kd256.log
python tools/profile_kd.py 64 2 > gpurun_out/plain_kd64.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_kd64.csv python tools/profile_kd.py 64 1 > gpurun_out/ncu_kd64.log 2>&1
tail -n 3 gpurun_out/plain_kd64.log
