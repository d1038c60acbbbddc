mkdir -p gpurun_out
CRFR_WGRAD_STREAM=0 python tools/profile_step.py 128 2 > gpurun_out/plain_step128.log 2>&1 && CRFR_WGRAD_STREAM=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python tools/profile_step.py 128 1 > gpurun_out/ncu_step128.log 2>&1
tail -n 3 gpurun_out/plain_step128.log
