mkdir -p gpurun_out
python tools/bench_norm.py 128 > gpurun_out/plain_norm128.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"norm_act_fwd_kernel|norm_act_bwd_reduce_kernel|norm_act_bwd_apply_kernel|stats_partial_kernel" -s 8 -c 4 -f -o gpurun_out/norm_r1 python tools/bench_norm.py 128 > gpurun_out/ncu_norm.log 2>&1
tail -n 4 gpurun_out/plain_norm128.log; tail -n 3 gpurun_out/ncu_norm.log
