mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "instance_norm or batch_norm" --timeout 600 > gpurun_out/t_norm.log 2>&1; echo "exit $?" >> gpurun_out/t_norm.log; tail -n 25 gpurun_out/t_norm.log
timeout 300 python tools/bench_norm.py 128 > gpurun_out/bench_norm128.log 2>&1; tail -n 4 gpurun_out/bench_norm128.log
CRFR_FUSED_NORM_BWD=0 timeout 300 python tools/bench_norm.py 128 > gpurun_out/bench_norm128_old.log 2>&1; tail -n 4 gpurun_out/bench_norm128_old.log
timeout 900 python -m pytest tests/test_fsrnet_gpu.py -q -m gpu --timeout 600 > gpurun_out/t_fsr.log 2>&1; echo "exit $?" >> gpurun_out/t_fsr.log; tail -n 5 gpurun_out/t_fsr.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fused.log 2>&1; tail -c 900 gpurun_out/bench_fused.log | head -c 500
