mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "deconv" --timeout 600 > gpurun_out/t_deconv.log 2>&1; echo "exit $?" >> gpurun_out/t_deconv.log; tail -n 12 gpurun_out/t_deconv.log
timeout 900 python -m pytest tests/test_fsrnet_gpu.py -q -m gpu --timeout 600 > gpurun_out/t_fsr.log 2>&1; echo "exit $?" >> gpurun_out/t_fsr.log; tail -n 12 gpurun_out/t_fsr.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_deconv.log 2>&1; tail -c 1300 gpurun_out/bench_deconv.log | head -c 300; echo
