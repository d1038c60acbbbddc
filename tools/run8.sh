mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "norm" --timeout 600 > gpurun_out/t_norm.log 2>&1; echo "exit $?" >> gpurun_out/t_norm.log; tail -n 8 gpurun_out/t_norm.log
timeout 900 python -m pytest tests/test_fsrnet_gpu.py tests/test_resnet_gpu.py -q -m gpu --timeout 600 > gpurun_out/t_fsr.log 2>&1; echo "exit $?" >> gpurun_out/t_fsr.log; tail -n 5 gpurun_out/t_fsr.log
python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_nodz.log 2>&1; tail -c 1300 gpurun_out/bench_nodz.log | head -c 300; echo
