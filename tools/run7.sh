mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log; tail -n 15 gpurun_out/pytest_gpu.log
python bench.py --steps 6 --warmup 3 > gpurun_out/bench_default.log 2>&1; tail -c 2200 gpurun_out/bench_default.log
