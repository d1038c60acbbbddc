mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?" >> gpurun_out/pytest_gpu.log; tail -n 15 gpurun_out/pytest_gpu.log
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_default.log 2>&1; tail -c 2600 gpurun_out/bench_default.log
