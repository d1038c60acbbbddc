mkdir -p gpurun_out
bash tools/gpu_checks.sh quick > gpurun_out/gpu_checks.log 2>&1; grep -E "^===|passed|failed|error|exit" gpurun_out/gpu_checks.log | head -60
python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c128.log 2>&1; tail -c 1500 gpurun_out/bench_c128.log
