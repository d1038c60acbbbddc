mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -q -m gpu -k "matcher" --timeout 300 > gpurun_out/t_match.log 2>&1; echo "exit $?" >> gpurun_out/t_match.log; tail -n 6 gpurun_out/t_match.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "eval" --timeout 300 2>&1 | tail -n 3
python tools/bench_match.py 4096 262144 2>&1 | tail -n 1
python tools/bench_match.py 10000 1000000 2>&1 | tail -n 1
