mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tc_gpu.py -q -m gpu -x -k "rowconv or conv_fwd or conv_dgrad or conv_wgrad or full_size or views" --timeout 300 > gpurun_out/tc_rowconv.log 2>&1; echo "exit $?" >> gpurun_out/tc_rowconv.log
tail -n 5 gpurun_out/tc_rowconv.log
timeout 300 python tools/bench_conv.py 32 > gpurun_out/bench_conv32.log 2>&1; tail -n 3 gpurun_out/bench_conv32.log
timeout 300 python tools/bench_conv.py 128 > gpurun_out/bench_conv128.log 2>&1; tail -n 3 gpurun_out/bench_conv128.log
