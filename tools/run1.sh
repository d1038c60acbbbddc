mkdir -p gpurun_out
./build_tmp/mma_bench > gpurun_out/mma_bench.log 2>&1
python tools/bench_conv.py 16 > gpurun_out/bench_conv16.log 2>&1
python tools/bench_conv.py 32 > gpurun_out/bench_conv32.log 2>&1
python tools/profile_step.py 16 3 > gpurun_out/step16.log 2>&1
python tools/profile_step.py 32 3 > gpurun_out/step32.log 2>&1
python tools/profile_step.py 64 3 > gpurun_out/step64.log 2>&1
python bench.py --chunk 32 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c32.log 2>&1
python tools/profile_step.py 32 1 > gpurun_out/plain_step32.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python tools/profile_step.py 32 1 > gpurun_out/ncu_step32.log 2>&1
tail -n 30 gpurun_out/mma_bench.log gpurun_out/bench_conv16.log gpurun_out/bench_conv32.log gpurun_out/step16.log gpurun_out/step32.log gpurun_out/step64.log
tail -c 1500 gpurun_out/bench_c32.log
