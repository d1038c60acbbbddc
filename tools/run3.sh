mkdir -p gpurun_out
python tools/bench_conv.py 64 > gpurun_out/plain_conv64.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:rowconv_kernel -s 4 -c 1 -f -o gpurun_out/rowconv_r1b python tools/bench_conv.py 64 > gpurun_out/ncu_rowconv_b.log 2>&1
tail -n 5 gpurun_out/plain_conv64.log gpurun_out/ncu_rowconv_b.log
