mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "edge_conv" --timeout 600 > gpurun_out/t_edge.log 2>&1; echo "exit $?" >> gpurun_out/t_edge.log; tail -n 8 gpurun_out/t_edge.log
timeout 1200 python -m pytest tests/test_resnet_gpu.py -q -m gpu --timeout 900 > gpurun_out/t_resnet.log 2>&1; echo "exit $?" >> gpurun_out/t_resnet.log; tail -n 40 gpurun_out/t_resnet.log
