#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons only its own group) and keeps the logs
# under gpurun_out/.  Usage on the GPU box: bash tools/gpu_checks.sh [quick]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import crfr_b200; crfr_b200.build()" > gpurun_out/build.log 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log; }
run diag_fwd python tools/tc_diag.py fwd
run diag_dgrad python tools/tc_diag.py dgrad
run diag_wgrad python tools/tc_diag.py wgrad
run diag_match python tools/tc_diag.py match
run kernels python -m pytest tests/test_kernels_gpu.py -q -m gpu -x --timeout 300
run fsrnet_direct python -m pytest tests/test_fsrnet_gpu.py -q -m gpu -k "direct or rejects" --timeout 500
run tc_fwd python -m pytest tests/test_tc_gpu.py -q -m gpu -k "conv_fwd" --timeout 300
run tc_dgrad python -m pytest tests/test_tc_gpu.py -q -m gpu -k "conv_dgrad" --timeout 300
run tc_wgrad python -m pytest tests/test_tc_gpu.py -q -m gpu -k "conv_wgrad" --timeout 300
run tc_misc python -m pytest tests/test_tc_gpu.py -q -m gpu -k "views or full_size" --timeout 300
run tc_match python -m pytest tests/test_tc_gpu.py -q -m gpu -k "matcher" --timeout 300
run fsrnet_auto python -m pytest tests/test_fsrnet_gpu.py -q -m gpu -k "not direct" --timeout 500
