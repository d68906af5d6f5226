#!/bin/bash
# round 2, GPU call 41 (last of the budget): the whole GPU suite on the tree with the watermark test and EstimateMultiple
set -u
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_two_view_gpu.py tests/test_op_gpu.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2c41_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c41_gpu_tests.log
tail -25 gpurun_out/r2c41_gpu_tests.log | cut -c1-250
