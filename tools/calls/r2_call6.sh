#!/bin/bash
# round 2, GPU call 6: op shim (shared context, validation, staging, begin/wait), overhead breakdown, op benchmark
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_op_gpu.py -x -q -m gpu > gpurun_out/r2c6_op_tests.log 2>&1; echo "op tests rc=$?" >> gpurun_out/r2c6_op_tests.log
tail -12 gpurun_out/r2c6_op_tests.log
timeout 300 python tools/overhead_case.py 100 > gpurun_out/r2c6_overhead.log 2>&1; cat gpurun_out/r2c6_overhead.log
timeout 600 python tools/op_bench.py 100 8192 10 > gpurun_out/r2c6_op_bench.log 2>&1; cat gpurun_out/r2c6_op_bench.log
