#!/bin/bash
# round 2, GPU call 11: op tests with GPU verification, 16-epilogue-warp variant, compute-sanitizer racecheck
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_op_gpu.py tests/test_two_view_gpu.py -x -q -m gpu > gpurun_out/r2c11_op_tests.log 2>&1; echo "op tests rc=$?" >> gpurun_out/r2c11_op_tests.log
tail -8 gpurun_out/r2c11_op_tests.log
{
for lib in tools/bin/libsmb_epi16.so; do
  timeout 300 python tools/variant_case.py $lib 20 4
  SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py $lib 20 4
  timeout 300 python tools/variant_case.py $lib 100 4
done
} > gpurun_out/r2c11_variants.log 2>&1
cat gpurun_out/r2c11_variants.log | tail -4
timeout 120 python tools/sanitize_case.py 20 4096 > gpurun_out/r2c11_sanitize_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 50 python tools/sanitize_case.py 20 4096 > gpurun_out/r2_racecheck.log 2>&1
echo "racecheck rc=$?"; tail -15 gpurun_out/r2_racecheck.log | cut -c1-200
