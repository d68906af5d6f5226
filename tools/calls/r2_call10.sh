#!/bin/bash
# round 2, GPU call 10: GPU two-view geometry verification (first kernel) + 16-epilogue-warp variant + timing
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_two_view_gpu.py -x -q -m gpu > gpurun_out/r2c10_tvg_tests.log 2>&1; echo "tvg tests rc=$?" >> gpurun_out/r2c10_tvg_tests.log
tail -25 gpurun_out/r2c10_tvg_tests.log
timeout 300 python tools/verify_case.py > gpurun_out/r2c10_verify_case.log 2>&1; tail -5 gpurun_out/r2c10_verify_case.log
{
for lib in tools/bin/libsmb_epi16.so scanner_colmap_b200/libsmb.so; do
  timeout 300 python tools/variant_case.py $lib 20 4
  SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py $lib 20 4
  timeout 300 python tools/variant_case.py $lib 100 4
done
} > gpurun_out/r2c10_variants.log 2>&1
cat gpurun_out/r2c10_variants.log
