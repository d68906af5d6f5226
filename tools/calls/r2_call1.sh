#!/bin/bash
# round 2, GPU call 1: parity of the alternating-group epilogue + A/B against the round-1 layout + trace
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2c1_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2c1_parity.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2c1_parity.log
tail -3 gpurun_out/r2c1_parity.log
{
for lib in scanner_colmap_b200/libsmb.so tools/bin/libsmb_classic.so; do
  for n in 20 100; do
    timeout 300 python tools/variant_case.py $lib $n 4
    SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py $lib $n 4
  done
done
} > gpurun_out/r2c1_variants.log 2>&1
cat gpurun_out/r2c1_variants.log
timeout 300 python tools/trace_case.py 20 > gpurun_out/r2c1_trace.log 2>&1
grep -E "EPI|INS|TOPS" gpurun_out/r2c1_trace.log | head -40
