#!/bin/bash
# round 2, GPU call 4: direct survivor handling in the epilogue warps (no mailbox) and scout barrier proxies
set -u
mkdir -p gpurun_out
SMB_LIB=$PWD/tools/bin/libsmb_scout.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2c4_parity_scout.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2c4_parity_scout.log
tail -3 gpurun_out/r2c4_parity_scout.log
SMB_LIB=$PWD/tools/bin/libsmb_direct.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2c4_parity_direct.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2c4_parity_direct.log
tail -3 gpurun_out/r2c4_parity_direct.log
{
for lib in tools/bin/libsmb_direct.so tools/bin/libsmb_scout.so tools/bin/libsmb_classic.so; do
  timeout 300 python tools/variant_case.py $lib 20 4
  SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py $lib 20 4
  timeout 300 python tools/variant_case.py $lib 100 4
done
} > gpurun_out/r2c4_variants.log 2>&1
cat gpurun_out/r2c4_variants.log
SMB_TRACE_LIB=libsmb_scout_trace.so timeout 300 python tools/trace_case.py 20 > gpurun_out/r2c4_trace_scout.log 2>&1
for f in gpurun_out/r2c4_trace_scout.log; do echo "== $f"; grep -E "^tile|EPI  0|EPI  4|TOPS" $f | tail -16 | cut -c1-260; done
