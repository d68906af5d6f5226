#!/bin/bash
# round 2, GPU call 3: half-commit variant (two N=128 halves per tile, separate barriers): parity, A/B, timeline
set -u
mkdir -p gpurun_out
tools/bin/commit_latency > gpurun_out/r2c3_commit_latency.log 2>&1; cat gpurun_out/r2c3_commit_latency.log
SMB_LIB=$PWD/tools/bin/libsmb_half.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2c3_parity_half.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2c3_parity_half.log
tail -3 gpurun_out/r2c3_parity_half.log
{
for lib in tools/bin/libsmb_half.so tools/bin/libsmb_classic.so; do
  timeout 300 python tools/variant_case.py $lib 20 4
  SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py $lib 20 4
  timeout 300 python tools/variant_case.py $lib 100 4
done
} > gpurun_out/r2c3_variants.log 2>&1
cat gpurun_out/r2c3_variants.log
SMB_TRACE_LIB=libsmb_half_trace.so timeout 300 python tools/trace_case.py 20 > gpurun_out/r2c3_trace_half.log 2>&1
SMB_TRACE_LIB=libsmb_classic_trace.so timeout 300 python tools/trace_case.py 20 > gpurun_out/r2c3_trace_classic.log 2>&1
for f in gpurun_out/r2c3_trace_half.log gpurun_out/r2c3_trace_classic.log; do echo "== $f"; grep -E "^tile|EPI  0|EPI  4|INS 0|TOPS" $f | tail -17 | cut -c1-260; done
