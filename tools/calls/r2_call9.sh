#!/bin/bash
# round 2, GPU call 9: anti-phase sibling epilogue warps (SMB_EPI_STAGGER): parity, A/B, timeline
set -u
mkdir -p gpurun_out
SMB_LIB=$PWD/tools/bin/libsmb_stagger.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tcgen05 or config1 or full_size or overflow or ragged or adversarial" > gpurun_out/r2c9_parity_stagger.log 2>&1; echo "parity rc=$?" >> gpurun_out/r2c9_parity_stagger.log
tail -3 gpurun_out/r2c9_parity_stagger.log
{
for lib in tools/bin/libsmb_stagger.so scanner_colmap_b200/libsmb.so; do
  timeout 300 python tools/variant_case.py $lib 20 4
  SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py $lib 20 4
  timeout 300 python tools/variant_case.py $lib 100 4
done
} > gpurun_out/r2c9_variants.log 2>&1
cat gpurun_out/r2c9_variants.log
SMB_TRACE_LIB=libsmb_stagger_trace.so timeout 300 python tools/trace_case.py 20 > gpurun_out/r2c9_trace_stagger.log 2>&1
grep -E "^tile|EPI  0|EPI  1|TOPS" gpurun_out/r2c9_trace_stagger.log | tail -16 | cut -c1-200
