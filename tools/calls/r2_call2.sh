#!/bin/bash
# round 2, GPU call 2: barrier wait policy of the alternating epilogue (sleeping try_wait vs polling) + timeline trace
set -u
mkdir -p gpurun_out
{
for lib in tools/bin/libsmb_poll.so tools/bin/libsmb_hint100.so tools/bin/libsmb_sleep32.so tools/bin/libsmb_classic_poll.so; do
  timeout 300 python tools/variant_case.py $lib 20 4
  SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py $lib 20 4
done
timeout 300 python tools/variant_case.py tools/bin/libsmb_poll.so 100 4
} > gpurun_out/r2c2_variants.log 2>&1
cat gpurun_out/r2c2_variants.log
SMB_TRACE_LIB=libsmb_trace.so timeout 300 python tools/trace_case.py 20 > gpurun_out/r2c2_trace_sleep.log 2>&1
SMB_TRACE_LIB=libsmb_poll_trace.so timeout 300 python tools/trace_case.py 20 > gpurun_out/r2c2_trace_poll.log 2>&1
for f in gpurun_out/r2c2_trace_sleep.log gpurun_out/r2c2_trace_poll.log; do echo "== $f"; grep -E "^tile|EPI  0|EPI  8|INS 0|TOPS" $f | head -34; done
