#!/bin/bash
# round 2, GPU call 32: thin survivor records + insert warps recomputing scores with dp4a (-DSMB_RECOMPUTE): parity, A/B
set -u
mkdir -p gpurun_out
SMB_LIB=$PWD/tools/bin/libsmb_rc.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not dp4a and not rejects and not engines_agree" > gpurun_out/r2c32_parity_rc.log 2>&1; tail -3 gpurun_out/r2c32_parity_rc.log
{
for rep in 1 2; do
for n in x64 rc; do
  timeout 300 python tools/variant_case.py tools/bin/libsmb_$n.so 100 4
done
done
SMB_DEBUG_FLAGS=8 timeout 300 python tools/variant_case.py tools/bin/libsmb_rc.so 100 4
SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py tools/bin/libsmb_rc.so 100 4
timeout 300 python tools/variant_case.py tools/bin/libsmb_rc.so 20 4
timeout 300 python tools/variant_case.py tools/bin/libsmb_x64.so 20 4
} > gpurun_out/r2c32_variants.log 2>&1
cat gpurun_out/r2c32_variants.log
