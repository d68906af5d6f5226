#!/bin/bash
# round 2, GPU call 40: fast first-lane posting + one tcgen05.ld.x128 (-DSMB_FAST_POST -DSMB_LD_X128): A/B, then parity
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do
for n in x64 fpx; do
  timeout 100 python tools/variant_case.py tools/bin/libsmb_$n.so 100 3
done
done
} > gpurun_out/r2c40_variants.log 2>&1
cat gpurun_out/r2c40_variants.log
SMB_LIB=$PWD/tools/bin/libsmb_fpx.so timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not dp4a and not rejects and not engines_agree" > gpurun_out/r2c40_parity_fpx.log 2>&1; tail -3 gpurun_out/r2c40_parity_fpx.log
