#!/bin/bash
# round 2, GPU call 30: two tcgen05.ld.x64 per tile half (release as early as possible) vs four x32
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do
for n in t1 x64; do
  timeout 300 python tools/variant_case.py tools/bin/libsmb_$n.so 100 4
done
done
SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py tools/bin/libsmb_t1.so 100 4
SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py tools/bin/libsmb_x64.so 100 4
} > gpurun_out/r2c30_variants.log 2>&1
cat gpurun_out/r2c30_variants.log
