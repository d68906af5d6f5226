#!/bin/bash
# round 2, GPU call 29: early (non-blocking) test of the next tile's t_full, several positions
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do
for n in t1 e0 e3 e3f0 e5 e6; do
  timeout 300 python tools/variant_case.py tools/bin/libsmb_$n.so 100 4
done
done
} > gpurun_out/r2c29_variants.log 2>&1
cat gpurun_out/r2c29_variants.log
