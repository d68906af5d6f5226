#!/bin/bash
# round 2, GPU call 33: one tcgen05.ld.x128 per tile half, MMA thread spinning on t_empty, insert-warp idle sleep
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do
for n in x64 x128 mmaspin idle500 idle100; do
  timeout 300 python tools/variant_case.py tools/bin/libsmb_$n.so 100 4
done
done
} > gpurun_out/r2c33_variants.log 2>&1
cat gpurun_out/r2c33_variants.log
