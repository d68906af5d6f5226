#!/bin/bash
# round 2, GPU call 35: early t_full test issued inside the posting path
set -u
mkdir -p gpurun_out
{
for rep in 1 2 3; do
for n in x64 pe; do
  timeout 300 python tools/variant_case.py tools/bin/libsmb_$n.so 100 4
done
done
} > gpurun_out/r2c35_variants.log 2>&1
cat gpurun_out/r2c35_variants.log
