#!/bin/bash
# round 2, GPU call 38: compact insert loop (instruction-cache footprint 30 KB -> 20 KB): A/B and posting-path length
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do
for n in x64 ci; do
  timeout 300 python tools/variant_case.py tools/bin/libsmb_$n.so 100 4
done
done
SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py tools/bin/libsmb_ci.so 100 4
timeout 300 python tools/variant_case.py tools/bin/libsmb_cipc.so 20 1 | grep -v "summed" | cut -c1-170
} > gpurun_out/r2c38_variants.log 2>&1
cat gpurun_out/r2c38_variants.log
