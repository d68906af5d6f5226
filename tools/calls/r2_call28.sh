#!/bin/bash
# round 2, GPU call 28: wait-policy variants (plain try_wait polls before the hinted wait; busy polling)
set -u
mkdir -p gpurun_out
{
for rep in 1 2; do
for n in classic t1 t2 t4 all1 all2 busy; do
  timeout 300 python tools/variant_case.py tools/bin/libsmb_$n.so 100 4
done
done
} > gpurun_out/r2c28_variants.log 2>&1
cat gpurun_out/r2c28_variants.log
