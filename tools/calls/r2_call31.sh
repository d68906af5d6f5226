#!/bin/bash
# round 2, GPU call 31: where does survivor posting cost its time?  (timing ablations through SMB_DEBUG_FLAGS; results are wrong by design)
set -u
mkdir -p gpurun_out
{
for f in 0 8 16 24 32 40 4; do
  SMB_DEBUG_FLAGS=$f timeout 300 python tools/variant_case.py tools/bin/libsmb_x64.so 100 4
done
} > gpurun_out/r2c31_post_ablation.log 2>&1
cat gpurun_out/r2c31_post_ablation.log
