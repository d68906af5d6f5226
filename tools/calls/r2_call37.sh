#!/bin/bash
# round 2, GPU call 37: length of the posting path in the untraced kernel (two clock reads inside the path only)
set -u
mkdir -p gpurun_out
timeout 300 python tools/variant_case.py tools/bin/libsmb_postclk.so 20 1 > gpurun_out/r2c37_postclk.log 2>&1
cat gpurun_out/r2c37_postclk.log | cut -c1-200
