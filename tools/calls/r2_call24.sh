#!/bin/bash
# round 2, GPU call 24: shared-memory port microbenchmark (does TMA / TMEM-read / mailbox traffic slow the tensor pipe?)
set -u
mkdir -p gpurun_out
timeout 300 tools/bin/smem_port > gpurun_out/r2c24_smem_port.log 2>&1; echo "rc=$?" >> gpurun_out/r2c24_smem_port.log
cat gpurun_out/r2c24_smem_port.log
