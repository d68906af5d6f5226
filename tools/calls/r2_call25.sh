#!/bin/bash
# round 2, GPU call 25: TS-mode functional probe (A operand from tensor memory) + shared-memory latency probes under MMA load
set -u
mkdir -p gpurun_out
timeout 120 tools/bin/ts_probe > gpurun_out/r2c25_ts_probe.log 2>&1; echo "rc=$?" >> gpurun_out/r2c25_ts_probe.log
cat gpurun_out/r2c25_ts_probe.log
timeout 300 tools/bin/smem_port > gpurun_out/r2c25_smem_port.log 2>&1; echo "rc=$?" >> gpurun_out/r2c25_smem_port.log
grep -E "probe|rc=" gpurun_out/r2c25_smem_port.log
