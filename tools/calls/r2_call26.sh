#!/bin/bash
# round 2, GPU call 26: mbarrier raw layout + dependent latencies of try_wait / test_wait / plain LDS
set -u
mkdir -p gpurun_out
timeout 120 tools/bin/mbar_probe > gpurun_out/r2c26_mbar_probe.log 2>&1; echo "rc=$?" >> gpurun_out/r2c26_mbar_probe.log
cat gpurun_out/r2c26_mbar_probe.log
