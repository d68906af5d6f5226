#!/bin/bash
# round 2, GPU call 14: ballot-based posting variant, per-CTA busy time (load balance), e2e chunk sweep, ragged ncu capture
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "config1 or plan_reuse or halo_style or sub_batches" > gpurun_out/r2c14_tests.log 2>&1; tail -3 gpurun_out/r2c14_tests.log
SMB_LIB=$PWD/tools/bin/libsmb_postballot.so timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not dp4a and not rejects and not engines_agree" > gpurun_out/r2c14_parity_postballot.log 2>&1; tail -3 gpurun_out/r2c14_parity_postballot.log
{
for lib in tools/bin/libsmb_postballot.so scanner_colmap_b200/libsmb.so tools/bin/libsmb_postballot.so scanner_colmap_b200/libsmb.so; do
  timeout 300 python tools/variant_case.py $lib 20 4
  timeout 300 python tools/variant_case.py $lib 100 4
done
} > gpurun_out/r2c14_variants.log 2>&1
cat gpurun_out/r2c14_variants.log
timeout 300 python tools/e2e_chunks.py > gpurun_out/r2c14_e2e_chunks.log 2>&1; cat gpurun_out/r2c14_e2e_chunks.log
timeout 300 python tools/ragged_case.py 400 > gpurun_out/r2c14_ragged.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:score_tcgen05 -s 1 -c 1 -o gpurun_out/r2_prof_score_ragged python tools/ragged_case.py 400 > gpurun_out/r2c14_ncu_ragged.log 2>&1
echo "ncu ragged rc=$?"; cat gpurun_out/r2c14_ragged.log
