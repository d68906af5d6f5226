#!/bin/bash
# round 2, GPU call 27: FMA-pipe pre-filter (-DSMB_FMA_FILTER) and plain try_wait on t_full (-DSMB_FAST_TFULL): parity, A/B
set -u
mkdir -p gpurun_out
SMB_LIB=$PWD/tools/bin/libsmb_fmafast.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not dp4a and not rejects and not engines_agree" > gpurun_out/r2c27_parity_fmafast.log 2>&1; tail -3 gpurun_out/r2c27_parity_fmafast.log
{
for rep in 1 2; do
for lib in tools/bin/libsmb_classic.so tools/bin/libsmb_fma.so tools/bin/libsmb_fast.so tools/bin/libsmb_fmafast.so; do
  timeout 300 python tools/variant_case.py $lib 100 4
done
done
for lib in tools/bin/libsmb_classic.so tools/bin/libsmb_fmafast.so; do
  SMB_DEBUG_FLAGS=4 timeout 300 python tools/variant_case.py $lib 100 4
  timeout 300 python tools/variant_case.py $lib 20 4
done
} > gpurun_out/r2c27_variants.log 2>&1
cat gpurun_out/r2c27_variants.log
