#!/bin/bash
# round 2, GPU call 16: host uploads waited for inside the score kernel (one launch for a call whose images are still arriving)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c16_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c16_gpu_tests.log
tail -6 gpurun_out/r2c16_gpu_tests.log
timeout 300 python tools/verify_case.py > gpurun_out/r2c16_verify_case.log 2>&1; tail -2 gpurun_out/r2c16_verify_case.log
timeout 300 python tools/e2e_chunks.py > gpurun_out/r2c16_e2e_chunks.log 2>&1; cat gpurun_out/r2c16_e2e_chunks.log
timeout 600 python tools/op_bench.py 100 8192 10 > gpurun_out/r2c16_op_bench.log 2>&1; cat gpurun_out/r2c16_op_bench.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/r2c16_bench_n1.json 2> gpurun_out/r2c16_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c16_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
for k in ('value','ms_per_step','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'], d['parity'])
PY
