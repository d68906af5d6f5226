#!/bin/bash
# round 2, GPU call 15: decide_kernel clears the accumulators it read (no per-call memset): full suite + bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c15_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c15_gpu_tests.log
tail -4 gpurun_out/r2c15_gpu_tests.log
timeout 300 python tools/overhead_case.py 100 > gpurun_out/r2c15_overhead.log 2>&1; cat gpurun_out/r2c15_overhead.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c15_bench_n1.json 2> gpurun_out/r2c15_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c15_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
for k in ('value','ms_per_step','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['other_kernels_ms_per_step'])
for k in ('parity','strong','ragged','exhaustive'):
    print(k, json.dumps(d.get(k))[:420])
PY
