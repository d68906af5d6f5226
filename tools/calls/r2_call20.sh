#!/bin/bash
# round 2, GPU call 20: sub-batches software-pipelined over two streams (decide of k under score of k+1)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c20_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c20_gpu_tests.log
tail -4 gpurun_out/r2c20_gpu_tests.log
for w in 1 2 4 8; do echo "SMB_WAVES=$w"; SMB_WAVES=$w timeout 300 python tools/overhead_case.py 100 2>&1 | head -1; done > gpurun_out/r2c20_waves.log 2>&1; cat gpurun_out/r2c20_waves.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c20_bench_n1.json 2> gpurun_out/r2c20_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c20_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
for k in ('value','ms_per_step','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'], d['parity']['pairs_checked'], d['parity']['ok'])
print('roofline', d['roofline']['achieved'], d['roofline']['launch_ms'], d['roofline']['launches_per_step'], d['roofline']['other_kernels_ms_per_step'])
for k in ('strong','ragged','exhaustive'): print(k, d[k]['pairs_per_s'], d[k]['top_per_s'], d[k]['ms_per_step'])
PY
