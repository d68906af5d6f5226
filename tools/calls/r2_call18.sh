#!/bin/bash
# round 2, GPU call 18: cheaper planning (one lookup per image run, no plan compare after a layout change): suite + e2e sweep + bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c18_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c18_gpu_tests.log
tail -4 gpurun_out/r2c18_gpu_tests.log
timeout 300 python tools/e2e_chunks.py > gpurun_out/r2c18_e2e_chunks.log 2>&1; cat gpurun_out/r2c18_e2e_chunks.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c18_bench_n1.json 2> gpurun_out/r2c18_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c18_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
for k in ('value','ms_per_step','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'], d['parity'])
for k in ('strong','ragged','exhaustive'): print(k, d[k]['pairs_per_s'], d[k]['top_per_s'])
PY
