#!/bin/bash
# round 2, GPU call 19: full GPU suite + smoke + default bench on the final tree
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c19_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c19_gpu_tests.log
tail -4 gpurun_out/r2c19_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c19_smoke.log 2>&1; tail -1 gpurun_out/r2c19_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c19_bench_n1.json 2> gpurun_out/r2c19_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c19_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
for k in ('value','ms_per_step','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'], d['parity']['pairs_checked'], d['parity']['ok'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['frac_of_spec_4500'])
for k in ('strong','ragged','exhaustive'): print(k, d[k]['pairs_per_s'], d[k]['top_per_s'])
print(d['cpu_baseline'])
PY
