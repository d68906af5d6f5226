#!/bin/bash
# round 2, GPU call 5: refactored host side (zero-copy results, begin/wait, plan reuse), full GPU suite, bench with extras
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c5_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c5_gpu_tests.log
tail -15 gpurun_out/r2c5_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c5_smoke.log 2>&1; tail -2 gpurun_out/r2c5_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c5_bench.json 2> gpurun_out/r2c5_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2c5_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2c5_bench.json').read().strip().splitlines()[-1])
    for k in ('value','ms_per_step','e2e','roofline','parity','strong','ragged','exhaustive','cpu_baseline','gpu_launches'):
        print(k, json.dumps(d.get(k))[:600])
except Exception as e:
    print("no json", e)
PY
