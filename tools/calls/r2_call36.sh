#!/bin/bash
# round 2, GPU call 36 (2 GPUs): N=2 bench line on the final tree (two-stream waves, x64 loads) + op bench
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c36_bench_n2.json 2> gpurun_out/r2c36_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2c36_bench_n2.err | cut -c1-300
timeout 600 python tools/op_bench.py 100 8192 10 > gpurun_out/r2c36_op_bench.log 2>&1; cat gpurun_out/r2c36_op_bench.log
python - <<'PY'
import json
for f in ('gpurun_out/r2c36_bench_n2.json',):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        for k in ('value','ms_per_step','n_gpus','gpu_launches'): print(' ',k, d.get(k))
        print('  e2e', d['e2e']['value'])
        print('  roofline', d['roofline']['achieved'], d['roofline']['launch_ms'], d['roofline']['other_kernels_ms_per_step'])
        for k in ('parity','strong','ragged','exhaustive'):
            if d.get(k): print(' ',k, json.dumps(d.get(k))[:330])
    except Exception as e:
        print(f, "no json", e)
PY
