#!/bin/bash
# round 2, GPU call 13 (8 GPUs): the default bench under torchrun at N=8 and N=4
set -u
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2c13_topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2c13_bench_n8.json 2> gpurun_out/r2c13_bench_n8.err; echo "bench n8 rc=$?"
tail -3 gpurun_out/r2c13_bench_n8.err | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2c13_bench_n4.json 2> gpurun_out/r2c13_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2c13_bench_n8.json','gpurun_out/r2c13_bench_n4.json'):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f)
        for k in ('value','ms_per_step','n_gpus','gpu_launches'): print(' ',k, d.get(k))
        print('  e2e', d['e2e']['value'])
        print('  roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['other_kernels_ms_per_step'])
        for k in ('parity','strong','ragged','exhaustive'):
            if d.get(k): print(' ',k, json.dumps(d.get(k))[:520])
    except Exception as e:
        print(f, "no json", e)
PY
