#!/bin/bash
# round 2, GPU call 7 (2 GPUs): the multi-GPU bench path after the host rewrite + both arms under torchrun
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "plan_reuse or halo_style or sub_batches or upload" > gpurun_out/r2c7_tests.log 2>&1; tail -3 gpurun_out/r2c7_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2c7_bench_n2.json 2> gpurun_out/r2c7_bench_n2.err; echo "bench n2 rc=$?"
tail -3 gpurun_out/r2c7_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-extra > gpurun_out/r2c7_bench_n1.json 2> gpurun_out/r2c7_bench_n1.err; echo "bench n1 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c7_ref_n2.json 2> gpurun_out/r2c7_ref_n2.err; echo "ref n2 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2c7_bench_n2.json','gpurun_out/r2c7_bench_n1.json','gpurun_out/r2c7_ref_n2.json'):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f)
        for k in ('value','ms_per_step','n_gpus','gpu_launches'): print(' ',k, d.get(k))
        print('  e2e', d['e2e']['value'])
        if 'roofline' in d: print('  roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['other_kernels_ms_per_step'])
        for k in ('parity','strong','ragged','exhaustive','cpu_baseline'): 
            if d.get(k): print(' ',k, json.dumps(d.get(k))[:420])
    except Exception as e:
        print(f, "no json", e)
PY
