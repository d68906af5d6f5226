#!/bin/bash
# round 2, GPU call 39 (8 GPUs): the N=8 line on the final tree (headline + e2e + parity sample only)
set -u
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/r2c39_bench_n8.json 2> gpurun_out/r2c39_bench_n8.err; echo "bench n8 rc=$?"
tail -3 gpurun_out/r2c39_bench_n8.err | cut -c1-300
python - <<'PY'
import json
for f in ('gpurun_out/r2c39_bench_n8.json',):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f)
        for k in ('value','ms_per_step','n_gpus','gpu_launches'): print(' ',k, d.get(k))
        print('  e2e', d['e2e']['value'])
        print('  roofline', d['roofline']['achieved'], d['roofline']['launch_ms'], d['roofline']['launches_per_step'], d['roofline']['other_kernels_ms_per_step'])
        print('  parity', d['parity']['pairs_checked'], d['parity']['halo_pairs_checked'], d['parity']['ok'])
    except Exception as e:
        print(f, "no json", e)
PY
