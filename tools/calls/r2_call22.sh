#!/bin/bash
# round 2, GPU call 22 (8 GPUs): does the adaptive wave count hide the contended result delivery at N=8?
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2c22_bench_n8.json 2> gpurun_out/r2c22_bench_n8.err; echo "bench n8 rc=$?"
tail -3 gpurun_out/r2c22_bench_n8.err | cut -c1-300
SMB_WAVES=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 8 --steps 10 --warmup 3 --no-extra > gpurun_out/r2c22_bench_n8_waves1.json 2> gpurun_out/r2c22_bench_n8_waves1.err; echo "bench n8 waves1 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2c22_bench_n8.json','gpurun_out/r2c22_bench_n8_waves1.json'):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        print(f)
        for k in ('value','ms_per_step','n_gpus','gpu_launches'): print(' ',k, d.get(k))
        print('  e2e', d['e2e']['value'])
        print('  roofline', d['roofline']['achieved'], d['roofline']['launch_ms'], d['roofline']['launches_per_step'], d['roofline']['other_kernels_ms_per_step'])
        print('  parity', d['parity']['pairs_checked'], d['parity']['halo_pairs_checked'], d['parity']['ok'])
        for k in ('strong','ragged','exhaustive'):
            if d.get(k): print(' ',k, d[k]['pairs_per_s'], d[k]['ms_per_step'], d[k]['rank_step_ms_max_over_mean'])
    except Exception as e:
        print(f, "no json", e)
PY
