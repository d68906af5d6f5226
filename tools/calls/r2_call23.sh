#!/bin/bash
# round 2, GPU call 23 (4 GPUs): the default bench at N=4 on the final tree + the waves test
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "waves or sub_batches or overflow" > gpurun_out/r2c23_tests.log 2>&1; tail -3 gpurun_out/r2c23_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2c23_bench_n4.json 2> gpurun_out/r2c23_bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2c23_bench_n4.json',):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith('{')][-1])
        for k in ('value','ms_per_step','n_gpus','gpu_launches'): print(' ',k, d.get(k))
        print('  e2e', d['e2e']['value'])
        print('  roofline', d['roofline']['achieved'], d['roofline']['launch_ms'], d['roofline']['launches_per_step'], d['roofline']['other_kernels_ms_per_step'])
        print('  parity', d['parity']['pairs_checked'], d['parity']['halo_pairs_checked'], d['parity']['ok'])
        for k in ('strong','ragged','exhaustive'):
            if d.get(k): print(' ',k, d[k]['pairs_per_s'], d[k]['ms_per_step'], d[k]['rank_step_ms_max_over_mean'])
    except Exception as e:
        print(f, "no json", e)
PY
