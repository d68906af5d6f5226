#!/bin/bash
# round 2, GPU call 8: full GPU suite on the staged decide kernel, N=1 bench, ncu launch list + full captures
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c8_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c8_gpu_tests.log
tail -4 gpurun_out/r2c8_gpu_tests.log
CMD="python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline"
$CMD > gpurun_out/r2c8_bench_plain.json 2> gpurun_out/r2c8_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2c8_ncu1.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/r2c8_bench_plain2.json 2>> gpurun_out/r2c8_bench_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:score_tcgen05 -s 4 -c 1 -o gpurun_out/r2_prof_score $CMD > gpurun_out/r2c8_ncu2.log 2>&1
echo "ncu score rc=$?"
ncu --set full --clock-control none --import-source on -k regex:decide_kernel -s 4 -c 1 -o gpurun_out/r2_prof_decide $CMD > gpurun_out/r2c8_ncu3.log 2>&1
echo "ncu decide rc=$?"
ncu --set full --clock-control none --import-source on -k regex:runner_up -s 4 -c 1 -o gpurun_out/r2_prof_runner_up $CMD > gpurun_out/r2c8_ncu4.log 2>&1
echo "ncu runner_up rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c8_bench_plain.json').read().splitlines() if l.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['achieved'], d['roofline']['other_kernels_ms_per_step'], d['parity'])
PY
ls -la gpurun_out/r2_prof_* gpurun_out/r2_launches.csv
