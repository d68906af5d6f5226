#!/bin/bash
# round 2, GPU call 34: full GPU suite + smoke + default bench on the x64 / plain-try_wait tree, then the ncu launch
# list and one full capture of the score kernel
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2c34_gpu_tests.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/r2c34_gpu_tests.log
tail -4 gpurun_out/r2c34_gpu_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c34_smoke.log 2>&1; tail -1 gpurun_out/r2c34_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c34_bench_n1.json 2> gpurun_out/r2c34_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c34_bench_n1.json').read().splitlines() if l.startswith('{')][-1])
for k in ('value','ms_per_step','gpu_launches'): print(k, d.get(k))
print('e2e', d['e2e']['value'], d['parity']['pairs_checked'], d['parity']['ok'])
print('roofline', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['frac_of_spec_4500'])
for k in ('strong','ragged','exhaustive'): print(k, d[k]['pairs_per_s'], d[k]['top_per_s'])
print(d['cpu_baseline'])
PY
CMD="python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline"
$CMD > gpurun_out/r2c34_bench_plain.json 2> gpurun_out/r2c34_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv $CMD > gpurun_out/r2c34_ncu1.log 2>&1
echo "ncu list rc=$?"
$CMD > gpurun_out/r2c34_bench_plain2.json 2>> gpurun_out/r2c34_bench_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:score_tcgen05 -s 4 -c 1 -o gpurun_out/r2b_prof_score $CMD > gpurun_out/r2c34_ncu2.log 2>&1
echo "ncu score rc=$?"
ls -la gpurun_out/r2b_prof_* gpurun_out/r2b_launches.csv
