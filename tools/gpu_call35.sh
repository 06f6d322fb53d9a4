#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu.log
t0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','ms_per_step_isolated','gpu_launches')}, {k:v for k,v in d['e2e'].items() if k!='d2h'})
print('roof', d['roofline']['frac'], d['roofline']['aggregate']['gconv_fwd_dgrad_frac'], d['roofline']['aggregate']['gwgrad_frac'])
print('disp', d['dispatch']['reference_point_frac'], d['dispatch']['best_frac'], [(p['T'], p['plan_us']) for p in d['dispatch']['points']])
PY
python -c "
import __graft_entry__ as g
g.smoke(); print('smoke ok')" 2>&1 | tail -2
