#!/bin/bash
# final single-GPU pass of round 2: GPU suite, default bench line (archived), full dispatch sweep of BASELINE configs[4]
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu.log
t0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<'PY' > gpurun_out/r2_dispatch_full_sweep.json 2> gpurun_out/r2_dispatch_full_sweep.err
import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device("cuda", 0)
peaks = bench.load_peaks(); flush = bench.L2Flusher(dev)
d = bench.dispatch_sweep(dev, peaks, flush, full=True)
print(json.dumps(d))
PY
echo "full sweep rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','ms_per_step_isolated','gpu_launches')}, d['e2e'])
print('roof', d['roofline']['frac'], d['roofline']['aggregate']['gconv_fwd_dgrad_frac'], d['roofline']['aggregate']['gwgrad_frac'])
print('disp', d['dispatch']['reference_point_frac'], d['dispatch']['best_frac'])
s=json.load(open('gpurun_out/r2_dispatch_full_sweep.json'))
fr=[p['dispatch_combine_GBs']/s['peak'] for p in s['points']]
print('full sweep points', len(fr), 'min %.3f median %.3f max %.3f' % (min(fr), sorted(fr)[len(fr)//2], max(fr)), 'n>=0.7:', sum(f>=0.7 for f in fr))
PY
