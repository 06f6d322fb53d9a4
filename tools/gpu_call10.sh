#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py -q --no-header 2>&1 | tail -25
timeout 1500 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_parity.py tests/test_gpu_router_trunk.py -q --no-header 2>&1 | tail -30 > gpurun_out/c10_t.log; echo "e2e+parity rc=${PIPESTATUS[0]}"
tail -n 12 gpurun_out/c10_t.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-sampler > gpurun_out/c10_bench.log 2> gpurun_out/c10_bench.err; echo "bench rc=$?"
grep -v Warning gpurun_out/c10_bench.err | grep -v "run_backward\|^$" | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c10_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
PY
timeout 600 python tools/prof_step.py > gpurun_out/c10_prof_step.txt 2>&1; echo "prof rc=$?"
