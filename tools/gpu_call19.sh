#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_router_trunk.py -q --no-header -x 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-sampler > gpurun_out/c19_bench.log 2> gpurun_out/c19_bench.err; echo "bench rc=$?"
grep -v Warning gpurun_out/c19_bench.err | grep -v "run_backward\|^$" | tail -8
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c19_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
PY
timeout 600 python tools/timeline.py 500 > gpurun_out/c19_timeline.txt 2>&1; echo "timeline rc=$?"
