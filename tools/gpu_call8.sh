#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py -q --no-header -rA 2>&1 | tail -40 > gpurun_out/c8_t_optim.log; echo "optim rc=${PIPESTATUS[0]}"
grep -E "PASSED|FAILED|passed|failed|Error|error|^E " gpurun_out/c8_t_optim.log | head -30
timeout 900 python bench.py --steps 10 --warmup 3 --no-sampler > gpurun_out/c8_bench.log 2> gpurun_out/c8_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/c8_bench.err | grep -v Warning | tail -15
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c8_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
PY
