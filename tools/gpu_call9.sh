#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_optim.py tests/test_gpu_glue.py -q --no-header -rA 2>&1 | tail -60 > gpurun_out/c9_t_new.log; echo "new tests rc=${PIPESTATUS[0]}"
grep -E "PASSED|FAILED|passed|failed|Error|^E " gpurun_out/c9_t_new.log | head -40
timeout 1500 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_parity.py -q --no-header -x 2>&1 | tail -30 > gpurun_out/c9_t_e2e.log; echo "e2e+parity rc=${PIPESTATUS[0]}"
tail -n 12 gpurun_out/c9_t_e2e.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-sampler > gpurun_out/c9_bench.log 2> gpurun_out/c9_bench.err; echo "bench rc=$?"
grep -v Warning gpurun_out/c9_bench.err | grep -v "run_backward\|^$" | tail -12
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c9_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
PY
