#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_router_trunk.py -q --no-header -rA -x 2>&1 | tail -60 > gpurun_out/c5_t_trunk.log; echo "trunk rc=${PIPESTATUS[0]}"
grep -E "PASSED|FAILED|passed|failed|Error|error" gpurun_out/c5_t_trunk.log | head -30
timeout 900 python -m pytest tests/test_gpu_gconv.py -q -k "gwgrad" 2>&1 | tail -3
timeout 1500 python -m pytest tests/test_gpu_e2e.py -q --no-header -rA 2>&1 | tail -80 > gpurun_out/c5_t_e2e.log; echo "e2e rc=${PIPESTATUS[0]}"
grep -E "PASSED|FAILED|passed|failed" gpurun_out/c5_t_e2e.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 --no-sampler > gpurun_out/c5_bench.log 2> gpurun_out/c5_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/c5_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c5_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
PY
