#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q --no-header -k "dispatch or permute or moe_layer or full_config or golden" 2>&1 | tail -8
timeout 1500 python -m pytest tests/test_gpu_e2e.py -q --no-header 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 --no-sampler > gpurun_out/c11_bench.log 2> gpurun_out/c11_bench.err; echo "bench rc=$?"
grep -v Warning gpurun_out/c11_bench.err | grep -v "run_backward\|^$" | tail -8
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c11_bench.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
for p in d['dispatch']['points']: print(p)
print(d['dispatch']['best_frac'], d['dispatch']['reference_point_frac'])
PY
