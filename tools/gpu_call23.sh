#!/bin/bash
# gconv2: producer-decoded tile geometry, runtime ring depth (64 KB where it fits)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gconv.py -q --no-header -x 2>&1 | tail -3
timeout 300 python tools/perf_shapes.py 2 --no-cudnn > gpurun_out/c23_shapes_new.json 2> gpurun_out/c23_shapes_new.txt; echo "shapes rc=$?"
for cfg in "64 64 32" "64 64 16" "32 32 32" "128 64 16"; do
  set -- $cfg
  CIN=$1 COUT=$2 GVER=2 timeout 120 python tools/trace_gconv2.py $3 > gpurun_out/c23_trace_g2_$1_$2_$3.txt 2>&1
done
head -9 gpurun_out/c23_shapes_new.txt; grep -A6 fwd_dgrad_frac gpurun_out/c23_shapes_new.txt
head -8 gpurun_out/c23_trace_g2_32_32_32.txt; head -6 gpurun_out/c23_trace_g2_64_64_32.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-sampler > gpurun_out/c23_bench.log 2> gpurun_out/c23_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c23_bench.log').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','ms_per_step_isolated','gpu_launches')}, d['e2e'])
PY
