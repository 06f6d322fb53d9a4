#!/bin/bash
# converged (uniform-register) MMA issue in gconv2 / gwgrad2: parity, per-shape A/B against the previous build, traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gconv.py -q --no-header -x 2>&1 | tail -4
for v in base new ring64; do
  if [ $v = new ]; then unset HDMOE_B200_LIB; else export HDMOE_B200_LIB=tools/variants/libhdmoe_$v.so; fi
  timeout 300 python tools/perf_shapes.py 2 --no-cudnn > gpurun_out/c21_shapes_$v.json 2> gpurun_out/c21_shapes_$v.txt; echo "shapes $v rc=$?"
done
unset HDMOE_B200_LIB
for cfg in "64 64 32" "64 64 16" "32 32 32" "64 32 32" "128 64 16"; do
  set -- $cfg
  CIN=$1 COUT=$2 GVER=2 timeout 120 python tools/trace_gconv2.py $3 > gpurun_out/c21_trace_g2_$1_$2_$3.txt 2>&1
done
cat gpurun_out/c21_shapes_new.txt | head -16
timeout 900 python bench.py --steps 10 --warmup 3 --no-sampler > gpurun_out/c21_bench.log 2> gpurun_out/c21_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c21_bench.log').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','ms_per_step_isolated','gpu_launches')}, d['e2e'])
PY
