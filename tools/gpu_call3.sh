#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gconv.py -q -k "gconv3" > gpurun_out/c3_t_gconv3.log 2>&1; echo "gconv3 tests rc=$?"
tail -n 3 gpurun_out/c3_t_gconv3.log
timeout 300 python tools/perf_shapes.py 3 --no-cudnn > gpurun_out/c3_shapes_v3.json 2> gpurun_out/c3_shapes_v3.txt; echo "shapes v3 rc=$?"
cat gpurun_out/c3_shapes_v3.txt
timeout 1500 python -m pytest tests/test_gpu_e2e.py -q --no-header -rA -k sampler 2>&1 | tail -80 > gpurun_out/c3_t_e2e.log; echo "e2e rc=${PIPESTATUS[0]}"
grep -E "PASSED|FAILED|passed|failed" gpurun_out/c3_t_e2e.log
cat gpurun_out/parity_e2e.json | tail -60
