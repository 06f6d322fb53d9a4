#!/bin/bash
# round-2 GPU call 1: first hardware run of gconv3, kernel-level parity tests, per-shape tables, regression of the suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 300 python tools/dbg_gconv.py v3 > gpurun_out/c1_dbg_gconv_v3.txt 2>&1; echo "dbg v3 rc=$?"
timeout 900 python -m pytest tests/test_gpu_gconv.py -q -k "not gconv3" > gpurun_out/c1_t_gconv2.log 2>&1; echo "gconv2 tests rc=$?"
timeout 900 python -m pytest tests/test_gpu_gconv.py -q -k "gconv3" > gpurun_out/c1_t_gconv3.log 2>&1; echo "gconv3 tests rc=$?"
timeout 300 python tools/perf_shapes.py 2 > gpurun_out/c1_shapes_v2.json 2> gpurun_out/c1_shapes_v2.txt; echo "shapes v2 rc=$?"
timeout 300 python tools/perf_shapes.py 3 --no-cudnn > gpurun_out/c1_shapes_v3.json 2> gpurun_out/c1_shapes_v3.txt; echo "shapes v3 rc=$?"
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_gconv.py > gpurun_out/c1_t_all.log 2>&1; echo "suite rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/c1_bench.log 2> gpurun_out/c1_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/c1_t_gconv2.log gpurun_out/c1_t_gconv3.log gpurun_out/c1_t_all.log
tail -5 gpurun_out/c1_dbg_gconv_v3.txt
