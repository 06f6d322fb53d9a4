#!/bin/bash
# gwgrad2 with kernel rows stacked in N: parity, cycle accounting, per-shape table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gconv.py -q --no-header -x 2>&1 | tail -4
WGLIB=tools/libwg2trace.so timeout 300 python tools/dbg_wgrad.py time > gpurun_out/c37_wgrad_trace.txt 2>&1; echo "wgrad trace rc=$?"
tail -22 gpurun_out/c37_wgrad_trace.txt
timeout 300 python tools/perf_shapes.py --no-cudnn > gpurun_out/c37_shapes.json 2> gpurun_out/c37_shapes.txt; echo "shapes rc=$?"
head -14 gpurun_out/c37_shapes.txt | cut -c60-110; grep -A8 fwd_dgrad_frac gpurun_out/c37_shapes.txt
