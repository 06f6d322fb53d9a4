#!/bin/bash
# gconv2: tile decode under the MMAs, parallel barrier set-up, early scheduler re-arm; gwgrad2 cycle accounting
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gconv.py -q --no-header -x 2>&1 | tail -3
timeout 300 python tools/perf_shapes.py 2 --no-cudnn > gpurun_out/c22_shapes_new.json 2> gpurun_out/c22_shapes_new.txt; echo "shapes rc=$?"
for cfg in "64 64 32" "64 64 16" "32 32 32" "128 64 16"; do
  set -- $cfg
  CIN=$1 COUT=$2 GVER=2 timeout 120 python tools/trace_gconv2.py $3 > gpurun_out/c22_trace_g2_$1_$2_$3.txt 2>&1
done
WGLIB=tools/libwg2trace.so timeout 300 python tools/dbg_wgrad.py time > gpurun_out/c22_wgrad_trace.txt 2>&1; echo "wgrad trace rc=$?"
head -16 gpurun_out/c22_shapes_new.txt; grep -A6 fwd_dgrad_frac gpurun_out/c22_shapes_new.txt
tail -30 gpurun_out/c22_wgrad_trace.txt
