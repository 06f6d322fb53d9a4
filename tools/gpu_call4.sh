#!/bin/bash
mkdir -p gpurun_out
for H in 32 16; do
GVER=3 timeout 120 python tools/trace_gconv2.py $H > gpurun_out/c4_trace_g3_$H.txt 2>&1
GVER=2 timeout 120 python tools/trace_gconv2.py $H > gpurun_out/c4_trace_g2_$H.txt 2>&1
done
GVER=3 CIN=32 COUT=32 timeout 120 python tools/trace_gconv2.py 32 > gpurun_out/c4_trace_g3_32_c32.txt 2>&1
GVER=2 CIN=32 COUT=32 timeout 120 python tools/trace_gconv2.py 32 > gpurun_out/c4_trace_g2_32_c32.txt 2>&1
timeout 600 python tools/dbg_numerics.py > gpurun_out/c4_numerics.txt 2>&1
cat gpurun_out/c4_trace_g3_32.txt gpurun_out/c4_numerics.txt
