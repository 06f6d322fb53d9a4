#!/bin/bash
# refresh of the ncu --set full capture of gconv2 / gwgrad2 after the gwgrad2 rework, per-shape table, cycle trace
mkdir -p gpurun_out
timeout 300 python tools/ncu_targets.py > /dev/null 2>&1; echo "targets plain rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_gconv2_gwgrad2_full \
   --kernel-name 'regex:^(gconv2|gwgrad2)' -c 2 python tools/ncu_targets.py > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/r2_gconv2_gwgrad2_full.ncu-rep
python tools/ncu_summary.py gpurun_out/r2_gconv2_gwgrad2_full.ncu-rep gpurun_out/r2_gconv2_gwgrad2_table.md; cat gpurun_out/r2_gconv2_gwgrad2_table.md
ncu -i gpurun_out/r2_gconv2_gwgrad2_full.ncu-rep --page details > gpurun_out/r2_ncu_details_gconv2_gwgrad2.txt 2>&1
timeout 300 python tools/perf_shapes.py > gpurun_out/r2_gconv2_shape_table.json 2> gpurun_out/r2_gconv2_shape_table.txt; echo "shapes rc=$?"
WGLIB=tools/libwg2trace.so timeout 300 python tools/dbg_wgrad.py time > gpurun_out/r2_gwgrad2_cycle_trace.txt 2>&1; echo "trace rc=$?"
find gpurun_out -size +30M -print -delete
