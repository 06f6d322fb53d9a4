#!/bin/bash
# profiles of round 2: ncu launch list of one timed (eager) step, ncu --set full of every hand-written kernel, shape tables
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --profile-step --no-graph --no-sampler > gpurun_out/r2_plain_step.log 2> gpurun_out/r2_plain_step.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_train_step.csv python bench.py --steps 3 --warmup 3 --profile-step --no-graph --no-sampler > gpurun_out/r2_ncu_step.log 2>&1; echo "ncu launch list rc=$?"
timeout 300 python tools/ncu_targets.py > gpurun_out/r2_targets_plain.log 2>&1; echo "targets plain rc=$?"
timeout 1500 ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r2_kernels -f python tools/ncu_targets.py > gpurun_out/r2_ncu_targets.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r2_kernels.ncu-rep
timeout 300 python tools/perf_shapes.py 2 > gpurun_out/r2_shapes_v2.json 2> gpurun_out/r2_shapes_v2.txt; echo "shapes v2 rc=$?"
timeout 300 python tools/perf_shapes.py 3 --no-cudnn > gpurun_out/r2_shapes_v3.json 2> gpurun_out/r2_shapes_v3.txt; echo "shapes v3 rc=$?"
GVER=2 timeout 120 python tools/trace_gconv2.py 32 > gpurun_out/r2_trace_g2_32.txt 2>&1
GVER=3 timeout 120 python tools/trace_gconv2.py 32 > gpurun_out/r2_trace_g3_32.txt 2>&1
tail -3 gpurun_out/r2_ncu_targets.log
