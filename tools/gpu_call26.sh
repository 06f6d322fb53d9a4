#!/bin/bash
# round-2 evidence pass on the current code: GPU suite, default bench + reference arm, step profile / timeline,
# ncu launch list of one train step, ncu --set full of every hand-written kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu.log
t0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
t0=$(date +%s)
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$? wall=$(( $(date +%s) - t0 ))s"
timeout 600 python tools/prof_step.py 140 > gpurun_out/r2_prof_step.txt 2>&1; echo "prof rc=$?"
timeout 600 python tools/timeline.py 500 > gpurun_out/r2_timeline.txt 2>&1; echo "timeline rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_train_step.csv \
   python bench.py --steps 3 --warmup 3 --profile-step --no-graph --no-sampler > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r2_launches_train_step.csv
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_kernels \
   --kernel-name 'regex:^(analytic|attn_d4|axpby|cat_k|combine|gain_silu|gate_|gconv2|gn1_|gwgrad2|heun|lin32|mt_|nchw_|nhwc_|permute|pixnorm|plan_|precond|router_gate|scale2|scale_pair|scaling_router|silu_bwd|split_k|sqerr|swap_|vit_block|wprep)' \
   python tools/ncu_targets.py > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/r2_kernels.ncu-rep
python tools/ncu_summary.py gpurun_out/r2_kernels.ncu-rep gpurun_out/r2_kernels_ncu_table.md; echo "summary rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','ms_per_step_isolated','gpu_launches')}, d['e2e'])
print('sampler', json.dumps(d.get('sampler'))[:600]); print('config_c', json.dumps(d.get('config_c'))[:600]); print('ref_cuda', json.dumps(d.get('ref_cuda_eager'))[:400])
print('agg', d['roofline'].get('aggregate')); print('disp ref frac', d['dispatch']['reference_point_frac'], d['dispatch']['best_frac'])
PY
