#!/bin/bash
# round-2 evidence pass, second half (call 26 produced the ncu table; its 137 MB report exceeded the 64 MiB pull limit)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_glue.py tests/test_gpu_optim.py -m gpu -q --no-header -x > gpurun_out/r2_pytest_glue.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_glue.log
t0=$(date +%s)
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$? wall=$(( $(date +%s) - t0 ))s"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
timeout 600 python tools/prof_step.py 140 > gpurun_out/r2_prof_step.txt 2>&1; echo "prof rc=$?"
timeout 600 python tools/timeline.py 500 > gpurun_out/r2_timeline.txt 2>&1; echo "timeline rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_train_step.csv \
   python bench.py --steps 3 --warmup 3 --profile-step --no-graph --no-sampler > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"; wc -l gpurun_out/r2_launches_train_step.csv
gzip -f gpurun_out/r2_launches_train_step.csv
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r2_gconv2_gwgrad2_full \
   --kernel-name 'regex:^(gconv2|gwgrad2)' -c 2 python tools/ncu_targets.py > gpurun_out/r2_ncu_full.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/r2_gconv2_gwgrad2_full.ncu-rep
python tools/ncu_summary.py gpurun_out/r2_gconv2_gwgrad2_full.ncu-rep gpurun_out/r2_gconv2_gwgrad2_table.md; echo "summary rc=$?"
ncu -i gpurun_out/r2_gconv2_gwgrad2_full.ncu-rep --page details > gpurun_out/r2_ncu_details_gconv2_gwgrad2.txt 2>&1
find gpurun_out -size +30M -print -delete
du -sh gpurun_out
head -60 gpurun_out/r2_prof_step.txt | cut -c1-150
