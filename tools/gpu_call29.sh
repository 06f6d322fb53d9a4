#!/bin/bash
# 2 GPUs: the NCCL expert-parallel parity test, then the default bench at N = 2 (DP headline, sampler split over the ranks,
# config C: DP graph / EP eager / DP eager / exchange timed alone)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -x > gpurun_out/r2_pytest_multi.log 2>&1; echo "multi pytest rc=$?"; tail -3 gpurun_out/r2_pytest_multi.log
t0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
print('sampler', json.dumps(d.get('sampler'))[:500])
print('config_c', json.dumps(d.get('config_c'), indent=1)[:3000])
PY
grep -v "Warning\|warn\|^$" gpurun_out/r2_bench_n2.err | tail -15
