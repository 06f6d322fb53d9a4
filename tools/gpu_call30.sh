#!/bin/bash
# 2 GPUs: peer-memory transport of the expert-parallel exchange: parity vs NCCL transport / local experts, then the
# default bench at N = 2 (config C: DP graph, EP peer graph, EP NCCL eager, DP eager, exchange timed alone)
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_gpu_multi.py -m gpu -q --no-header -x -s > gpurun_out/r2_pytest_multi.log 2>&1; echo "multi pytest rc=$?"; grep -E "EP vs local|passed|failed|Error|error" gpurun_out/r2_pytest_multi.log | tail -8
t0=$(date +%s)
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$? wall=$(( $(date +%s) - t0 ))s"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
    c=d.get('config_c') or {}
    for k,v in c.items():
        print(k, json.dumps(v)[:400])
except Exception as e:
    print("no line", e)
PY
grep -v "Warning\|warn\|^$\|run_backward" gpurun_out/r2_bench_n2.err | tail -25
