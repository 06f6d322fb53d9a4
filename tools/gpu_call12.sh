#!/bin/bash
# 2-GPU call: EP NCCL parity test, DP + extras bench at N=2 (sampler split, config C DP vs EP)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q --no-header -s 2>&1 | tail -15 > gpurun_out/c12_t_multi.log; echo "multi rc=${PIPESTATUS[0]}"
tail -n 8 gpurun_out/c12_t_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c12_bench_n2.log 2> gpurun_out/c12_bench_n2.err; echo "bench n2 rc=$?"
grep -v "Warning\|warn\|run_backward\|^$" gpurun_out/c12_bench_n2.err | tail -25
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c12_bench_n2.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
    print(json.dumps(d['sampler'])[:800]); print(json.dumps(d['config_c'])[:1200])
except Exception as e: print('no line', e)
PY
