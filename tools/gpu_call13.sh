#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c13_bench_n2.log 2> gpurun_out/c13_bench_n2.err; echo "bench n2 rc=$?"
grep -v "Warning\|warn\|run_backward\|^$" gpurun_out/c13_bench_n2.err | tail -25
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/c13_bench_n2.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
    print(json.dumps(d['config_c'])[:1500])
except Exception as e: print('no line', e)
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --parallelism ep --no-sampler > gpurun_out/c13_bench_ep2.log 2> gpurun_out/c13_bench_ep2.err; echo "bench ep2 rc=$?"
grep -v "Warning\|warn\|run_backward\|^$" gpurun_out/c13_bench_ep2.err | tail -12
tail -c 900 gpurun_out/c13_bench_ep2.log | head -c 900
