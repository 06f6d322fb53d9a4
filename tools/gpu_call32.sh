#!/bin/bash
# N GPUs (arg 1): default bench through torchrun; config C: DP graph, EP (peer transport, graph), EP NCCL eager, DP eager
N=${1:-4}
mkdir -p gpurun_out
t0=$(date +%s)
timeout -k 10 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench n$N rc=$? wall=$(( $(date +%s) - t0 ))s"
python - $N <<'PY'
import json, sys
N=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2_bench_n{N}.json').read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ('value','ms_per_step','n_gpus')}, d['e2e'])
    print('sampler', json.dumps(d.get('sampler'))[:300])
    c=d.get('config_c') or {}
    for k,v in c.items():
        print(k, json.dumps(v)[:700])
except Exception as e:
    print("no line", e)
PY
grep -v "Warning\|warn\|^$\|run_backward\|OMP_NUM\|\*\*\*\*\|Producer process" gpurun_out/r2_bench_n$N.err | tail -25
