#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -6
timeout 600 python - <<'PY' 2>&1 | tail -12
import sys, torch, json
sys.path.insert(0,'.')
import bench
dev=torch.device('cuda',0); peaks=bench.load_peaks(); flush=bench.L2Flusher(dev)
d=bench.dispatch_sweep(dev, peaks, flush)
for p in d['points']: print(p)
print(d['best_frac'], d['reference_point_frac'])
PY
