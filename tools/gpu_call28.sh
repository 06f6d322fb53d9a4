#!/bin/bash
# per-expert parallel plan scan + folded tail fill: full GPU suite, dispatch sweep (full), default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -x > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
python - <<'PY' > gpurun_out/r2_dispatch_sweep.json 2> gpurun_out/r2_dispatch_sweep.err
import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device("cuda", 0)
peaks = bench.load_peaks(); flush = bench.L2Flusher(dev)
d = bench.dispatch_sweep(dev, peaks, flush)
print(json.dumps(d, indent=1))
PY
echo "sweep rc=$?"
grep -E '"T"|plan_us|permute_us|combine_us|dispatch_combine_GBs|reference_point|best_frac' gpurun_out/r2_dispatch_sweep.json | head -48
