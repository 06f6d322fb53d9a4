#!/bin/bash
# full GPU test-suite on the new gconv2 / gwgrad2 / combine; gconv3 (N-stacked taps, converged issue) per-shape table; dispatch sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -5
timeout 300 python tools/perf_shapes.py 3 --no-cudnn > gpurun_out/c25_shapes_v3.json 2> gpurun_out/c25_shapes_v3.txt; echo "shapes v3 rc=$?"
head -9 gpurun_out/c25_shapes_v3.txt; grep -A6 fwd_dgrad_frac gpurun_out/c25_shapes_v3.txt
GVER=3 timeout 120 python tools/trace_gconv2.py 32 > gpurun_out/c25_trace_g3_64_64_32.txt 2>&1
GVER=3 CIN=32 COUT=32 timeout 120 python tools/trace_gconv2.py 32 > gpurun_out/c25_trace_g3_32_32_32.txt 2>&1
head -12 gpurun_out/c25_trace_g3_64_64_32.txt; tail -5 gpurun_out/c25_trace_g3_64_64_32.txt
python - <<'PY' > gpurun_out/c25_dispatch.txt 2>&1
import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device("cuda", 0)
peaks = bench.load_peaks(); flush = bench.L2Flusher(dev)
d = bench.dispatch_sweep(dev, peaks, flush)
print(json.dumps(d, indent=1))
PY
grep -E '"T"|plan_us|permute_us|combine_us|dispatch_combine_GBs|reference_point' gpurun_out/c25_dispatch.txt | head -40
