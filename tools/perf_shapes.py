"""Per-shape table of the grouped convolution kernels (forward, data gradient, weight gradient) over every U-Net expert
layer of SURVEY Appendix D at the bench routing: python tools/perf_shapes.py [--no-cudnn] -> JSON on stdout."""
import json
import sys
import torch
sys.path.insert(0, '.')
import bench
import hdmoe_b200  # noqa: F401
from hdmoe_b200 import ops

impl = 2
dev = torch.device("cuda", 0)
peaks = bench.load_peaks()
flush = bench.L2Flusher(dev)
t = bench.gconv_shape_table(dev, peaks, flush, iters=10, with_cudnn="--no-cudnn" not in sys.argv)
t["impl"] = impl
print(json.dumps(t))
for r in t["layers"]:
    sys.stderr.write("%4d->%-4d %2dx%-2d %-7s x%-2d  fwd %7.1f us %.3f | dgrad %7.1f us %.3f | wgrad %7.1f us %.3f | cudnn fwd %s bwd %s\n" % (
        r["cin"], r["cout"], r["hw"], r["hw"], r["k"], r["mult"], r["fwd_us"], r["fwd_frac"], r["dgrad_us"], r["dgrad_frac"],
        r["wgrad_us"], r["wgrad_frac"], r.get("cudnn_bf16_fwd_us"), r.get("cudnn_bf16_bwd_us")))
sys.stderr.write(json.dumps(t["aggregate"], indent=1) + "\n")
