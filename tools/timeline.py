"""Kernel timeline of ONE replay of the captured train step (torch.profiler / CUPTI timestamps, per stream):
prints per-bin concurrency and the dominant kernels, to find the critical path.  Guidance only."""
import sys, collections, re, torch
sys.path.insert(0, '.')
import bench, hdmoe_b200
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
hdmoe_b200.set_expert_dtype(torch.bfloat16)
r = bench.TrainRunner(1, 32, 256, 0, 1, dev, "dp", use_graph=True, warmup=3)      # exactly bench.py's captured step
g = r.graphed
for _ in range(3): g(None)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    g(None); torch.cuda.synchronize()
evs = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r'<.*', '', ev.name); n = re.sub(r'\(.*', '', n); n = n.replace("void ", "").replace("at::native::", "")[:40]
        evs.append((ev.time_range.start, ev.time_range.end, getattr(ev, "device_resource_id", getattr(ev, "device_index", 0)), n))
evs.sort()
t0 = evs[0][0]; t1 = max(e[1] for e in evs)
print(f"replay span {(t1 - t0)/1e3:.3f} ms, {len(evs)} kernels, sum of durations {sum(e[1]-e[0] for e in evs)/1e3:.3f} ms")
binw = float(sys.argv[1]) if len(sys.argv) > 1 else 500.0     # us
nb = int((t1 - t0) / binw) + 1
for i in range(nb):
    lo, hi = t0 + i * binw, t0 + (i + 1) * binw
    act = [(min(e[1], hi) - max(e[0], lo), e) for e in evs if e[1] > lo and e[0] < hi]
    busy = sum(a for a, _ in act)
    # union coverage
    iv = sorted((max(e[0], lo), min(e[1], hi)) for _, e in act)
    cov = 0; cur = None
    for a, c in iv:
        if cur is None or a > cur[1]:
            if cur: cov += cur[1] - cur[0]
            cur = [a, c]
        else:
            cur[1] = max(cur[1], c)
    if cur: cov += cur[1] - cur[0]
    names = collections.Counter()
    for a, e in act: names[e[3]] += a
    streams = len({e[2] for _, e in act})
    top = ", ".join(f"{k}:{v/binw:.2f}" for k, v in names.most_common(4))
    print(f"{i*binw/1e3:6.2f} ms  cover {cov/binw:4.2f}  load {busy/binw:4.2f}  streams {streams}  n={len(act):3d} | {top}")
print("\n--- per stream ---")
by = collections.defaultdict(list)
for e in evs: by[e[2]].append(e)
for sid, lst in sorted(by.items(), key=lambda kv: kv[1][0][0]):
    busy = sum(e[1] - e[0] for e in lst)
    names = collections.Counter()
    for e in lst: names[e[3]] += e[1] - e[0]
    print(f"stream {sid}: n={len(lst)} first {(lst[0][0]-t0)/1e3:.2f} ms last {(max(e[1] for e in lst)-t0)/1e3:.2f} ms busy {busy/1e3:.2f} ms | " + ", ".join(f"{k}:{v/1e3:.2f}" for k, v in names.most_common(5)))
    # activity windows (gaps > 300 us split)
    wins = []; cur = [lst[0][0], lst[0][1], lst[0][1] - lst[0][0]]
    for e in lst[1:]:
        if e[0] - cur[1] > 300: wins.append(cur); cur = [e[0], e[1], 0.0]
        cur[1] = max(cur[1], e[1]); cur[2] += e[1] - e[0]
    wins.append(cur)
    print("    windows: " + "  ".join(f"[{(a-t0)/1e3:.2f}-{(b-t0)/1e3:.2f} busy {c/1e3:.2f}]" for a, b, c in wins))
print("\n--- totals by kernel name (device time inside the replay) ---")
tot = collections.Counter(); cnt = collections.Counter()
for e in evs: tot[e[3]] += e[1] - e[0]; cnt[e[3]] += 1
T = sum(tot.values())
for k, v in tot.most_common(45):
    print(f"{v/1e3:8.3f} ms {100*v/T:5.1f}% n={cnt[k]:5d}  {k}")
