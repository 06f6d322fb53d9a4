"""torch.profiler breakdown of one eager train step (kernel-level CUDA time; guidance only, ncu is the evidence)."""
import sys, collections, re, torch
sys.path.insert(0, '.')
import bench, hdmoe_b200
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
hdmoe_b200.set_expert_dtype(torch.bfloat16)
r = bench.TrainRunner(1, 32, 256, 0, 1, dev, "dp", use_graph=False, warmup=0)     # bench.py's step, eager
step = lambda: r.eager_step(r.dev_batch)
for _ in range(3): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
tot = collections.Counter(); cnt = collections.Counter()
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        n = re.sub(r'<.*', '', ev.name); n = re.sub(r'\(.*', '', n)[:70]
        tot[n] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time; cnt[n] += 1
T = sum(tot.values())
print("CUDA kernels: %d launches, %.2f ms device time" % (sum(cnt.values()), T / 1e3))
for k, v in tot.most_common(34):
    print(f"{v/1e3:8.3f} ms {100*v/T:5.1f}% n={cnt[k]:5d}  {k}")
print("\n--- by ATen op (self device time) ---")
ka = prof.key_averages(group_by_input_shape=True)
rows = []
for e in ka:
    t = getattr(e, "self_device_time_total", None)
    if t is None: t = e.self_cuda_time_total
    if t > 0: rows.append((t, e.count, e.key, str(e.input_shapes)[:90]))
rows.sort(reverse=True)
for t, c, k, sh in rows[:int(sys.argv[1]) if len(sys.argv) > 1 else 45]:
    print(f"{t/1e3:8.3f} ms n={c:4d} {k:38s} {sh}")
