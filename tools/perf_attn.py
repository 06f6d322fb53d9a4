"""Time the d=4 attention kernels (CUDA events) on the trunk shapes of the bench workload."""
import sys, torch
sys.path.insert(0, '.')
import hdmoe_b200
from hdmoe_b200 import ops
B, H = 256, 8
for Sq, Sk in ((1024, 1024), (1024, 77)):
    q, k, v = (torch.randn(B, s, H * 4, device="cuda", requires_grad=True) for s in (Sq, Sk, Sk))
    gy = torch.randn(B, Sq, H * 4, device="cuda")
    for impl, tf32 in (("cc", False), ("tc", False), ("tc", True)):
        ops.set_attention_impl(impl); torch.backends.cuda.matmul.allow_tf32 = tf32
        for _ in range(2):
            o = ops.attention_d4(q, k, v, H, 0.5); o.backward(gy)
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        for _ in range(5): o = ops.attention_d4(q, k, v, H, 0.5)
        e[1].record()
        for _ in range(5): o.backward(gy, retain_graph=True)
        e[2].record(); torch.cuda.synchronize()
        pairs = B * H * Sq * Sk
        f, b = e[0].elapsed_time(e[1]) / 5, e[1].elapsed_time(e[2]) / 5
        print(f"Sq={Sq} Sk={Sk} {impl} tf32={tf32}: fwd {f:.3f} ms ({pairs/f/1e6:.1f} Gpair/s)  bwd {b:.3f} ms")
