"""Times the grouped conv kernel on the U-Net expert layer shapes (B=256 routed 36/48/75/97, k=3,3,5,5)."""
import ctypes as C
import sys
import torch
sys.path.insert(0, '.')
import hdmoe_b200
from hdmoe_b200 import _lib as L
lib = L.lib()
FN = lib.hdmoe_gconv2_fwd
dev = "cuda"
counts = [36, 48, 75, 97]
ks = [3, 3, 5, 5]
R = sum(counts)
row_e = sum(([e] * c for e, c in enumerate(counts)), [])
re_d = torch.tensor(row_e, dtype=torch.int32, device=dev)
nr_d = torch.tensor([R], dtype=torch.int32, device=dev)
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tot_ms = tot_fl = 0
for (H, Cin, Cout, mult) in [(32, 64, 32, 1), (32, 32, 32, 8), (16, 32, 32, 2), (16, 64, 64, 11), (16, 128, 64, 2),
                             (16, 96, 64, 1), (32, 64, 64, 2), (32, 96, 32, 1), (32, 64, 32, 2)]:
    x = torch.randn(R, H, H, Cin, device=dev).to(torch.bfloat16)
    tot = 0
    wrow = []
    for k in ks:
        wrow.append(tot)
        tot += k * k * Cout
    wt = (torch.randn(tot, Cin, device=dev) / 30).to(torch.bfloat16)
    y = torch.empty(R, H, H, Cout, dtype=torch.bfloat16, device=dev)
    ks_h = (C.c_int32 * 4)(*ks)
    wr_h = (C.c_int32 * 4)(*wrow)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def call():
        L.check(FN(p(x), p(wt), p(y), R, H, H, Cin, Cout, tot, p(re_d), p(nr_d), 4, ks_h, wr_h, None, 0,
                                    None, 0.0, 0.0, st), "gconv")
    for _ in range(3):
        call()
    ms = 0
    for _ in range(10):
        flush.add_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); call(); b.record(); torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    ms /= 10
    fl = sum(2.0 * c * H * H * Cout * Cin * k * k for c, k in zip(counts, ks))
    print(f"{H}x{H} Cin={Cin} Cout={Cout}: {ms*1e3:8.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  (x{mult} per expert fwd)", flush=True)
    tot_ms += ms * mult
    tot_fl += fl * mult
print(f"U-Net expert conv layers, forward, B=256: {tot_ms:.3f} ms, {tot_fl/tot_ms/1e9:.1f} TFLOP/s aggregate")
