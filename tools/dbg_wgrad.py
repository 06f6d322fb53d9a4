"""Bring-up check of the tcgen05 grouped weight-gradient kernel against torch.nn.grad.conv2d_weight."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import hdmoe_b200
from hdmoe_b200 import _lib as L
lib = L.lib()
import os
if os.environ.get('WGLIB'):
    lib = C.CDLL(os.environ['WGLIB'])
    lib.hdmoe_gconv_wgrad.restype = C.c_int
    lib.hdmoe_gconv_wgrad.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
dev = "cuda"
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())

def run(R, H, W, Cin, Cout, ks, counts, time_it=False):
    gen = torch.Generator().manual_seed(R + H + Cin + Cout)
    E = len(ks)
    row_e = sum(([e] * c for e, c in enumerate(counts)), [])
    n_rows = len(row_e)
    row_e += [-1] * (R - n_rows)
    x = torch.randn(R, H, W, Cin, generator=gen).to(torch.bfloat16)
    dy = torch.randn(R, H, W, Cout, generator=gen).to(torch.bfloat16)
    wrow, tot = [], 0
    for k in ks:
        wrow.append(tot); tot += k * k * Cout
    xd, dyd = x.to(dev), dy.to(dev)
    dW = torch.zeros(tot, Cin, dtype=torch.float32, device=dev)
    re_d = torch.tensor(row_e, dtype=torch.int32, device=dev); nr_d = torch.tensor([n_rows], dtype=torch.int32, device=dev)
    ks_h = (C.c_int32 * E)(*ks); wr_h = (C.c_int32 * E)(*wrow)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    FN = lib.hdmoe_gconv_wgrad
    call = lambda: L.check(FN(p(xd), p(dyd), p(dW), R, H, W, Cin, Cout, tot, p(re_d), p(nr_d), E, ks_h, wr_h, st), "wgrad")
    call()
    torch.cuda.synchronize()
    got = dW.cpu()
    worst = 0.0
    lo = 0
    for e, c in enumerate(counts):
        if c == 0: continue
        k = ks[e]
        xin = x[lo:lo + c].float().permute(0, 3, 1, 2)
        g = dy[lo:lo + c].float().permute(0, 3, 1, 2)
        ref = torch.nn.grad.conv2d_weight(xin, (Cout, Cin, k, k), g, padding=(k - 1) // 2)   # [Cout,Cin,k,k]
        ref_t = ref.permute(2, 3, 0, 1).reshape(k * k * Cout, Cin)
        blk = got[wrow[e]:wrow[e] + k * k * Cout]
        err = (blk - ref_t).abs().max().item() / (ref_t.abs().max().item() + 1e-6)
        worst = max(worst, err)
        lo += c
    msg = f"R={R} {H}x{W} Cin={Cin} Cout={Cout} ks={ks} counts={counts}: max rel err {worst:.5f}"
    if time_it:
        for _ in range(2): dW.zero_(); call()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): call()
        b.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        fl = sum(2.0 * c * H * W * Cout * Cin * k * k for c, k in zip(counts, ks))
        msg += f"   {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s"
        if os.environ.get('WGLIB'):
            import numpy as np
            buf = (C.c_longlong * (148 * 16))()
            lib.hdmoe_wg_trace_read.argtypes = [C.c_void_p]
            lib.hdmoe_wg_trace_read(buf)
            a = np.array(buf[:], dtype=np.int64).reshape(148, 16)[:, :7]
            names = ["wait_item", "decode", "wait_tma", "issue", "flush/acc_wait", "n_items", "n_flush"]
            msg += "\n    issuer-0 cycles per CTA (mean / max): " + "  ".join(f"{n} {a[:, i].mean():.0f}/{a[:, i].max()}" for i, n in enumerate(names))
            msg += f"\n    total {a[:, :5].sum(1).mean():.0f} / {a[:, :5].sum(1).max()}"
    print(msg, flush=True)
    return worst

bad = 0
for a in [dict(R=2, H=16, W=16, Cin=64, Cout=64, ks=[1], counts=[2]),
          dict(R=2, H=16, W=16, Cin=64, Cout=64, ks=[3], counts=[2]),
          dict(R=6, H=32, W=32, Cin=32, Cout=32, ks=[3, 3, 5, 5], counts=[1, 2, 0, 2]),
          dict(R=5, H=16, W=16, Cin=128, Cout=64, ks=[3, 5], counts=[2, 3]),
          dict(R=5, H=16, W=16, Cin=96, Cout=64, ks=[3, 5], counts=[2, 3]),
          dict(R=4, H=32, W=32, Cin=96, Cout=32, ks=[5, 3], counts=[2, 2]),
          dict(R=4, H=32, W=32, Cin=64, Cout=32, ks=[3, 5], counts=[2, 1]),
          dict(R=40, H=32, W=32, Cin=64, Cout=64, ks=[3, 3, 5, 5], counts=[5, 10, 12, 13]),
          dict(R=256, H=32, W=32, Cin=64, Cout=64, ks=[3, 3, 5, 5], counts=[36, 48, 75, 97], time_it=True),
          dict(R=256, H=16, W=16, Cin=64, Cout=64, ks=[3, 3, 5, 5], counts=[36, 48, 75, 97], time_it=True),
          dict(R=256, H=32, W=32, Cin=32, Cout=32, ks=[3, 3, 5, 5], counts=[36, 48, 75, 97], time_it=True),
          dict(R=256, H=16, W=16, Cin=128, Cout=64, ks=[3, 3, 5, 5], counts=[36, 48, 75, 97], time_it=True),
          dict(R=256, H=32, W=32, Cin=96, Cout=32, ks=[3, 3, 5, 5], counts=[36, 48, 75, 97], time_it=True),
          dict(R=7, H=24, W=20, Cin=32, Cout=64, ks=[7, 1, 3], counts=[3, 2, 2]),
          dict(R=3, H=64, W=64, Cin=64, Cout=32, ks=[5, 3], counts=[2, 1]),
          dict(R=5, H=64, W=64, Cin=32, Cout=32, ks=[3, 5], counts=[2, 3]),
          dict(R=3, H=12, W=10, Cin=32, Cout=32, ks=[3], counts=[3])]:
    if run(**a) > 0.02: bad += 1
print("BAD" if bad else "ALL OK", bad)
