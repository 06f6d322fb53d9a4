"""Bring-up check of the tcgen05 grouped conv kernel against torch conv2d (GPU, fp32 on bf16-rounded data)."""
import ctypes as C
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
import hdmoe_b200
from hdmoe_b200 import _lib as L

lib = L.lib()
FN = lib.hdmoe_gconv2_fwd
dev = "cuda"


def run(R, H, W, Cin, Cout, ks, counts, scale=False, act=0, res=False, cin_real=None):
    gen = torch.Generator().manual_seed(R * 131 + H + Cin + Cout)
    E = len(ks)
    cin_real = cin_real or Cin
    assert sum(counts) <= R
    row_e = []
    for e, c in enumerate(counts):
        row_e += [e] * c
    n_rows = len(row_e)
    row_e += [-1] * (R - n_rows)
    x = torch.randn(R, cin_real, H, W, generator=gen)
    ws = [torch.randn(Cout, cin_real, k, k, generator=gen) / (cin_real * k * k) ** 0.5 for k in ks]
    xb = x.to(torch.bfloat16)
    wb = [w.to(torch.bfloat16) for w in ws]
    # NHWC, channel padded
    xn = torch.zeros(R, H, W, Cin, dtype=torch.bfloat16)
    xn[..., :cin_real] = xb.permute(0, 2, 3, 1)
    wt_rows, wrow = [], []
    tot = 0
    for w, k in zip(wb, ks):
        t = torch.zeros(k * k, Cout, Cin, dtype=torch.bfloat16)
        t[..., :cin_real] = w.permute(2, 3, 0, 1).reshape(k * k, Cout, cin_real)
        wrow.append(tot)
        tot += k * k * Cout
        wt_rows.append(t.reshape(-1, Cin))
    wt = torch.cat(wt_rows).contiguous()
    sc = (torch.rand(R, Cout, generator=gen) + 0.5) if scale else None
    rs = torch.randn(R, H, W, Cout, generator=gen).to(torch.bfloat16) if res else None
    xd, wd = xn.to(dev), wt.to(dev)
    y = torch.full((R, H, W, Cout), float("nan"), dtype=torch.bfloat16, device=dev)
    re_d = torch.tensor(row_e, dtype=torch.int32, device=dev)
    nr_d = torch.tensor([n_rows], dtype=torch.int32, device=dev)
    ks_h = (C.c_int32 * E)(*ks)
    wr_h = (C.c_int32 * E)(*wrow)
    scd = sc.to(dev) if scale else None
    rsd = rs.to(dev) if res else None
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    rc = FN(p(xd), p(wd), p(y), R, H, W, Cin, Cout, tot, p(re_d), p(nr_d), E, ks_h, wr_h, p(scd), act,
                             p(rsd), 0.6, 0.8, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(rc, "gconv_fwd")
    torch.cuda.synchronize()
    yh = y.float().cpu()
    worst = 0.0
    for r in range(n_rows):
        e = row_e[r]
        k = ks[e]
        ref = F.conv2d(xb[r:r + 1].float(), wb[e].float(), padding=(k - 1) // 2)[0].permute(1, 2, 0)
        if scale:
            ref = ref * sc[r]
        if act == 1:
            ref = F.silu(ref) / 0.596
        if res:
            ref = 0.6 * rs[r].float() + 0.8 * ref
        err = (yh[r] - ref).abs().max().item() / (ref.abs().max().item() + 1e-6)
        worst = max(worst, err)
    untouched = bool(torch.isnan(yh[n_rows:]).all()) if n_rows < R else True
    print(f"R={R} {H}x{W} Cin={Cin}({cin_real}) Cout={Cout} ks={ks} counts={counts} scale={scale} act={act} res={res}: "
          f"max rel err {worst:.4f} tail_untouched={untouched}", flush=True)
    return worst


bad = 0
for args in [
    dict(R=2, H=16, W=16, Cin=64, Cout=64, ks=[1], counts=[2]),
    dict(R=2, H=16, W=16, Cin=64, Cout=64, ks=[3], counts=[2]),
    dict(R=3, H=32, W=32, Cin=32, Cout=32, ks=[3], counts=[3]),
    dict(R=6, H=32, W=32, Cin=32, Cout=32, ks=[3, 3, 5, 5], counts=[1, 2, 0, 2]),
    dict(R=5, H=16, W=16, Cin=128, Cout=64, ks=[3, 5], counts=[2, 3]),
    dict(R=5, H=16, W=16, Cin=96, Cout=64, ks=[3, 5], counts=[2, 3]),
    dict(R=4, H=32, W=32, Cin=96, Cout=32, ks=[5, 3], counts=[2, 2]),
    dict(R=4, H=32, W=32, Cin=64, Cout=32, ks=[3, 5], counts=[2, 1], cin_real=33),
    dict(R=4, H=32, W=32, Cin=64, Cout=64, ks=[3, 5], counts=[2, 2], scale=True, act=1),
    dict(R=4, H=32, W=32, Cin=32, Cout=32, ks=[3, 5], counts=[1, 3], res=True),
    dict(R=3, H=64, W=64, Cin=32, Cout=128, ks=[3], counts=[3]),
    dict(R=300, H=32, W=32, Cin=64, Cout=64, ks=[3, 3, 5, 5], counts=[40, 60, 90, 100]),
    dict(R=3, H=64, W=64, Cin=64, Cout=64, ks=[5, 3], counts=[2, 1], scale=True, res=True),
    dict(R=5, H=20, W=12, Cin=32, Cout=96, ks=[7, 1, 3], counts=[2, 1, 2], act=1),
    dict(R=4, H=9, W=40, Cin=64, Cout=128, ks=[5, 3], counts=[1, 2], res=True),
    dict(R=300, H=16, W=16, Cin=64, Cout=64, ks=[3, 3, 5, 5], counts=[40, 60, 90, 100], scale=True, act=1, res=True),
]:
    if run(**args) > 0.02:
        bad += 1
print("BAD" if bad else "ALL OK", bad)
