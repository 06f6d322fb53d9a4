"""One gconv launch shape, few iterations (for ncu)."""
import ctypes as C, sys, torch
sys.path.insert(0, '.')
import hdmoe_b200
from hdmoe_b200 import _lib as L
lib = L.lib()
FN = lib.hdmoe_gconv2_fwd if (len(sys.argv) > 1 and sys.argv[1] == 'v2') else lib.hdmoe_gconv_fwd
dev = "cuda"
counts, ks = [36, 48, 75, 97], [3, 3, 5, 5]
R = sum(counts); H = 32; Cin = 64; Cout = 64
row_e = sum(([e] * c for e, c in enumerate(counts)), [])
re_d = torch.tensor(row_e, dtype=torch.int32, device=dev); nr_d = torch.tensor([R], dtype=torch.int32, device=dev)
p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
x = torch.randn(R, H, H, Cin, device=dev).to(torch.bfloat16)
tot = 0; wrow = []
for k in ks:
    wrow.append(tot); tot += k * k * Cout
wt = (torch.randn(tot, Cin, device=dev) / 30).to(torch.bfloat16)
y = torch.empty(R, H, H, Cout, dtype=torch.bfloat16, device=dev)
ks_h = (C.c_int32 * 4)(*ks); wr_h = (C.c_int32 * 4)(*wrow)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(4):
    L.check(FN(p(x), p(wt), p(y), R, H, H, Cin, Cout, tot, p(re_d), p(nr_d), 4, ks_h, wr_h, None, 0, None, 0.0, 0.0, st), "gconv")
torch.cuda.synchronize()
print("done")
