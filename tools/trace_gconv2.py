import ctypes as C, sys, torch, numpy as np
import os
VER = "2"
lib = C.CDLL(os.environ.get("G2LIB", "tools/libg%strace.so" % VER))
dev = "cuda"
H = int(sys.argv[1]) if len(sys.argv) > 1 else 16
counts, ks = [36, 48, 75, 97], [3, 3, 5, 5]
R = sum(counts); Cin = int(os.environ.get('CIN', 64)); Cout = int(os.environ.get('COUT', 64))
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
fn = getattr(lib, 'hdmoe_gconv%s_fwd' % VER)
fn.restype = C.c_int
fn.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_void_p]
for _ in range(3):
    rc = fn(p(x), p(wt), p(y), R, H, H, Cin, Cout, tot, p(re_d), p(nr_d), 4, ks_h, wr_h, None, 0, None, 0.0, 0.0, None)
    assert rc == 0
    torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(10):
    fn(p(x), p(wt), p(y), R, H, H, Cin, Cout, tot, p(re_d), p(nr_d), 4, ks_h, wr_h, None, 0, None, 0.0, 0.0, None)
ev[1].record(); torch.cuda.synchronize()
print("gconv", VER, "Cin", Cin, "Cout", Cout, "H", H, "us per launch (back-to-back, warm L2): %.1f" % (ev[0].elapsed_time(ev[1]) * 100))
buf = (C.c_longlong * (148 * 64))()
rd = getattr(lib, 'hdmoe_g%s_trace_read' % VER)
rd.argtypes = [C.c_void_p]
rd(buf)
a = np.array(buf[:], dtype=np.int64).reshape(148, 8, 8)
for b in (0, 147):
    print("CTA", b)
    t00 = a[b, 0, 0]
    for t in range(8):
        s = a[b, t, :6]
        if s[3] == 0: break
        print(f"  tile {t}: start +{s[0]-t00:7d}  wait_acc {s[1]-s[0]:6d}  wait_A {s[2]-s[1]:6d}  mma_issue {s[3]-s[2]:7d}  | mma_done-issue_end {s[4]-s[3]:7d}  epilogue {s[5]-s[4]:6d}")
# SM clock from the cycle counter vs the nanosecond global timer, and the wall-clock span of the launch
g_start = min(a[b, 0, 6] for b in range(148) if a[b, 0, 6] > 0)
g_end = max(a[b, t, 7] for b in range(148) for t in range(8))
print("launch span by globaltimer (first tile start -> last traced epilogue end): %.1f us" % ((g_end - g_start) / 1e3))
for b in (0, 73, 147):
    last = max(t for t in range(8) if a[b, t, 5] > 0)
    cyc = a[b, last, 5] - a[b, 0, 0]; ns = a[b, last, 7] - a[b, 0, 6]
    print(f"CTA {b}: {cyc} cycles in {ns} ns -> {cyc / ns:.3f} GHz; start offset {(a[b, 0, 6] - g_start) / 1e3:.1f} us, end offset {(a[b, last, 7] - g_start) / 1e3:.1f} us")
ends_ns = sorted((max(a[b, t, 7] for t in range(8)) - g_start) / 1e3 for b in range(148))
print("CTA end times (us): min %.1f median %.1f max %.1f" % (ends_ns[0], ends_ns[74], ends_ns[-1]))

if VER == "2" and hasattr(lib, "hdmoe_g2_span_read"):
    sp = (C.c_longlong * (148 * 2))()
    lib.hdmoe_g2_span_read.argtypes = [C.c_void_p]
    lib.hdmoe_g2_span_read(sp)
    sp = np.array(sp[:], dtype=np.int64).reshape(148, 2)
    k0, k1 = sp[:, 0].min(), sp[:, 1].max()
    first_tile = np.array([a[b, 0, 6] for b in range(148)])
    last_epi = np.array([max(a[b, t, 7] for t in range(8)) for b in range(148)])
    print("kernel span by globaltimer (first CTA entry -> last CTA exit): %.1f us" % ((k1 - k0) / 1e3))
    print("CTA entry spread: %.2f us; entry -> first tile start (prologue): median %.2f us max %.2f us" % (
        (sp[:, 0].max() - k0) / 1e3, np.median(first_tile - sp[:, 0]) / 1e3, (first_tile - sp[:, 0]).max() / 1e3))
    print("last traced epilogue end -> CTA exit (teardown): median %.2f us; CTA exit spread: min %.1f median %.1f max %.1f us" % (
        np.median(sp[:, 1] - last_epi) / 1e3, (sp[:, 1].min() - k0) / 1e3, np.median(sp[:, 1] - k0) / 1e3, (k1 - k0) / 1e3))
