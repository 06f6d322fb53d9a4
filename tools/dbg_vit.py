"""Fused ViT block kernel vs the Vit_block module (fp32, eval): forward (and backward when available)."""
import sys, ctypes as C, torch
sys.path.insert(0, '.')
import hdmoe_b200
from hdmoe_b200 import model_components as mc, _lib as L
torch.manual_seed(0)
dev = "cuda"
E, Ss = 4, [64, 16, 16, 4]
blocks = [mc.Vit_block(num_heads=8, num_groups=4, num_channels=32, seq_ln=S, emb_dim=32, time_dim=64).to(dev).eval() for S in Ss]
lnf = [torch.nn.LayerNorm(32).to(dev) for _ in Ss]
with torch.no_grad():
    for b, ln in zip(blocks, lnf):
        for n, p in list(b.named_parameters()) + list(ln.named_parameters()):
            if "rel_pos_bias" in n or "bias" in n: p.copy_(torch.randn_like(p) * 0.3)
            elif "weight" in n and p.ndim == 1: p.copy_(1 + 0.2 * torch.randn_like(p))
R = 37
row_e = torch.randint(0, E, (R,), device=dev, dtype=torch.int32).sort().values
row_e[-1] = -1
tok = torch.randn(R, 64, 32, device=dev)
time = torch.randn(R, 64, device=dev)
def pack(b, ln):
    t = b.TMSA
    mods = [b.linear1, t.q_proj, t.k_proj, t.v_proj, t.out_proj, t.q_time, t.k_time, t.v_time, b.linear2, b.linear3]
    w = torch.cat([m.prepared_weight(1.0, torch.float32).flatten() for m in mods])
    a = torch.cat([b.GN.weight, b.GN.bias, b.norm1.weight, b.norm1.bias, b.norm2.weight, b.norm2.bias, ln.weight, ln.bias,
                   t.rel_pos_bias.flatten()])
    return w, a
ws, as_ = zip(*[pack(b, ln) for b, ln in zip(blocks, lnf)])
w_off, a_off, o1, o2 = [], [], 0, 0
for w, a in zip(ws, as_):
    w_off.append(o1); a_off.append(o2); o1 += w.numel(); o2 += a.numel()
w_all, a_all = torch.cat(ws).contiguous(), torch.cat(as_).contiguous()
I64, I32 = C.c_int64 * E, C.c_int32 * E
for final_ln in (0, 1):
    out = torch.empty_like(tok)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    L.check(L.lib().hdmoe_vit_block_fwd(p(tok), p(time), p(row_e), p(w_all), p(a_all), I64(*w_off), I64(*a_off), I32(*Ss), E, R,
                                        final_ln, p(out), st), "vit_block_fwd")
    torch.cuda.synchronize()
    worst = 0.0
    for r in range(R):
        e = int(row_e[r])
        if e < 0:
            assert float(out[r].abs().max()) == 0
            continue
        S = Ss[e]
        with torch.no_grad():
            ref = blocks[e](tok[r:r + 1, :S], time_embedding=time[r:r + 1])
            if final_ln: ref = lnf[e](ref)
        err = float((out[r, :S] - ref[0]).norm() / ref.norm())
        worst = max(worst, err)
        assert float(out[r, S:].abs().max()) == 0 if S < 64 else True
    print("final_ln", final_ln, "worst rel err", worst)
