"""Fused ViT-expert path (vit_fused.py / csrc/vit_block.cu) vs the composite torch path through one MoE layer:
output, input gradients and every parameter gradient (fp32 expert dtype)."""
import sys, copy, torch
sys.path.insert(0, '.')
import hdmoe_b200
from hdmoe_b200 import model_components as mc, _denoiser as D, vit_fused
torch.manual_seed(0)
dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
hdmoe_b200.set_expert_dtype(torch.float32)
E, patches = 4, [4, 8, 8, 16]
def make():
    torch.manual_seed(1)
    ex = torch.nn.ModuleList([mc.Vit_expert(num_heads=8, num_groups=4, in_channels=32, seq_ln=(32 // p) ** 2, emb_dim=32, num_blocks=4,
                                            patch_size=p, time_dim=64, text_dim=768) for p in patches]).to(dev)
    with torch.no_grad():
        for n, p in ex.named_parameters():
            if "rel_pos_bias" in n or "pos_emb" in n or n.endswith("bias"): p.copy_(torch.randn_like(p) * 0.3)
            elif n.endswith("weight") and p.ndim == 1: p.copy_(1 + 0.2 * torch.randn_like(p))
    return ex
for train in (False, True):
    B = 41
    gen = torch.Generator().manual_seed(5)
    x0 = torch.randn(B, 32, 32, 32, generator=gen).to(dev)
    t0 = torch.randn(B, 64, generator=gen).to(dev)
    tx0 = torch.randn(B, 77, 768, generator=gen).to(dev)
    idx = torch.randint(0, E, (B,), generator=gen)
    wr = torch.zeros(B, E).scatter_(1, idx[:, None], 1.0).to(dev)
    gy = torch.randn(B, 32, 32, 32, generator=gen).to(dev)
    res = {}
    for fused in (False, True):
        vit_fused.set_fused_vit(fused)
        ex = make()
        ex.train(train)
        x, t, tx = (v.clone().requires_grad_(True) for v in (x0, t0, tx0))
        out = D.router_to_unet_experts(x, ex, wr, t, tx, top_k=1)
        out.backward(gy)
        torch.cuda.synchronize()
        res[fused] = dict(out=out.detach(), dx=x.grad, dt=t.grad, dtx=tx.grad, **{"p." + n: p.grad for n, p in ex.named_parameters()},
                          **{"w." + n: p.detach().clone() for n, p in ex.named_parameters() if n.endswith("weights")})
    worst, worst_k = 0.0, None
    for k, a in res[False].items():
        b = res[True][k]
        if a is None or b is None:
            if (a is None) != (b is None):
                nz = float((a if b is None else b).abs().max())
                print(f"  {k}: one side has no gradient (other max abs {nz:.3e})")
            continue
        den = float(a.norm())
        err = float((a - b).norm()) / (den + 1e-12) if den > 0 else float(b.norm())
        if err > worst: worst, worst_k = err, k
        if err > 2e-4: print(f"  MISMATCH {k}: rel {err:.3e} (ref norm {den:.3e})")
    print(f"train={train}: {len(res[False])} tensors compared, worst rel-L2 {worst:.3e} at {worst_k}")
