"""Device-time breakdown of the train step by model section (sum of kernel durations from torch.profiler over one
eager forward+backward of each section run standalone on the bench workload; guidance only)."""
import sys, collections, re, torch
sys.path.insert(0, '.')
import bench, hdmoe_b200
from hdmoe_b200 import _denoiser as D
from hdmoe_b200.utils import EDM_LOSS
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
hdmoe_b200.set_expert_dtype(torch.bfloat16)
D.set_trunk_weight_prep(False)        # sections run outside HDMOEM._forward
B = 256
model = bench.build_model(1, dev); model.train()
net = model.net
b = {k: v.to(dev) for k, v in bench.synth_batch(B, 32, 0, dev).items()}
top = int(sys.argv[1]) if len(sys.argv) > 1 else 12

def dev_time(fn, label):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    tot = collections.Counter(); cnt = collections.Counter()
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            n = re.sub(r'<.*', '', ev.name); n = re.sub(r'\(.*', '', n)[:60]
            tot[n] += ev.device_time; cnt[n] += 1
    T = sum(tot.values())
    print(f"== {label}: {T/1e3:.3f} ms, {sum(cnt.values())} launches")
    for k, v in tot.most_common(top):
        print(f"   {v/1e3:7.3f} ms n={cnt[k]:4d} {k}")
    return T

feats = torch.randn(B, 32, 32, 32, device=dev)
te = torch.randn(B, 64, device=dev)
text = b["text"]

def router(r, mask):
    def f():
        x = feats.clone().requires_grad_(True); t = te.clone().requires_grad_(True)
        w, p, raw = r(x=x, time_emb=t, zeta=2.0, mask=mask, noise=None)
        (w.sum() + p.sum() + raw.sum()).backward()
    return f
dev_time(router(net.Unet_router, b["um"]), "Unet_router fwd+bwd")
dev_time(router(net.vit_router, b["vm"]), "vit_router fwd+bwd")
with torch.no_grad():
    w_un, _, _ = net.Unet_router(x=feats, time_emb=te, zeta=2.0, mask=b["um"], noise=None)
    w_vit, _, _ = net.vit_router(x=feats, time_emb=te, zeta=2.0, mask=b["vm"], noise=None)
print("unet counts", (w_un > 0).sum(0).tolist(), "vit counts", (w_vit > 0).sum(0).tolist())
def moe(experts, w):
    def f():
        x = feats.clone().requires_grad_(True); t = te.clone().requires_grad_(True)
        ww = w.clone().requires_grad_(True)
        out = D.router_to_unet_experts(x, experts, ww, t, text, top_k=net.top_k)
        out.square().mean().backward()
    return f
dev_time(moe(net.Unet_experts, w_un), "U-Net MoE layer fwd+bwd")
dev_time(moe(net.VIT_experts, w_vit), "ViT MoE layer fwd+bwd")
def tail():
    out_u = feats.clone().requires_grad_(True); out_v = (feats * 0.5).clone().requires_grad_(True)
    uf = out_u.flatten(2).transpose(1, 2); vf = out_v.flatten(2).transpose(1, 2)
    a = net.cross_attn(query=uf, context=vf, gain_s=1.0, gain_t=1.0)
    bb = net.cross_attn_text(query=a, context=text, gain_s=1.0, gain_t=1.0)
    fin = a + net.alpha_txt * (bb - a)
    img = fin.transpose(1, 2).reshape(B, 32, 32, 32)
    import torch.nn.functional as F
    from hdmoe_b200 import model_internals as util
    g = net.gate2(util.mp_silu(net.gate1(util.mp_cat(out_u, img, dim=1))))
    g = F.softmax(g, dim=1)
    gated = g[:, 0:1] * out_u + g[:, 1:2] * img
    out = net.output_proj(util.mp_sum(out_u, gated, t=0.5))
    out.square().mean().backward()
dev_time(tail, "trunk tail (2x cross-attn, gate, output_proj) fwd+bwd")
D.set_trunk_weight_prep(True)
crit = EDM_LOSS(**bench.LOSS)
params = list(model.parameters())
opt = torch.optim.AdamW(params, lr=5e-4, fused=True)
def full():
    out = model(x=b["x"], sigma=b["sigma"], text_emb=b["text"], Unet_router_mask=b["um"], Vit_router_mask=b["vm"], zeta=2.0, return_log_var=True)
    loss = crit(b["sigma"], b["x0"], b["sigma"], out)
    opt.zero_grad(set_to_none=True)
    loss["loss"].backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt.step()
dev_time(full, "FULL train step")
def optim():
    torch.nn.utils.clip_grad_norm_(params, 1.0); opt.step()
dev_time(optim, "clip + AdamW")
