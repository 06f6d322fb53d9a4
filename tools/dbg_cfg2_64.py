"""BASELINE configs[2]: model_config2 at 4x64x64 -- one train step (fwd + EDM_LOSS + bwd) with the bf16 grouped expert
path vs the fp32 per-expert path on one GPU: loss, output and gradient agreement, and step time."""
import sys, time, torch
sys.path.insert(0, '.')
import bench, hdmoe_b200
from hdmoe_b200.utils import EDM_LOSS
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
full = dict(bench.FULL, IN_img_resolution=64)
torch.manual_seed(0)
model = hdmoe_b200.model_config2.preconditioned_HDMOEM(**full)
gen = torch.Generator().manual_seed(100)
with torch.no_grad():
    for p in model.parameters():
        if float(p.abs().max()) == 0: p.copy_(torch.randn(p.shape, generator=gen) * 0.3)
model.to(dev).train()
for mod in model.modules():
    if isinstance(mod, torch.nn.Dropout): mod.p = 0.0
    if hasattr(mod, "dropout") and not isinstance(mod, torch.nn.Dropout): mod.dropout = 0
print("params", sum(p.numel() for p in model.parameters()))
crit = EDM_LOSS(**bench.LOSS)
b = {k: v.to(dev) for k, v in bench.synth_batch(B, 64, 0, dev).items()}
noise = {"vit": torch.randn(B, 4, device=dev), "unet": torch.randn(B, 4, device=dev)}
state0 = {k: v.clone() for k, v in model.state_dict().items()}
res = {}
for mode, dt, grouped in (("bf16 grouped", torch.bfloat16, True), ("fp32 loop", torch.float32, False)):
    hdmoe_b200.set_expert_dtype(dt); hdmoe_b200.set_grouped_experts(grouped)
    model.load_state_dict(state0); model.zero_grad(set_to_none=True)
    def step():
        out = model(x=b["x"], sigma=b["sigma"], text_emb=b["text"], Unet_router_mask=b["um"], Vit_router_mask=b["vm"], zeta=2.0,
                    transition_point=-1.2, softness=1.6, return_log_var=True, noise=noise)
        loss = crit(b["sigma"], b["x0"], b["sigma"], out)["loss"]
        loss.backward()
        return out, loss
    out, loss = step()
    torch.cuda.synchronize()
    res[mode] = dict(out=out["denoised"].detach().float(), loss=float(loss),
                     g={n: p.grad.detach().float().clone() for n, p in model.named_parameters() if p.grad is not None})
    model.load_state_dict(state0); model.zero_grad(set_to_none=True)
    t0 = time.perf_counter()
    for _ in range(3):
        model.zero_grad(set_to_none=True); step()
    torch.cuda.synchronize()
    print(f"{mode}: loss {float(loss):.6f}  eager step {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms (B={B}, 64x64)")
a, c = res["bf16 grouped"], res["fp32 loop"]
rel = lambda x, y: float((x - y).norm() / (y.norm() + 1e-12))
print("denoised rel-L2 bf16 vs fp32:", rel(a["out"], c["out"]))
num = sum(float(((a["g"][n] - c["g"][n]) ** 2).sum()) for n in c["g"] if n in a["g"])
den = sum(float((c["g"][n] ** 2).sum()) for n in c["g"])
print("all-parameter gradient rel-L2:", (num / den) ** 0.5, " missing grads:", [n for n in c["g"] if n not in a["g"]][:5])
