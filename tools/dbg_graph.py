import sys, traceback, torch
sys.path.insert(0, '.')
import bench, hdmoe_b200
from hdmoe_b200.utils import EDM_LOSS
from hdmoe_b200.train_step import GraphedTrainStep
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
hdmoe_b200.set_expert_dtype(torch.bfloat16)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = bench.build_model(1, dev); model.train()
crit = EDM_LOSS(**bench.LOSS)
params = list(model.parameters())
opt = torch.optim.AdamW(params, lr=5e-4, fused=True, capturable=True)
b = {k: v.to(dev) for k, v in bench.synth_batch(B, 32, 0, dev).items()}
def step(b):
    out = model(x=b["x"], sigma=b["sigma"], text_emb=b["text"], Unet_router_mask=b["um"], Vit_router_mask=b["vm"], zeta=2.0, return_log_var=True)
    loss = crit(b["sigma"], b["x0"], b["sigma"], out)
    opt.zero_grad(set_to_none=True)
    loss["loss"].backward()
    torch.nn.utils.clip_grad_norm_(params, 1.0)
    opt.step()
    return loss["loss"]
try:
    g = GraphedTrainStep(step, b).capture()
    print("captured"); l = [float(g().item()) for _ in range(3)]; print(l)
except Exception:
    traceback.print_exc()
