"""Where does the bf16-mode error of the raw denoiser output come from?  D(x; sigma) at fixed sigmas under four numeric
configurations (fp32/bf16 experts x strict-fp32/TF32 trunk) against the fp32 CPU oracle."""
import sys
import torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import hdmoe_b200
from conftest import FULL
from oracle import hdmoe_oracle as O
from test_gpu_e2e import _model, rel_l2

model = _model(2, seed=1)
sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
B = 32
gen = torch.Generator().manual_seed(77)
noise = torch.randn(B, 4, 32, 32, generator=gen)
text = torch.randn(B, 77, 768, generator=gen)
ones = torch.ones(B, 4)
model.cuda().eval()
for sigma in (80.0, 30.0, 10.0, 2.0, 0.3):
    x = noise * sigma
    s = torch.tensor(sigma)
    cap = {}
    with torch.no_grad():
        ref = O.preconditioned(sd, FULL, x, s, text, ones, ones, 0.0, -1.2, 1.6, variant=2, capture=cap)
    line = f"sigma {sigma:5.1f}: "
    for dt in (torch.float32, torch.bfloat16):
        for tf32 in (False, True):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32
            hdmoe_b200.set_expert_dtype(dt)
            with torch.no_grad():
                out = model(x=x.cuda(), sigma=s.cuda(), text_emb=text.cuda(), Unet_router_mask=ones.cuda(),
                            Vit_router_mask=ones.cuda(), zeta=0, transition_point=-1.2, softness=1.6)
            same = torch.equal(model.net.Unet_router.last["topk_idx"].cpu().long().flatten(), ref["Unet_raw"].argmax(1)) and \
                torch.equal(model.net.vit_router.last["topk_idx"].cpu().long().flatten(), ref["vit_raw"].argmax(1))
            line += f"{'bf16' if dt == torch.bfloat16 else 'fp32'}/{'tf32' if tf32 else 'strict'} {rel_l2(out['denoised'].cpu(), ref['denoised']):.2e} gate {rel_l2(out['out_gate'].cpu(), ref['out_gate']):.2e} {'' if same else 'ROUTING DIFFERS'} | "
    print(line, flush=True)
print("oracle capture keys:", list(cap.keys()))
hdmoe_b200.set_expert_dtype(torch.float32)
