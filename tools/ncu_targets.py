"""One launch of every hand-written hot kernel at the bench workload's shapes, bracketed by cudaProfilerStart/Stop:
    ncu --set full --import-source on --clock-control none --profile-from-start off -o gpurun_out/r1_kernels \
        python tools/ncu_targets.py
The same shapes as bench.py (roofline_gconv, dispatch_sweep) and the B = 256 trunk attention / router trunk."""
import sys, torch
sys.path.insert(0, '.')
import hdmoe_b200
from hdmoe_b200 import ops, nhwc
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.manual_seed(0)

def gconv_case():
    counts, ks = [36, 48, 75, 97], [3, 3, 5, 5]
    R, H, Cin, Cout = sum(counts), 32, 64, 64
    row_e = torch.tensor(sum(([e] * c for e, c in enumerate(counts)), []), dtype=torch.int32, device=dev)
    n_rows = torch.tensor([R], dtype=torch.int32, device=dev)
    x = torch.randn(R, H, H, Cin, device=dev).to(torch.bfloat16)
    dy = torch.randn(R, H, H, Cout, device=dev).to(torch.bfloat16)
    wrow, tot = [], 0
    for k in ks:
        wrow.append(tot); tot += k * k * Cout
    w = (torch.randn(tot, Cin, device=dev) / 30).to(torch.bfloat16)
    dw = torch.zeros(tot, Cin, device=dev)
    def run():
        ops.gconv_raw(x, w, Cout, tot, row_e, n_rows, ks, wrow)
        ops.gconv_wgrad_raw(x, dy, dw, row_e, n_rows, ks, wrow)
    return run

def attn_case():
    B, H = 256, 8
    q, k, v = (torch.randn(B, s, H * 4, device=dev, requires_grad=True) for s in (1024, 1024, 1024))
    gy = torch.randn(B, 1024, H * 4, device=dev)
    def run():
        o = ops.attention_d4(q, k, v, H, 0.5); o.backward(gy)
    return run

def gn_case():
    x = torch.randn(256, 128, 32, 32, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.ones(128, device=dev, requires_grad=True); b = torch.zeros(128, device=dev, requires_grad=True)
    gy = torch.randn(256, 128, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
    gp = torch.randn(256, 128, device=dev)
    def run():
        ops.gn1_relu(x, g, b, 1e-5).backward(gy)
        ops.gn1_relu(x, g, b, 1e-5, pool=True).backward(gp)
    return run

def dispatch_case():
    T, E, K = 256, 4, 1
    logits = torch.randn(T, E, device=dev)
    sparse, probs, lg, idx, tw, st = ops.router_gate_from_logits(logits, K) if False else (None,) * 6
    w = torch.zeros(T, E, device=dev); w.scatter_(1, logits.argmax(1, keepdim=True), 1.0)
    x = torch.randn(T, 32, 32, 32, device=dev).to(torch.bfloat16)
    te = torch.randn(T, 64, device=dev).to(torch.bfloat16)
    def run():
        plan = ops.dispatch_plan(w, K)
        rows = ops.permute(plan, x, te)
        ops.combine(rows[0], w, plan, base=None, out_dtype=torch.float32)
    return run

def router_case():
    B, Cn, E = 256, 128, 4
    pooled = torch.randn(B, Cn, device=dev); cond = torch.randn(B, 2 * Cn, device=dev)
    w_hat = torch.randn(E, Cn, device=dev) / 11; noise = torch.randn(B, E, device=dev)
    def run():
        ops.router_gate(pooled, cond, w_hat, 1, noise=noise, zeta=2.0, mask=None)
    return run

def glue_case():
    x = torch.randn(256, 32, 32, 64, device=dev).to(torch.bfloat16).requires_grad_(True)
    gain = torch.rand(256, 64, device=dev) + 0.5
    rows = torch.randn(256, 32, 32, 32, device=dev).to(torch.bfloat16)
    def run():
        xn, a = nhwc.pixnorm_silu(x); (xn + a).backward(torch.ones_like(xn))
        y = nhwc.gain_silu(x, gain); y.backward(torch.ones_like(y))
        nhwc.mp_cat(nhwc.mp_sum(x.detach(), x.detach(), 0.5), x.detach(), 0.5)
        nhwc.nhwc_to_rows(nhwc.rows_to_nhwc(rows, 64)[..., :32].contiguous())
    return run

def vit_case():
    from hdmoe_b200 import model_components as mc, _denoiser as D
    hdmoe_b200.set_expert_dtype(torch.bfloat16)
    ex = torch.nn.ModuleList([mc.Vit_expert(num_heads=8, num_groups=4, in_channels=32, seq_ln=(32 // p) ** 2, emb_dim=32, num_blocks=4,
                                            patch_size=p, time_dim=64, text_dim=768) for p in (4, 8, 8, 16)]).to(dev).train()
    B = 256
    x = torch.randn(B, 32, 32, 32, device=dev, requires_grad=True); t = torch.randn(B, 64, device=dev); tx = torch.randn(B, 77, 768, device=dev)
    idx = torch.tensor(sum(([e] * c for e, c in enumerate([36, 48, 75, 97])), []), device=dev)
    wr = torch.zeros(B, 4, device=dev).scatter_(1, idx[:, None], 1.0)
    gy = torch.randn(B, 32, 32, 32, device=dev)
    def run():
        D.router_to_unet_experts(x, ex, wr, t, tx, top_k=1).backward(gy)
    return run

def trunk_case():
    """router trunks of both routers on the tcgen05 kernels (dense N = 64 / 128 layers) + bf16 GroupNorm kernels"""
    from hdmoe_b200 import model_components as mc
    from hdmoe_b200.router_trunk import GroupedRouterTrunk
    rs = [mc.Router(in_channels=32, time_dim=64, top_k=1, num_experts=4).to(dev).train() for _ in range(2)]
    runner = GroupedRouterTrunk(rs)
    feats = torch.randn(256, 32, 32, 32, device=dev, requires_grad=True)
    scaling = torch.rand(256, 2, device=dev) + 0.5
    gp = torch.randn(256, 128, device=dev)
    def run():
        a, b, t = ops.scale_pair(feats, scaling, want_trunk=True)
        pv, pu = runner(None, True, pre_nhwc=t)
        ((pv * gp).sum() + (pu * gp).sum() + a.sum() + b.sum()).backward()
    return run

def tail_case():
    """channels-last HDMOEM tail: cfg1 swap + text-blend / gate / mix kernels, loss data term, fused optimizer"""
    from hdmoe_b200.optim import FusedAdamW
    B, S, C = 256, 1024, 32
    u, v, a, b = (torch.randn(B, S, C, device=dev, requires_grad=True) for _ in range(4))
    w = torch.rand(B, device=dev, requires_grad=True)
    alpha = torch.tensor(0.3, device=dev, requires_grad=True)
    W1 = (torch.randn(32, 64, 1, 1, device=dev) / 8).requires_grad_(True)
    W2 = (torch.randn(2, 32, 1, 1, device=dev) / 6).requires_grad_(True)
    d, x0 = torch.randn(B, 4, 32, 32, device=dev, requires_grad=True), torch.randn(B, 4, 32, 32, device=dev)
    params = [torch.nn.Parameter(torch.randn(n, device=dev)) for n in (2_000_000, 3_000_000, 4_000_000, 38_821)]
    opt = FusedAdamW(params, lr=5e-4, max_grad_norm=1.0)
    for p_ in params:
        p_.grad = torch.randn_like(p_)
    def run():
        q, c = ops.trunk_swap(u, v, w)
        mix, g = ops.trunk_gate(q, a, b, alpha, W1, W2, 32, 32)
        (mix.sum() + c.sum() + ops.sqerr_rows(d, x0).sum()).backward()
        opt.step()
    return run

def plan_case():
    T, E, K = 1048576, 64, 1
    idx = torch.randint(0, E, (T, K), device=dev, dtype=torch.int32)
    tw = torch.ones(T, K, device=dev)
    x = torch.randn(T, 128, device=dev).to(torch.bfloat16)
    sp = torch.zeros(T, E, device=dev).scatter_(1, idx.long(), 1.0)
    def run():
        plan = ops.dispatch_plan_from_topk(idx, tw, E)
        rows = ops.permute(plan, x)
        ops.combine(rows[0], sp, plan)
    return run

cases = [gconv_case(), attn_case(), gn_case(), dispatch_case(), router_case(), glue_case(), vit_case(), trunk_case(),
         tail_case(), plan_case()]
for c in cases:
    c(); c()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for c in cases:
    c()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")
