"""Fused HDMOEM glue kernels (csrc/trunk_glue.cu) against the oracle's op-by-op arithmetic of models/model_config2.py:
291-301 (text blend, mp_cat, gate1, mp_silu, gate2, pixel softmax, blend, mp_sum) and of the cfg1 soft swap
(models/model_config1.py:277-283), forward and every gradient, in float64.  fp32 kernels: rel-L2 <= 1e-5."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from oracle import hdmoe_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,W", [(3, 8, 8), (5, 32, 32), (2, 20, 12)])
def test_trunk_gate_matches_oracle_chain(B, H, W):
    from hdmoe_b200 import ops
    C = 32
    gen = torch.Generator().manual_seed(B * H + W)
    u, a, b = (torch.randn(B, H * W, C, generator=gen) for _ in range(3))
    alpha = torch.tensor(0.37)
    W1 = torch.randn(C, 2 * C, 1, 1, generator=gen) / (2 * C) ** 0.5
    W2 = torch.randn(2, C, 1, 1, generator=gen) / C ** 0.5
    gm = torch.randn(B, H * W, C, generator=gen)
    gg = torch.randn(B, 2, H, W, generator=gen)
    # oracle chain on NCHW tensors, float64
    leaves = [t.double().requires_grad_(True) for t in (u, a, b, alpha, W1, W2)]
    ur, ar, br, alr, w1r, w2r = leaves
    nchw = lambda t: t.transpose(1, 2).reshape(B, C, H, W)
    out_u = nchw(ur)
    fin = ar + alr * (br - ar)
    img = nchw(fin)
    g = F.conv2d(O.mp_silu(F.conv2d(O.mp_cat(out_u, img, dim=1), w1r)), w2r)
    g = F.softmax(g, dim=1)
    gated = g[:, 0:1] * out_u + g[:, 1:2] * img
    mix_ref = O.mp_sum(out_u, gated, t=0.5)
    ((mix_ref.flatten(2).transpose(1, 2) * gm.double()).sum() + (g * gg.double()).sum()).backward()
    dl = [t.cuda().requires_grad_(True) for t in (u, a, b, alpha, W1, W2)]
    mix, gate = ops.trunk_gate(dl[0], dl[1], dl[2], dl[3], dl[4], dl[5], H, W)
    ((mix * gm.cuda()).sum() + (gate * gg.cuda()).sum()).backward()
    assert rel_l2(mix.cpu(), mix_ref.flatten(2).transpose(1, 2)) < 1e-5
    assert rel_l2(gate.cpu(), g) < 1e-5
    for got, ref, name in zip(dl, leaves, ("u", "a", "b", "alpha_txt", "gate1", "gate2")):
        assert rel_l2(got.grad.cpu(), ref.grad) < 2e-5, name
    # out_gate gradient absent (the usual case: out_gate is only logged)
    dl2 = [t.detach().clone().requires_grad_(True) for t in dl]
    mix2, _ = ops.trunk_gate(dl2[0], dl2[1], dl2[2], dl2[3], dl2[4], dl2[5], H, W)
    (mix2 * gm.cuda()).sum().backward()
    assert torch.isfinite(dl2[0].grad).all() and float(dl2[4].grad.abs().sum()) > 0


@pytest.mark.parametrize("B,S", [(4, 64), (7, 1024)])
def test_trunk_swap_matches_oracle(B, S):
    from hdmoe_b200 import ops
    gen = torch.Generator().manual_seed(S)
    u, v = torch.randn(B, S, 32, generator=gen), torch.randn(B, S, 32, generator=gen)
    w = torch.rand(B, generator=gen)
    gq, gc = torch.randn(B, S, 32, generator=gen), torch.randn(B, S, 32, generator=gen)
    ur, vr, wr = (t.double().requires_grad_(True) for t in (u, v, w))
    st = wr.view(-1, 1, 1)
    q_ref = st * vr + (1 - st) * ur                 # models/model_config1.py:281-283
    c_ref = st * ur + (1 - st) * vr
    ((q_ref * gq.double()).sum() + (c_ref * gc.double()).sum()).backward()
    ud, vd, wd = (t.cuda().requires_grad_(True) for t in (u, v, w))
    q, c = ops.trunk_swap(ud, vd, wd)
    ((q * gq.cuda()).sum() + (c * gc.cuda()).sum()).backward()
    assert rel_l2(q.cpu(), q_ref) < 1e-6 and rel_l2(c.cpu(), c_ref) < 1e-6
    assert rel_l2(ud.grad.cpu(), ur.grad) < 1e-6 and rel_l2(vd.grad.cpu(), vr.grad) < 1e-6
    assert rel_l2(wd.grad.cpu(), wr.grad) < 1e-5


@pytest.mark.parametrize("B,H,W", [(3, 32, 32), (2, 16, 16), (2, 20, 12)])
@pytest.mark.parametrize("want_trunk", [False, True])
def test_scale_pair_matches_reference_products(B, H, W, want_trunk):
    """in_vit = s_vit * feats, in_unet = s_unet * feats (models/model_config2.py:250-251) + the channels-last bf16 copy
    for the tcgen05 router trunk, forward and all gradients (d feats, d scaling) against float64 autograd."""
    from hdmoe_b200 import ops
    gen = torch.Generator().manual_seed(B + H)
    feats = torch.randn(B, 32, H, W, generator=gen)
    scaling = torch.rand(B, 2, generator=gen) * 2
    gv, gu = torch.randn(B, 32, H, W, generator=gen), torch.randn(B, 32, H, W, generator=gen)
    gt = torch.randn(2 * B, H, W, 32, generator=gen).to(torch.bfloat16)
    fr, sr = feats.double().requires_grad_(True), scaling.double().requires_grad_(True)
    iv = sr[:, 0].view(-1, 1, 1, 1) * fr
    iu = sr[:, 1].view(-1, 1, 1, 1) * fr
    loss = (iv * gv.double()).sum() + (iu * gu.double()).sum()
    if want_trunk:
        tr = torch.cat([iv, iu]).permute(0, 2, 3, 1)
        loss = loss + (tr * gt.double()).sum()
    loss.backward()
    fd, sd = feats.cuda().requires_grad_(True), scaling.cuda().requires_grad_(True)
    a, b, t = ops.scale_pair(fd, sd, want_trunk=want_trunk)
    l2 = (a * gv.cuda()).sum() + (b * gu.cuda()).sum()
    if want_trunk:
        assert t.dtype == torch.bfloat16 and t.shape == (2 * B, H, W, 32)
        assert rel_l2(t.float().cpu(), torch.cat([iv, iu]).permute(0, 2, 3, 1).detach()) < 4e-3       # bf16 rounding
        l2 = l2 + (t.float() * gt.cuda().float()).sum()
    else:
        assert t is None
    l2.backward()
    assert torch.equal(a.cpu(), (scaling[:, 0].view(-1, 1, 1, 1) * feats)) and torch.equal(b.cpu(), scaling[:, 1].view(-1, 1, 1, 1) * feats)
    assert rel_l2(fd.grad.cpu(), fr.grad) < 1e-5
    assert rel_l2(sd.grad.cpu(), sr.grad) < 1e-5


def test_analytic_scaling_matches_reference():
    from hdmoe_b200 import ops
    t = torch.linspace(-2.0, 1.2, 37)
    w = torch.sigmoid((t * 4 - (-1.2)) / 1.6)
    ref = torch.stack([(w + 1e-2) * 2, ((1.0 - w) + 1e-2) * 2], dim=1)         # models/model_config2.py:244-249
    got = ops.analytic_scaling(t.cuda(), -1.2, 1.6).cpu()
    assert rel_l2(got, ref) < 1e-6


@pytest.mark.parametrize("train", [False, True])
def test_scaling_router_fused_matches_oracle(train):
    """Scaling_router.forward (models/model_components.py:41-66) as one kernel per direction against oracle.scaling_router
    (fp64): output, input gradient, all weight / GroupNorm gradients; the module path (prepared weights, forced weight
    norm in training) is what runs."""
    from hdmoe_b200 import model_components as mc
    torch.manual_seed(3)
    sr = mc.Scaling_router(emb_dim=64, num_experts=2, dropout=0.0)
    gen = torch.Generator().manual_seed(4)
    with torch.no_grad():
        for i in (1, 4):
            sr.soft_route[i].weight.copy_(1 + 0.2 * torch.randn(sr.soft_route[i].weight.shape, generator=gen))
            sr.soft_route[i].bias.copy_(0.2 * torch.randn(sr.soft_route[i].bias.shape, generator=gen))
    B = 37
    x = torch.randn(B, 64, generator=gen)
    nz = torch.randn(B, 2, generator=gen)
    gy = torch.randn(B, 2, generator=gen)
    sd = {"s." + k: v.detach().double().clone().requires_grad_(True) for k, v in sr.state_dict().items()}
    xr = x.double().requires_grad_(True)
    import contextlib
    with (O.training_mode() if train else contextlib.nullcontext()):
        ref = O.scaling_router(sd, "s.", xr, noise=nz.double() if train else None, zeta=0.7 if train else 0.0)
    (ref * gy.double()).sum().backward()
    sr.cuda().train(train)
    xd = x.cuda().requires_grad_(True)
    out = sr(xd, zeta=0.7, noise=nz.cuda())
    (out * gy.cuda()).sum().backward()
    assert rel_l2(out.cpu(), ref) < 1e-5
    assert rel_l2(xd.grad.cpu(), xr.grad) < 2e-5
    for k, p in sr.named_parameters():
        assert rel_l2(p.grad.cpu(), sd["s." + k].grad) < 5e-5, k
        if train and k.endswith("weights"):
            assert rel_l2(p.detach().cpu(), sd["s." + k].detach()) < 1e-6, k        # forced weight norm (Q6)
    assert torch.allclose(out.sum(1).cpu(), torch.full((B,), 2.0), atol=1e-5)      # rows sum to 2 (test_routers.py:28-29)


def _away_from_band_edge(sigma, mg, step, min_active, margin=1e-6):
    """rows whose mask is decided by more than `margin`: no |dist - bandwidth| and no top-min_active tie inside it"""
    pct = 0.5 * (1 + torch.erf((torch.log(sigma.flatten().double()) - mg.p_mean) / (mg.p_std * 2 ** 0.5)))
    dist = (pct.view(-1, 1) - mg.expert_centers.double().view(1, -1)).abs()
    srt = dist.sort(dim=1).values
    ok = (dist - mg.bandwidth_scheduler(step)).abs().min(dim=1).values > margin
    if min_active < dist.shape[1]:
        ok &= (srt[:, min_active] - srt[:, min_active - 1]) > margin
    return ok


@pytest.mark.gpu
def test_train_inputs_producer_matches_reference_fixture_and_oracle():
    """Fused on-device producers (csrc/edm_step.cu train_inputs_kernel): noise add bit-exact, both band masks equal
    to the reference's MaskGenerator fixtures (tests/golden/producers.npz) at four schedule steps and to the oracle on
    4 096 random noise levels; a sample whose percentile sits within 1e-6 of a band edge may legitimately fall on
    either side (device erff / logf vs the CPU's), so those are excluded -- and counted."""
    from conftest import load_golden
    from hdmoe_b200.utils import MaskGenerator, make_train_inputs
    g = load_golden("producers")
    dev = torch.device("cuda")
    gens = {}
    for tag, attrs, rng in (("unet", [3, 3, 5, 5], (0.0, 0.6)), ("vit", [4, 8, 8, 16], (0.4, 1.0))):
        gens[tag] = MaskGenerator(expert_attributes=attrs, p_mean=-1.2, p_std=1.6, bandwidth=0.3, max_bandwidth=0.8,
                                  min_active=1, total_steps=5000, step_size=0.1, noise_range=rng, strat_band="step")
    sigma = g["mask.sigma"]
    B = sigma.numel()
    gen = torch.Generator().manual_seed(11)
    x0, eps = torch.randn(B, 4, 8, 8, generator=gen) * 0.5, torch.randn(B, 4, 8, 8, generator=gen)
    for step in (0, 700, 2600, 6000):
        x, mu, mv = make_train_inputs(x0.to(dev), sigma.to(dev), gens["unet"], gens["vit"], step=step, eps=eps.to(dev))
        assert torch.equal(x.cpu(), x0 + eps * sigma.view(-1, 1, 1, 1))
        for tag, got in (("unet", mu), ("vit", mv)):
            ok = _away_from_band_edge(sigma, gens[tag], step, 1)
            assert int(ok.sum()) >= B - 1
            assert torch.equal(got.cpu()[ok], g[f"mask.{tag}.{step}"][ok])
    # wide random sweep against the oracle, min_active = 2, one generator only
    sig = torch.exp(torch.randn(4096, generator=gen) * 1.6 - 1.2).clamp(0.002, 80.0)
    mg = MaskGenerator(expert_attributes=[1, 2, 3, 4, 5, 6, 7, 8], p_mean=-1.2, p_std=1.6, bandwidth=0.2, min_active=2)
    x0b = torch.randn(4096, 4, generator=gen)
    xb, ma, mb = make_train_inputs(x0b.to(dev), sig.to(dev), mg, None, step=0, eps=torch.zeros(4096, 4, device=dev))
    assert mb is None and torch.equal(xb.cpu(), x0b)
    want = O.band_mask(sig, mg.expert_centers, mg.bandwidth_scheduler(0), -1.2, 1.6, min_active=2)
    safe = _away_from_band_edge(sig, mg, 0, 2)
    assert int(safe.sum()) > 4000
    assert torch.equal(ma.cpu()[safe], want[safe])
    assert bool((ma.sum(1) >= 2).all())
