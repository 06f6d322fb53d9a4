"""Router.hard_route on the tcgen05 grouped convolution kernels (router_trunk.py) against the oracle's fp32 trunk
(oracle.router_trunk, models/model_components.py:100-112,141-143): pooled features, input gradient, convolution-weight
and GroupNorm gradients of BOTH routers from the single grouped launch per layer, the train-mode forced weight norm, and
routing decisions identical to the oracle on >= 4096 random samples (margin rule)."""
import contextlib

import pytest
import torch

from conftest import rel_l2
from oracle import hdmoe_oracle as O

pytestmark = pytest.mark.gpu


def _routers(seed=0, E=4, k=1):
    from hdmoe_b200 import model_components as mc
    torch.manual_seed(seed)
    rs = [mc.Router(in_channels=32, time_dim=64, top_k=k, num_experts=E, dropout=0.0) for _ in range(2)]
    gen = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for r in rs:
            for i in (1, 4, 7):
                r.hard_route[i].weight.copy_(1 + 0.2 * torch.randn(r.hard_route[i].weight.shape, generator=gen))
                r.hard_route[i].bias.copy_(0.2 * torch.randn(r.hard_route[i].bias.shape, generator=gen))
    return rs


class _RoundBF(torch.autograd.Function):
    """bf16 rounding point of the tcgen05 trunk (activations are stored in bf16 between kernels, forward and backward)"""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _trunk_with_rounding_points(sd, x):
    """oracle.router_trunk with the path's bf16 rounding points made explicit (input, prepared weights, every convolution
    output, the GroupNorm+ReLU outputs that are written back; statistics / pool in fp32) -- float64 arithmetic otherwise.
    ReLU masks are decided on bf16-rounded pre-activations, which the plain fp32 oracle cannot reproduce."""
    import torch.nn.functional as F
    rb = _RoundBF.apply
    h = rb(x.double())
    for i in (0, 3, 6):
        w = rb(O.mp_weight(sd[f"hard_route.{i}.weights"].double()))       # straight-through rounding of the operand
        h = rb(F.conv2d(h, w, padding=1))
        h = F.relu(F.group_norm(h, 1, sd[f"hard_route.{i + 1}.weight"].double(), sd[f"hard_route.{i + 1}.bias"].double()))
        if i != 6:
            h = rb(h)
    return h.mean(dim=(2, 3))


def _sd(r):
    return {k: v.detach().cpu().clone().requires_grad_(True) for k, v in r.state_dict().items()}


@pytest.mark.parametrize("train", [False, True])
@pytest.mark.parametrize("B,H", [(6, 32), (3, 64), (5, 16)])
def test_router_trunk_tcgen05_vs_oracle(B, H, train):
    from hdmoe_b200.router_trunk import GroupedRouterTrunk
    rs = _routers()
    gen = torch.Generator().manual_seed(B + H)
    xs = [torch.randn(B, 32, H, H, generator=gen) * s for s in (1.9, 0.12)]      # the two sigma-scaled router inputs
    gys = [torch.randn(B, 128, generator=gen) for _ in range(2)]
    # (a) the fp32 oracle; (b) the same arithmetic with the path's bf16 rounding points (input gradient / GroupNorm
    # gradients of a ReLU network are only comparable when the ReLU masks agree: CPU simulation of the rounding points
    # alone moves grad_x by 8 % against fp32, tools note in DESIGN §2)
    refs = []
    for r, x, gy in zip(rs, xs, gys):
        sd = _sd(r)
        xr = x.clone().requires_grad_(True)
        with (O.training_mode() if train else contextlib.nullcontext()):
            pooled = O.router_trunk(sd, "", xr)
        (pooled * gy).sum().backward()
        xq = x.clone().requires_grad_(True)
        sdq = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}      # after the forced weight norm
        pooled_q = _trunk_with_rounding_points(sdq, xq)
        (pooled_q * gy.double()).sum().backward()
        refs.append((pooled.detach(), xr.grad, sd, pooled_q.detach(), xq.grad, sdq))
    for r in rs:
        r.cuda().train(train)
    assert GroupedRouterTrunk.supported(rs, xs[0].cuda())
    runner = GroupedRouterTrunk(rs)
    xd = [x.cuda().requires_grad_(True) for x in xs]
    pooled = runner(xd, training=train)
    sum((p * gy.cuda()).sum() for p, gy in zip(pooled, gys)).backward()
    torch.cuda.synchronize()
    for e, (r, (p_ref, gx_ref, sd, p_q, gx_q, sdq)) in enumerate(zip(rs, refs)):
        assert pooled[e].dtype == torch.float32
        assert rel_l2(pooled[e].cpu(), p_ref) < 1e-2, ("pooled vs fp32 oracle", e, rel_l2(pooled[e].cpu(), p_ref))
        assert rel_l2(pooled[e].cpu(), p_q) < 2e-3, ("pooled vs rounding-aware oracle", e, rel_l2(pooled[e].cpu(), p_q))
        assert rel_l2(xd[e].grad.cpu(), gx_q) < 2e-2, ("grad x", e, rel_l2(xd[e].grad.cpu(), gx_q))
        assert rel_l2(xd[e].grad.cpu(), gx_ref) < 0.3, ("grad x vs fp32 oracle", e, rel_l2(xd[e].grad.cpu(), gx_ref))
        for i in (0, 3, 6):
            g = r.hard_route[i].weights.grad
            assert g is not None
            assert rel_l2(g.cpu(), sdq[f"hard_route.{i}.weights"].grad) < 2e-2, (f"conv {i}", e)
            assert rel_l2(g.cpu(), sd[f"hard_route.{i}.weights"].grad) < 0.3, (f"conv {i} vs fp32 oracle", e)
            if train:       # forced weight norm (Q6) rewrote the master weights exactly like the reference
                assert rel_l2(r.hard_route[i].weights.detach().cpu(), sd[f"hard_route.{i}.weights"].detach()) < 1e-6
        for i in (1, 4, 7):
            assert rel_l2(r.hard_route[i].weight.grad.cpu(), sdq[f"hard_route.{i}.weight"].grad) < 2e-2, (f"gamma {i}", e)
            assert rel_l2(r.hard_route[i].bias.grad.cpu(), sdq[f"hard_route.{i}.bias"].grad) < 2e-2, (f"beta {i}", e)


def test_router_routing_identical_to_oracle_4096_samples():
    """Routing through the tcgen05 trunk + fused gate: top-1 indices identical to the fp32 oracle on every sample whose
    oracle margin exceeds 2e-2 (>= 4096 random samples, both routers); logits within 1e-2."""
    from hdmoe_b200.router_trunk import GroupedRouterTrunk
    rs = _routers(seed=3)
    for r in rs:
        r.cuda().eval()
    runner = GroupedRouterTrunk(rs)
    sds = [{k: v.detach().cpu() for k, v in r.state_dict().items()} for r in rs]
    gen = torch.Generator().manual_seed(11)
    tot = kept = 0
    worst = 0.0
    for it in range(16):
        B = 128
        xs = [torch.randn(B, 32, 32, 32, generator=gen) * s for s in (1.0, 0.7)]
        te = torch.randn(B, 64, generator=gen)
        with torch.no_grad():
            pooled = runner([x.cuda() for x in xs], training=False)
            for r, sd, x, p in zip(rs, sds, xs, pooled):
                sp, gp, lg = r(x=x.cuda(), time_emb=te.cuda(), zeta=0, pooled=p)
                sp_r, gp_r, lg_r, idx_r = O.router(sd, "", x, te, 1)
                v = torch.sort(lg_r, dim=-1, descending=True).values
                ok = (v[:, 0] - v[:, 1]) > 2e-2
                got = r.last["topk_idx"].cpu().long().flatten()
                assert torch.equal(got[ok], idx_r.flatten()[ok])
                worst = max(worst, rel_l2(lg.cpu(), lg_r))
                tot += 2 * 0 + B
                kept += int(ok.sum())
    assert tot >= 4096 and kept > 0.8 * tot, (tot, kept)
    assert worst < 1e-2, worst
