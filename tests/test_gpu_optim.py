"""hdmoe_b200.optim.FusedAdamW (csrc/optim.cu: gradient-norm clip + AdamW in three launches) against the library pair the
reference's loop calls, torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW (Utils/training.py:195-197): parameters,
moments, clipped gradients and the reported norm over several steps; odd sizes around the chunk size, unaligned views,
a parameter without gradient, two param groups; replay inside a CUDA graph."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu

SIZES = [(1,), (3,), (8191,), (8192,), (8193,), (64, 33, 3, 3), (100003,), (7, 5), (2, 3, 4, 5)]


def _make(seed, offset=1):
    """parameters as views of ONE flat buffer starting `offset` floats in: most are not 16-byte aligned"""
    gen = torch.Generator().manual_seed(seed)
    base = torch.randn(sum(torch.Size(s).numel() for s in SIZES) + 8, generator=gen).cuda()
    ps, o = [], offset
    for s in SIZES:
        n = torch.Size(s).numel()
        ps.append(torch.nn.Parameter(base[o:o + n].view(s)))
        o += n
    return ps


def _grads_like(ps, gen, scale, offset=2):
    flat = torch.empty(sum(p.numel() for p in ps) + 8, device="cuda")
    out, o = [], offset
    for p in ps:
        n = p.numel()
        g = flat[o:o + n].view(p.shape)
        g.copy_((torch.randn(p.shape, generator=gen) * scale).cuda())
        out.append(g)
        o += n
    return out


@pytest.mark.parametrize("max_norm", [1.0, 1e9, None])
def test_fused_adamw_matches_torch(max_norm):
    from hdmoe_b200.optim import FusedAdamW
    pa, pb = _make(0), _make(0)
    groups = lambda ps: [dict(params=ps[:4], lr=2e-3, weight_decay=0.0), dict(params=ps[4:], lr=5e-4)]
    ref = torch.optim.AdamW(groups(pa), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2)
    opt = FusedAdamW(groups(pb), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2, max_grad_norm=max_norm)
    gen = torch.Generator().manual_seed(1)
    for it in range(5):
        gs = _grads_like(pb, gen, 10.0 if it == 1 else 0.1)
        for i, (a, b, g) in enumerate(zip(pa, pb, gs)):
            if i == 2 and it < 4:                      # a parameter without a gradient is skipped
                a.grad = b.grad = None
                continue
            a.grad, b.grad = g.clone(), g
        if max_norm is not None:
            n_ref = torch.nn.utils.clip_grad_norm_(pa, max_norm)
        ref.step()
        opt.step()
        if max_norm is not None:
            assert abs(float(opt.last_grad_norm) - float(n_ref)) < 1e-5 * float(n_ref)
            for a, b in zip(pa, pb):
                if a.grad is not None:
                    assert rel_l2(b.grad, a.grad) < 1e-6          # clipped in place like clip_grad_norm_
        for a, b in zip(pa, pb):
            assert rel_l2(b.detach(), a.detach()) < 2e-6, it
    for i, (a, b) in enumerate(zip(pa, pb)):
        assert float(opt.state[b]["step"]) == float(ref.state[a]["step"]) == (1.0 if i == 2 else 5.0)   # per-tensor steps
        assert rel_l2(opt.state[b]["exp_avg"], ref.state[a]["exp_avg"]) < 1e-6
        assert rel_l2(opt.state[b]["exp_avg_sq"], ref.state[a]["exp_avg_sq"]) < 1e-6
    # torch-format state round trip
    sd = opt.state_dict()
    pc = _make(0)
    opt2 = FusedAdamW(groups(pc), lr=1e-3, betas=(0.9, 0.99), eps=1e-8, weight_decay=1e-2, max_grad_norm=max_norm)
    opt2.load_state_dict(sd)
    assert float(opt2._state[0]) == 5.0
    assert rel_l2(opt2.state[pc[5]]["exp_avg"], opt.state[pb[5]]["exp_avg"]) == 0.0


def test_fused_adamw_replays_in_cuda_graph():
    """warm-up step eagerly, capture one step, replay it three times with new gradients in the static buffers: the same
    four updates as the eager library pair."""
    from hdmoe_b200.optim import FusedAdamW
    pa, pb = _make(3), _make(3)
    ref = torch.optim.AdamW(pa, lr=1e-3)
    opt = FusedAdamW(pb, lr=1e-3, max_grad_norm=0.5)
    gen = torch.Generator().manual_seed(4)
    static_g = _grads_like(pb, gen, 1.0)
    seq = [[g.clone() for g in static_g]] + [[g.clone() for g in _grads_like(pb, gen, 1.0)] for _ in range(3)]
    for p, g in zip(pb, static_g):
        p.grad = g
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        opt.step()                                     # eager warm-up step on seq[0] (allocates the state)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=s):
        opt.step()                                     # recorded, not executed
    opt._pinned.zero_()                                # the capture must not depend on the eager staging buffer
    for gs in seq[1:]:
        for sg, g in zip(static_g, gs):
            sg.copy_(g)
        graph.replay()
    for gs in seq:
        for p, g in zip(pa, gs):
            p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(pa, 0.5)
        ref.step()
    torch.cuda.synchronize()
    assert float(opt._state[0]) == 4.0
    for a, b in zip(pa, pb):
        assert rel_l2(b.detach(), a.detach()) < 2e-6
