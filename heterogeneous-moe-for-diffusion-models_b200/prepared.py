"""Multi-tensor weight preparation for arbitrary groups of MP_Conv modules (trunk, routers, ViT experts).

The reference normalises every MP_Conv weight inside its own forward with ~8 small launches (and ~12 more in the
backward); with ~200 such modules outside the grouped U-Net path this is ~4 000 launches per train step.  A
PreparedGroup prepares all weights of a group with ONE W-PREP launch (forced in-place weight norm in training,
gated per expert by a device flag; normalise; scale; cast), hands the prepared tensors to the modules through a
registry that `MP_Conv.prepared_weight` consults, and converts all weight-operand gradients back to
master-weight gradients with ONE launch in the backward.  All buffers are persistent, so the launches replay inside
a CUDA graph."""
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops

_REG: Dict[int, Tuple[torch.Tensor, float]] = {}

# Gradient hand-off without AccumulateGrad's clone.  The weight gradients of a group are views of ONE persistent
# buffer; returned from backward() they are cloned one by one (445 device-to-device copies per train step, ~1 ms at the
# end of the backward).  With direct hand-off a parameter whose .grad is None receives the view itself and backward()
# returns None for it; a parameter that already holds a gradient takes the normal autograd route.  Parameter hooks do
# not fire for handed-off gradients (none are used on this path); keep zero_grad(set_to_none=True), the default.
_DIRECT_GRADS = [True]


def set_direct_grads(enabled: bool) -> None:
    _DIRECT_GRADS[0] = bool(enabled)


def unalias_grads(params: Sequence[torch.Tensor], buf: torch.Tensor) -> None:
    """Before `buf` is overwritten: a .grad left over from an earlier backward that still aliases it is cloned
    (gradient accumulation across backward passes keeps its meaning)."""
    lo, hi = buf.data_ptr(), buf.data_ptr() + buf.numel() * buf.element_size()
    for p_ in params:
        g = p_.grad
        if g is not None and lo <= g.data_ptr() < hi:
            p_.grad = g.clone()


def deliver_grads(params: Sequence[torch.Tensor], views: Sequence[torch.Tensor]) -> List[Optional[torch.Tensor]]:
    if not _DIRECT_GRADS[0] or torch.is_grad_enabled():
        return list(views)
    out = []
    for p_, v in zip(params, views):
        if p_.grad is None and p_.requires_grad and p_.is_leaf:
            p_.grad = v
            out.append(None)
        else:
            out.append(v)
    return out


def lookup(module, gain) -> Optional[torch.Tensor]:
    ent = _REG.get(id(module))
    if ent is None or torch.is_tensor(gain) or float(gain) != ent[1]:
        return None
    return ent[0]


class _PrepFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, group, training, *weights):
        group.generation += 1
        group._run_fwd(training)
        ctx.group, ctx.gen = group, group.generation
        return tuple(v.view_as(v) for v in group.w_hat_views)

    @staticmethod
    def backward(ctx, *grads):
        if ctx.gen != ctx.group.generation:
            # the prepared weights are overwritten in place by the next forward's W-PREP launch: the views autograd
            # saved for this backward no longer hold this forward's values
            raise RuntimeError("hdmoe_b200 PreparedGroup: backward of forward #%d after forward #%d ran (one backward "
                               "per forward; call backward before the next forward, or set_trunk_weight_prep(False))"
                               % (ctx.gen, ctx.group.generation))
        return (None, None, *ctx.group._run_bwd(grads))


class PreparedGroup:
    def __init__(self, modules: Sequence, gains: Sequence[float], expert_of: Optional[Sequence[int]] = None,
                 n_experts: int = 0, out_dtype=torch.float32):
        """modules: MP_Conv instances; gains: the (float) gain each one is called with; expert_of[i]: index of the
        expert the module belongs to (activity flag), or None for always-active groups."""
        self.mods = list(modules)
        self.gains = [float(g) for g in gains]
        self.expert_of = list(expert_of) if expert_of is not None else None
        self.n_experts = n_experts
        self.out_dtype = out_dtype
        self._dev = None
        self.generation = 0          # forwards through the autograd path (buffers hold ONE forward's state)

    def _build(self, dev):
        n = sum(m.weights.numel() for m in self.mods)
        self.w_hat_flat = torch.empty(n, dtype=self.out_dtype, device=dev)
        self.g_flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.dw_flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.active = torch.ones(max(self.n_experts, 1), dtype=torch.int32, device=dev)
        self.w_hat_views, self.g_views, self.dw_views = [], [], []
        o = 0
        for m in self.mods:
            k = m.weights.numel()
            self.w_hat_views.append(self.w_hat_flat[o:o + k].view_as(m.weights))
            self.g_views.append(self.g_flat[o:o + k].view_as(m.weights))
            self.dw_views.append(self.dw_flat[o:o + k].view_as(m.weights))
            o += k
        self._wp = self._wpb = None
        self._sig = self._sigb = None
        self._dev = dev

    def _entries(self):
        ent = []
        for i, m in enumerate(self.mods):
            act = self.active[self.expert_of[i]:self.expert_of[i] + 1] if self.expert_of is not None else None
            ent.append(dict(w=m.weights, out=self.w_hat_views[i], gain=self.gains[i], active=act))
        return ent

    def _run_fwd(self, training):
        ent = self._entries()
        if self._wp is None:
            self._wp = ops.WeightPrep(ent, self._dev)
        self._wp.entries = ent
        sig = self._wp.signature()
        if sig != self._sig:
            self._wp.upload()
            self._sig = sig
        self._wp.run_uploaded(force=training)

    def _run_bwd(self, grads):
        unalias_grads([m.weights for m in self.mods], self.dw_flat)
        dsts, srcs = [], []
        for v, g in zip(self.g_views, grads):
            if g is None:
                v.zero_()
            else:
                dsts.append(v)
                srcs.append(g.reshape(v.shape))
        if dsts:
            torch._foreach_copy_(dsts, srcs)
        ent = [dict(w=m.weights, d_w_hat=self.g_views[i], d_w=self.dw_views[i], gain=self.gains[i])
               for i, m in enumerate(self.mods)]
        if self._wpb is None:
            self._wpb = ops.WeightPrepBackward(ent, self._dev)
        self._wpb.entries = ent
        sig = self._wpb.signature()
        if sig != self._sigb:
            self._wpb.upload()
            self._sigb = sig
        self._wpb.run_uploaded()
        return deliver_grads([m.weights for m in self.mods], self.dw_views)

    class _Ctx:
        def __init__(self, group, tensors):
            self.group, self.tensors = group, tensors

        def __enter__(self):
            for m, t, g in zip(self.group.mods, self.tensors, self.group.gains):
                _REG[id(m)] = (t, g)
            return self

        def __exit__(self, *a):
            for m in self.group.mods:
                _REG.pop(id(m), None)

    def prepared(self, training: bool, counts: Optional[torch.Tensor] = None):
        """Context manager: inside it the group's modules use the jointly prepared weights."""
        dev = self.mods[0].weights.device
        if self._dev != dev:
            self._build(dev)
        if counts is not None and self.expert_of is not None:
            self.active.copy_(counts)
        if torch.is_grad_enabled() and any(m.weights.requires_grad for m in self.mods):
            tensors = _PrepFn.apply(self, training, *[m.weights for m in self.mods])
        else:
            self._run_fwd(training)
            tensors = self.w_hat_views
        return PreparedGroup._Ctx(self, tensors)


def vit_expert_group(experts, out_dtype) -> PreparedGroup:
    """All MP_Conv modules of a ModuleList of Vit_expert with the gain each is called with
    (models/model_components.py:533-557, 687, 696; models/model_internals.py:364-372,407)."""
    mods, gains, owner = [], [], []
    for e, ex in enumerate(experts):
        def add(m, g):
            if m is not None:
                mods.append(m)
                gains.append(g)
                owner.append(e)
        add(ex.map_txt, 1.0)
        for blk in ex.diffit:
            gs, gt = blk.gain_s, blk.gain_t
            add(blk.skip_proj, gs)
            add(blk.linear1, gs)
            for nm in ("q_proj", "k_proj", "v_proj", "out_proj"):
                add(getattr(blk.TMSA, nm), gs)
            for nm in ("q_time", "k_time", "v_time"):
                add(getattr(blk.TMSA, nm), gt)
            add(blk.linear2, gs)
            add(blk.linear3, gs)
        add(ex.unpatch_proj, 1.0)
    return PreparedGroup(mods, gains, owner, len(experts), out_dtype)


def trunk_group(net, router_convs: bool = True) -> PreparedGroup:
    """The always-active MP_Conv modules of HDMOEM outside the experts (gain 1 everywhere).  router_convs=False leaves
    out the routers' trunk convolutions (router_trunk.py prepares them itself as tcgen05 operands)."""
    mods = [net.input_proj, net.out_fourier1, net.out_fourier2]
    if hasattr(net, "scaling_net"):
        mods += [net.scaling_net.soft_route[0], net.scaling_net.soft_route[3], net.scaling_net.linear]
    for r in (net.vit_router, net.Unet_router):
        if router_convs:
            mods += [r.hard_route[0], r.hard_route[3], r.hard_route[6]]
        mods += [r.time_linear, r.linear]
    for a in (net.cross_attn, net.cross_attn_text):
        mods += [a.q_proj, a.k_proj, a.v_proj, a.out_proj]
    mods += [net.gate1, net.gate2, net.output_proj]
    return PreparedGroup(mods, [1.0] * len(mods))
