"""Router.hard_route (models/model_components.py:100-112,141) of the ViT router and the U-Net router on the tcgen05
grouped convolution kernels -- the B200-native replacement of the three cuDNN TF32 convolutions + GroupNorm + ReLU
chains that make up 1/3 - 1/2 of the denoiser's per-sample FLOPs (2 x 491 MFLOP per sample at 32x32).

Both routers have the same trunk shape (C -> 2C -> 4C -> 4C, 3x3), so they run as ONE grouped launch per layer: the
rows of the two router inputs are concatenated ([in_vit rows | in_unet rows], "expert" = router), exactly the layout the
dispatch plan gives the heterogeneous experts, and the dense (one kernel size, N = 64 / 128) layers go through
`gconv2_fwd_kernel` (forward and data gradient) and `gwgrad2_kernel` (weight gradient; Cout = 128 as two 64-channel
passes).  Activations are NHWC bf16 between the layers; GroupNorm(1, C) + ReLU (+ the global average pool after the last
layer) is one kernel per layer with fp32 statistics and an fp32 pooled output, so the router tail (adaLN, logits,
softmax, top-k: ops.router_gate) sees fp32 features.  Weights: one multi-tensor W-PREP launch (forced weight norm in
training, normalise, scale, cast to the tap-major bf16 operands of the forward and of the data gradient); the weight
gradients come back through one multi-tensor launch, like the U-Net experts (grouped.py).

Used when the expert path runs in bf16 (the benchmarked configuration); the fp32 configuration keeps the library
convolutions + fp32 GroupNorm kernels (model_components.Router._fused_trunk) for 1e-5-grade parity.
"""
from typing import List

import torch

from . import nhwc
from . import ops
from .grouped import _ConvLayer, _GConvFn

_ENABLED = [True]


def set_router_tcgen05_trunk(enabled: bool) -> None:
    _ENABLED[0] = bool(enabled)


def enabled() -> bool:
    return _ENABLED[0]


class _TrunkPlan:
    """The fixed 'dispatch plan' of the grouped trunk: router e owns rows [e * B, (e + 1) * B)."""

    def __init__(self, E: int, B: int, device):
        self.row_expert = torch.arange(E, dtype=torch.int32, device=device).repeat_interleave(B).contiguous()
        self.n_rows_dev = torch.tensor([E * B], dtype=torch.int32, device=device)


class _PrepTrunk(torch.autograd.Function):
    """Forward: the W-PREP launch of the 3 x E trunk convolution weights.  Backward (runs after every convolution's
    backward through the token dependency): accumulated operand gradients -> master-weight gradients, one launch."""

    @staticmethod
    def forward(ctx, runner, training, *params):
        runner._run_prep(training)
        ctx.runner, ctx.gen = runner, runner.generation
        return torch.zeros(1, device=params[0].device)

    @staticmethod
    def backward(ctx, g_token):
        ctx.runner.check_generation(ctx.gen)
        return (None, None, *ctx.runner._run_prep_backward())


class GroupedRouterTrunk:
    CONVS, NORMS = (0, 3, 6), (1, 4, 7)

    def __init__(self, routers):
        self.routers = list(routers)
        self.E = len(self.routers)
        self.layers: List[_ConvLayer] = [
            _ConvLayer(f"hard_route.{i}", [r.hard_route[i] for r in self.routers], 1.0) for i in self.CONVS]
        self._built_for = None
        self._plans = {}
        self._wg_stream, self._wg_pending = None, False
        self.plan = None
        self.generation = 0

    @staticmethod
    def supported(routers, x) -> bool:
        """Shape constraints of the tcgen05 kernels for this trunk (include/hdmoe_gemm.h)."""
        r0 = routers[0]
        c = r0.hard_route[0].weights.shape[1]
        H, W = x.shape[-2], x.shape[-1]
        same = all(tuple(r.hard_route[i].weights.shape) == tuple(r0.hard_route[i].weights.shape)
                   for r in routers for i in GroupedRouterTrunk.CONVS)
        ks = all(r0.hard_route[i].weights.shape[-1] == 3 and r0.hard_route[i].weights.shape[-2] == 3
                 for i in GroupedRouterTrunk.CONVS)
        norms = all(isinstance(r0.hard_route[i], torch.nn.GroupNorm) and r0.hard_route[i].num_groups == 1
                    and ops.gn1_relu_supported(r0.hard_route[i].num_channels) for i in GroupedRouterTrunk.NORMS)
        return (x.is_cuda and same and ks and norms and c == 32 and H % 4 == 0 and H <= 248 and W <= 232)

    def check_generation(self, gen: int) -> None:
        if gen != self.generation:
            raise RuntimeError("hdmoe_b200 router trunk: backward of forward #%d after forward #%d ran (one backward per "
                               "forward)" % (gen, self.generation))

    # ------------------------------------------------------------------------------------------ buffers
    def _params(self):
        return [c.weights for L_ in self.layers for c in L_.convs]

    def _build(self, device):
        bf = dict(dtype=torch.bfloat16, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        n_fwd = sum(L_.rows_total * L_.cin_pad for L_ in self.layers)
        n_bwd = sum(L_.rows_total_t * L_.cout for L_ in self.layers)
        self.w_fwd_all, self.w_bwd_all = torch.zeros(n_fwd, **bf), torch.zeros(n_bwd, **bf)
        self.dw_all = torch.zeros(n_fwd, **f32)
        o1 = o2 = 0
        for L_ in self.layers:
            n1, n2 = L_.rows_total * L_.cin_pad, L_.rows_total_t * L_.cout
            L_.w_fwd = self.w_fwd_all[o1:o1 + n1].view(L_.rows_total, L_.cin_pad)
            L_.dw = self.dw_all[o1:o1 + n1].view(L_.rows_total, L_.cin_pad)
            L_.w_bwd = self.w_bwd_all[o2:o2 + n2].view(L_.rows_total_t, L_.cout)
            o1 += n1
            o2 += n2
        self.grad_flat = torch.zeros(sum(p_.numel() for p_ in self._params()), **f32)
        self._wp = self._wpb = None
        self._wp_sig = self._wpb_sig = None
        self._built_for = device

    def _entries(self):
        ent = []
        for L_ in self.layers:
            for e, c in enumerate(L_.convs):
                k = L_.ks[e]
                ent.append(dict(w=c.weights, gain=1.0, layout="taps", cin_pad=L_.cin_pad,
                                out=L_.w_fwd[L_.wrow[e]:L_.wrow[e] + k * k * L_.cout],
                                layout2="taps_t", cin_rows=L_.cin_rows, cout_pad=L_.cout,
                                out2=L_.w_bwd[L_.wrow_t[e]:L_.wrow_t[e] + k * k * L_.cin_rows]))
        return ent

    def _run_prep(self, training):
        dev = self._built_for
        if self._wp is None:
            self._wp = ops.WeightPrep(self._entries(), dev)
        self._wp.entries = self._entries()
        sig = self._wp.signature()
        if sig != self._wp_sig:
            self._wp.upload()
            self._wp_sig = sig
        self._wp.run_uploaded(force=training)
        self.dw_all.zero_()

    def _run_prep_backward(self):
        from . import prepared
        if self._wg_pending:
            torch.cuda.current_stream().wait_stream(self._wg_stream)
            self._wg_pending = False
        params = self._params()
        prepared.unalias_grads(params, self.grad_flat)
        views, o = [], 0
        for p_ in params:
            views.append(self.grad_flat[o:o + p_.numel()].view_as(p_))
            o += p_.numel()
        ent, it = [], iter(views)
        for L_ in self.layers:
            for e, c in enumerate(L_.convs):
                k = L_.ks[e]
                ent.append(dict(w=c.weights, d_w_hat=L_.dw[L_.wrow[e]:L_.wrow[e] + k * k * L_.cout], d_w=next(it), gain=1.0,
                                layout="taps", cin_pad=L_.cin_pad))
        if self._wpb is None:
            self._wpb = ops.WeightPrepBackward(ent, self._built_for)
        self._wpb.entries = ent
        sig = self._wpb.signature()
        if sig != self._wpb_sig:
            self._wpb.upload()
            self._wpb_sig = sig
        self._wpb.run_uploaded()
        return prepared.deliver_grads(params, views)

    # ------------------------------------------------------------------------------------------ weight gradient
    def weight_grad(self, layer, x, dy, plan=None):
        p = plan if plan is not None else self.plan
        ops.gconv_wgrad_raw(x, dy, layer.dw, p.row_expert, p.n_rows_dev, layer.ks, layer.wrow)

    def weight_grad_async(self, layer, x, dy, plan=None):
        from .grouped import _WGRAD_STREAM
        if not _WGRAD_STREAM[0]:
            return self.weight_grad(layer, x, dy, plan)
        cur = torch.cuda.current_stream()
        if self._wg_stream is None:
            self._wg_stream = torch.cuda.Stream(device=x.device)
        side = self._wg_stream
        side.wait_stream(cur)
        x.record_stream(side)
        dy.record_stream(side)
        with torch.cuda.stream(side):
            self.weight_grad(layer, x, dy, plan)
        self._wg_pending = True

    # ------------------------------------------------------------------------------------------ forward
    def __call__(self, xs, training: bool, pre_nhwc=None):
        """xs: one [B, C, H, W] router input per router (any float dtype) -> list of fp32 pooled features [B, 4C].
        pre_nhwc: the inputs already concatenated and channels-last, bf16 [E * B, H, W, C] (ops.scale_pair writes it in
        the same pass that scales the features); xs is ignored then."""
        if pre_nhwc is not None:
            xs = [pre_nhwc]
        dev = xs[0].device
        B = xs[0].shape[0] // (self.E if pre_nhwc is not None else 1)
        if self._built_for != dev:
            self._build(dev)
        key = (B, dev)
        if key not in self._plans:
            self._plans[key] = _TrunkPlan(self.E, B, dev)
        plan = self.plan = self._plans[key]
        need_grad = torch.is_grad_enabled() and (any(x.requires_grad for x in xs)
                                                 or any(p_.requires_grad for p_ in self._params()))
        if need_grad:
            self.generation += 1
            token = _PrepTrunk.apply(self, training, *self._params())
        else:
            self._run_prep(training)
            token = None
        if pre_nhwc is not None:
            h = pre_nhwc
        else:
            rows = torch.cat([x.to(torch.bfloat16) for x in xs], dim=0).contiguous()
            h = nhwc.rows_to_nhwc(rows, self.layers[0].cin_pad, ones_channel=False)
        for li, (L_, ni) in enumerate(zip(self.layers, self.NORMS)):
            if need_grad:
                h = _GConvFn.apply(h, token, self, li)
            else:
                h = ops.gconv_raw(h, L_.w_fwd, L_.cout, L_.rows_total, plan.row_expert, plan.n_rows_dev, L_.ks, L_.wrow)
            gns = [r.hard_route[ni] for r in self.routers]
            gamma = torch.stack([g.weight for g in gns])
            beta = torch.stack([g.bias for g in gns])
            h = ops.gn1_relu_nhwc(h, gamma, beta, gns[0].eps, pool=(li == len(self.layers) - 1))
        return list(h.view(self.E, B, -1).unbind(0))
