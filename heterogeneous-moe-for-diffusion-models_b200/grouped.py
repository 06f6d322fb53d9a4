"""Grouped execution of the heterogeneous U-Net experts (the B200-native replacement of the reference's
`for expert in experts` loop, models/model_config2.py:25-37, for `Unet_expert` modules).

All experts of the MoE layer share one layer graph (models/model_components.py:281-433) and differ in the
kernel size of their k x k convolutions.  Rows arrive in the dispatch plan's expert-major order, so every layer
is ONE launch over all rows: the tcgen05 grouped implicit-GEMM convolution looks up each 128-pixel tile's
expert (kernel size, weight block) on the device.  Launch shapes are static (cap = T*k rows) and the live row
count is read on the device, so the path needs no host synchronisation.

Weights: one multi-tensor W-PREP launch per forward prepares every expert weight (forced weight norm in
training, quirk Q6, only for experts that received rows; normalise; scale by gain/sqrt(fan_in); cast to bf16;
write the tap-major K-major operand layouts for the forward and the data-gradient convolution).  Gradients
flow back through a single autograd node (`_PrepWeights`) that turns the accumulated weight-operand gradients
into master-weight gradients with one multi-tensor launch.

Activations are NHWC bf16 inside the expert; elementwise glue (pixel-norm, mp_silu, mp_sum, mp_cat, resample)
runs on whole [cap, H, W, C] tensors.  In eval mode (sampler) the emb-gain * mp_silu and the mp_sum residual
are fused into the convolution epilogue.
"""
import math
from typing import List, Optional

import torch
import torch.nn.functional as F

from . import model_components as mc
from . import model_internals as m
from . import nhwc
from . import ops

EPS = 1e-4
_WGRAD_STREAM = [True]      # weight-gradient launches on a side stream (off the data-gradient critical path)


def set_wgrad_stream(enabled: bool) -> None:
    _WGRAD_STREAM[0] = bool(enabled)


def _round_up(x, k):
    return (x + k - 1) // k * k


class _ConvLayer:
    """One convolution position of the shared layer graph: the E per-expert MP_Conv modules behind it."""

    def __init__(self, name: str, convs: List[m.MP_Conv], gain=1.0):
        self.name = name
        self.convs = convs
        w0 = convs[0].weights
        self.cout, self.cin = w0.shape[0], w0.shape[1]
        self.ks = [int(c.weights.shape[-1]) for c in convs]
        self.cin_pad = _round_up(self.cin, 32) if self.cin % 32 == 0 else _round_up(self.cin, 64)
        self.cin_rows = self.cin if self.cin % 32 == 0 else (self.cin // 32) * 32   # data-gradient channels
        self.gain = gain                      # float, or list of per-expert 0-dim Parameters (out_gain)
        self.wrow, self.wrow_t = [], []
        r = rt = 0
        for k in self.ks:
            self.wrow.append(r)
            self.wrow_t.append(rt)
            r += k * k * self.cout
            rt += k * k * self.cin_rows
        self.rows_total, self.rows_total_t = r, rt
        self.w_fwd = self.w_bwd = None        # bf16 operand buffers (views)
        self.dw = None                        # fp32 gradient accumulator [rows_total, cin_pad], tap-major blocks


class _GConvFn(torch.autograd.Function):
    """y = conv(x, W_e(row)) for every row; linear op (training path).  backward: data gradient through the same
    tcgen05 kernel with the transposed/flipped operand; weight gradient accumulated into the layer's buffers."""

    @staticmethod
    def forward(ctx, x, token, runner, li):
        layer = runner.layers[li]
        plan = runner.plan
        y = ops.gconv_raw(x, layer.w_fwd, layer.cout, layer.rows_total, plan.row_expert, plan.n_rows_dev, layer.ks,
                          layer.wrow)
        ctx.save_for_backward(x)
        ctx.runner, ctx.li = runner, li
        ctx.plan, ctx.gen = plan, runner.generation
        return y

    @staticmethod
    def backward(ctx, dy):
        # The `token` input only orders the backward pass: its edge makes the W-PREP backward node wait for this node
        # (autograd counts dependencies on edges, whatever flows through them), so its gradient is None -- a zeros tensor
        # per convolution cost 44 fill + 38 accumulate launches per step inside the expert backward.
        (x,) = ctx.saved_tensors
        runner, layer = ctx.runner, ctx.runner.layers[ctx.li]
        runner.check_generation(ctx.gen)
        plan = ctx.plan
        dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.gconv_raw(dy, layer.w_bwd, layer.cin_rows, layer.rows_total_t, plan.row_expert, plan.n_rows_dev,
                               layer.ks, layer.wrow_t)
            if layer.cin_rows < layer.cin_pad:
                dx = F.pad(dx, (0, layer.cin_pad - layer.cin_rows))
        runner.weight_grad_async(layer, x, dy, plan)
        return dx, None, None, None


class _PrepWeights(torch.autograd.Function):
    """Forward: the multi-tensor W-PREP launch.  Backward: runs after every consumer (autograd dependency
    order) and converts the accumulated operand gradients into master-weight gradients in one launch."""

    @staticmethod
    def forward(ctx, runner, training, *params):
        runner._run_prep(training)
        ctx.runner, ctx.gen = runner, runner.generation
        token = torch.zeros(1, device=params[0].device)
        outs = (token, runner.lin_noise.clone(), runner.lin_text.clone(), runner.lin_emb.clone())
        return outs

    @staticmethod
    def backward(ctx, g_token, g_noise, g_text, g_emb):
        ctx.runner.check_generation(ctx.gen)
        grads = ctx.runner._run_prep_backward(g_noise, g_text, g_emb)
        return (None, None, *grads)


class GroupedUnetExperts:
    """Runs a ModuleList of structurally identical `Unet_expert`s layer by layer with grouped kernels."""

    def __init__(self, experts):
        assert len(experts) >= 1 and all(isinstance(e, mc.Unet_expert) for e in experts)
        self.experts = list(experts)
        self.E = len(self.experts)
        e0 = self.experts[0]
        names_enc, names_dec = list(e0.encoders.keys()), list(e0.decoders.keys())
        for e in self.experts:
            assert list(e.encoders.keys()) == names_enc and list(e.decoders.keys()) == names_dec
        self.layers: List[_ConvLayer] = []
        self.program = []          # ('conv', li) | ('block', spec)
        self.emb_slices = []       # (offset, cout) of every block's emb_layer inside the stacked emb buffer
        emb_off = 0

        def add_conv(name, getter, gain=1.0):
            self.layers.append(_ConvLayer(name, [getter(e) for e in self.experts], gain))
            return len(self.layers) - 1

        def add_block(group, name):
            nonlocal emb_off
            b0 = getattr(e0, group)[name]
            spec = dict(name=f"{group}.{name}", type=b0.type, resample=b0.resample, t=b0.residual_balance,
                        dropout=b0.dropout, emb_gain=b0.emb_gain, cat="block" in name and group == "decoders")
            spec["skip"] = (add_conv(f"{group}.{name}.conv_skip", lambda e: getattr(e, group)[name].conv_skip)
                            if b0.conv_skip is not None else None)
            spec["res1"] = add_conv(f"{group}.{name}.conv_res1", lambda e: getattr(e, group)[name].conv_res1,
                                    b0.conv_gain1)
            spec["res2"] = add_conv(f"{group}.{name}.conv_res2", lambda e: getattr(e, group)[name].conv_res2,
                                    b0.conv_gain2)
            spec["emb"] = (emb_off, b0.out_channels)
            spec["emb_mods"] = [getattr(e, group)[name].emb_layer for e in self.experts]
            emb_off += b0.out_channels
            self.program.append(("block", spec))

        for name in names_enc:
            if "conv" in name:
                self.program.append(("conv", add_conv(f"encoders.{name}", lambda e: e.encoders[name])))
            else:
                add_block("encoders", name)
        for name in names_dec:
            add_block("decoders", name)
        self.out_li = add_conv("out_conv", lambda e: e.out_conv, [e.out_gain for e in self.experts])
        self.emb_total = emb_off
        self.emb_size = e0.emb_size
        self.has_text = e0.map_text is not None
        self._built_for = None
        self._wg_stream, self._wg_pending = None, False
        self.plan = None
        # The operand buffers, the weight-gradient accumulator and the W-PREP tables are persistent (pointer-stable for
        # CUDA-graph replay) and therefore hold the state of ONE forward: `generation` counts forwards that took the
        # autograd path, and every backward node checks that it belongs to the latest one.
        self.generation = 0

    def check_generation(self, gen: int) -> None:
        if gen != self.generation:
            raise RuntimeError(
                "hdmoe_b200 grouped experts: backward of forward #%d after forward #%d ran -- the grouped path keeps the "
                "prepared weights and gradient accumulators of one forward at a time (one backward per forward; for "
                "several micro-batches call backward before the next forward, or use set_grouped_experts(False))"
                % (gen, self.generation))

    # ------------------------------------------------------------------------------------------ buffers
    def _params(self):
        ps = []
        for L_ in self.layers:
            ps += [c.weights for c in L_.convs]
        ps += [e.out_gain for e in self.experts]
        ps += [e.map_noise.weights for e in self.experts]
        if self.has_text:
            ps += [e.map_text.weights for e in self.experts]
        for kind, spec in self.program:
            if kind == "block":
                ps += [md.weights for md in spec["emb_mods"]]
        return ps

    def _build(self, device):
        E = self.E
        bf = dict(dtype=torch.bfloat16, device=device)
        f32 = dict(dtype=torch.float32, device=device)
        n_fwd = sum(L_.rows_total * L_.cin_pad for L_ in self.layers)
        n_bwd = sum(L_.rows_total_t * L_.cout for L_ in self.layers)
        self.w_fwd_all = torch.zeros(n_fwd, **bf)
        self.w_bwd_all = torch.zeros(n_bwd, **bf)
        self.dw_all = torch.zeros(n_fwd, **f32)
        o1 = o2 = o3 = 0
        for L_ in self.layers:
            L_.w_fwd = self.w_fwd_all[o1:o1 + L_.rows_total * L_.cin_pad].view(L_.rows_total, L_.cin_pad)
            o1 += L_.rows_total * L_.cin_pad
            L_.w_bwd = self.w_bwd_all[o2:o2 + L_.rows_total_t * L_.cout].view(L_.rows_total_t, L_.cout)
            o2 += L_.rows_total_t * L_.cout
            L_.dw = self.dw_all[o3:o3 + L_.rows_total * L_.cin_pad].view(L_.rows_total, L_.cin_pad)
            o3 += L_.rows_total * L_.cin_pad
        e0 = self.experts[0]
        tdim = e0.map_noise.weights.shape[1]
        self.lin_noise = torch.empty(E, self.emb_size, tdim, **f32)
        xdim = e0.map_text.weights.shape[1] if self.has_text else 1
        self.lin_text = torch.empty(E, self.emb_size, xdim, **f32)
        self.lin_emb = torch.empty(E, self.emb_total, self.emb_size, **f32)
        # persistent buffers so that every launch argument is pointer-stable (CUDA-graph replay)
        self.g_noise, self.g_text, self.g_emb = (torch.zeros_like(t) for t in (self.lin_noise, self.lin_text, self.lin_emb))
        self.active_buf = torch.ones(E, dtype=torch.int32, device=device)
        self.gain_grads = torch.zeros(E, **f32)
        self.grad_flat = torch.zeros(sum(p_.numel() for p_ in self._params()), **f32)
        self._wp = self._wpb = None
        self._wp_sig = self._wpb_sig = None
        self._built_for = device

    def _entries(self):
        """W-PREP descriptor entries (pointer-stable: the per-expert activity flags live in a persistent buffer)."""
        act = [self.active_buf[e:e + 1] for e in range(self.E)]
        ent = []
        for L_ in self.layers:
            for e, c in enumerate(L_.convs):
                k = L_.ks[e]
                g = L_.gain[e] if isinstance(L_.gain, list) else L_.gain
                ent.append(dict(w=c.weights, gain=g, active=act[e], layout="taps", cin_pad=L_.cin_pad,
                                out=L_.w_fwd[L_.wrow[e]:L_.wrow[e] + k * k * L_.cout],
                                layout2="taps_t", cin_rows=L_.cin_rows, cout_pad=L_.cout,
                                out2=L_.w_bwd[L_.wrow_t[e]:L_.wrow_t[e] + k * k * L_.cin_rows]))
        for e, ex in enumerate(self.experts):
            ent.append(dict(w=ex.map_noise.weights, out=self.lin_noise[e], active=act[e]))
            if self.has_text:
                ent.append(dict(w=ex.map_text.weights, out=self.lin_text[e], active=act[e]))
        for kind, spec in self.program:
            if kind == "block":
                off, co = spec["emb"]
                for e, md in enumerate(spec["emb_mods"]):
                    ent.append(dict(w=md.weights, out=self.lin_emb[e, off:off + co], gain=spec["emb_gain"], active=act[e]))
        return ent

    def _run_prep(self, training):
        dev = self.plan.counts.device
        self.active_buf.copy_(self.plan.counts)
        wp = ops.WeightPrep(self._entries(), dev) if self._wp is None else self._wp
        sig = None
        if self._wp is not None:
            self._wp.entries = self._entries()
            sig = self._wp.signature()
        if self._wp is None or sig != self._wp_sig:
            # (re)upload the descriptor table: first call, or parameters moved (.to(), optimizer state swap)
            self._wp = wp
            wp.upload()
            self._wp_sig = wp.signature()
        self._wp.run_uploaded(force=training)
        self.dw_all.zero_()

    def _bwd_entries(self):
        ent, o = [], 0
        params = self._params()
        views = []
        for p_ in params:
            views.append(self.grad_flat[o:o + p_.numel()].view_as(p_))
            o += p_.numel()
        it = iter(views)
        for L_ in self.layers:
            for e, c in enumerate(L_.convs):
                g = L_.gain[e] if isinstance(L_.gain, list) else L_.gain
                k = L_.ks[e]
                ent.append(dict(w=c.weights, d_w_hat=L_.dw[L_.wrow[e]:L_.wrow[e] + k * k * L_.cout], d_w=next(it), gain=g,
                                layout="taps", cin_pad=L_.cin_pad,
                                d_gain=self.gain_grads[e:e + 1] if isinstance(L_.gain, list) else None))
        self._gain_views = [next(it) for _ in range(self.E)]
        for e, ex in enumerate(self.experts):
            ent.append(dict(w=ex.map_noise.weights, d_w_hat=self.g_noise[e], d_w=next(it)))
        if self.has_text:
            for e, ex in enumerate(self.experts):
                ent.append(dict(w=ex.map_text.weights, d_w_hat=self.g_text[e], d_w=next(it)))
        for kind, spec in self.program:
            if kind == "block":
                off, co = spec["emb"]
                for e, md in enumerate(spec["emb_mods"]):
                    ent.append(dict(w=md.weights, d_w_hat=self.g_emb[e, off:off + co], d_w=next(it), gain=spec["emb_gain"]))
        return ent, views

    def _run_prep_backward(self, g_noise, g_text, g_emb):
        """Accumulated operand gradients -> master-weight gradients, ONE multi-tensor launch.  The returned
        gradients are views of a persistent buffer (call optimizer.zero_grad(set_to_none=True), the default)."""
        dev = self.plan.counts.device
        if self._wg_pending:
            torch.cuda.current_stream().wait_stream(self._wg_stream)
            self._wg_pending = False
        from . import prepared
        prepared.unalias_grads(self._params(), self.grad_flat)
        for buf, g in ((self.g_noise, g_noise), (self.g_text, g_text), (self.g_emb, g_emb)):
            if g is None:
                buf.zero_()
            else:
                buf.copy_(g)
        self.gain_grads.zero_()
        ent, views = self._bwd_entries()
        if self._wpb is None:
            self._wpb = ops.WeightPrepBackward(ent, dev)
        self._wpb.entries = ent
        sig = self._wpb.signature()
        if sig != self._wpb_sig:
            self._wpb.upload()
            self._wpb_sig = sig
        self._wpb.run_uploaded()
        for e in range(self.E):
            self._gain_views[e].copy_(self.gain_grads[e])
        return prepared.deliver_grads(self._params(), views)

    # ------------------------------------------------------------------------------------------ wgrad
    def weight_grad(self, layer: _ConvLayer, x, dy, plan=None):
        """layer.dw (fp32, tap-major blocks) += grouped weight gradient: one tcgen05 launch for all experts."""
        p = plan if plan is not None else self.plan
        ops.gconv_wgrad_raw(x, dy, layer.dw, p.row_expert, p.n_rows_dev, layer.ks, layer.wrow)

    def weight_grad_async(self, layer, x, dy, plan=None):
        """Weight gradient on a side stream: nothing on the data-gradient chain depends on it; the join is in
        _run_prep_backward (the W-PREP backward node runs after every convolution's backward)."""
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
    def _conv(self, x, li, token, training, scale=None, act=0, residual=None, res_a=0.0, res_b=1.0):
        layer = self.layers[li]
        if training:
            assert scale is None and act == 0 and residual is None
            return _GConvFn.apply(x, token, self, li)
        p = self.plan
        return ops.gconv_raw(x, layer.w_fwd, layer.cout, layer.rows_total, p.row_expert, p.n_rows_dev, layer.ks,
                             layer.wrow, scale=scale, act=act, residual=residual, res_a=res_a, res_b=res_b)

    @staticmethod
    def _resample(x, mode):
        if mode == "keep":
            return x
        R, H, W, C = x.shape
        if mode == "down":
            return x.view(R, H // 2, 2, W // 2, 2, C).float().mean(dim=(2, 4)).to(x.dtype)
        return x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)

    def _select(self, out_all, idx):
        return out_all[torch.arange(out_all.shape[0], device=out_all.device), idx]

    def __call__(self, plan, x_rows, time_rows, text_rows, training: bool, nhwc_out: bool = False):
        """x_rows [cap, C, H, W] (dispatch payload, any float dtype) -> [cap, C, H, W] bf16 ([cap, H, W, C] when
        nhwc_out: the channels-last trunk consumes the expert output without a layout change)."""
        dev = x_rows.device
        if self._built_for != dev:
            self._build(dev)
        self.plan = plan
        # autograd path whenever a gradient can flow (independent of .training: eval-mode forwards under autograd
        # propagate to the input and the expert parameters like the reference); `training` only selects the forced
        # weight normalisation (Q6) and dropout
        need_grad = torch.is_grad_enabled() and (x_rows.requires_grad or time_rows.requires_grad
                                                 or any(p_.requires_grad for p_ in self._params()))
        if need_grad:
            self.generation += 1
            token, Wn, Wt, We = _PrepWeights.apply(self, training, *self._params())
        else:
            self._run_prep(training)
            token, Wn, Wt, We = None, self.lin_noise, self.lin_text, self.lin_emb
        idx = plan.row_expert.clamp(min=0).long()
        # conditioning: emb = mp_silu(mp_sum(map_noise(t), map_text(txt), label_balance)); every block's 1+emb_layer
        t32, e0 = time_rows.float(), self.experts[0]
        emb = self._select(torch.einsum("rk,eok->reo", t32, Wn), idx)
        if self.has_text and text_rows is not None:
            txt = self._select(torch.einsum("rk,eok->reo", text_rows.float(), Wt), idx)
            emb = m.mp_sum(emb, txt, t=e0.label_balance)
        emb = m.mp_silu(emb)
        gains = 1 + self._select(torch.einsum("rk,eok->reo", emb, We), idx)        # [cap, sum Cout_b] fp32
        # input: NHWC, ones channel appended (models/model_components.py:416), zero-padded to the K chunk
        x = nhwc.rows_to_nhwc(x_rows.to(torch.bfloat16).contiguous(), self.layers[0].cin_pad)
        use = need_grad
        skips = []
        for kind, item in self.program:
            if kind == "conv":
                x = self._conv(x, item, token, use)
                skips.append(x)
                continue
            s = item
            if s["cat"]:
                x = nhwc.mp_cat(x.contiguous(), skips.pop().contiguous(), e0.concat_balance)
            x = self._resample(x, s["resample"])
            off, co = s["emb"]
            g = gains[:, off:off + co]
            t = s["t"]
            c = math.sqrt((1 - t) ** 2 + t ** 2)
            x = x.contiguous()
            if s["type"] == "enc":
                if s["skip"] is not None:
                    x = self._conv(x, s["skip"], token, use)
                x, a = nhwc.pixnorm_silu(x)
            else:
                a = nhwc.gain_silu(x)
            if use:
                y = self._conv(a, s["res1"], token, True)
                y = nhwc.gain_silu(y, g)
                if training and s["dropout"]:
                    y = F.dropout(y, p=s["dropout"])
                y = self._conv(y, s["res2"], token, True)
                if s["type"] == "dec" and s["skip"] is not None:
                    x = self._conv(x, s["skip"], token, True)
                x = nhwc.mp_sum(x, y, t)
            else:   # fused epilogues: mp_silu(conv * (1+emb)) and mp_sum(x, conv, t)
                y = self._conv(a, s["res1"], token, False, scale=g, act=1)
                if training and s["dropout"]:             # train mode under no_grad still drops (as the reference)
                    y = F.dropout(y, p=s["dropout"])
                if s["type"] == "dec" and s["skip"] is not None:
                    x = self._conv(x, s["skip"], token, False)
                x = self._conv(y, s["res2"], token, False, residual=x.contiguous(), res_a=(1 - t) / c, res_b=t / c)
            if "encoders" in s["name"]:
                skips.append(x)
        x = self._conv(x, self.out_li, token, use)
        return x if nhwc_out else nhwc.nhwc_to_rows(x)
