"""Autograd wrappers of the fused NHWC bf16 elementwise kernels (csrc/nhwc_ops.cu) that glue the grouped U-Net
expert convolutions together: the elementwise part of Unet_block.forward (models/model_components.py:232-253)
and the layout changes at the dispatch / combine boundary.  CUDA only; every function raises on CPU tensors."""
import ctypes as C
import math

import torch

from . import _lib as L
from .ops import _cuda, _p, _st

_BF = torch.bfloat16


def _chk(*ts):
    _cuda(*ts)
    for t in ts:
        if t is not None and (t.dtype != _BF or not t.is_contiguous()):
            raise RuntimeError("hdmoe_b200.nhwc: tensors must be contiguous bfloat16")


class _PixnormSilu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _chk(x)
        Cn = x.shape[-1]
        xn, a = torch.empty_like(x), torch.empty_like(x)
        L.check(L.lib().hdmoe_nhwc_pixnorm_silu_fwd(_p(x), _p(xn), _p(a), x.numel() // Cn, Cn, _st()), "pixnorm_silu_fwd")
        ctx.save_for_backward(x)
        return xn, a

    @staticmethod
    def backward(ctx, g_xn, g_a):
        (x,) = ctx.saved_tensors
        Cn = x.shape[-1]
        if g_a is None:
            g_a = torch.zeros_like(x)
        g_a = g_a.contiguous()
        g_xn = None if g_xn is None else g_xn.contiguous()
        dx = torch.empty_like(x)
        L.check(L.lib().hdmoe_nhwc_pixnorm_silu_bwd(_p(x), _p(g_xn), _p(g_a), _p(dx), x.numel() // Cn, Cn, _st()),
                "pixnorm_silu_bwd")
        return dx


def pixnorm_silu(x):
    """(normalize(x, dim=C), mp_silu(normalize(x))) for NHWC bf16 x; ref models/model_components.py:238,240."""
    return _PixnormSilu.apply(x)


class _GainSilu(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, gain):
        _chk(z)
        R, Cn = z.shape[0], z.shape[-1]
        if gain is not None:
            _cuda(gain)
            gain = gain.float().contiguous()
            assert gain.shape == (R, Cn)
        y = torch.empty_like(z)
        L.check(L.lib().hdmoe_nhwc_gain_silu_fwd(_p(z), _p(gain), _p(y), R, z.numel() // (R * Cn), Cn, _st()), "gain_silu_fwd")
        ctx.save_for_backward(z, gain)
        return y

    @staticmethod
    def backward(ctx, dy):
        z, gain = ctx.saved_tensors
        R, Cn = z.shape[0], z.shape[-1]
        dy = dy.contiguous()
        dz = torch.empty_like(z)
        dgain = torch.empty_like(gain) if gain is not None and ctx.needs_input_grad[1] else None
        L.check(L.lib().hdmoe_nhwc_gain_silu_bwd(_p(z), _p(gain), _p(dy), _p(dz), _p(dgain), R, z.numel() // (R * Cn), Cn,
                                                 _st()), "gain_silu_bwd")
        return dz, dgain


def gain_silu(z, gain=None):
    """mp_silu(z * gain[row, None, None, :]) for NHWC bf16 z, fp32 gain [R, C] (None: plain mp_silu);
    ref models/model_components.py:242-243."""
    return _GainSilu.apply(z, gain)


class _Axpby(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, ca, cb):
        _chk(x, y)
        assert x.shape == y.shape
        out = torch.empty_like(x)
        L.check(L.lib().hdmoe_nhwc_axpby(_p(x), _p(y), ca, cb, _p(out), x.numel(), _st()), "axpby")
        ctx.c = (ca, cb)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        gx, gy = torch.empty_like(g), torch.empty_like(g)
        L.check(L.lib().hdmoe_nhwc_scale2(_p(g), ctx.c[0], ctx.c[1], _p(gx), _p(gy), g.numel(), _st()), "scale2")
        return gx, gy, None, None


def mp_sum(x, y, t: float):
    """magnitude-preserving sum, ref models/model_internals.py:61-66."""
    c = math.sqrt((1 - t) ** 2 + t ** 2)
    return _Axpby.apply(x, y, (1 - t) / c, t / c)


class _Cat(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, wa, wb):
        _chk(a, b)
        Ca, Cb = a.shape[-1], b.shape[-1]
        assert a.shape[:-1] == b.shape[:-1]
        out = torch.empty(*a.shape[:-1], Ca + Cb, dtype=_BF, device=a.device)
        L.check(L.lib().hdmoe_nhwc_cat(_p(a), _p(b), wa, wb, Ca, Cb, _p(out), a.numel() // Ca, _st()), "cat")
        ctx.meta = (wa, wb, Ca, Cb)
        return out

    @staticmethod
    def backward(ctx, g):
        wa, wb, Ca, Cb = ctx.meta
        g = g.contiguous()
        ga = torch.empty(*g.shape[:-1], Ca, dtype=_BF, device=g.device)
        gb = torch.empty(*g.shape[:-1], Cb, dtype=_BF, device=g.device)
        L.check(L.lib().hdmoe_nhwc_split(_p(g), wa, wb, Ca, Cb, _p(ga), _p(gb), g.numel() // (Ca + Cb), _st()), "split")
        return ga, gb, None, None


def mp_cat(a, b, t: float):
    """magnitude-preserving channel concatenation, ref models/model_internals.py:69-92."""
    na, nb = a.shape[-1], b.shape[-1]
    c = math.sqrt((na + nb) / ((1 - t) ** 2 + t ** 2))
    return _Cat.apply(a, b, c * (1 - t) / math.sqrt(na), c * t / math.sqrt(nb))


def _to_nhwc(x, Cd, one):
    R, Cs, H, W = x.shape
    out = torch.empty(R, H, W, Cd, dtype=_BF, device=x.device)
    L.check(L.lib().hdmoe_nchw_to_nhwc(_p(x), _p(out), R, Cs, Cd, H * W, int(one), _st()), "nchw_to_nhwc")
    return out


def _to_nchw(x, Cd):
    R, H, W, Cs = x.shape
    out = torch.empty(R, Cd, H, W, dtype=_BF, device=x.device)
    L.check(L.lib().hdmoe_nhwc_to_nchw(_p(x), _p(out), R, Cs, Cd, H * W, _st()), "nhwc_to_nchw")
    return out


class _RowsToNhwc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, Cd, one):
        _chk(x)
        ctx.Cs = x.shape[1]
        return _to_nhwc(x, Cd, one)

    @staticmethod
    def backward(ctx, g):
        return _to_nchw(g.contiguous(), ctx.Cs), None, None


class _NhwcToRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _chk(x)
        return _to_nchw(x, x.shape[-1])

    @staticmethod
    def backward(ctx, g):
        return _to_nhwc(g.contiguous(), g.shape[1], False)


def rows_to_nhwc(x_rows, c_pad: int, ones_channel: bool = True):
    """[R, C, H, W] bf16 -> [R, H, W, c_pad] with a ones channel at index C (models/model_components.py:416) and
    zero padding up to c_pad."""
    return _RowsToNhwc.apply(x_rows, c_pad, ones_channel)


def nhwc_to_rows(x):
    """[R, H, W, C] bf16 -> contiguous [R, C, H, W]."""
    return _NhwcToRows.apply(x)
