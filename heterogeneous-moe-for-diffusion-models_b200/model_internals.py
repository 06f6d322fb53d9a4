"""L0 primitives of the denoiser, B200 host side (mirrors the public names of the reference's
models/model_internals.py so the path drops in: normalize, mp_silu, mp_sum, mp_cat, resample, MP_Fourier,
MP_Conv, MP_Attention).  Dense trunk math uses stock torch CUDA ops; the MoE hot path (router gate,
permute, grouped expert convolutions, combine, EDM step) runs the kernels in csrc/ through ops.py."""
import math
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

EPS = 1e-4
_SILU_GAIN = 1.0 / 0.596

# Device-side "this expert received rows" flag (0-dim bool tensor) for sync-free MoE execution: the reference
# only runs -- and therefore only force-normalises the weights of -- experts with at least one routed sample.
_ACTIVE = [None]
_FUSED_ATTN = [True]     # trunk head_dim-4 attention through csrc/attention_tc.cu (False: library SDPA, for A/B tests)


def set_fused_attention(enabled: bool) -> None:
    _FUSED_ATTN[0] = bool(enabled)


class active_flag:
    def __init__(self, flag):
        self.flag = flag

    def __enter__(self):
        self.old = _ACTIVE[0]
        _ACTIVE[0] = self.flag

    def __exit__(self, *a):
        _ACTIVE[0] = self.old


def normalize(x: torch.Tensor, dim: Optional[Sequence[int]] = None, eps: float = EPS) -> torch.Tensor:
    """RMS-normalise: x / (eps + ||x|| * sqrt(n_norm / n_x)); ref models/model_internals.py:8-30."""
    dims = list(range(1, x.ndim)) if dim is None else list(dim)
    n = torch.linalg.vector_norm(x, dim=dims, keepdim=True, dtype=torch.float32)
    n = torch.add(eps, n, alpha=math.sqrt(n.numel() / x.numel()))
    return x / n.to(x.dtype)


def mp_silu(x: torch.Tensor) -> torch.Tensor:
    """ref models/model_internals.py:33-47"""
    return F.silu(x) * _SILU_GAIN if x.dtype != torch.float32 else F.silu(x) / 0.596


def mp_sum(a: torch.Tensor, b: torch.Tensor, t: float = 0.5) -> torch.Tensor:
    """ref models/model_internals.py:50-66"""
    return torch.lerp(a, b, t) / math.sqrt((1 - t) ** 2 + t ** 2)


def mp_cat(a: torch.Tensor, b: torch.Tensor, dim: int = 1, t: float = 0.5) -> torch.Tensor:
    """ref models/model_internals.py:69-92"""
    na, nb = a.shape[dim], b.shape[dim]
    c = math.sqrt((na + nb) / ((1 - t) ** 2 + t ** 2))
    return torch.cat([a * (c * (1 - t) / math.sqrt(na)), b * (c * t / math.sqrt(nb))], dim=dim)


def resample(x: torch.Tensor, f=(1, 1), mode: str = "keep") -> torch.Tensor:
    """2x box resampling; ref models/model_internals.py:95-127.  With the only filter the model uses
    (f = [1, 1]) 'down' is a 2x2 mean and 'up' replicates each pixel into a 2x2 block."""
    if mode == "keep":
        return x
    if list(f) != [1, 1]:
        raise ValueError("hdmoe_b200.resample implements the [1, 1] box filter used by the denoiser")
    if mode == "down":
        return F.avg_pool2d(x, 2)
    if mode == "up":
        return x.repeat_interleave(2, dim=2).repeat_interleave(2, dim=3)
    raise ValueError(f"Invalid mode: {mode}")


class MP_Fourier(nn.Module):
    """ref models/model_internals.py:130-175"""

    def __init__(self, num_channels: int, bandwidth: float = 1):
        super().__init__()
        self.register_buffer("freqs", 2 * torch.pi * torch.randn(num_channels) * bandwidth)
        self.register_buffer("phases", 2 * torch.pi * torch.rand(num_channels))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = torch.outer(x.to(torch.float32), self.freqs.to(torch.float32)) + self.phases.to(torch.float32)
        return (y.cos() * math.sqrt(2)).to(x.dtype)


class MP_Conv(nn.Module):
    """Magnitude-preserving conv / linear (ref models/model_internals.py:209-275): weight rows are
    normalised on every call (and, in training mode, rewritten in place first -- quirk Q6), scaled by
    gain/sqrt(fan_in) in fp32 and cast to the activation dtype; 'same' padding is asymmetric for even k."""

    def __init__(self, in_channels: int, out_channels: int, kernel: tuple, stride: int = 1):
        super().__init__()
        self.out_channels = out_channels
        self.weights = nn.Parameter(torch.randn(out_channels, in_channels, *kernel))
        assert self.weights.numel() != 0
        self.kernel = kernel
        self.stride = stride

    def prepared_weight(self, gain=1.0, dtype=torch.float32) -> torch.Tensor:
        from . import prepared
        pw = prepared.lookup(self, gain)      # jointly prepared by a PreparedGroup (one launch for many layers)
        if pw is not None:
            return pw if pw.dtype == dtype else pw.to(dtype)
        w = self.weights.to(torch.float32)
        if self.training:
            with torch.no_grad():
                wn = normalize(w)
                self.weights.copy_(wn if _ACTIVE[0] is None else torch.where(_ACTIVE[0], wn, w))
        w = normalize(w)
        w = w * (gain / math.sqrt(w[0].numel()))
        return w.to(dtype)

    def forward(self, x: torch.Tensor, gain=1.0) -> torch.Tensor:
        w = self.prepared_weight(gain, x.dtype)
        if x.ndim == 2:
            return F.linear(x, w)
        assert x.ndim == 4
        k = w.shape[-1]
        if x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
            w = w.contiguous(memory_format=torch.channels_last)
        if self.stride != 1:
            return F.conv2d(x, w, padding=k // 2, stride=self.stride)
        lo = (k - 1) // 2
        hi = k - 1 - lo
        if lo == hi:
            return F.conv2d(x, w, padding=lo)
        return F.conv2d(F.pad(x, (lo, hi, lo, hi)), w)


class MP_Attention(nn.Module):
    """Multi-head attention over the sequence with MP projections (ref models/model_internals.py:279-409).
    Scores are S_q x S_k per head (quirk Q17); self-attention adds a learned rel_pos_bias (sliced or
    bicubically resized, :382-399) and time projections."""

    def __init__(self, num_heads: int, emb_dim: int, seq_ln: int, time_dim: Optional[int] = 0,
                 context_dim: Optional[int] = None, attn_balance: Optional[float] = 0.5,
                 is_cross_attn: Optional[bool] = False):
        super().__init__()
        assert emb_dim % num_heads == 0
        self.num_heads, self.emb_dim, self.head_dim = num_heads, emb_dim, emb_dim // num_heads
        self.time_emb = time_dim
        context_dim = emb_dim if context_dim is None else context_dim
        self.is_cross = is_cross_attn
        self.attn_balance = attn_balance
        self.time_dependent = time_dim > 0
        self.rel_pos_bias = nn.Parameter(torch.zeros(num_heads, seq_ln, seq_ln)) if not is_cross_attn else None
        self.q_proj = MP_Conv(emb_dim, emb_dim, kernel=(1, 1))
        self.k_proj = MP_Conv(context_dim, emb_dim, kernel=(1, 1))
        self.v_proj = MP_Conv(context_dim, emb_dim, kernel=(1, 1))
        self.q_time = MP_Conv(time_dim, emb_dim, kernel=(1, 1)) if self.time_dependent else None
        self.k_time = MP_Conv(time_dim, emb_dim, kernel=(1, 1)) if self.time_dependent and not is_cross_attn else None
        self.v_time = MP_Conv(time_dim, emb_dim, kernel=(1, 1)) if self.time_dependent and not is_cross_attn else None
        self.out_proj = MP_Conv(emb_dim, emb_dim, kernel=(1, 1))

    def _proj(self, conv: MP_Conv, x: torch.Tensor, gain) -> torch.Tensor:
        # a 1x1 convolution over (B, C, S, 1) is a linear map over the channel axis of (B, S, C)
        w = conv.prepared_weight(gain, x.dtype).flatten(1)
        if x.is_cuda and w.shape[0] == 32 and w.shape[1] == 32:
            from . import ops
            return ops.linear32(x, w)          # streaming weight-gradient kernel for the thin trunk projections
        return F.linear(x, w)

    def forward(self, query: torch.Tensor, gain_s: float, gain_t: float, context: Optional[torch.Tensor] = None,
                time_embedding: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, S, D = query.shape
        assert D == self.emb_dim
        ctx = query if context is None else context
        q = self._proj(self.q_proj, query, gain_s)
        k = self._proj(self.k_proj, ctx, gain_s)
        v = self._proj(self.v_proj, ctx, gain_s)
        if self.time_dependent and time_embedding is not None:
            te = time_embedding.reshape(B, 1, -1).to(query.dtype)
            q = q + self._proj(self.q_time, te, gain_t)
            if not self.is_cross:
                k = k + self._proj(self.k_time, te, gain_t)
                v = v + self._proj(self.v_time, te, gain_t)
        H, hd = self.num_heads, self.head_dim
        if self.is_cross and hd == 4 and H <= 8 and q.is_cuda and q.dtype == torch.float32 and _FUSED_ATTN[0]:
            # trunk cross-attention: fused flash-style kernel, no (B, heads, S_q, S_k) score tensor
            from . import ops
            o = ops.attention_d4(q, k, v, H, 1.0 / math.sqrt(hd))
            o = self._proj(self.out_proj, o, gain_s)
            return mp_sum(query, o, self.attn_balance)
        q = q.view(B, -1, H, hd).transpose(1, 2)
        k = k.view(B, -1, H, hd).transpose(1, 2)
        v = v.view(B, -1, H, hd).transpose(1, 2)
        bias = None
        if not self.is_cross:
            bias = self.rel_pos_bias
            if S <= bias.shape[1]:
                bias = bias[:, :S, :S]
            else:
                bias = F.interpolate(bias.unsqueeze(0), size=(S, S), mode="bicubic", align_corners=False).squeeze(0)
            bias = bias.to(q.dtype)
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=bias, scale=1.0 / math.sqrt(hd))
        o = o.transpose(1, 2).reshape(B, S, D)
        o = self._proj(self.out_proj, o, gain_s)
        return mp_sum(query, o, self.attn_balance)
