"""L1 components (same public names, constructor kwargs, forward signatures and state_dict keys as the
reference's models/model_components.py): Scaling_router, Router, Unet_block, Unet_expert, Vit_block,
Vit_expert.  Router.forward's tail is ONE fused sm_100a kernel (ops.router_gate)."""
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import model_internals as m
from . import ops


_CHANNELS_LAST = [False]   # measured: no gain on B200 (cuDNN picks NHWC kernels either way)


_FUSED_TRUNK = [True]       # fused GroupNorm(1)+ReLU(+pool) kernels and channels-last convolutions in Router.hard_route


_FUSED_SCALING = [True]     # Scaling_router as one kernel per direction (csrc/trunk_glue.cu)


def set_fused_scaling_router(enabled: bool) -> None:
    _FUSED_SCALING[0] = bool(enabled)


def set_router_fused_trunk(enabled: bool) -> None:
    _FUSED_TRUNK[0] = bool(enabled)


def set_router_channels_last(enabled: bool) -> None:
    _CHANNELS_LAST[0] = bool(enabled)


class Scaling_router(nn.Module):
    """Soft 2-way gain router of model_config1 (ref models/model_components.py:7-66)."""

    def __init__(self, emb_dim: Optional[int] = 3, num_experts: Optional[int] = 2, dropout: Optional[float] = 0.2):
        super().__init__()
        self.soft_route = nn.Sequential(
            m.MP_Conv(in_channels=emb_dim, out_channels=emb_dim * 2, kernel=()),
            nn.GroupNorm(1, emb_dim * 2),
            nn.ReLU(),
            m.MP_Conv(in_channels=emb_dim * 2, out_channels=emb_dim * 4, kernel=()),
            nn.GroupNorm(1, emb_dim * 4),
            nn.ReLU(),
            nn.Dropout(dropout),
        )
        self.linear = m.MP_Conv(in_channels=emb_dim * 4, out_channels=num_experts, kernel=())

    def _fused_ok(self, x) -> bool:
        sr = self.soft_route
        return (_FUSED_SCALING[0] and x.is_cuda and x.dtype == torch.float32 and x.ndim == 2 and x.shape[1] == 64
                and tuple(sr[0].weights.shape) == (128, 64) and tuple(sr[3].weights.shape) == (256, 128)
                and tuple(self.linear.weights.shape) == (2, 256) and sr[1].num_groups == 1 and sr[4].num_groups == 1
                and sr[1].eps == sr[4].eps)

    def forward(self, x: torch.Tensor, zeta: Optional[float] = 1e-2, noise: Optional[torch.Tensor] = None):
        if x.ndim == 3:
            x = x.squeeze(1)
        if self._fused_ok(x):
            # the whole chain (two MP linears + GroupNorm(1) + ReLU, dropout mask, logits, exploration noise, softmax * 2)
            # as ONE kernel forward and one backward (csrc/trunk_glue.cu) on the prepared weights
            sr, B = self.soft_route, x.shape[0]
            W1 = sr[0].prepared_weight(1.0, torch.float32)
            W2 = sr[3].prepared_weight(1.0, torch.float32)
            W3 = self.linear.prepared_weight(1.0, torch.float32)
            keep = None
            p = sr[6].p
            if self.training and p > 0:
                keep = (torch.rand(B, 256, device=x.device) >= p).float() / (1.0 - p)
            nz = None
            if self.training:
                nz = torch.randn(B, 2, device=x.device) if noise is None else noise
            return ops.scaling_router(x, W1, sr[1].weight, sr[1].bias, W2, sr[4].weight, sr[4].bias, W3, noise=nz,
                                      zeta=float(zeta), keep=keep, eps=sr[1].eps)
        x = self.linear(self.soft_route(x))
        if self.training:
            x = x + (torch.randn_like(x) if noise is None else noise) * zeta
        return F.softmax(x, dim=-1) * 2


class Router(nn.Module):
    """Sparse top-k router (ref models/model_components.py:68-168).

    forward(x, time_emb, mask=None, zeta=1e-2) -> (sparse_gate_weights, gate_probs, logits), as the
    reference.  The conv trunk uses stock ops; everything from the pooled features on -- adaLN modulation,
    the 128->E linear, zeta-scaled exploration noise, the mask pattern, softmax, top-k, softmax(top-k), the
    sparse scatter and the load-balance / z-loss partial sums -- is one kernel.  Extra, kernel-side
    outputs of the last call are kept in `self.last` (topk_idx int32, topk_w, stats) for the dispatch
    plan and the fused loss.  `noise=` lets a caller supply the randn draw (parity runs)."""

    def __init__(self, in_channels: Optional[int] = 3, time_dim: Optional[int] = 256, top_k: Optional[int] = 1,
                 num_experts: Optional[int] = 5, dropout: Optional[float] = 0.2):
        super().__init__()
        c = in_channels
        self.hard_route = nn.Sequential(
            m.MP_Conv(in_channels=c, out_channels=c * 2, kernel=(3, 3)),
            nn.GroupNorm(1, c * 2),
            nn.ReLU(),
            m.MP_Conv(in_channels=c * 2, out_channels=c * 4, kernel=(3, 3)),
            nn.GroupNorm(1, c * 4),
            nn.ReLU(),
            m.MP_Conv(in_channels=c * 4, out_channels=c * 4, kernel=(3, 3)),
            nn.GroupNorm(1, c * 4),
            nn.ReLU(),
            nn.AdaptiveAvgPool2d((1, 1)),
            nn.Dropout(dropout),
        )
        self.out_router = c * 4
        self.time_linear = m.MP_Conv(in_channels=time_dim, out_channels=self.out_router * 2, kernel=())
        self.linear = m.MP_Conv(in_channels=c * 4, out_channels=num_experts, kernel=())
        self.k = top_k
        self.last = None

    def _fusable(self) -> bool:
        seq = self.hard_route
        return all(isinstance(seq[i + 1], nn.GroupNorm) and seq[i + 1].num_groups == 1
                   and ops.gn1_relu_supported(seq[i + 1].num_channels) for i in (0, 3, 6))

    def _fused_trunk(self, x: torch.Tensor) -> torch.Tensor:
        """hard_route with channels-last library convolutions (no layout-conversion kernels) and the fused
        GroupNorm(1) + ReLU (+ average pool for the last layer) kernels; same modules, parameters and result."""
        seq = self.hard_route
        h = x.contiguous(memory_format=torch.channels_last)
        for i in (0, 3, 6):
            h = seq[i](h).contiguous(memory_format=torch.channels_last)
            gn = seq[i + 1]
            h = ops.gn1_relu(h, gn.weight, gn.bias, gn.eps, pool=(i == 6))
        B, Cn = h.shape
        return seq[10](h.view(B, Cn, 1, 1)).reshape(B, Cn)

    def forward(self, x: torch.Tensor, time_emb: torch.Tensor, mask: Optional[torch.Tensor] = None,
                zeta: Optional[float] = 1e-2, noise: Optional[torch.Tensor] = None,
                pooled: Optional[torch.Tensor] = None):
        """`pooled` (B200 extra): the trunk's pooled features when they were computed outside (router_trunk.py runs the
        trunks of both routers as grouped tcgen05 launches); only the Dropout of hard_route is applied to them here."""
        B = x.shape[0]
        if pooled is not None:
            pooled = self.hard_route[10](pooled.reshape(B, -1, 1, 1)).reshape(B, -1).float()
        elif x.is_cuda and x.dtype == torch.float32 and _FUSED_TRUNK[0] and self._fusable():
            pooled = self._fused_trunk(x).float()
        else:
            if x.is_cuda and _CHANNELS_LAST[0]:
                x = x.contiguous(memory_format=torch.channels_last)
            pooled = self.hard_route(x).reshape(B, -1).float()
        if time_emb.ndim == 3:
            time_emb = time_emb.squeeze(1)
        cond = self.time_linear(m.mp_silu(time_emb)).float()
        w_hat = self.linear.prepared_weight(1.0, torch.float32)
        if self.training:
            if noise is None:
                noise = torch.randn(B, w_hat.shape[0], device=x.device, dtype=torch.float32)
        else:
            noise = None
        sparse, probs, logits, idx, tw, stats = ops.router_gate(pooled, cond, w_hat, self.k, noise=noise,
                                                                zeta=float(zeta), mask=mask)
        self.last = {"topk_idx": idx, "topk_w": tw, "stats": stats}
        # the load-balance / z-loss partial sums travel with the probabilities so that EDM_LOSS can consume them without
        # a change of the reference's return structure (hdmoe_b200.utils.EDM_LOSS reads the attribute)
        probs._hdmoe_stats = stats
        return sparse, probs, logits


class Unet_block(nn.Module):
    """EDM2-style block with a per-expert kernel size (ref models/model_components.py:171-253)."""

    def __init__(self, in_channels: int, out_channels: int, kernel: tuple, emb_size: int,
                 resample: Optional[str] = "keep", Type: Optional[str] = "enc",
                 residual_balance: Optional[float] = 0.5, Dropout: Optional[float] = 0.2,
                 emb_gain: Optional[float] = 1.0, conv_gain: Optional[float] = 1.0):
        super().__init__()
        self.in_channels, self.out_channels, self.emb_size = in_channels, out_channels, emb_size
        self.residual_balance = residual_balance
        self.type = Type
        self.resample = resample
        self.kernel = kernel
        self.dropout = Dropout
        self.emb_gain = emb_gain
        self.conv_gain1 = self.conv_gain2 = conv_gain
        self.conv_skip = m.MP_Conv(in_channels, out_channels, kernel=(1, 1)) if in_channels != out_channels else None
        self.emb_layer = m.MP_Conv(emb_size, out_channels, kernel=())
        self.conv_res1 = m.MP_Conv(out_channels if Type == "enc" else in_channels, out_channels, kernel=kernel)
        self.conv_res2 = m.MP_Conv(out_channels, out_channels, kernel=kernel)

    def forward(self, x: torch.Tensor, embedding: torch.Tensor) -> torch.Tensor:
        emb = 1 + self.emb_layer(embedding, gain=self.emb_gain)
        x = m.resample(x, mode=self.resample)
        if self.type == "enc":
            if self.conv_skip is not None:
                x = self.conv_skip(x)
            x = m.normalize(x, dim=[1])
        y = self.conv_res1(m.mp_silu(x), gain=self.conv_gain1)
        y = m.mp_silu(y * emb[:, :, None, None].to(x.dtype))
        if self.training and self.dropout != 0:
            y = F.dropout(y, p=self.dropout)
        y = self.conv_res2(y, gain=self.conv_gain2)
        if self.type == "dec" and self.conv_skip is not None:
            x = self.conv_skip(x)
        return m.mp_sum(x, y, t=self.residual_balance)


class Unet_expert(nn.Module):
    """Magnitude-preserving U-Net expert (ref models/model_components.py:255-433)."""

    def __init__(self, img_resolution: int, img_channels: int, time_emb_dim: int, text_emb_dim: int,
                 channel_mult: list, model_channels: Optional[int] = 192, channel_mult_emb: Optional[int] = None,
                 num_blocks: Optional[int] = 3, kernel_size: Optional[tuple] = (3, 3),
                 label_balance: Optional[float] = 0.5, concat_balance: Optional[float] = 0.5):
        super().__init__()
        self.block_channels = [model_channels * i for i in channel_mult]
        self.emb_size = model_channels * channel_mult_emb if channel_mult_emb is not None else max(self.block_channels)
        self.label_balance, self.concat_balance = label_balance, concat_balance
        self.kernel_size = tuple(kernel_size)
        self.out_gain = nn.Parameter(torch.zeros([]))
        self.map_noise = m.MP_Conv(time_emb_dim, self.emb_size, kernel=())
        self.map_text = m.MP_Conv(text_emb_dim, self.emb_size, kernel=()) if text_emb_dim > 0 else None
        blk = dict(kernel=kernel_size, emb_size=self.emb_size)
        self.encoders = nn.ModuleDict()
        cout = img_channels + 1
        for level, ch in enumerate(self.block_channels):
            res = img_resolution >> level
            if level == 0:
                cin, cout = cout, ch
                self.encoders[f"{res}x{res}_conv"] = m.MP_Conv(cin, cout, kernel=kernel_size)
            else:
                self.encoders[f"{res}x{res}_down"] = Unet_block(cout, cout, Type="enc", resample="down", **blk)
            for i in range(num_blocks):
                cin, cout = cout, ch
                self.encoders[f"{res}x{res}_block{i}"] = Unet_block(cin, cout, Type="enc", resample="keep", **blk)
        self.decoders = nn.ModuleDict()
        skips = [b.out_channels for b in self.encoders.values()]
        for level, ch in reversed(list(enumerate(self.block_channels))):
            res = img_resolution >> level
            if level == len(self.block_channels) - 1:
                self.decoders[f"{res}x{res}_in0"] = Unet_block(cout, cout, Type="dec", resample="keep", **blk)
                self.decoders[f"{res}x{res}_in1"] = Unet_block(cout, cout, Type="dec", resample="keep", **blk)
            else:
                self.decoders[f"{res}x{res}_up"] = Unet_block(cout, cout, Type="dec", resample="up", **blk)
            for i in range(num_blocks + 1):
                cin, cout = cout + skips.pop(), ch
                self.decoders[f"{res}x{res}_block{i}"] = Unet_block(cin, cout, Type="dec", resample="keep", **blk)
        self.out_channels = cout
        self.out_conv = m.MP_Conv(cout, img_channels, kernel=kernel_size)

    def embed(self, time_emb: torch.Tensor, text_emb: Optional[torch.Tensor]) -> torch.Tensor:
        emb = self.map_noise(time_emb)
        if self.map_text is not None and text_emb is not None:
            if text_emb.ndim == 3:
                text_emb = text_emb.mean(dim=1)
            emb = m.mp_sum(emb, self.map_text(text_emb), t=self.label_balance)
        return m.mp_silu(emb)

    def forward(self, x: torch.Tensor, time_emb: torch.Tensor, text_emb: torch.Tensor) -> torch.Tensor:
        emb = self.embed(time_emb.to(x.dtype), None if text_emb is None else text_emb.to(x.dtype))
        x = torch.cat([x, torch.ones_like(x[:, :1])], dim=1)
        skips = []
        for name, block in self.encoders.items():
            x = block(x) if "conv" in name else block(x, embedding=emb)
            skips.append(x)
        for name, block in self.decoders.items():
            if "block" in name:
                x = m.mp_cat(x, skips.pop(), t=self.concat_balance)
            x = block(x, embedding=emb)
        return self.out_conv(x, gain=self.out_gain)


class Vit_block(nn.Module):
    """DiffiT-style block (ref models/model_components.py:435-562)."""

    def __init__(self, num_heads: int, num_groups: int, num_channels: int, seq_ln: int, emb_dim: int,
                 resample: Optional[str] = "keep", time_dim: Optional[int] = 0, res_balance: Optional[float] = 0.5,
                 attn_balance: Optional[float] = 0.5, gain_s: Optional[float] = 1.0, gain_t: Optional[float] = 1.0):
        super().__init__()
        self.res_balance, self.gain_s, self.gain_t = res_balance, gain_s, gain_t
        self.emb_dim = emb_dim
        self.resample = resample
        self.GN = nn.GroupNorm(num_groups=num_groups, num_channels=num_channels)
        self.skip_proj = m.MP_Conv(num_channels, emb_dim, kernel=()) if num_channels != emb_dim else None
        self.linear1 = m.MP_Conv(num_channels, emb_dim, kernel=())
        self.norm1 = nn.LayerNorm(emb_dim)
        self.norm2 = nn.LayerNorm(emb_dim)
        self.TMSA = m.MP_Attention(num_heads=num_heads, emb_dim=emb_dim, seq_ln=seq_ln, time_dim=time_dim,
                                   attn_balance=attn_balance)
        self.linear2 = m.MP_Conv(emb_dim, emb_dim * 4, kernel=())
        self.linear3 = m.MP_Conv(emb_dim * 4, emb_dim, kernel=())

    @staticmethod
    def _ln(norm: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
        return F.layer_norm(x, norm.normalized_shape, norm.weight.to(x.dtype), norm.bias.to(x.dtype), norm.eps)

    def forward(self, x: torch.Tensor, time_embedding: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = m.resample(x, mode=self.resample)
        B, S, Cin = x.shape
        res_main = x
        h = F.group_norm(x.transpose(1, 2), self.GN.num_groups, self.GN.weight.to(x.dtype), self.GN.bias.to(x.dtype),
                         self.GN.eps)
        h = m.mp_silu(h).transpose(1, 2).reshape(B * S, Cin)
        h = self.linear1(h, gain=self.gain_s)
        res_attn = h
        y = self._ln(self.norm1, h).reshape(B, S, self.emb_dim)
        if time_embedding is not None and time_embedding.ndim == 2:
            time_embedding = time_embedding[:, None, :]
        y = self.TMSA(y, time_embedding=time_embedding, gain_s=self.gain_s, gain_t=self.gain_t)
        y = m.mp_sum(y.reshape(B * S, self.emb_dim), res_attn, t=self.res_balance)
        h = self._ln(self.norm2, y)
        h = m.mp_silu(self.linear2(h, gain=self.gain_s))
        h = self.linear3(h, gain=self.gain_s)
        h = m.mp_sum(h, y, t=self.res_balance).reshape(B, S, self.emb_dim)
        if self.skip_proj is not None:
            r = self.skip_proj(res_main.reshape(B * S, Cin), gain=self.gain_s).reshape(B, S, self.emb_dim)
            return m.mp_sum(r, h, t=self.res_balance)
        return m.mp_sum(res_main, h, t=self.res_balance)


class Vit_expert(nn.Module):
    """Isotropic ViT expert: strided-conv patchify, DiffiT blocks, linear + PixelShuffle unpatchify
    (ref models/model_components.py:564-706)."""

    def __init__(self, num_heads: int, num_groups: int, in_channels: int, seq_ln: int, emb_dim: int, num_blocks: int,
                 patch_size: int, time_dim: Optional[int] = 0, text_dim: Optional[int] = 0,
                 res_balance: Optional[float] = 0.5, attn_balance: Optional[float] = 0.5,
                 emb_balance: Optional[float] = 0.5, gain_s: Optional[float] = 1.0, gain_t: Optional[float] = 1.0):
        super().__init__()
        self.seq_ln, self.emb_balance, self.emb_dim = seq_ln, emb_balance, emb_dim
        self.patch = nn.Conv2d(in_channels, emb_dim, kernel_size=patch_size, stride=patch_size)
        self.map_txt = m.MP_Conv(text_dim, time_dim, kernel=()) if text_dim != time_dim and text_dim != 0 else None
        self.pos_emb = nn.Parameter(torch.zeros(1, seq_ln, emb_dim))
        self.diffit = nn.ModuleList(
            Vit_block(num_heads=num_heads, num_groups=num_groups, num_channels=emb_dim, seq_ln=seq_ln, emb_dim=emb_dim,
                      resample="keep", time_dim=time_dim, res_balance=res_balance, attn_balance=attn_balance,
                      gain_s=gain_s, gain_t=gain_t) for _ in range(num_blocks))
        self.norm = nn.LayerNorm(emb_dim)
        self.unpatch_proj = m.MP_Conv(emb_dim, in_channels * patch_size ** 2, kernel=())
        self.unpatch = nn.PixelShuffle(upscale_factor=patch_size)

    def forward(self, x: torch.Tensor, time_emb: torch.Tensor = None, text_emb: Optional[torch.Tensor] = None):
        B, _, H, W = x.shape
        p = self.patch.kernel_size[0]
        ph, pw = (p - H % p) % p, (p - W % p) % p
        if ph or pw:
            x = F.pad(x, (0, pw, 0, ph))
        if x.is_cuda:
            # kernel == stride: the patchify convolution is a linear map over non-overlapping patches.  As a GEMM its
            # backward avoids the library's strided-convolution dgrad / wgrad kernels (1.5-2 ms per train step).
            Hp, Wp = x.shape[-2:]
            hp, wp = Hp // p, Wp // p
            cols = x.reshape(B, -1, hp, p, wp, p).permute(0, 2, 4, 1, 3, 5).reshape(B, hp * wp, -1)
            x = F.linear(cols, self.patch.weight.to(x.dtype).flatten(1), self.patch.bias.to(x.dtype))
        else:
            x = F.conv2d(x, self.patch.weight.to(x.dtype), self.patch.bias.to(x.dtype), stride=p)
            _, D, hp, wp = x.shape
            x = x.flatten(2).transpose(1, 2)
        assert hp * wp == self.seq_ln, f"Sequence length mismatch: Got {hp * wp}, expected {self.seq_ln}, shape: {x.shape}"
        x = x + self.pos_emb.to(x.dtype)
        if time_emb is not None:
            time_emb = time_emb.to(x.dtype)
        if text_emb is not None:
            text_emb = text_emb.to(x.dtype)
            if self.map_txt is not None:
                if text_emb.ndim == 3:
                    text_emb = text_emb.mean(dim=1)
                text_emb = self.map_txt(text_emb)
            time_emb = m.mp_sum(time_emb, text_emb, t=self.emb_balance)
        for block in self.diffit:
            x = block(x, time_embedding=time_emb)
        x = Vit_block._ln(self.norm, x).reshape(B * self.seq_ln, self.emb_dim)
        x = self.unpatch_proj(x).reshape(B, self.seq_ln, -1).transpose(1, 2).reshape(B, -1, hp, wp)
        x = self.unpatch(x)
        if ph or pw:
            x = x[:, :, :H, :W]
        return x
