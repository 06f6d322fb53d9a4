"""CPU oracle for the heterogeneous-MoE hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a functional, state_dict-driven restatement (torch CPU tensors, fp32 or fp64)
of the reference algorithm for the path named in BASELINE.json's north_star.  It is the
checker for `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py`; nothing under `heterogeneous-moe-for-diffusion-models_b200/` may import
it, and the product path never routes through it.

Why torch-on-CPU and not numpy: the arithmetic of the reference lives in a third-party
dependency that is not vendored under /root/reference -- PyTorch (pinned torch==2.6.0+cu124
in the reference's requirements.txt; this image has 2.11.0+cu128).  Restating conv2d /
matmul / softmax through the same library's CPU kernels keeps the oracle within an ulp or
two of the reference, and autograd gives the backward that the train-step baseline needs.
The integer parts (top-k indices, dispatch order, offsets) are restated explicitly.

Parity pinning: every function below is checked against fixtures produced by importing the
unmodified reference in the build container (`tools/make_golden.py`, outputs committed under
`tests/golden/`).  The reference's own tests hold no golden vectors, only
analytic invariants; those are re-checked in tests/test_oracle_cpu.py as well.

All `file:line` citations are relative to /root/reference.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# --------------------------------------------------------------------------------------
# L0 primitives  (models/model_internals.py)
# --------------------------------------------------------------------------------------


def normalize(x: Tensor, dim: Optional[Sequence[int]] = None, eps: float = 1e-4) -> Tensor:
    """x / (eps + ||x||_dim * sqrt(norm.numel()/x.numel())).  models/model_internals.py:26-30."""
    if dim is None:
        dim = list(range(1, x.ndim))
    nrm = torch.linalg.vector_norm(x, dim=list(dim), keepdim=True, dtype=torch.float32 if x.dtype != torch.float64 else torch.float64)
    alpha = math.sqrt(nrm.numel() / x.numel())
    nrm = eps + alpha * nrm
    return x / nrm.to(x.dtype)


def mp_silu(x: Tensor) -> Tensor:
    """silu(x)/0.596.  models/model_internals.py:47."""
    return F.silu(x) / 0.596


def mp_sum(a: Tensor, b: Tensor, t: float = 0.5) -> Tensor:
    """lerp(a,b,t)/sqrt((1-t)^2+t^2).  models/model_internals.py:66."""
    return a.lerp(b, t) / math.sqrt((1 - t) ** 2 + t ** 2)


def mp_cat(a: Tensor, b: Tensor, dim: int = 1, t: float = 0.5) -> Tensor:
    """Magnitude-preserving concat.  models/model_internals.py:87-92."""
    na, nb = a.shape[dim], b.shape[dim]
    c = math.sqrt((na + nb) / ((1 - t) ** 2 + t ** 2))
    return torch.cat([(c * (1 - t) / math.sqrt(na)) * a, (c * t / math.sqrt(nb)) * b], dim=dim)


def resample(x: Tensor, mode: str) -> Tensor:
    """2x box filter.  models/model_internals.py:107-127 with f=[1,1]:
    'down' = depthwise conv, 0.25 weights, stride 2, pad 0; 'up' = conv_transpose with
    weights 0.25*4 = 1, stride 2, pad 0 (each input pixel replicated to a 2x2 block)."""
    if mode == "keep":
        return x
    c = x.shape[1]
    k = torch.full((c, 1, 2, 2), 0.25, dtype=x.dtype)
    if mode == "down":
        return F.conv2d(x, k, stride=2, groups=c)
    if mode == "up":
        return F.conv_transpose2d(x, k * 4, stride=2, groups=c)
    raise ValueError(mode)


def mp_fourier(x: Tensor, freqs: Tensor, phases: Tensor) -> Tensor:
    """sqrt(2)*cos(x (outer) freqs + phases).  models/model_internals.py:171-175."""
    y = x.to(torch.float32 if x.dtype != torch.float64 else torch.float64)
    y = torch.outer(y, freqs.to(y.dtype)) + phases.to(y.dtype)
    return (y.cos() * math.sqrt(2)).to(x.dtype)


def mp_weight(w: Tensor, gain=1.0) -> Tensor:
    """normalize(w) * gain / sqrt(fan_in).  models/model_internals.py:258-259."""
    w = normalize(w)
    return w * (gain / math.sqrt(w[0].numel()))


def forced_weight_norm_(sd: SD, keys: Optional[Sequence[str]] = None) -> None:
    """Training-mode side effect of MP_Conv.forward: weights <- normalize(weights), in place,
    under no_grad.  models/model_internals.py:254-256 (quirk Q6)."""
    with torch.no_grad():
        for k, v in sd.items():
            if k.endswith(".weights") and (keys is None or k in keys):
                v.copy_(normalize(v.detach()))


def mp_conv(x: Tensor, w: Tensor, gain=1.0) -> Tensor:
    """MP_Conv.forward without the in-place side effect.  models/model_internals.py:258-271:
    2-D input -> linear; 4-D -> asymmetric 'same' pad (left (k-1)//2, right the rest), conv2d."""
    wh = mp_weight(w, gain).to(x.dtype)
    if x.ndim == 2:
        return F.linear(x, wh)
    assert x.ndim == 4
    k = wh.shape[-1]
    tot = k - 1
    lo = tot // 2
    x = F.pad(x, (lo, tot - lo, lo, tot - lo))
    return F.conv2d(x, wh)


_TRAIN = [False]


class training_mode:
    """Context manager: inside it every parameterised MP_Conv call first overwrites its weight with
    normalize(weight) under no_grad, exactly when (and only if) that layer's forward runs --
    models/model_internals.py:254-256 (quirk Q6).  Experts that receive no rows keep their weights."""

    def __enter__(self):
        self._old = _TRAIN[0]
        _TRAIN[0] = True

    def __exit__(self, *a):
        _TRAIN[0] = self._old


def pconv(sd: SD, key: str, x: Tensor, gain=1.0) -> Tensor:
    """MP_Conv.forward of the layer whose parameter is sd[key] (with the train-mode side effect)."""
    if _TRAIN[0]:
        with torch.no_grad():
            sd[key].copy_(normalize(sd[key].detach()))
    return mp_conv(x, sd[key], gain)


def mp_attention(sd: SD, pfx: str, query: Tensor, num_heads: int, gain_s: float, gain_t: float,
                 context: Optional[Tensor] = None, time_embedding: Optional[Tensor] = None,
                 attn_balance: float = 0.5) -> Tensor:
    """MP_Attention.forward.  models/model_internals.py:354-409.  Cross-attention iff the
    module has no rel_pos_bias (`is_cross_attn`, :324)."""
    B, S, D = query.shape
    hd = D // num_heads
    is_cross = (pfx + "rel_pos_bias") not in sd
    ctx = query if context is None else context
    q_in = query.permute(0, 2, 1).unsqueeze(-1)
    c_in = ctx.permute(0, 2, 1).unsqueeze(-1)
    q = pconv(sd, pfx + "q_proj.weights", q_in, gain_s)
    k = pconv(sd, pfx + "k_proj.weights", c_in, gain_s)
    v = pconv(sd, pfx + "v_proj.weights", c_in, gain_s)
    if (pfx + "q_time.weights") in sd and time_embedding is not None:
        te = time_embedding.reshape(B, -1, 1, 1)
        q = q + pconv(sd, pfx + "q_time.weights", te, gain_t)
        if not is_cross:
            k = k + pconv(sd, pfx + "k_time.weights", te, gain_t)
            v = v + pconv(sd, pfx + "v_time.weights", te, gain_t)
    q = q.reshape(B, num_heads, hd, -1).transpose(-1, -2)
    k = k.reshape(B, num_heads, hd, -1).transpose(-1, -2)
    v = v.reshape(B, num_heads, hd, -1).transpose(-1, -2)
    s = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd)
    if not is_cross:
        bias = sd[pfx + "rel_pos_bias"]
        if S <= bias.shape[1]:
            bias = bias[:, :S, :S]
        else:  # models/model_internals.py:389-397
            bias = F.interpolate(bias.unsqueeze(0), size=(S, S), mode="bicubic", align_corners=False).squeeze(0)
        s = s + bias
    p = s.softmax(dim=-1)
    o = torch.matmul(p, v)
    o = o.transpose(1, 2).contiguous().reshape(B, S, D)
    o = pconv(sd, pfx + "out_proj.weights", o.permute(0, 2, 1).unsqueeze(-1), gain_s)
    o = o.squeeze(-1).permute(0, 2, 1)
    return mp_sum(query, o, attn_balance)


# --------------------------------------------------------------------------------------
# Router (models/model_components.py:68-168) and Scaling_router (:7-66)
# --------------------------------------------------------------------------------------


def topk_lowest_index(x: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """Deterministic top-k: descending value, lowest index wins ties (torch.topk leaves the tie
    order unspecified, quirk Q4; the bit-exact contract is stated for this rule)."""
    # stable sort of -x keeps ascending index order among equal keys
    order = torch.sort(-x, dim=-1, stable=True).indices[..., :k]
    return torch.gather(x, -1, order), order


def router_tail(pooled: Tensor, time_emb: Tensor, w_time: Tensor, w_lin: Tensor, top_k: int,
                noise: Optional[Tensor] = None, zeta: float = 0.0, mask: Optional[Tensor] = None):
    """Router.forward from the pooled trunk features on.  models/model_components.py:143-168.
    `noise` is the randn_like draw of :156 supplied by the caller (None = eval mode).
    Returns (sparse_gate_weights, gate_probs, logits, topk_indices)."""
    if time_emb.ndim == 3:
        time_emb = time_emb.squeeze(1)
    cond = mp_conv(mp_silu(time_emb), w_time)
    gamma, beta = cond.chunk(2, dim=1)
    x = pooled * (1 + gamma) + beta
    x = mp_conv(x, w_lin)
    if noise is not None:
        x = x + noise * zeta
    if mask is not None:
        x = x.masked_fill(mask == 0, float("-inf"))
    return router_gate_from_logits(x, top_k)


def router_gate_from_logits(x: Tensor, top_k: int):
    """models/model_components.py:163-168 on already masked logits."""
    gate_probs = F.softmax(x, dim=-1)
    vals, idx = topk_lowest_index(x, top_k)
    w = F.softmax(vals, dim=-1)
    sparse = torch.zeros_like(x).scatter(-1, idx, w)
    return sparse, gate_probs, x, idx


def router_trunk(sd: SD, pfx: str, x: Tensor) -> Tensor:
    """hard_route: 3x[MP_Conv 3x3 -> GroupNorm(1,C) -> ReLU] -> global avg pool (dropout off).
    models/model_components.py:100-112,141-143."""
    for i in (0, 3, 6):
        x = pconv(sd, f"{pfx}hard_route.{i}.weights", x)
        x = F.group_norm(x, 1, sd[f"{pfx}hard_route.{i + 1}.weight"], sd[f"{pfx}hard_route.{i + 1}.bias"])
        x = F.relu(x)
    return x.mean(dim=(2, 3))


def router(sd: SD, pfx: str, x: Tensor, time_emb: Tensor, top_k: int, mask=None, noise=None, zeta=0.0):
    pooled = router_trunk(sd, pfx, x)
    if _TRAIN[0]:
        forced_weight_norm_(sd, (pfx + "time_linear.weights", pfx + "linear.weights"))
    return router_tail(pooled, time_emb, sd[pfx + "time_linear.weights"], sd[pfx + "linear.weights"],
                       top_k, noise=noise, zeta=zeta, mask=mask)


def scaling_router(sd: SD, pfx: str, x: Tensor, noise: Optional[Tensor] = None, zeta: float = 0.0) -> Tensor:
    """Scaling_router.forward (cfg1).  models/model_components.py:56-66, dropout off."""
    if x.ndim == 3:
        x = x.squeeze(1)
    for i in (0, 3):
        x = pconv(sd, f"{pfx}soft_route.{i}.weights", x)
        x = F.group_norm(x, 1, sd[f"{pfx}soft_route.{i + 1}.weight"], sd[f"{pfx}soft_route.{i + 1}.bias"])
        x = F.relu(x)
    x = pconv(sd, pfx + "linear.weights", x)
    if noise is not None:
        x = x + noise * zeta
    return F.softmax(x, dim=-1) * 2


# --------------------------------------------------------------------------------------
# Dispatch plan / permute / combine  (models/model_config2.py:11-39)
# --------------------------------------------------------------------------------------


def dispatch_plan(out_router: Tensor):
    """Index build implied by the per-expert boolean-mask loop (models/model_config2.py:25-33):
    expert-major, ascending sample index inside an expert, criterion `weight > 0` (quirk Q3;
    NaN > 0 is False).  Integer outputs (numpy int32): counts[E], offsets[E+1], src_row[R]
    (token of each permuted row), expert_of_row[R], and the per-row gate weight."""
    w = out_router.detach()
    T, E = w.shape
    sel = (w > 0).cpu().numpy()
    counts = sel.sum(axis=0).astype(np.int32)
    offsets = np.zeros(E + 1, dtype=np.int32)
    offsets[1:] = np.cumsum(counts)
    src, exp = [], []
    for e in range(E):
        rows = np.nonzero(sel[:, e])[0]
        src.append(rows)
        exp.append(np.full(rows.shape, e))
    src_row = np.concatenate(src).astype(np.int32) if src else np.zeros(0, np.int32)
    expert_of_row = np.concatenate(exp).astype(np.int32) if exp else np.zeros(0, np.int32)
    return counts, offsets, src_row, expert_of_row


def permute_rows(x: Tensor, src_row: np.ndarray) -> Tensor:
    """x[mask] for every expert, concatenated expert-major (models/model_config2.py:31)."""
    return x[torch.as_tensor(src_row, dtype=torch.long)]


def combine_rows(rows: Tensor, out_router: Tensor, src_row: np.ndarray, expert_of_row: np.ndarray,
                 T: int, base: Optional[Tensor] = None) -> Tensor:
    """output = zeros; output[mask] += out_e * w[mask, e] for e ascending
    (models/model_config2.py:23,35-37).  `base` (not in the reference: it starts from zeros)
    is the optional residual of north_star item (4)."""
    out = torch.zeros((T,) + tuple(rows.shape[1:]), dtype=rows.dtype) if base is None else base.clone()
    src = torch.as_tensor(src_row, dtype=torch.long)
    exp = torch.as_tensor(expert_of_row, dtype=torch.long)
    w = out_router[src, exp].reshape((-1,) + (1,) * (rows.ndim - 1)).to(rows.dtype)
    contrib = rows * w
    E = out_router.shape[1]
    for e in range(E):  # ascending expert order == reference summation order
        m = exp == e
        if m.any():
            out.index_add_(0, src[m], contrib[m])  # unique tokens inside one expert
    return out


def moe_layer(x: Tensor, out_router: Tensor, time_emb: Tensor, text_emb: Optional[Tensor], expert_fn):
    """router_to_unet_experts.  models/model_config2.py:11-39.  expert_fn(e, x_e, t_e, txt_e)."""
    txt = text_emb.mean(dim=1) if (text_emb is not None and text_emb.ndim == 3) else text_emb
    counts, offsets, src_row, expert_of_row = dispatch_plan(out_router)
    xr = permute_rows(x, src_row)
    tr = permute_rows(time_emb, src_row)
    txr = permute_rows(txt, src_row) if txt is not None else None
    outs = []
    for e in range(out_router.shape[1]):
        lo, hi = int(offsets[e]), int(offsets[e + 1])
        if hi == lo:
            continue
        outs.append(expert_fn(e, xr[lo:hi], tr[lo:hi], None if txr is None else txr[lo:hi]))
    rows = torch.cat(outs, dim=0) if outs else xr
    return combine_rows(rows, out_router, src_row, expert_of_row, x.shape[0])


# --------------------------------------------------------------------------------------
# Experts  (models/model_components.py:171-706)
# --------------------------------------------------------------------------------------


def unet_block(sd: SD, pfx: str, x: Tensor, emb: Tensor, typ: str, resample_mode: str,
               residual_balance: float = 0.5) -> Tensor:
    """Unet_block.forward, dropout off.  models/model_components.py:232-253."""
    e = 1 + pconv(sd, pfx + "emb_layer.weights", emb)
    x = resample(x, resample_mode)
    has_skip = (pfx + "conv_skip.weights") in sd
    if typ == "enc":
        if has_skip:
            x = pconv(sd, pfx + "conv_skip.weights", x)
        x = normalize(x, dim=[1])
    y = pconv(sd, pfx + "conv_res1.weights", mp_silu(x))
    y = mp_silu(y * e.unsqueeze(2).unsqueeze(3).to(x.dtype))
    y = pconv(sd, pfx + "conv_res2.weights", y)
    if typ == "dec" and has_skip:
        x = pconv(sd, pfx + "conv_skip.weights", x)
    return mp_sum(x, y, residual_balance)


def _module_order(sd: SD, pfx: str) -> List[str]:
    """Child names under `pfx` in state_dict (== registration) order."""
    seen: List[str] = []
    for k in sd:
        if k.startswith(pfx):
            name = k[len(pfx):].split(".")[0]
            if name not in seen:
                seen.append(name)
    return seen


def unet_expert(sd: SD, pfx: str, x: Tensor, time_emb: Tensor, text_emb: Optional[Tensor],
                label_balance: float = 0.5, concat_balance: float = 0.5) -> Tensor:
    """Unet_expert.forward.  models/model_components.py:406-433."""
    emb = pconv(sd, pfx + "map_noise.weights", time_emb)
    if (pfx + "map_text.weights") in sd and text_emb is not None:
        if text_emb.ndim == 3:
            text_emb = text_emb.mean(dim=1)
        emb = mp_sum(emb, pconv(sd, pfx + "map_text.weights", text_emb), label_balance)
    emb = mp_silu(emb)
    x = torch.cat([x, torch.ones_like(x[:, :1])], dim=1)
    skips = []
    for name in _module_order(sd, pfx + "encoders."):
        p = f"{pfx}encoders.{name}."
        if "conv" in name:
            x = pconv(sd, p + "weights", x)
        else:
            x = unet_block(sd, p, x, emb, "enc", "down" if "down" in name else "keep")
        skips.append(x)
    for name in _module_order(sd, pfx + "decoders."):
        p = f"{pfx}decoders.{name}."
        if "block" in name:
            x = mp_cat(x, skips.pop(), t=concat_balance)
        x = unet_block(sd, p, x, emb, "dec", "up" if "up" in name else "keep")
    return pconv(sd, pfx + "out_conv.weights", x, gain=sd[pfx + "out_gain"])


def vit_block(sd: SD, pfx: str, x: Tensor, time_embedding: Optional[Tensor], num_heads: int,
              num_groups: int, res_balance: float = 0.5, gain_s: float = 1.0, gain_t: float = 1.0) -> Tensor:
    """Vit_block.forward (num_channels == emb_dim, no skip_proj in the shipped configs).
    models/model_components.py:525-562."""
    B, S, C = x.shape
    res_main = x
    h = F.group_norm(x.transpose(1, 2), num_groups, sd[pfx + "GN.weight"], sd[pfx + "GN.bias"])
    h = mp_silu(h).transpose(1, 2).reshape(B * S, C)
    h = pconv(sd, pfx + "linear1.weights", h, gain_s)
    D = h.shape[-1]
    res_attn = h
    y = F.layer_norm(h, (D,), sd[pfx + "norm1.weight"], sd[pfx + "norm1.bias"]).reshape(B, S, D)
    if time_embedding is not None and time_embedding.ndim == 2:
        time_embedding = time_embedding[:, None, :]
    y = mp_attention(sd, pfx + "TMSA.", y, num_heads, gain_s, gain_t, time_embedding=time_embedding)
    y = mp_sum(y.reshape(B * S, D), res_attn, res_balance)
    h = F.layer_norm(y, (D,), sd[pfx + "norm2.weight"], sd[pfx + "norm2.bias"])
    h = mp_silu(pconv(sd, pfx + "linear2.weights", h, gain_s))
    h = pconv(sd, pfx + "linear3.weights", h, gain_s)
    h = mp_sum(h, y, res_balance).reshape(B, S, D)
    if (pfx + "skip_proj.weights") in sd:
        r = pconv(sd, pfx + "skip_proj.weights", res_main.reshape(B * S, C), gain_s).reshape(B, S, D)
        return mp_sum(r, h, res_balance)
    return mp_sum(res_main, h, res_balance)


def vit_expert(sd: SD, pfx: str, x: Tensor, time_emb: Tensor, text_emb: Optional[Tensor],
               num_heads: int, num_groups: int, emb_balance: float = 0.5) -> Tensor:
    """Vit_expert.forward.  models/model_components.py:666-706."""
    B, C, H, W = x.shape
    pw = sd[pfx + "patch.weight"]
    p = pw.shape[-1]
    ph, pwid = (p - H % p) % p, (p - W % p) % p
    if ph or pwid:
        x = F.pad(x, (0, pwid, 0, ph))
    x = F.conv2d(x, pw, sd[pfx + "patch.bias"], stride=p)
    _, D, hp, wp = x.shape
    S = hp * wp
    x = x.flatten(2).transpose(1, 2) + sd[pfx + "pos_emb"]
    if text_emb is not None:
        if (pfx + "map_txt.weights") in sd:
            if text_emb.ndim == 3:
                text_emb = text_emb.mean(dim=1)
            text_emb = pconv(sd, pfx + "map_txt.weights", text_emb)
        time_emb = mp_sum(time_emb, text_emb, emb_balance)
    nb = len(_module_order(sd, pfx + "diffit."))
    for b in range(nb):
        x = vit_block(sd, f"{pfx}diffit.{b}.", x, time_emb, num_heads, num_groups)
    x = F.layer_norm(x, (D,), sd[pfx + "norm.weight"], sd[pfx + "norm.bias"])
    x = pconv(sd, pfx + "unpatch_proj.weights", x.reshape(B * S, D))
    x = x.reshape(B, S, -1).transpose(1, 2).reshape(B, -1, hp, wp)
    x = F.pixel_shuffle(x, p)
    if ph or pwid:
        x = x[:, :, :H, :W]
    return x


# --------------------------------------------------------------------------------------
# Denoiser assembly (models/model_config1.py / model_config2.py) + EDM preconditioning
# --------------------------------------------------------------------------------------


def hdmoem(sd: SD, cfg: dict, x: Tensor, time_vec: Tensor, text_emb: Tensor,
           unet_mask: Optional[Tensor], vit_mask: Optional[Tensor], zeta: float = 0.0,
           transition_point: Optional[float] = None, softness: Optional[float] = None,
           noise: Optional[Dict[str, Tensor]] = None, variant: int = 2, alpha_routing: float = 10.0,
           pfx: str = "net.", capture: Optional[dict] = None):
    """HDMOEM.forward.  variant=2: models/model_config2.py:239-303; variant=1:
    models/model_config1.py:241-309.  `noise` holds the train-mode randn draws
    {'vit','unet'[, 'scaling']} (None = eval).  Returns the reference's 7-tuple."""
    noise = noise or {}
    B, C, H, W = x.shape
    nh, ng, k = cfg["VIT_num_heads"], cfg["VIT_num_groups"], cfg["top_k"]
    te = mp_fourier(time_vec, sd[pfx + "Fourier_emb.freqs"], sd[pfx + "Fourier_emb.phases"])
    te = pconv(sd, pfx + "out_fourier1.weights", te)
    te = pconv(sd, pfx + "out_fourier2.weights", mp_silu(te))
    feats = pconv(sd, pfx + "input_proj.weights", x)
    if variant == 2:
        vw = torch.sigmoid((time_vec * 4 - transition_point) / softness).view(-1, 1, 1, 1)
        s_vit = (vw + 1e-2) * 2
        s_unet = ((1.0 - vw) + 1e-2) * 2
        scaling = torch.cat([s_vit, s_unet], dim=1).view(-1, 2)
    else:
        scaling = scaling_router(sd, pfx + "scaling_net.", te, noise=noise.get("scaling"), zeta=zeta)
        s_vit = scaling[:, 0:1].view(-1, 1, 1, 1)
        s_unet = scaling[:, 1:2].view(-1, 1, 1, 1)
    in_unet = s_unet * feats
    in_vit = s_vit * feats
    # the ViT router runs first (quirk Q2)
    w_vit, p_vit, raw_vit, idx_vit = router(sd, pfx + "vit_router.", in_vit, te, k, mask=vit_mask,
                                            noise=noise.get("vit"), zeta=zeta)
    w_un, p_un, raw_un, idx_un = router(sd, pfx + "Unet_router.", in_unet, te, k, mask=unet_mask,
                                        noise=noise.get("unet"), zeta=zeta)
    out_u = moe_layer(in_unet, w_un, te, text_emb,
                      lambda e, xe, t, tx: unet_expert(sd, f"{pfx}Unet_experts.{e}.", xe, t, tx,
                                                       cfg.get("Unet_label_balance", 0.5),
                                                       cfg.get("Unet_concat_balance", 0.5)))
    out_v = moe_layer(in_vit, w_vit, te, text_emb,
                      lambda e, xe, t, tx: vit_expert(sd, f"{pfx}VIT_experts.{e}.", xe, t, tx, nh, ng))
    uf = out_u.flatten(2).transpose(1, 2)
    vf = out_v.flatten(2).transpose(1, 2)
    if variant == 2:
        q, ctx = uf, vf
    else:  # models/model_config1.py:277-283
        stronger = torch.sigmoid(alpha_routing * (s_vit - s_unet)).view(-1, 1, 1)
        q = stronger * vf + (1 - stronger) * uf
        ctx = stronger * uf + (1 - stronger) * vf
    a = mp_attention(sd, pfx + "cross_attn.", q, nh, 1.0, 1.0, context=ctx)
    b = mp_attention(sd, pfx + "cross_attn_text.", a, nh, 1.0, 1.0, context=text_emb)
    fin = a + sd[pfx + "alpha_txt"] * (b - a)
    img = fin.transpose(1, 2).reshape(B, -1, H, W)
    g = pconv(sd, pfx + "gate1.weights", mp_cat(out_u, img, dim=1))
    g = pconv(sd, pfx + "gate2.weights", mp_silu(g))
    g = F.softmax(g, dim=1)
    gated = g[:, 0:1] * out_u + g[:, 1:2] * img
    out = pconv(sd, pfx + "output_proj.weights", mp_sum(out_u, gated, 0.5))
    if capture is not None:
        capture.update(time_embed=te, in_unet=in_unet, in_vit=in_vit, w_unet=w_un, w_vit=w_vit,
                       idx_unet=idx_un, idx_vit=idx_vit, out_unet=out_u, out_vit=out_v)
    return out, p_un, raw_un, p_vit, raw_vit, scaling, g


def edm_coefficients(sigma: Tensor, sigma_data: float):
    """models/model_config2.py:431-435."""
    sigma = sigma.to(torch.float32 if sigma.dtype != torch.float64 else torch.float64)
    c_skip = sigma_data ** 2 / (sigma ** 2 + sigma_data ** 2)
    c_out = sigma * sigma_data / (sigma ** 2 + sigma_data ** 2).sqrt()
    c_in = 1 / (sigma_data ** 2 + sigma ** 2).sqrt()
    c_noise = sigma.flatten().log() / 4
    return c_skip, c_out, c_in, c_noise


def preconditioned(sd: SD, cfg: dict, x: Tensor, sigma: Tensor, text_emb: Tensor, unet_mask, vit_mask,
                   zeta: float = 0.0, transition_point: Optional[float] = None, softness: Optional[float] = None,
                   return_log_var: bool = False, noise=None, variant: int = 2, capture=None) -> Dict[str, Tensor]:
    """preconditioned_HDMOEM.forward.  models/model_config2.py:431-468.  Quirk Q1: `x` is
    overwritten by x*c_in before the skip connection."""
    c_skip, c_out, c_in, c_noise = edm_coefficients(sigma, cfg["sigma_data"])
    B = x.shape[0]
    if c_noise.shape[0] == 1 and B > 1:
        c_noise = c_noise.expand(B)
    x = x * c_in
    out, p_un, raw_un, p_vit, raw_vit, scaling, gate = hdmoem(
        sd, cfg, x, c_noise, text_emb, unet_mask, vit_mask, zeta, transition_point, softness,
        noise=noise, variant=variant, capture=capture)
    D_x = c_skip * x + c_out * out
    log_var = None
    if return_log_var:
        lv = mp_fourier(c_noise, sd["log_var_fourier.freqs"], sd["log_var_fourier.phases"])
        log_var = pconv(sd, "log_var_linear.weights", lv).reshape(-1, 1, 1, 1)
    return {"denoised": D_x, "Unet_router_loss": p_un, "Unet_raw": raw_un, "vit_router_loss": p_vit,
            "vit_raw": raw_vit, "scaling_net_out": scaling, "out_gate": gate, "log_var": log_var}


# --------------------------------------------------------------------------------------
# Loss-side router statistics (Utils/utils.py:158-172) and the full EDM loss (:127-156)
# --------------------------------------------------------------------------------------


def load_balance(gate_probs: Tensor, num_experts: int) -> Tensor:
    """E * sum_e (mean_b p_be)^2.  Utils/utils.py:158-161."""
    return num_experts * torch.sum(gate_probs.mean(dim=0) ** 2)


def z_loss(logits: Tensor) -> Tensor:
    """mean_b min(logsumexp(clamp(l,-50,50))^2, 100).  Utils/utils.py:167-172."""
    z = torch.logsumexp(logits.clamp(min=-50, max=50), dim=-1) ** 2
    return torch.mean(z.clamp(max=100))


def edm_loss(x0: Tensor, out: Dict[str, Tensor], num_experts: int, unet_bal: float, vit_bal: float, z_bal: float):
    """EDM_LOSS.__call__ with lambda = 1, prior loss disabled.  Utils/utils.py:134-156."""
    err = (out["denoised"] - x0) ** 2
    if out["log_var"] is None:
        pure = torch.mean(err)
    else:
        lv = out["log_var"].clamp(min=-10, max=10)
        pure = torch.mean(err / lv.exp() + lv)
    pure = pure.clamp(max=50)
    bal = (unet_bal * load_balance(out["Unet_router_loss"], num_experts)
           + vit_bal * load_balance(out["vit_router_loss"], num_experts)).clamp(max=50)
    zl = (z_bal * z_loss(out["Unet_raw"]) + z_bal * z_loss(out["vit_raw"])).clamp(max=50)
    return {"loss": (pure + zl + bal).clamp(max=50), "denoising": torch.mean(err), "balance": bal,
            "z_loss": zl, "pure_loss": pure}


# --------------------------------------------------------------------------------------
# Host-side producers (Utils/utils.py:175-330)
# --------------------------------------------------------------------------------------


def zeta_schedule(step: int, total_steps: int, max_zeta: float, min_zeta: float = 0.0,
                  strategy: str = "cos", alpha: float = 4.0, warmup_ratio: float = 0.05) -> float:
    """ZetaScheduler.get_zeta.  Utils/utils.py:201-225."""
    warm = int(total_steps * warmup_ratio)
    if step < warm:
        return max_zeta
    if step >= total_steps:
        return min_zeta
    cur, tot = step - warm, total_steps - warm
    if strategy == "cos":
        return float(min_zeta + (max_zeta - min_zeta) * 0.5 * (1 + np.cos(np.pi * cur / tot)))
    if strategy == "exp":
        term = max(min(-alpha * (cur - (max_zeta / tot)), 10), -10)
        z = (max_zeta - min_zeta) * np.exp(term) + min_zeta
        return float(max(min(z, max_zeta), min_zeta))
    raise ValueError(strategy)


def expert_centers(attrs: Sequence[float], noise_range=(0.0, 1.0)) -> Tensor:
    """MaskGenerator.__init__ rank-spaced centres.  Utils/utils.py:262-277."""
    a = torch.tensor(list(attrs), dtype=torch.float32)
    order = torch.sort(a, stable=True).indices
    pts = torch.linspace(noise_range[0], noise_range[1], steps=len(a))
    c = torch.zeros_like(a)
    c[order] = pts
    return c


def mask_bandwidth(step: int, bandwidth: float, max_bw: float, total_steps: int, step_size: float,
                   strat: str = "step") -> float:
    """MaskGenerator.bandwidth_scheduler.  Utils/utils.py:311-329."""
    if step >= total_steps:
        return max_bw
    if strat == "linear":
        return bandwidth + (max_bw - bandwidth) * (step / float(total_steps))
    cur = int(step / (total_steps * step_size))
    tot = int(1.0 / step_size)
    return bandwidth + (max_bw - bandwidth) * min(cur / tot, 1.0)


def band_mask(sigma: Tensor, centers: Tensor, bw: float, p_mean: float, p_std: float, min_active: int = 1) -> Tensor:
    """MaskGenerator.__call__.  Utils/utils.py:290-309."""
    ls = torch.log(sigma.flatten())
    pct = (0.5 * (1 + torch.erf((ls - p_mean) / (p_std * np.sqrt(2))))).clamp(0, 1)
    dist = torch.abs(pct.view(-1, 1) - centers.view(1, -1))
    mask = (dist <= bw).float()
    _, top = torch.topk(-dist, k=min_active, dim=-1)
    mask.scatter_(1, top, 1.0)
    return mask


# --------------------------------------------------------------------------------------
# EDM Heun sampler  (Utils/EDM_sampler.py)
# --------------------------------------------------------------------------------------


def edm_schedule(num_steps: int, sigma_min: float, sigma_max: float, rho: float, dtype=torch.float32) -> Tensor:
    """Karras rho-schedule with a trailing 0.  Utils/EDM_sampler.py:82-87."""
    i = torch.arange(num_steps, dtype=dtype)
    t = (sigma_max ** (1 / rho) + i / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
    return torch.cat([t, torch.zeros_like(t[:1])])


def edm_sample(denoise_fn, noise: Tensor, num_steps: int = 32, sigma_min: float = 0.002, sigma_max: float = 80.0,
               rho: float = 7, S_churn: float = 0.0, S_min: float = 0.0, S_max: float = float("inf"),
               S_noise: float = 1.0, step_noise: Optional[List[Tensor]] = None, dtype=torch.float32,
               trace: Optional[list] = None) -> Tensor:
    """EDM_Sampler.sample.  Utils/EDM_sampler.py:81-108.  denoise_fn(x, sigma_0dim) -> D(x).
    `step_noise[i]` is the per-step randn_like draw of :99 (None -> zeros: with S_churn == 0
    the draw is multiplied by exactly 0, quirk Q14)."""
    t_steps = edm_schedule(num_steps, sigma_min, sigma_max, rho, dtype)
    x_next = noise.to(dtype) * t_steps[0]
    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_cur = x_next
        if S_churn > 0 and S_min <= float(t_cur) <= S_max:
            gamma = min(S_churn / num_steps, math.sqrt(2) - 1)
        else:
            gamma = 0
        t_hat = t_cur + gamma * t_cur
        eps = step_noise[i] if step_noise is not None else torch.zeros_like(x_cur)
        x_hat = x_cur + (t_hat ** 2 - t_cur ** 2).sqrt() * S_noise * eps
        den = denoise_fn(x_hat, t_hat)
        d_cur = (x_hat - den) / t_hat
        x_next = x_hat + (t_next - t_hat) * d_cur
        if i < num_steps - 1:
            den2 = denoise_fn(x_next, t_next)
            d_prime = (x_next - den2) / t_next
            x_next = x_hat + (t_next - t_hat) * (0.5 * d_cur + 0.5 * d_prime)
        if trace is not None:
            trace.append(x_next.clone())
    return x_next


def cfg_denoise(d_cond: Tensor, d_ref: Tensor, guidance: float) -> Tensor:
    """ref.lerp(cond, guidance).  Utils/EDM_sampler.py:70."""
    return d_ref.lerp(d_cond, guidance)


def make_denoiser(sd: SD, cfg: dict, text_emb: Tensor, transition_mean: float, softness: float,
                  guidance: float = 1.0, uncond_text_emb: Optional[Tensor] = None, variant: int = 2):
    """EDM_Sampler.denoise bound to a model.  Utils/EDM_sampler.py:34-70: zeta=0, masks=ones."""
    E = cfg["num_experts"]

    def fn(x, sigma):
        ones = torch.ones((x.shape[0], E), dtype=x.dtype)
        d = preconditioned(sd, cfg, x, sigma, text_emb, ones, ones, 0.0, transition_mean, softness,
                           variant=variant)["denoised"]
        if guidance == 1.0:
            return d
        t2 = uncond_text_emb if uncond_text_emb is not None else text_emb
        r = preconditioned(sd, cfg, x, sigma, t2, ones, ones, 0.0, transition_mean, softness,
                           variant=variant)["denoised"]
        return cfg_denoise(d, r, guidance)

    return fn
