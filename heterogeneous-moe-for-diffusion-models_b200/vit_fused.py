"""ViT experts of one MoE layer through the fused DiffiT-block kernels (csrc/vit_block.cu).

The reference runs every `Vit_block` (models/model_components.py:525-562, with `MP_Attention`,
models/model_internals.py:354-409) as ~85 launches forward and ~170 backward, per expert; the sync-free port ran
~600-kernel chains per expert that became the critical path of the expert phases of the train step (5.4 ms of the
backward at batch 256).  Here the four blocks of ALL experts of the layer are four launches forward and four
backward: one CTA per dispatched row, the row's expert read on the device, every intermediate in shared memory,
the backward recomputes the block from its saved input.  Patchify / un-patchify (GEMMs whose shape depends on the
expert's patch size) stay per-expert torch ops on all rows, selected by the row's expert on the device, so the path
needs no host synchronisation and replays inside a CUDA graph.
"""
import ctypes
from typing import List, Sequence

import torch
import torch.nn.functional as F

from . import _lib as L
from . import model_components as mc
from . import model_internals as m
from . import prepared

_FUSED_VIT = [True]
_TOK = 64          # token slots per row (kVS)
_EMB = 32
_TIME = 64
_WBLOCK = 19456    # prepared-weight floats of one (expert, block): W_TOTAL in vit_block.cu


def set_fused_vit(enabled: bool) -> None:
    _FUSED_VIT[0] = bool(enabled)


def fusable(experts: Sequence, xr: torch.Tensor) -> bool:
    if not (_FUSED_VIT[0] and xr.is_cuda and len(experts) <= 8):
        return False
    for ex in experts:
        if not isinstance(ex, mc.Vit_expert) or ex.emb_dim != _EMB or ex.seq_ln > _TOK or ex.map_txt is None:
            return False
        if ex.map_txt.weights.shape[0] != _TIME or ex.emb_balance != 0.5:
            return False
        for blk in ex.diffit:
            t = blk.TMSA
            if (blk.skip_proj is not None or blk.resample != "keep" or blk.GN.num_groups != 4 or t.num_heads != 8
                    or blk.res_balance != 0.5 or t.attn_balance != 0.5 or t.q_time is None
                    or t.q_time.weights.shape[1] != _TIME or blk.GN.eps != 1e-5 or blk.norm1.eps != 1e-5
                    or blk.linear2.weights.shape[0] != 4 * _EMB):
                return False
    return True


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class _Meta:
    """Offset tables of the blocks inside the prepared-weight buffer and the aux buffer."""

    def __init__(self, tokens: List[int], w_off: List[List[int]], a_off: List[List[int]]):
        E = len(tokens)
        self.E, self.nb = E, len(w_off)
        self.tokens = (ctypes.c_int32 * E)(*tokens)
        self.w_off = [(ctypes.c_int64 * E)(*o) for o in w_off]
        self.a_off = [(ctypes.c_int64 * E)(*o) for o in a_off]


class _VitBlocksFn(torch.autograd.Function):
    """All DiffiT blocks (+ the final LayerNorm) of every expert: tok [R, 64, 32] fp32 -> [R, 64, 32]."""

    @staticmethod
    def forward(ctx, tok, time, row_expert, aux, meta, w_flat, *w_views):
        lib = L.lib()
        R = tok.shape[0]
        xs = [tok.contiguous()]
        time = time.contiguous()
        for b in range(meta.nb):
            out = torch.empty_like(xs[-1])
            L.check(lib.hdmoe_vit_block_fwd(_p(xs[-1]), _p(time), _p(row_expert), _p(w_flat), _p(aux), meta.w_off[b],
                                            meta.a_off[b], meta.tokens, meta.E, R, int(b == meta.nb - 1), _p(out), _st()),
                    "vit_block_fwd")
            xs.append(out)
        ctx.meta, ctx.w_flat, ctx.shapes = meta, w_flat, [v.shape for v in w_views]
        ctx.save_for_backward(time, row_expert, aux, *xs[:-1])
        return xs[-1]

    @staticmethod
    def backward(ctx, g):
        lib = L.lib()
        meta, w_flat = ctx.meta, ctx.w_flat
        time, row_expert, aux, *xs = ctx.saved_tensors
        R = xs[0].shape[0]
        g = g.contiguous()
        d_w = torch.zeros_like(w_flat)
        d_aux = torch.zeros_like(aux)
        d_time = None
        for b in reversed(range(meta.nb)):
            d_tok = torch.empty_like(g)
            dt = torch.empty_like(time)
            L.check(lib.hdmoe_vit_block_bwd(_p(xs[b]), _p(time), _p(row_expert), _p(w_flat), _p(aux), meta.w_off[b],
                                            meta.a_off[b], meta.tokens, meta.E, R, int(b == meta.nb - 1), _p(g), _p(d_tok),
                                            _p(dt), _p(d_w), _p(d_aux), _st()), "vit_block_bwd")
            g = d_tok
            d_time = dt if d_time is None else d_time + dt
        # gradient of every prepared weight = its slice of d_w (modules outside the blocks get zeros)
        d_views, o = [], 0
        for shp in ctx.shapes:
            n = 1
            for s in shp:
                n *= s
            d_views.append(d_w[o:o + n].view(shp))
            o += n
        return (g, d_time, None, d_aux, None, None, *d_views)


class FusedVitExperts:
    """Runner for one ModuleList of Vit_expert (cached on the list, like GroupedUnetExperts)."""

    def __init__(self, experts):
        self.experts = list(experts)
        self.group = prepared.vit_expert_group(self.experts, torch.float32)
        self.meta, self._meta_base = None, None

    def _build_meta(self):
        grp = self.group
        base = grp.w_hat_flat.data_ptr()
        pos = {id(mod): (v.data_ptr() - base) // 4 for mod, v in zip(grp.mods, grp.w_hat_views)}
        nb = len(self.experts[0].diffit)
        w_off = [[pos[id(ex.diffit[b].linear1)] for ex in self.experts] for b in range(nb)]
        for b in range(nb):                     # the ten matrices of a block must lie in the kernel's order
            for e, ex in enumerate(self.experts):
                blk, t = ex.diffit[b], ex.diffit[b].TMSA
                mods = [blk.linear1, t.q_proj, t.k_proj, t.v_proj, t.out_proj, t.q_time, t.k_time, t.v_time, blk.linear2,
                        blk.linear3]
                o = w_off[b][e]
                for mod in mods:
                    assert pos[id(mod)] == o, "prepared-weight layout does not match vit_block.cu"
                    o += mod.weights.numel()
                assert o - w_off[b][e] == _WBLOCK
        a_off, o = [[0] * len(self.experts) for _ in range(nb)], 0
        for e, ex in enumerate(self.experts):
            for b in range(nb):
                a_off[b][e] = o
                o += 256 + 8 * ex.seq_ln * ex.seq_ln
        self.meta = _Meta([ex.seq_ln for ex in self.experts], w_off, a_off)

    def _aux(self) -> torch.Tensor:
        parts = []
        for ex in self.experts:
            for blk in ex.diffit:
                parts += [blk.GN.weight, blk.GN.bias, blk.norm1.weight, blk.norm1.bias, blk.norm2.weight, blk.norm2.bias,
                          ex.norm.weight, ex.norm.bias, blk.TMSA.rel_pos_bias.reshape(-1)]
        return torch.cat([p.float() for p in parts])

    def __call__(self, plan, xr, tr, txr, nhwc_out: bool = False):
        experts, training = self.experts, self.experts[0].training
        R = xr.shape[0]
        x32, t32 = xr.float(), tr.float()
        tx32 = None if txr is None else txr.float()
        row_e = plan.row_expert
        with self.group.prepared(training, plan.counts) as pc:
            if self.meta is None or self._meta_base != self.group.w_hat_flat.data_ptr():
                self._build_meta()          # first call, or the group's buffers were rebuilt (model moved)
                self._meta_base = self.group.w_hat_flat.data_ptr()
            # per expert: patchify (+ pos_emb) and the block conditioning vector, selected by the row's expert
            tok, cond = None, t32.new_zeros(R, _TIME)
            zero = torch.zeros((), dtype=torch.float32, device=x32.device)
            geo = []
            for e, ex in enumerate(experts):
                sel = (row_e == e)
                B, _, H, W = x32.shape
                p = ex.patch.kernel_size[0]
                ph, pw = (p - H % p) % p, (p - W % p) % p
                xe = F.pad(x32, (0, pw, 0, ph)) if (ph or pw) else x32
                hp, wp = xe.shape[-2] // p, xe.shape[-1] // p
                assert hp * wp == ex.seq_ln, f"Sequence length mismatch: Got {hp * wp}, expected {ex.seq_ln}"
                cols = xe.reshape(B, -1, hp, p, wp, p).permute(0, 2, 4, 1, 3, 5).reshape(B, hp * wp, -1)
                te = F.linear(cols, ex.patch.weight.float().flatten(1), ex.patch.bias.float()) + ex.pos_emb.float()
                te = torch.where(sel.view(-1, 1, 1), F.pad(te, (0, 0, 0, _TOK - ex.seq_ln)), zero)
                tok = te if tok is None else tok + te
                ce = t32
                if tx32 is not None:
                    with m.active_flag(plan.counts[e] > 0):
                        ce = m.mp_sum(t32, ex.map_txt(tx32), t=ex.emb_balance)
                cond = torch.where(sel.view(-1, 1), ce, cond)
                geo.append((p, hp, wp, ph, pw, H, W, sel))
            out = _VitBlocksFn.apply(tok, cond, row_e, self._aux(), self.meta, self.group.w_hat_flat, *pc.tensors)
            # per expert: un-patchify its rows
            res = None
            for e, ex in enumerate(experts):
                p, hp, wp, ph, pw, H, W, sel = geo[e]
                with m.active_flag(plan.counts[e] > 0):
                    y = ex.unpatch_proj(out[:, :ex.seq_ln].reshape(R * ex.seq_ln, _EMB))
                if nhwc_out:
                    # PixelShuffle straight into channels-last: out[r, h*p+i, w*p+j, c] = y[r, (h, w), c*p*p + i*p + j]
                    Cn = y.shape[-1] // (p * p)
                    y = y.reshape(R, hp, wp, Cn, p, p).permute(0, 1, 4, 2, 5, 3).reshape(R, hp * p, wp * p, Cn)
                    if ph or pw:
                        y = y[:, :H, :W, :]
                else:
                    y = y.reshape(R, ex.seq_ln, -1).transpose(1, 2).reshape(R, -1, hp, wp)
                    y = ex.unpatch(y)
                    if ph or pw:
                        y = y[:, :, :H, :W]
                y = torch.where(sel.view(-1, 1, 1, 1), y, torch.zeros((), dtype=y.dtype, device=y.device))
                res = y if res is None else res + y
        return res.to(xr.dtype)
