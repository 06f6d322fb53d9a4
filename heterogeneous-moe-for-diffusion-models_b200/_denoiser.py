"""Denoiser assembly shared by model_config1 / model_config2 (ref models/model_config{1,2}.py):
router_to_unet_experts (dispatch -> experts -> combine), HDMOEM and the EDM-preconditioned wrapper."""
import math
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import model_components as mc
from . import model_internals as util
from . import ops

# dtype of the expert path (dispatch payload, expert activations, combine input).  The reference only runs
# fp32 (quirks Q16); bf16 here is the B200 training configuration of BASELINE.json.
_EXPERT_DTYPE = [torch.float32]


def set_expert_dtype(dtype: torch.dtype) -> None:
    assert dtype in (torch.float32, torch.bfloat16)
    _EXPERT_DTYPE[0] = dtype


def get_expert_dtype() -> torch.dtype:
    return _EXPERT_DTYPE[0]


# grouped tcgen05 execution of the U-Net experts (bf16 expert path only); the switch exists for A/B parity tests
_GROUPED = [True]


def set_grouped_experts(enabled: bool) -> None:
    _GROUPED[0] = bool(enabled)


# one multi-tensor W-PREP launch for all trunk / router MP_Conv weights (prepared.py)
_TRUNK_PREP = [True]


def set_trunk_weight_prep(enabled: bool) -> None:
    _TRUNK_PREP[0] = bool(enabled)


# ViT experts without host synchronisation (see _run_experts_on_rows); off = reference-style row slicing
_SYNC_FREE = [True]


def set_sync_free_vit(enabled: bool) -> None:
    _SYNC_FREE[0] = bool(enabled)


# Independent branches on separate CUDA streams: the four ViT experts (tiny launch-latency-bound kernels, 5-13
# MFLOP per sample) and the whole ViT router + MoE branch run beside the U-Net branch.  Inside a captured step the
# fork/join becomes parallel graph branches; backward follows automatically (autograd replays every node on the
# stream of its forward).  Program order on the host is unchanged, so RNG consumption order (quirk Q2) is kept.
_BRANCH_STREAMS = [True]
_side_streams = {}


def set_branch_streams(enabled: bool) -> None:
    _BRANCH_STREAMS[0] = bool(enabled)


def _streams(device, tag: str, n: int):
    key = (device, tag)
    lst = _side_streams.setdefault(key, [])
    while len(lst) < n:
        lst.append(torch.cuda.Stream(device=device))
    return lst[:n]


class _fork:
    """with _fork(stream, inputs): ... runs the body on `stream` after everything queued on the current stream."""

    def __init__(self, stream, inputs=()):
        self.s, self.inputs = stream, inputs

    def __enter__(self):
        self.main = torch.cuda.current_stream()
        self.s.wait_stream(self.main)
        for t in self.inputs:
            if t is not None:
                t.record_stream(self.s)
        self.ctx = torch.cuda.stream(self.s)
        self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        return self.ctx.__exit__(*exc)

    def join(self, *outputs):
        main = torch.cuda.current_stream()
        main.wait_stream(self.s)
        for t in outputs:
            if t is not None:
                t.record_stream(main)


# channels-last tail of HDMOEM.forward with the fused swap / gate kernels (csrc/trunk_glue.cu); off = the reference's
# op-by-op chain (kept for A/B parity tests)
_FUSED_GLUE = [True]


def set_fused_glue(enabled: bool) -> None:
    _FUSED_GLUE[0] = bool(enabled)


# ---------------------------------------------------------------------------------------------------------
# Per-model options.  The switches above are process-wide defaults; `set_model_options(model, ...)` pins values on ONE
# model, applied for the duration of that model's forward (and therefore recorded into its autograd graph / CUDA graph),
# so two models with different settings -- an fp32 teacher next to a bf16 student, a grouped and a per-expert copy in an
# A/B test -- coexist in one process.
# ---------------------------------------------------------------------------------------------------------
_OPTION_CELLS = {"expert_dtype": _EXPERT_DTYPE, "grouped_experts": _GROUPED, "trunk_weight_prep": _TRUNK_PREP,
                 "sync_free_vit": _SYNC_FREE, "branch_streams": _BRANCH_STREAMS, "fused_glue": _FUSED_GLUE}


def set_model_options(model: nn.Module, **options) -> None:
    """Pin switches on one model (a `preconditioned_HDMOEM` or its `.net`): expert_dtype, grouped_experts,
    trunk_weight_prep, sync_free_vit, branch_streams, fused_glue.  `None` removes a pin (the process default applies)."""
    net = getattr(model, "net", model)
    pins = dict(getattr(net, "_hdmoe_options", None) or {})
    for k, v in options.items():
        if k not in _OPTION_CELLS:
            raise ValueError(f"unknown option {k!r}; known: {sorted(_OPTION_CELLS)}")
        if k == "expert_dtype" and v is not None and v not in (torch.float32, torch.bfloat16):
            raise ValueError("expert_dtype must be torch.float32 or torch.bfloat16")
        if v is None:
            pins.pop(k, None)
        else:
            pins[k] = v if k == "expert_dtype" else bool(v)
    net.__dict__["_hdmoe_options"] = pins


class _model_options:
    """Context manager: the pins of `net` replace the process defaults inside the block."""

    def __init__(self, net):
        self.pins = net.__dict__.get("_hdmoe_options") or {}

    def __enter__(self):
        self.saved = {k: _OPTION_CELLS[k][0] for k in self.pins}
        for k, v in self.pins.items():
            _OPTION_CELLS[k][0] = v
        return self

    def __exit__(self, *exc):
        for k, v in self.saved.items():
            _OPTION_CELLS[k][0] = v
        return False


# expert parallelism for the U-Net MoE layer; None = every rank runs all experts (pure DP)
_EP = {"placement": None, "group": None, "capacity_factor": None, "transport": "nccl"}


def enable_expert_parallel(kernel_sizes=None, group=None, placement=None, capacity_factor=None, transport: str = "nccl") -> None:
    """Shard the U-Net experts over the ranks of `group` (cost-balanced); ViT experts stay replicated.
    capacity_factor: rows a rank's experts may receive per layer in units of T*k (None = exact worst case, G*T*k);
    an overflow sets a device flag that expert_parallel.check_overflow() raises on.
    transport: "nccl" (torch.distributed all-to-all) or "peer" (peer-memory pull kernels behind a device-side barrier,
    one node; CUDA-graph capturable -- peer.py)."""
    if transport not in ("nccl", "peer"):
        raise ValueError("transport must be 'nccl' or 'peer'")
    import torch.distributed as dist
    from . import expert_parallel as EP
    if placement is None:
        placement = EP.ExpertPlacement.balanced(EP.unet_expert_costs(kernel_sizes), dist.get_world_size(group))
    _EP["placement"], _EP["group"], _EP["capacity_factor"], _EP["transport"] = placement, group, capacity_factor, transport


def disable_expert_parallel() -> None:
    _EP["placement"] = _EP["group"] = _EP["capacity_factor"] = None
    _EP["transport"] = "nccl"


def _run_experts_on_rows(experts, plan, xr, tr, txr, nhwc_out: bool = False):
    """experts on expert-major rows: grouped tcgen05 path when possible, per-expert loop otherwise.  nhwc_out: rows come
    back channels-last [cap, H, W, C] (natively from the grouped / fused runners, converted otherwise)."""
    if nhwc_out:
        native = None
        if (xr.dtype == torch.bfloat16 and _GROUPED[0] and len(experts) > 0
                and all(isinstance(ex, mc.Unet_expert) for ex in experts) and _groupable(experts, xr)):
            native = "unet"
        elif _SYNC_FREE[0] and all(isinstance(ex, mc.Vit_expert) for ex in experts):
            from . import vit_fused
            if vit_fused.fusable(experts, xr):
                native = "vit"
        if native is None:
            return _run_experts_on_rows(experts, plan, xr, tr, txr).permute(0, 2, 3, 1).contiguous()
    if (xr.dtype == torch.bfloat16 and _GROUPED[0] and len(experts) > 0
            and all(isinstance(ex, mc.Unet_expert) for ex in experts) and _groupable(experts, xr)):
        from .grouped import GroupedUnetExperts
        holder = experts.__dict__ if isinstance(experts, nn.ModuleList) else experts[0].__dict__
        key = "_hdmoe_grouped_" + "_".join(str(id(e)) for e in experts)
        runner = holder.get(key)
        if runner is None:
            runner = GroupedUnetExperts(experts)
            holder[key] = runner
        return runner(plan, xr, tr, txr, training=experts[0].training, nhwc_out=nhwc_out).contiguous()
    if _SYNC_FREE[0] and all(isinstance(ex, mc.Vit_expert) for ex in experts):
        from . import vit_fused
        if vit_fused.fusable(experts, xr):
            # fused DiffiT-block kernels: 4 launches forward / 4 backward for the blocks of all experts
            holder = experts.__dict__ if isinstance(experts, nn.ModuleList) else experts[0].__dict__
            runner = holder.get("_hdmoe_fused_vit")
            if runner is None:
                runner = vit_fused.FusedVitExperts(experts)
                holder["_hdmoe_fused_vit"] = runner
            return runner(plan, xr, tr, txr, nhwc_out=nhwc_out)
        # Sync-free (CUDA-graph-capturable) execution of the cheap ViT experts (5-13 MFLOP per sample):
        # every expert sees all rows with static shapes and its rows are selected on the device; the reference's
        # "only experts that received samples run" rule for the train-mode weight rewrite is kept by a device flag.
        from . import prepared
        holder = experts.__dict__ if isinstance(experts, nn.ModuleList) else experts[0].__dict__
        grp = holder.get("_hdmoe_prepared")
        if grp is None or grp.out_dtype != xr.dtype:
            grp = prepared.vit_expert_group(experts, xr.dtype)
            holder["_hdmoe_prepared"] = grp
        def one(e, expert):
            with util.active_flag(plan.counts[e] > 0):
                oe = expert(x=xr, time_emb=tr, text_emb=txr)
            sel = (plan.row_expert == e).view(-1, 1, 1, 1)
            return torch.where(sel, oe, torch.zeros((), dtype=oe.dtype, device=oe.device))

        outs = []
        with grp.prepared(experts[0].training, plan.counts):
            if _BRANCH_STREAMS[0] and xr.is_cuda:
                forks = []
                for e, (expert, st) in enumerate(zip(experts, _streams(xr.device, "vit", len(experts)))):
                    with _fork(st, (xr, tr, txr, plan.counts, plan.row_expert)) as fk:
                        outs.append(one(e, expert))
                    forks.append(fk)
                for fk, oe in zip(forks, outs):
                    fk.join(oe)
            else:
                outs = [one(e, expert) for e, expert in enumerate(experts)]
        out = outs[0]
        for oe in outs[1:]:
            out = out + oe
        return out
    off = plan.host_offsets()
    outs = []
    for e, expert in enumerate(experts):
        lo, hi = off[e], off[e + 1]
        if hi == lo:
            continue
        outs.append(expert(x=xr[lo:hi], time_emb=tr[lo:hi], text_emb=None if txr is None else txr[lo:hi]))
    R = off[-1]
    if R < plan.cap:   # unused tail rows (tokens dispatched to fewer than K experts)
        outs.append(xr.new_zeros((plan.cap - R,) + tuple(xr.shape[1:])))
    return torch.cat(outs, dim=0) if len(outs) != 1 else outs[0]


def router_to_unet_experts(x: torch.Tensor, experts: nn.ModuleList, out_router: torch.Tensor,
                           time_emb: torch.Tensor, text_emb: Optional[torch.Tensor],
                           top_k: Optional[int] = None, nhwc_out: bool = False, topk=None) -> torch.Tensor:
    """One MoE layer: same signature and result as the reference helper (models/model_config2.py:11-39).
    `nhwc_out` (B200 extra, used by HDMOEM's channels-last tail): the result comes back as [B, H, W, C].
    `topk` (B200 extra): the router kernel's (topk_idx, topk_w) of `out_router`; the plan is then built from these
    T*k pairs instead of two passes over the dense [T, E] matrix (same criterion on the same values: identical plan).

    dispatch plan (bit-exact, expert-major / token-ascending, criterion weight > 0) -> ONE fused gather of
    the image rows, time rows and mean-pooled text rows -> experts on contiguous row ranges -> ONE
    gate-weighted combine (ascending expert order, fp32 accumulate).  A single device->host copy of the
    E+1 offsets replaces the reference's >= 2E syncs."""
    if text_emb is not None and text_emb.ndim == 3:
        text_emb = text_emb.mean(dim=1)
    dt = get_expert_dtype()
    if _EP["placement"] is not None and all(isinstance(ex, mc.Unet_expert) for ex in experts):
        from . import expert_parallel as EP
        k = out_router.shape[1] if top_k is None else top_k

        def run_local(local_ids, lplan, xr_, tr_, txr_):
            return _run_experts_on_rows([experts[i] for i in local_ids], lplan, xr_, tr_, txr_)

        out = EP.ep_moe_layer(x, out_router, time_emb, text_emb, run_local, _EP["placement"], k, group=_EP["group"],
                              payload_dtype=dt, capacity_factor=_EP["capacity_factor"], transport=_EP["transport"])
        return out.permute(0, 2, 3, 1).contiguous() if nhwc_out else out
    if topk is not None and topk[0].shape[0] == out_router.shape[0]:
        plan = ops.dispatch_plan_from_topk(topk[0], topk[1], out_router.shape[1])
    else:
        plan = ops.dispatch_plan(out_router, top_k)
    srcs = [x.to(dt), time_emb.to(dt)] + ([text_emb.to(dt)] if text_emb is not None else [])
    rows = ops.permute(plan, *srcs)
    xr, tr = rows[0], rows[1]
    txr = rows[2] if text_emb is not None else None
    out_rows = _run_experts_on_rows(experts, plan, xr, tr, txr, nhwc_out=nhwc_out)
    return ops.combine(out_rows, out_router, plan, base=None, out_dtype=x.dtype)


def _groupable(experts, xr) -> bool:
    """Shape constraints of the tcgen05 grouped convolution (include/hdmoe_gemm.h)."""
    H, W = xr.shape[-2], xr.shape[-1]
    e0 = experts[0]
    levels = len(e0.block_channels)
    for lv in range(levels):
        h, w = H >> lv, W >> lv
        if w < 1 or 128 % w != 0 or (h * w) % 128 != 0 or h % (128 // w) != 0:
            return False
    chans = set(e0.block_channels) | {xr.shape[1]}
    return all(c in (32, 64, 128) for c in chans) and all(k[0] == k[1] and k[0] % 2 == 1 for k in
                                                          (ex.kernel_size for ex in experts))


class HDMOEM(nn.Module):
    """Hybrid diffusion MoE denoiser.  variant 2 = models/model_config2.py:42-303 (analytic sigma-sigmoid
    path scaling), variant 1 = models/model_config1.py:42-309 (learned Scaling_router + soft query/context
    swap).  Attribute names and state_dict keys follow the reference."""

    _variant = 2

    def __init__(self, IN_in_channels: int, IN_img_resolution: int, internal_channels: int, time_emb_dim: int,
                 text_emb_dim: int, num_experts: int, top_k: int, Fourier_bandwidth: float, VIT_num_blocks: int,
                 VIT_patch_sizes: List[int], VIT_num_groups: int, VIT_num_heads: int, VIT_emb_size: int,
                 Unet_num_blocks: int, Unet_channel_mult: List[int], Unet_kernel_sizes: List[Tuple[int, int]],
                 Unet_model_channels: Optional[int] = 192, Unet_channel_mult_emb: Optional[int] = None,
                 Unet_label_balance: Optional[float] = 0.5, Unet_concat_balance: Optional[float] = 0.5):
        super().__init__()
        C = self.internal_channels = internal_channels
        self.top_k = top_k
        self.input_proj = util.MP_Conv(IN_in_channels, C, kernel=(3, 3))
        self.Fourier_emb = util.MP_Fourier(num_channels=time_emb_dim // 2, bandwidth=Fourier_bandwidth)
        self.out_fourier1 = util.MP_Conv(time_emb_dim // 2, time_emb_dim * 2, kernel=())
        self.out_fourier2 = util.MP_Conv(time_emb_dim * 2, time_emb_dim, kernel=())
        if self._variant == 1:
            self.scaling_net = mc.Scaling_router(emb_dim=time_emb_dim, num_experts=2)
        self.Unet_router = mc.Router(in_channels=C, time_dim=time_emb_dim, top_k=top_k, num_experts=num_experts)
        self.vit_router = mc.Router(in_channels=C, time_dim=time_emb_dim, top_k=top_k, num_experts=num_experts)
        self.alpha_txt = nn.Parameter(torch.tensor(0.0))
        self.Unet_experts = nn.ModuleList(
            mc.Unet_expert(img_resolution=IN_img_resolution, img_channels=C, time_emb_dim=time_emb_dim,
                           text_emb_dim=text_emb_dim, num_blocks=Unet_num_blocks, channel_mult=Unet_channel_mult,
                           kernel_size=Unet_kernel_sizes[i], label_balance=Unet_label_balance,
                           concat_balance=Unet_concat_balance, model_channels=Unet_model_channels,
                           channel_mult_emb=Unet_channel_mult_emb) for i in range(num_experts))
        self.VIT_experts = nn.ModuleList(
            mc.Vit_expert(num_heads=VIT_num_heads, num_groups=VIT_num_groups, in_channels=C,
                          seq_ln=math.ceil(IN_img_resolution / VIT_patch_sizes[i]) ** 2, emb_dim=VIT_emb_size,
                          num_blocks=VIT_num_blocks, patch_size=VIT_patch_sizes[i], text_dim=text_emb_dim,
                          time_dim=time_emb_dim) for i in range(num_experts))
        self.cross_attn = util.MP_Attention(num_heads=VIT_num_heads, emb_dim=C, seq_ln=IN_img_resolution ** 2,
                                            context_dim=C, attn_balance=0.5, is_cross_attn=True)
        self.cross_attn_text = util.MP_Attention(num_heads=VIT_num_heads, emb_dim=C, seq_ln=IN_img_resolution ** 2,
                                                 context_dim=text_emb_dim, attn_balance=0.5, is_cross_attn=True)
        self.gate1 = util.MP_Conv(C * 2, C, kernel=(1, 1))
        self.gate2 = util.MP_Conv(C, 2, kernel=(1, 1))
        self.output_proj = util.MP_Conv(C, IN_in_channels, kernel=(3, 3))

    def _forward(self, x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta, transition_point=None,
                 softness=None, alpha_routing: float = 10, noise: Optional[dict] = None):
        with _model_options(self):            # this model's pinned switches (set_model_options) for the whole forward
            return self._forward_pinned(x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta, transition_point,
                                        softness, alpha_routing, noise)

    def _forward_pinned(self, x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta, transition_point=None,
                        softness=None, alpha_routing: float = 10, noise: Optional[dict] = None):
        noise = noise or {}
        if x.is_cuda and _TRUNK_PREP[0]:
            from . import prepared
            tc_trunk = self._router_trunk(x) is not None
            key = "_hdmoe_trunk_group" + ("_tc" if tc_trunk else "")
            grp = self.__dict__.get(key)
            if grp is None:
                grp = prepared.trunk_group(self, router_convs=not tc_trunk)
                self.__dict__[key] = grp
            with grp.prepared(self.training):
                return self._forward_body(x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta, transition_point,
                                          softness, alpha_routing, noise)
        return self._forward_body(x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta, transition_point,
                                  softness, alpha_routing, noise)

    @staticmethod
    def _topk_of(router, x):
        """(topk_idx, topk_w) the router's fused gate kernel produced in this forward (CUDA path only)."""
        last = getattr(router, "last", None)
        if not x.is_cuda or not last:
            return None
        return last["topk_idx"], last["topk_w"]

    def _router_trunk(self, x):
        """The grouped tcgen05 trunk runner of the two routers, or None when the configuration does not use it (fp32
        expert path, CPU tensors, unsupported shapes, switched off)."""
        from . import router_trunk as rt
        if not (x.is_cuda and rt.enabled() and get_expert_dtype() == torch.bfloat16):
            return None
        routers = [self.vit_router, self.Unet_router]
        if not rt.GroupedRouterTrunk.supported(routers, x):
            return None
        runner = self.__dict__.get("_hdmoe_router_trunk")
        if runner is None:
            runner = rt.GroupedRouterTrunk(routers)
            self.__dict__["_hdmoe_router_trunk"] = runner
        return runner

    def _forward_body(self, x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta, transition_point, softness,
                      alpha_routing, noise):
        B, _, H, W = x.shape
        te = self.out_fourier1(self.Fourier_emb(time_vec))
        te = self.out_fourier2(util.mp_silu(te))
        feats = self.input_proj(x)
        glue = _FUSED_GLUE[0] and x.is_cuda and x.dtype == torch.float32 and self.internal_channels == 32
        if self._variant == 2:      # models/model_config2.py:244-249
            if glue:
                scaling = ops.analytic_scaling(time_vec, transition_point, softness)      # one launch, [B, 2]
            else:
                vw = torch.sigmoid((time_vec * 4 - transition_point) / softness).view(-1, 1, 1, 1)
                scaling = torch.cat([(vw + 1e-2) * 2, ((1.0 - vw) + 1e-2) * 2], dim=1).view(-1, 2)
        else:                       # models/model_config1.py:246-249
            scaling = self.scaling_net(x=te, zeta=zeta, noise=noise.get("scaling"))
        s_vit = scaling[:, 0:1].view(-1, 1, 1, 1)
        s_unet = scaling[:, 1:2].view(-1, 1, 1, 1)
        # router trunks of both routers as grouped tcgen05 launches (bf16 configuration): pooled features up front
        trunk = self._router_trunk(x)
        pool_vit = pool_un = routed = None
        if glue:
            # one pass over feats: both scaled branch inputs and the channels-last bf16 router-trunk input
            in_vit, in_unet, trunk_in = ops.scale_pair(feats, scaling, want_trunk=trunk is not None)
            if trunk is not None and _BRANCH_STREAMS[0]:
                # The trunk runs on its own stream so that its BACKWARD does too (autograd replays a node on the stream of
                # its forward): the trunk's gradient only depends on the router-gate backward, which is available at the very
                # start of the backward pass, while its result is consumed at the very end (scale_pair backward) -- inside
                # the captured step the ~1.3 ms of trunk data / weight gradients become a graph branch beside the attention
                # backward instead of a serial tail.
                # The two router tails run on the same stream (ViT first: RNG order, quirk Q2): a gate backward on the main
                # stream would be enqueued behind the whole expert backward and hold the trunk's gradient back.
                with _fork(_streams(x.device, "trunk", 1)[0], (trunk_in, te, Vit_router_mask, Unet_router_mask)) as fkt:
                    pool_vit, pool_un = trunk(None, self.training, pre_nhwc=trunk_in)
                    routed = (self.vit_router(x=in_vit, time_emb=te, zeta=zeta, mask=Vit_router_mask,
                                              noise=noise.get("vit"), pooled=pool_vit),
                              self.Unet_router(x=in_unet, time_emb=te, zeta=zeta, mask=Unet_router_mask,
                                               noise=noise.get("unet"), pooled=pool_un))
                fkt.join(*routed[0], *routed[1], *(self._topk_of(self.vit_router, x) or ()),
                         *(self._topk_of(self.Unet_router, x) or ()))
            elif trunk is not None:
                pool_vit, pool_un = trunk(None, self.training, pre_nhwc=trunk_in)
        else:
            in_unet = s_unet * feats
            in_vit = s_vit * feats
            if trunk is not None:
                pool_vit, pool_un = trunk([in_vit, in_unet], self.training)
        # the ViT router is evaluated first (RNG order, quirk Q2)
        # (expert parallelism over NCCL keeps one stream: collectives of concurrently recorded branches must not reorder
        # between ranks; the peer-memory transport only issues its barriers on the U-Net branch, so the ViT branch may run
        # beside it)
        if _BRANCH_STREAMS[0] and x.is_cuda and (_EP["placement"] is None or _EP["transport"] == "peer"):
            # ViT branch (router + MoE layer) beside the U-Net branch; host program order as in the reference
            with _fork(_streams(x.device, "branch", 1)[0], (in_vit, te, text_emb, Vit_router_mask, pool_vit)) as fk:
                if routed is not None:
                    w_vit, p_vit, raw_vit = routed[0]
                else:
                    w_vit, p_vit, raw_vit = self.vit_router(x=in_vit, time_emb=te, zeta=zeta, mask=Vit_router_mask,
                                                            noise=noise.get("vit"), pooled=pool_vit)
            if routed is not None:
                w_un, p_un, raw_un = routed[1]
            else:
                w_un, p_un, raw_un = self.Unet_router(x=in_unet, time_emb=te, zeta=zeta, mask=Unet_router_mask,
                                                      noise=noise.get("unet"), pooled=pool_un)
            tk_u, tk_v = self._topk_of(self.Unet_router, x), self._topk_of(self.vit_router, x)
            out_u = router_to_unet_experts(in_unet, self.Unet_experts, w_un, te, text_emb, top_k=self.top_k, nhwc_out=glue,
                                           topk=tk_u)
            with torch.cuda.stream(fk.s):
                out_v = router_to_unet_experts(in_vit, self.VIT_experts, w_vit, te, text_emb, top_k=self.top_k,
                                               nhwc_out=glue, topk=tk_v)
            fk.join(w_vit, p_vit, raw_vit, out_v)
        else:
            if routed is not None:
                (w_vit, p_vit, raw_vit), (w_un, p_un, raw_un) = routed
            else:
                w_vit, p_vit, raw_vit = self.vit_router(x=in_vit, time_emb=te, zeta=zeta, mask=Vit_router_mask,
                                                        noise=noise.get("vit"), pooled=pool_vit)
                w_un, p_un, raw_un = self.Unet_router(x=in_unet, time_emb=te, zeta=zeta, mask=Unet_router_mask,
                                                      noise=noise.get("unet"), pooled=pool_un)
            tk_u, tk_v = self._topk_of(self.Unet_router, x), self._topk_of(self.vit_router, x)
            out_u = router_to_unet_experts(in_unet, self.Unet_experts, w_un, te, text_emb, top_k=self.top_k, nhwc_out=glue,
                                           topk=tk_u)
            out_v = router_to_unet_experts(in_vit, self.VIT_experts, w_vit, te, text_emb, top_k=self.top_k, nhwc_out=glue,
                                           topk=tk_v)
        if glue:
            # channels-last tail: out_u / out_v are [B, H, W, C]; swap (cfg1) and the whole text-blend / gate / mix chain
            # are one fused kernel each (csrc/trunk_glue.cu); output_proj reads the mix through a channels-last view
            C = self.internal_channels
            uf, vf = out_u.reshape(B, H * W, C), out_v.reshape(B, H * W, C)
            if self._variant == 2:
                q, ctx = uf, vf
            else:
                stronger = torch.sigmoid(alpha_routing * (s_vit - s_unet)).reshape(-1)
                q, ctx = ops.trunk_swap(uf, vf, stronger)
            a = self.cross_attn(query=q, context=ctx, gain_s=1.0, gain_t=1.0)
            b = self.cross_attn_text(query=a, context=text_emb, gain_s=1.0, gain_t=1.0)
            mix, g = ops.trunk_gate(uf, a, b, self.alpha_txt, self.gate1.prepared_weight(1.0, torch.float32),
                                    self.gate2.prepared_weight(1.0, torch.float32), H, W)
            out = self.output_proj(mix.reshape(B, H, W, C).permute(0, 3, 1, 2))
            return out, p_un, raw_un, p_vit, raw_vit, scaling, g
        uf = out_u.flatten(2).transpose(1, 2)
        vf = out_v.flatten(2).transpose(1, 2)
        if self._variant == 2:
            q, ctx = uf, vf
        else:                       # models/model_config1.py:277-283
            stronger = torch.sigmoid(alpha_routing * (s_vit - s_unet)).view(-1, 1, 1)
            q = stronger * vf + (1 - stronger) * uf
            ctx = stronger * uf + (1 - stronger) * vf
        a = self.cross_attn(query=q, context=ctx, gain_s=1.0, gain_t=1.0)
        b = self.cross_attn_text(query=a, context=text_emb, gain_s=1.0, gain_t=1.0)
        fin = a + self.alpha_txt * (b - a)
        img = fin.transpose(1, 2).reshape(B, self.internal_channels, H, W)
        g = self.gate2(util.mp_silu(self.gate1(util.mp_cat(out_u, img, dim=1))))
        g = F.softmax(g, dim=1)
        gated = g[:, 0:1] * out_u + g[:, 1:2] * img
        out = self.output_proj(util.mp_sum(out_u, gated, t=0.5))
        return out, p_un, raw_un, p_vit, raw_vit, scaling, g


class preconditioned_HDMOEM(nn.Module):
    """EDM preconditioning wrapper (ref models/model_config2.py:306-468).  c_in scaling and the
    c_skip/c_out output mix are fused elementwise kernels (ops.edm_precond_in/out); quirk Q1 (the skip uses
    the already scaled input) is preserved."""

    _net_cls = HDMOEM

    def __init__(self, IN_in_channels: int, IN_img_resolution: int, internal_channels: int, time_emb_dim: int,
                 text_emb_dim: int, num_experts: int, top_k: int, Fourier_bandwidth: float, VIT_num_blocks: int,
                 VIT_patch_sizes: List[int], VIT_num_groups: int, VIT_num_heads: int, VIT_emb_size: int,
                 Unet_num_blocks: int, Unet_channel_mult: List[int], Unet_kernel_sizes: List[Tuple[int, int]],
                 Unet_model_channels: Optional[int] = 192, Unet_channel_mult_emb: Optional[int] = None,
                 Unet_label_balance: Optional[float] = 0.5, Unet_concat_balance: Optional[float] = 0.5,
                 sigma_data: Optional[float] = 0.5, log_var_channels: Optional[int] = 128):
        super().__init__()
        self.sigma_data = sigma_data
        self.log_var_channels = log_var_channels
        self.num_experts = num_experts
        self.log_var_fourier = util.MP_Fourier(num_channels=log_var_channels)
        self.log_var_linear = util.MP_Conv(log_var_channels, 1, kernel=())
        self.net = self._net_cls(IN_in_channels=IN_in_channels, IN_img_resolution=IN_img_resolution,
                                 internal_channels=internal_channels, time_emb_dim=time_emb_dim,
                                 text_emb_dim=text_emb_dim, num_experts=num_experts, top_k=top_k,
                                 Fourier_bandwidth=Fourier_bandwidth, VIT_num_blocks=VIT_num_blocks,
                                 VIT_patch_sizes=VIT_patch_sizes, VIT_num_groups=VIT_num_groups,
                                 VIT_num_heads=VIT_num_heads, VIT_emb_size=VIT_emb_size,
                                 Unet_num_blocks=Unet_num_blocks, Unet_channel_mult=Unet_channel_mult,
                                 Unet_kernel_sizes=Unet_kernel_sizes, Unet_model_channels=Unet_model_channels,
                                 Unet_channel_mult_emb=Unet_channel_mult_emb, Unet_label_balance=Unet_label_balance,
                                 Unet_concat_balance=Unet_concat_balance)

    def _forward(self, x, sigma, text_emb, Unet_router_mask, Vit_router_mask, zeta, return_log_var=False,
                 precomputed_x_in: Optional[torch.Tensor] = None, raw_output: bool = False, **net_kw):
        sigma = sigma.to(torch.float32)
        c_noise = sigma.flatten().log() / 4
        B = x.shape[0]
        if c_noise.shape[0] == 1 and B > 1:
            c_noise = c_noise.expand(B)
        x_in = precomputed_x_in if precomputed_x_in is not None else ops.edm_precond_in(x, sigma, self.sigma_data)
        out_net, p_un, raw_un, p_vit, raw_vit, scaling, gate = self.net(
            x=x_in, text_emb=text_emb, time_vec=c_noise, Unet_router_mask=Unet_router_mask,
            Vit_router_mask=Vit_router_mask, zeta=zeta, **net_kw)
        # raw_output: the fused Heun kernels apply c_skip / c_out themselves (EDM_sampler.py fast path)
        D_x = out_net if raw_output else ops.edm_precond_out(x_in, out_net, sigma, self.sigma_data)
        log_var = None
        if return_log_var:
            log_var = self.log_var_linear(self.log_var_fourier(c_noise)).reshape(-1, 1, 1, 1)
        return {"denoised": D_x, "Unet_router_loss": p_un, "Unet_raw": raw_un, "vit_router_loss": p_vit,
                "vit_raw": raw_vit, "scaling_net_out": scaling, "out_gate": gate, "log_var": log_var}
