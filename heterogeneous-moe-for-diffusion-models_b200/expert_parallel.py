"""Expert parallelism for the MoE layers over one NVSwitch box.

One process per GPU (torch.distributed, NCCL).  The dense trunk stays data-parallel; for each MoE layer the
tokens of every rank are exchanged with ONE variable-split all-to-all before the experts (dispatch) and ONE
after (combine).  The reference has no distributed code at all; the single-GPU result on the
same global batch is the parity target.

Per layer:
  1. the local dispatch plan is built on the router output with the expert columns reordered by owner rank and
     spread into one fixed-size segment per destination rank: the gather kernel writes the all-to-all send buffers
     (image, time, text rows) directly;
  2. per-expert counts are all-gathered (G x E) and stay on the device;
  3. equal-split all-to-alls of the row buffers (static sizes: the layer is CUDA-graph capturable);
  4. received rows are compacted expert-major with one index gather computed from the counts on the device, the local
     experts run (grouped tcgen05 path when available), outputs return to their arrival slots with one index copy;
  5. all-to-all back; gate-weighted combine with the local plan, ascending expert order (the weights are applied at
     the token's home rank in fp32, preserving the reference's summation order).

Placement balances COST, not count: a 5x5 U-Net expert costs 2.7x a 3x3 one.
"""
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


# ------------------------------------------------------------------------------------------------- placement
@dataclass
class ExpertPlacement:
    owner: List[int]            # owner[e] = rank that runs expert e
    world: int

    @staticmethod
    def balanced(costs: Sequence[float], world: int) -> "ExpertPlacement":
        """Longest-processing-time greedy: heaviest expert to the least loaded rank (ties: lowest rank)."""
        load = [0.0] * world
        owner = [0] * len(costs)
        for e in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
            r = min(range(world), key=lambda j: (load[j], j))
            owner[e] = r
            load[r] += costs[e]
        return ExpertPlacement(owner, world)

    def order(self) -> List[int]:
        """Expert ids sorted by (owner, id): the column order that makes permuted rows destination-major."""
        return sorted(range(len(self.owner)), key=lambda e: (self.owner[e], e))

    def local(self, rank: int) -> List[int]:
        return [e for e in self.order() if self.owner[e] == rank]


def unet_expert_costs(kernel_sizes: Sequence[int]) -> List[float]:
    """Relative forward cost of a U-Net expert by kernel size (802 MF for 3x3, 2179 MF for 5x5)."""
    return [float(k * k) + 0.15 for k in kernel_sizes]


def split_sizes(counts_all: torch.Tensor, placement: ExpertPlacement, order: List[int], rank: int):
    """counts_all[s, j] = rows rank s routes to expert order[j].  -> (send_splits, recv_splits, recv_counts)
    where recv_counts[s] lists, per local expert (in `order`), the rows arriving from rank s."""
    G = placement.world
    c = counts_all.tolist()
    mine = [j for j, e in enumerate(order) if placement.owner[e] == rank]
    send = [sum(c[rank][j] for j, e in enumerate(order) if placement.owner[e] == g) for g in range(G)]
    recv_counts = [[c[s][j] for j in mine] for s in range(G)]
    recv = [sum(rc) for rc in recv_counts]
    return send, recv, recv_counts


# ------------------------------------------------------------------------------------------------- all-to-all
class _AllToAllEq(torch.autograd.Function):
    """Equal-split all-to-all of [G * C, D] buffers (segment g of the input goes to rank g; segment s of the output came
    from rank s).  Fixed sizes: no host-side split computation, so the call is CUDA-graph capturable; its own backward."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        x = x.contiguous()
        out = torch.empty_like(x)
        dist.all_to_all_single(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        out = torch.empty_like(g)
        dist.all_to_all_single(out, g, group=ctx.group)
        return out, None


def all_to_all_equal(x, group=None, transport: str = "nccl", key=None):
    """transport "nccl": torch.distributed all_to_all_single; "peer": pull over peer-mapped staging buffers behind a
    device-side barrier (peer.py / csrc/peer.cu) -- plain kernels, so the layer records into a CUDA graph."""
    if transport == "peer":
        from . import peer
        return peer.all_to_all_equal(x, key, group)
    return _AllToAllEq.apply(x, group)


# ------------------------------------------------------------------------------------------------- the layer
@dataclass
class LocalOps:
    """Device-local building blocks; the defaults are the sm_100a kernels, tests may inject CPU stand-ins."""
    plan: Callable          # (sparse_w, top_k) -> DispatchPlan
    permute: Callable       # (plan, *srcs) -> tuple of [cap, ...]
    combine: Callable       # (rows, sparse_w, plan, out_dtype) -> [T, ...]

    @staticmethod
    def cuda() -> "LocalOps":
        from . import ops
        return LocalOps(plan=ops.dispatch_plan, permute=ops.permute,
                        combine=lambda rows, w, plan, out_dtype: ops.combine(rows, w, plan, out_dtype=out_dtype))


@dataclass
class LocalRows:
    """What the local experts see: rows grouped expert-major (`row_expert`, -1 = unused), live count on the device."""
    cap: int
    E: int
    counts: torch.Tensor
    offsets: torch.Tensor
    row_expert: torch.Tensor
    n_rows_dev: torch.Tensor
    status: torch.Tensor
    _host: Optional[List[int]] = None

    def host_offsets(self) -> List[int]:          # per-expert fallback paths only (synchronises)
        if self._host is None:
            self._host = self.offsets.tolist()
        return self._host


_CONST = {}

# device tensors of the most recent layer call (read by benchmarks / diagnostics after a step; reading synchronises)
LAST_STATS = {}

# overflow flags of the fixed-capacity exchange: ONE persistent device word per (device, layer), OR-ed into by every layer
# call with an in-place op -- so a call recorded in a CUDA graph keeps reporting on every replay -- read and cleared by
# overflowed() / check_overflow()
_OVERFLOW: dict = {}


def overflowed() -> bool:
    """True if any expert-parallel layer of THIS rank since the last call received more rows than its capacity
    (synchronises, clears the flags).  Ranks can disagree: reduce the result over the group before acting on it."""
    bad = False
    for flag in _OVERFLOW.values():
        if int(flag):
            bad = True
            flag.zero_()
    return bad


def check_overflow() -> None:
    """Raise if any expert-parallel layer since the last call received more rows than its capacity (synchronises)."""
    if overflowed():
        raise RuntimeError("hdmoe_b200 expert parallelism: a rank received more rows than capacity_factor allows; "
                           "raise capacity_factor (None = exact worst case)")


def ep_moe_layer(x: torch.Tensor, out_router: torch.Tensor, time_emb: torch.Tensor, text_emb: Optional[torch.Tensor],
                 run_local_experts: Callable, placement: ExpertPlacement, top_k: int, group=None,
                 local_ops: Optional[LocalOps] = None, payload_dtype: Optional[torch.dtype] = None,
                 capacity_factor: Optional[float] = None, transport: str = "nccl", layer_key: str = "unet") -> torch.Tensor:
    """Expert-parallel `router_to_unet_experts` (models/model_config2.py:11-39 semantics on the global batch) with
    STATIC shapes: no host read of split sizes, so the whole layer (and the train step around it) records into a
    CUDA graph.

    Layout.  C = T * k rows is what one rank can send to one peer in the worst case.  The local (destination-major)
    dispatch plan is spread into G segments of C slots (holes carry row_src = -1 and are zero-filled by the gather
    kernel), so ONE gather writes the all-to-all send buffers of the image / time / text rows directly (no packing
    copy); equal-split all-to-alls move them; the G x E per-expert counts (one all-gather of E int64) tell every rank,
    on the device, where its experts' rows sit in the received segments.  The received rows are compacted expert-major
    with one index gather (capacity `capacity_factor * C` rows, default G * C = exact worst case; an overflow sets a
    flag that `check_overflow()` raises on), the local experts run with device-side counts, the outputs return to their
    arrival slots with one index copy, travel back, and the gate-weighted combine runs at the token's home rank in
    ascending expert order (fp32, the reference's summation order).

    run_local_experts(local_ids, rows: LocalRows, x_rows, time_rows, text_rows) -> out rows [rows.cap, C, H, W]."""
    from .ops import DispatchPlan
    lo = local_ops or LocalOps.cuda()
    rank = dist.get_rank(group)
    G = placement.world
    order = placement.order()
    E = len(order)
    dev = x.device
    if text_emb is not None and text_emb.ndim == 3:
        text_emb = text_emb.mean(dim=1)
    dt = payload_dtype or x.dtype
    T = x.shape[0]
    img_shape = tuple(x.shape[1:])
    n_img = x[0].numel()
    i64 = dict(dtype=torch.int64, device=dev)
    # 1. destination-major local plan, spread into G segments of C slots
    ck = (tuple(placement.owner), str(dev), rank)
    if ck not in _CONST:      # host -> device constants are created once, outside any stream capture (warm-up iterations)
        _CONST[ck] = (torch.tensor([placement.owner[e] for e in order], **i64), torch.tensor(order, **i64),
                      torch.tensor([j for j, e in enumerate(order) if placement.owner[e] == rank], **i64))
    owner_col, order_t, mine_t = _CONST[ck]
    w_perm = out_router.index_select(1, order_t)
    plan = lo.plan(w_perm, top_k)
    C = plan.cap
    cnt = plan.counts.to(torch.int64)
    send_cnt = torch.zeros(G, **i64).index_add_(0, owner_col, cnt)
    send_end = torch.cumsum(send_cnt, 0)
    send_off = send_end - send_cnt
    r = torch.arange(C, **i64)
    dest = torch.searchsorted(send_end, r, right=True).clamp_(max=G - 1)
    slot = torch.where(r < send_end[-1], dest * C + (r - send_off[dest]), torch.full_like(r, G * C))    # G*C = dump slot

    def spread(v, fill):
        out = torch.full((G * C + 1,), fill, dtype=v.dtype, device=dev).scatter_(0, slot, v)
        out[G * C:].fill_(fill)          # (fill_ is a kernel; indexed scalar assignment would copy from the host)
        return out[:G * C]

    tok_rows = plan.tok_rows.to(torch.int64)
    tok_rows_s = torch.where(tok_rows >= 0, slot[tok_rows.clamp(min=0)], tok_rows).to(torch.int32)
    offsets_s = plan.offsets.clone()
    offsets_s[E:E + 1].fill_(G * C)                          # every slot is visited; holes are zero-filled
    splan = DispatchPlan(T, E, plan.K, G * C, plan.counts, offsets_s, spread(plan.row_src, -1),
                         spread(plan.row_expert, -1), spread(plan.row_w, 0.0), tok_rows_s, plan.status)
    srcs = [x.reshape(T, -1).to(dt), time_emb.to(dt)] + ([text_emb.to(dt)] if text_emb is not None else [])
    send = lo.permute(splan, *srcs)                          # each [G * C, D_i]: the all-to-all send buffers
    # 2. per-expert counts of every rank (device)
    if transport == "peer":
        from . import peer
        pad = (-E) % 2                                       # 16-byte segments for the pull kernel
        cnt_p = torch.cat([cnt, cnt.new_zeros(pad)]) if pad else cnt
        counts_all = peer.all_gather(cnt_p, (layer_key, "counts"), group)[:, :E]
    else:
        counts_flat = torch.empty(G * E, **i64)
        dist.all_gather_into_tensor(counts_flat, cnt.contiguous(), group=group)
        counts_all = counts_flat.view(G, E)
    # 3. dispatch all-to-alls (equal splits)
    got = [all_to_all_equal(s_, group, transport, (layer_key, "dispatch", i_)) for i_, s_ in enumerate(send)]
    # 4. compact this rank's rows expert-major
    local_ids = placement.local(rank)
    mine = [j for j, e in enumerate(order) if placement.owner[e] == rank]
    n_loc = max(len(mine), 1)
    capR = G * C if capacity_factor is None else min(G * C, int(capacity_factor * C + 0.5))
    if mine:
        n_se = counts_all.index_select(1, mine_t)            # [G, n_loc] rows from rank s for my l-th expert
    else:
        n_se = torch.zeros(G, 1, **i64)
    src_prefix = torch.cumsum(n_se, 1) - n_se                # where they start inside segment s
    seg_sizes = n_se.t().reshape(-1)                         # (expert-major, source-ascending)
    seg_end = torch.cumsum(seg_sizes, 0)
    seg_off = seg_end - seg_sizes
    total = seg_end[-1]
    j = torch.arange(capR, **i64)
    seg = torch.searchsorted(seg_end, j, right=True).clamp_(max=n_loc * G - 1)
    le = torch.div(seg, G, rounding_mode="floor")
    sr = seg - le * G
    live = j < total
    src_index = torch.where(live, sr * C + src_prefix[sr, le] + (j - seg_off[seg]), torch.zeros_like(j))
    fk = (str(dev), layer_key)
    if fk not in _OVERFLOW:              # created by the first (warm-up) call, outside any stream capture
        _OVERFLOW[fk] = torch.zeros((), dtype=torch.int32, device=dev)
    _OVERFLOW[fk].copy_(torch.maximum(_OVERFLOW[fk], (total > capR).to(torch.int32)))
    LAST_STATS.update(recv_rows=total, capacity_rows=capR, sent_rows=send_end[-1], segment_rows=C)
    grows = [g_.index_select(0, src_index) for g_ in got]
    cnt_loc = n_se.sum(0)
    rows = LocalRows(cap=capR, E=n_loc, counts=cnt_loc.to(torch.int32),
                     offsets=torch.cat([torch.zeros(1, **i64), torch.cumsum(cnt_loc, 0)]).to(torch.int32),
                     row_expert=torch.where(live, le, torch.full_like(le, -1)).to(torch.int32),
                     n_rows_dev=total.clamp(max=capR).to(torch.int32).reshape(1),
                     status=torch.zeros(1, dtype=torch.int32, device=dev))
    xr = grows[0].reshape((capR,) + img_shape)
    tr = grows[1]
    txr = grows[2] if text_emb is not None else None
    if local_ids:
        out_rows = run_local_experts(local_ids, rows, xr, tr, txr)
        out_flat = out_rows.reshape(capR, -1).to(dt)
    else:
        out_flat = grows[0] * 0
    # 5. back to the arrival slots, home, gate-weighted combine
    back_idx = torch.where(live, src_index, torch.full_like(src_index, G * C))
    back = out_flat.new_zeros(G * C + 1, n_img).index_copy(0, back_idx, out_flat)[:G * C]
    home = all_to_all_equal(back, group, transport, (layer_key, "combine"))
    return lo.combine(home.reshape((G * C,) + img_shape), w_perm, splan, x.dtype)
