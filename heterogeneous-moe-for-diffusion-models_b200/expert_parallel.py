"""Expert parallelism for the MoE layers over one NVSwitch box.

One process per GPU (torch.distributed, NCCL).  The dense trunk stays data-parallel; for each MoE layer the
tokens of every rank are exchanged with ONE variable-split all-to-all before the experts (dispatch) and ONE
after (combine).  The reference has no distributed code at all; the single-GPU result on the
same global batch is the parity target.

Per layer:
  1. the local dispatch plan is built on the router output with the expert columns reordered by owner rank,
     so the permuted rows are already destination-major: the gather kernel writes the all-to-all send buffer
     directly (image | time | text packed per row);
  2. per-expert counts are all-gathered (G x E int32) -> send / receive split sizes;
  3. all-to-all-v of the packed rows;
  4. received rows are regrouped expert-major with a second (one-hot) dispatch plan, the local experts run
     (grouped tcgen05 path when available), the rows are put back in arrival order;
  5. all-to-all-v back; gate-weighted combine with the local plan, ascending expert order (the weights are
     applied at the token's home rank in fp32, preserving the reference's summation order).

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
class _AllToAllV(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, send, recv, group):
        ctx.send, ctx.recv, ctx.group = send, recv, group
        out = x.new_empty((sum(recv),) + tuple(x.shape[1:]))
        dist.all_to_all_single(out, x.contiguous(), output_split_sizes=recv, input_split_sizes=send, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        out = g.new_empty((sum(ctx.send),) + tuple(g.shape[1:]))
        dist.all_to_all_single(out, g.contiguous(), output_split_sizes=ctx.send, input_split_sizes=ctx.recv,
                               group=ctx.group)
        return out, None, None, None


def all_to_all_rows(x, send, recv, group=None):
    return _AllToAllV.apply(x, list(send), list(recv), group)


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


def ep_moe_layer(x: torch.Tensor, out_router: torch.Tensor, time_emb: torch.Tensor, text_emb: Optional[torch.Tensor],
                 run_local_experts: Callable, placement: ExpertPlacement, top_k: int, group=None,
                 local_ops: Optional[LocalOps] = None, payload_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Expert-parallel `router_to_unet_experts` (models/model_config2.py:11-39 semantics on the global batch).

    run_local_experts(local_ids, plan, x_rows, time_rows, text_rows) -> out rows [plan.cap, C, H, W]: runs this
    rank's experts (local_ids, in placement order) on rows grouped expert-major by `plan`."""
    lo = local_ops or LocalOps.cuda()
    rank = dist.get_rank(group)
    G = placement.world
    order = placement.order()
    if text_emb is not None and text_emb.ndim == 3:
        text_emb = text_emb.mean(dim=1)
    dt = payload_dtype or x.dtype
    T = x.shape[0]
    img_shape = tuple(x.shape[1:])
    n_img = x[0].numel()
    # 1. destination-major local plan
    w_perm = out_router[:, order]
    plan = lo.plan(w_perm, top_k)
    packed = torch.cat([x.reshape(T, -1).to(dt), time_emb.to(dt)] + ([text_emb.to(dt)] if text_emb is not None else []),
                       dim=1)
    (rows,) = lo.permute(plan, packed)
    # 2. counts -> splits (one G x E int32 all-gather + one host read)
    counts = plan.counts.to(torch.int64)
    gathered = [torch.empty_like(counts) for _ in range(G)]
    dist.all_gather(gathered, counts, group=group)
    counts_all = torch.stack(gathered).cpu()
    send, recv, recv_counts = split_sizes(counts_all, placement, order, rank)
    R = sum(send)
    # 3. dispatch all-to-all
    got = all_to_all_rows(rows[:R], send, recv, group)
    # 4. regroup expert-major on this rank, run the local experts, restore arrival order
    local_ids = placement.local(rank)
    n_loc = max(len(local_ids), 1)
    flat = [c for rc in recv_counts for c in rc]
    ids = torch.arange(len(local_ids), device=x.device).repeat(G) if local_ids else torch.zeros(0, dtype=torch.long,
                                                                                             device=x.device)
    row_e = torch.repeat_interleave(ids, torch.tensor(flat, device=x.device, dtype=torch.long)) if flat else ids
    Rr = got.shape[0]
    if Rr > 0:
        onehot = torch.zeros(Rr, n_loc, dtype=torch.float32, device=x.device)
        onehot[torch.arange(Rr, device=x.device), row_e] = 1.0
        lplan = lo.plan(onehot, 1)
        (grows,) = lo.permute(lplan, got)
        xr = grows[:, :n_img].reshape((-1,) + img_shape)
        tr = grows[:, n_img:n_img + time_emb.shape[1]]
        txr = grows[:, n_img + time_emb.shape[1]:] if text_emb is not None else None
        out_rows = run_local_experts(local_ids, lplan, xr, tr, txr)
        back = lo.combine(out_rows.reshape(out_rows.shape[0], -1).to(dt).contiguous(), onehot, lplan, dt)
    else:
        back = got[:, :n_img]
    # 5. combine all-to-all and gate-weighted sum at the home rank
    home = all_to_all_rows(back, recv, send, group)
    pad = plan.cap - R
    if pad > 0:
        home = torch.cat([home, home.new_zeros(pad, home.shape[1])], dim=0)
    out = lo.combine(home.reshape((plan.cap,) + img_shape).contiguous(), w_perm, plan, x.dtype)
    return out
