"""world_size-2 gloo tests (CPU) of the N > 1 host logic: expert placement, split sizes, the dispatch / combine
all-to-all pattern of expert parallelism (with CPU stand-ins for the device-local kernels taken from the
oracle) and the data-parallel gradient all-reduce used by bench.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hdmoe_b200 import expert_parallel as EP
from hdmoe_b200.ops import DispatchPlan
from oracle import hdmoe_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cpu_plan(sparse_w, top_k):
    counts, offsets, src, exp = O.dispatch_plan(sparse_w)
    T, E = sparse_w.shape
    K = min(top_k or E, E)
    cap = T * K
    R = int(offsets[-1])
    row_src = torch.full((cap,), -1, dtype=torch.int32)
    row_exp = torch.full((cap,), -1, dtype=torch.int32)
    row_w = torch.zeros(cap)
    row_src[:R] = torch.from_numpy(src)
    row_exp[:R] = torch.from_numpy(exp)
    row_w[:R] = sparse_w.detach()[torch.from_numpy(src).long(), torch.from_numpy(exp).long()]
    tok = -torch.ones(T, K, dtype=torch.int32)
    fill = [0] * T
    for r in np.lexsort((exp, src)):
        tok[src[r], fill[src[r]]] = int(r)
        fill[src[r]] += 1
    p = DispatchPlan(T, E, K, cap, torch.from_numpy(counts.copy()), torch.from_numpy(offsets.copy()), row_src, row_exp,
                     row_w, tok, torch.zeros(1, dtype=torch.int32))
    return p


def _cpu_permute(plan, *srcs):
    """gather stand-in: slot i takes row row_src[i]; holes (row_src < 0, the spread expert-parallel layout) and the
    unused tail are zero"""
    n = int(plan.offsets[plan.E])
    live = torch.zeros(plan.cap, dtype=torch.bool)
    live[:n] = plan.row_src[:n] >= 0
    idx = plan.row_src.long().clamp(min=0)
    outs = []
    for s in srcs:
        g = s[idx]
        outs.append(torch.where(live.view((-1,) + (1,) * (s.ndim - 1)), g, torch.zeros((), dtype=s.dtype)))
    return tuple(outs)


def _cpu_combine(rows, w, plan, out_dtype):
    """out[t] = sum over the token's rows (tok_rows: ascending expert) of w[t, expert] * rows[row], mul-then-add"""
    out = torch.zeros((plan.T,) + tuple(rows.shape[1:]), dtype=rows.dtype)
    tok = plan.tok_rows.long()
    for j in range(tok.shape[1]):
        r = tok[:, j]
        ok = r >= 0
        rc = r.clamp(min=0)
        e = plan.row_expert.long()[rc].clamp(min=0)
        wt = w[torch.arange(plan.T), e]
        term = rows[rc] * wt.view((-1,) + (1,) * (rows.ndim - 1)).to(rows.dtype)
        out = out + torch.where(ok.view((-1,) + (1,) * (rows.ndim - 1)), term, torch.zeros((), dtype=rows.dtype))
    return out.to(out_dtype)


CPU_OPS = EP.LocalOps(plan=_cpu_plan, permute=_cpu_permute, combine=_cpu_combine)


def _expert(e, x, t, txt):
    return x * float(e + 1) + t.mean(1).view(-1, 1, 1, 1) - 0.5 * txt.mean(1).view(-1, 1, 1, 1)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(100 + rank)
        T, E, k = 7, 4, 2
        x = torch.randn(T, 3, 4, 4, generator=gen, requires_grad=True)
        te = torch.randn(T, 6, generator=gen)
        txt = torch.randn(T, 5, 8, generator=gen)
        lg = torch.randn(T, E, generator=gen)
        if rank == 1:
            lg[:, 2] = float("-inf")                      # rank 1 never uses expert 2
        sp, _, _, _ = O.router_gate_from_logits(lg, k)
        sp = sp.clone().requires_grad_(True)
        placement = EP.ExpertPlacement.balanced(EP.unet_expert_costs([3, 3, 5, 5]), world)
        seen = []

        def run_local(local_ids, plan, xr, tr, txr):
            off = plan.offsets.tolist()
            outs = []
            for j, e in enumerate(local_ids):
                seen.append((e, off[j + 1] - off[j]))
                outs.append(_expert(e, xr[off[j]:off[j + 1]], tr[off[j]:off[j + 1]], txr[off[j]:off[j + 1]]))
            outs.append(xr.new_zeros((plan.cap - off[-1],) + tuple(xr.shape[1:])))
            return torch.cat(outs)

        out = EP.ep_moe_layer(x, sp, te, txt, run_local, placement, k, local_ops=CPU_OPS)
        ref_x = x.detach().clone().requires_grad_(True)
        ref_sp = sp.detach().clone().requires_grad_(True)
        ref = O.moe_layer(ref_x, ref_sp, te, txt, _expert)
        gy = torch.randn(out.shape, generator=gen)
        (out * gy).sum().backward()
        (ref * gy).sum().backward()
        ok = (torch.allclose(out, ref, atol=1e-6) and torch.allclose(x.grad, ref_x.grad, atol=1e-6)
              and torch.allclose(sp.grad, ref_sp.grad, atol=1e-5))
        # only experts this rank owns ever ran here
        ok = ok and all(placement.owner[e] == rank for e, _ in seen)
        # data-parallel gradient all-reduce (bench.py's N > 1 step): mean of per-rank gradients
        g = torch.full((5,), float(rank + 1))
        dist.all_reduce(g)
        g /= world
        ok = ok and torch.allclose(g, torch.full((5,), (1 + world) / 2))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_expert_parallel_all_to_all_matches_single_process():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_placement_balances_cost_not_count():
    p = EP.ExpertPlacement.balanced(EP.unet_expert_costs([3, 3, 5, 5]), 2)
    assert sorted(p.owner[2:]) == [0, 1] and sorted(p.owner[:2]) == [0, 1]     # one 5x5 and one 3x3 per rank
    p4 = EP.ExpertPlacement.balanced(EP.unet_expert_costs([3, 3, 5, 5]), 4)
    assert sorted(p4.owner) == [0, 1, 2, 3]
    order = p.order()
    assert [p.owner[e] for e in order] == sorted(p.owner)
    counts_all = torch.tensor([[1, 2, 3, 4], [5, 6, 7, 8]])
    send, recv, rc = EP.split_sizes(counts_all, p, order, 0)
    mine = [j for j, e in enumerate(order) if p.owner[e] == 0]
    assert sum(send) == 10 and recv == [sum(counts_all[s][j].item() for j in mine) for s in range(2)]


def _worker_capacity(rank, world, port, ret):
    """capacity_factor below the worst case: exact result while the rows fit, overflow flag raised otherwise"""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(7 + rank)
        T, E, k = 9, 4, 1
        x = torch.randn(T, 2, 2, 2, generator=gen)
        te, txt = torch.randn(T, 3, generator=gen), torch.randn(T, 2, 4, generator=gen)
        placement = EP.ExpertPlacement([0, 0, 1, 1], world)

        def run_local(local_ids, plan, xr, tr, txr):
            off = plan.offsets.tolist()
            outs = [_expert(e, xr[off[j]:off[j + 1]], tr[off[j]:off[j + 1]], txr[off[j]:off[j + 1]])
                    for j, e in enumerate(local_ids)]
            n = min(off[-1], plan.cap)
            return torch.cat(outs + [xr.new_zeros((max(plan.cap - off[-1], 0),) + tuple(xr.shape[1:]))])[:plan.cap]

        res = []
        for skew in (False, True):
            lg = torch.randn(T, E, generator=gen)
            if skew:
                lg[:, 0] += 100.0                     # every token of every rank goes to expert 0 (rank 0): 2*T rows > 1.5*T
            sp, _, _, _ = O.router_gate_from_logits(lg, k)
            out = EP.ep_moe_layer(x, sp, te, txt, run_local, placement, k, local_ops=CPU_OPS, capacity_factor=1.5)
            ref = O.moe_layer(x, sp, te, txt, _expert)
            try:
                EP.check_overflow()
                flagged = False
            except RuntimeError:
                flagged = True
            res.append((bool(torch.allclose(out, ref, atol=1e-6)), flagged))
        ret[rank] = res
    finally:
        dist.destroy_process_group()


def test_expert_parallel_capacity_factor_and_overflow_flag():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker_capacity, args=(world, port, ret), nprocs=world, join=True)
    r = dict(ret)
    assert r[0][0] == (True, False) and r[1][0] == (True, False)        # balanced routing fits: exact, no flag
    assert r[0][1][1] is True                                           # rank 0 overflowed and said so
    assert r[1][1] == (False, False) or r[1][1][1] is False             # rank 1 received nothing: no flag of its own
