"""torch.autograd bindings of the sm_100a kernels (C ABI in include/hdmoe_b200.h).

Every function here launches hand-written CUDA through ctypes on torch's current stream.  There is no
CPU / eager fallback: non-CUDA tensors raise.
"""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"hdmoe_b200: unsupported dtype {t.dtype} (float32 / bfloat16 only)") from None


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("hdmoe_b200 kernels need CUDA tensors; there is no CPU fallback")


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_ws_cache = {}


def _zero_workspace(nbytes: int, device) -> torch.Tensor:
    """Persistent zero-initialised workspace (the kernels leave it zeroed again)."""
    key = (device, torch.cuda.current_stream(device).cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


# ----------------------------------------------------------------------------------------------------
# (1) router gate
# ----------------------------------------------------------------------------------------------------
class _RouterGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pooled, cond, w_hat, noise, zeta, mask, logits_in, top_k):
        lib = L.lib()
        src = logits_in if logits_in is not None else pooled
        _cuda(src, cond, w_hat, noise, mask)
        T = src.shape[0]
        if logits_in is not None:
            E, Cc = logits_in.shape[1], 0
            logits_in = _f32c(logits_in)
        else:
            pooled, w_hat, cond = _f32c(pooled), _f32c(w_hat), _f32c(cond)
            E, Cc = w_hat.shape
            assert pooled.shape == (T, Cc) and (cond is None or cond.shape == (T, 2 * Cc))
        noise, mask = _f32c(noise), _f32c(mask)
        dev = src.device
        o = dict(dtype=torch.float32, device=dev)
        logits, probs, sparse = (torch.empty(T, E, **o) for _ in range(3))
        idx = torch.empty(T, top_k, dtype=torch.int32, device=dev)
        tw = torch.empty(T, top_k, **o)
        stats = torch.empty(2 * E + 1, **o)
        ws = _zero_workspace(lib.hdmoe_router_gate_workspace_bytes(T, E), dev)
        L.check(lib.hdmoe_router_gate_fwd(_p(pooled), _p(cond), _p(w_hat), _p(noise), float(zeta), _p(mask),
                                          _p(logits_in), T, Cc, E, top_k, _p(logits), _p(probs), _p(sparse),
                                          _p(idx), _p(tw), _p(stats), _p(ws), _st()), "router_gate_fwd")
        ctx.save_for_backward(pooled, cond, w_hat, logits, idx)
        ctx.dims = (T, Cc, E, top_k, logits_in is not None)
        ctx.mark_non_differentiable(idx, tw)
        return sparse, probs, logits, idx, tw, stats

    @staticmethod
    def backward(ctx, g_sparse, g_probs, g_logits, _gi, _gw, g_stats):
        pooled, cond, w_hat, logits, idx = ctx.saved_tensors
        T, Cc, E, k, teacher = ctx.dims
        lib = L.lib()
        g_sparse, g_probs, g_logits, g_stats = (_f32c(g) for g in (g_sparse, g_probs, g_logits, g_stats))
        if teacher:
            d_logits = torch.empty_like(logits)
            L.check(lib.hdmoe_router_gate_bwd(None, None, None, _p(logits), _p(idx), _p(g_sparse), _p(g_probs),
                                              _p(g_logits), _p(g_stats), T, 0, E, k, None, None, None,
                                              _p(d_logits), _st()), "router_gate_bwd")
            return None, None, None, None, None, None, d_logits, None
        d_pooled = torch.empty_like(pooled)
        d_cond = torch.empty_like(cond) if cond is not None else None
        d_w = torch.empty_like(w_hat)
        L.check(lib.hdmoe_router_gate_bwd(_p(pooled), _p(cond), _p(w_hat), _p(logits), _p(idx), _p(g_sparse),
                                          _p(g_probs), _p(g_logits), _p(g_stats), T, Cc, E, k, _p(d_pooled),
                                          _p(d_cond), _p(d_w), None, _st()), "router_gate_bwd")
        return d_pooled, d_cond, d_w, None, None, None, None, None


def router_gate(pooled, cond, w_hat, top_k: int, noise=None, zeta: float = 0.0, mask=None):
    """Fused router tail.  Returns (sparse_w, gate_probs, logits, topk_idx[int32], topk_w, stats);
    stats = [sum_t probs (E) | dispatch counts (E) | sum_t z-term (1)].  See include/hdmoe_b200.h §1."""
    return _RouterGate.apply(pooled, cond, w_hat, noise, zeta, mask, None, top_k)


def router_gate_from_logits(logits, top_k: int, mask=None):
    """Teacher-forced gate on given (already masked) logits: bit-exact index contract."""
    return _RouterGate.apply(None, None, None, None, 0.0, mask, logits, top_k)


# ----------------------------------------------------------------------------------------------------
# (2) dispatch plan, permute
# ----------------------------------------------------------------------------------------------------
@dataclass
class DispatchPlan:
    """Integer dispatch plan (device tensors; see include/hdmoe_b200.h §2)."""
    T: int
    E: int
    K: int
    cap: int
    counts: torch.Tensor
    offsets: torch.Tensor
    row_src: torch.Tensor
    row_expert: torch.Tensor
    row_w: torch.Tensor
    tok_rows: torch.Tensor
    status: torch.Tensor
    _host_offsets: Optional[List[int]] = None

    @property
    def n_rows_dev(self) -> torch.Tensor:
        return self.offsets[self.E:self.E + 1]

    def host_offsets(self) -> List[int]:
        """offsets as python ints -- ONE device->host copy (the reference syncs >= 2E times per layer)."""
        if self._host_offsets is None:
            both = torch.cat([self.offsets, self.status]).tolist()
            if both[-1] != 0:
                raise RuntimeError(f"dispatch plan overflow (status {both[-1]}): cap={self.cap}, K={self.K}")
            self._host_offsets = both[:-1]
        return self._host_offsets


def dispatch_plan(sparse_w: torch.Tensor, top_k: Optional[int] = None) -> DispatchPlan:
    """Expert-major, token-ascending dispatch plan of the entries with sparse_w > 0 (bit-exact)."""
    _cuda(sparse_w)
    lib = L.lib()
    w = _f32c(sparse_w.detach())
    T, E = w.shape
    K = E if top_k is None else min(int(top_k), E)
    cap = T * K
    dev = w.device
    i32 = dict(dtype=torch.int32, device=dev)
    counts, offsets = torch.empty(E, **i32), torch.empty(E + 1, **i32)
    row_src, row_expert = torch.empty(cap, **i32), torch.empty(cap, **i32)
    row_w = torch.empty(cap, dtype=torch.float32, device=dev)
    tok_rows = torch.empty(T, K, **i32)
    status = torch.empty(1, **i32)
    ws = torch.empty(lib.hdmoe_dispatch_plan_workspace_bytes(T, E), dtype=torch.uint8, device=dev)
    L.check(lib.hdmoe_dispatch_plan(_p(w), T, E, cap, K, _p(counts), _p(offsets), _p(row_src), _p(row_expert),
                                    _p(row_w), _p(tok_rows), _p(status), _p(ws), _st()), "dispatch_plan")
    return DispatchPlan(T, E, K, cap, counts, offsets, row_src, row_expert, row_w, tok_rows, status)


def dispatch_plan_from_topk(topk_idx: torch.Tensor, topk_w: torch.Tensor, num_experts: int) -> DispatchPlan:
    """The same plan built from the router kernel's top-k output (router_gate's topk_idx int32 [T, K], topk_w [T, K])
    instead of the dense [T, E] sparse-weight matrix: identical result (same criterion, weight > 0, on the same
    values), T*K*8 bytes read."""
    _cuda(topk_idx, topk_w)
    lib = L.lib()
    idx = topk_idx.detach().to(torch.int32).contiguous()
    tw = _f32c(topk_w.detach())
    T, K = idx.shape
    E = int(num_experts)
    cap = T * K
    dev = idx.device
    i32 = dict(dtype=torch.int32, device=dev)
    counts, offsets = torch.empty(E, **i32), torch.empty(E + 1, **i32)
    row_src, row_expert = torch.empty(cap, **i32), torch.empty(cap, **i32)
    row_w = torch.empty(cap, dtype=torch.float32, device=dev)
    tok_rows = torch.empty(T, K, **i32)
    status = torch.empty(1, **i32)
    ws = torch.empty(lib.hdmoe_dispatch_plan_workspace_bytes(T, E), dtype=torch.uint8, device=dev)
    L.check(lib.hdmoe_dispatch_plan_topk(_p(idx), _p(tw), T, E, K, cap, _p(counts), _p(offsets), _p(row_src), _p(row_expert),
                                         _p(row_w), _p(tok_rows), _p(status), _p(ws), _st()), "dispatch_plan_topk")
    return DispatchPlan(T, E, K, cap, counts, offsets, row_src, row_expert, row_w, tok_rows, status)


def _permute_raw(srcs: Sequence[torch.Tensor], plan: DispatchPlan) -> List[torch.Tensor]:
    lib = L.lib()
    outs, n = [], len(srcs)
    for lo in range(0, n, 4):
        grp = [s.contiguous() for s in srcs[lo:lo + 4]]
        dst = [torch.empty((plan.cap,) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device) for s in grp]
        m = len(grp)
        a_src = (C.c_void_p * m)(*[s.data_ptr() for s in grp])
        a_dst = (C.c_void_p * m)(*[d.data_ptr() for d in dst])
        a_rb = (C.c_int64 * m)(*[s[0].numel() * s.element_size() for s in grp])
        L.check(lib.hdmoe_permute_rows(a_src, a_dst, a_rb, m, _p(plan.row_src), _p(plan.n_rows_dev), plan.cap,
                                       _st()), "permute_rows")
        outs += dst
    return outs


def _combine_raw(rows, tok_rows, row_w, base, out_dtype, T, K):
    lib = L.lib()
    rows = rows.contiguous()
    D = rows[0].numel()
    out = torch.empty((T,) + tuple(rows.shape[1:]), dtype=out_dtype, device=rows.device)
    if base is not None:
        base = base.to(out_dtype).contiguous()
    L.check(lib.hdmoe_combine_rows(_p(rows), _dt(rows), _p(tok_rows), _p(row_w), _p(base), _p(out), _dt(out),
                                   T, K, D, _st()), "combine_rows")
    return out


class _Permute(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, *srcs):
        _cuda(*srcs)
        for s in srcs:
            if s.shape[0] != plan.T:
                raise RuntimeError(f"permute: leading dim {s.shape[0]} != plan.T {plan.T}")
        ctx.plan = plan
        return tuple(_permute_raw(srcs, plan))

    @staticmethod
    def backward(ctx, *grads):
        plan = ctx.plan
        out = []
        for g in grads:
            # d_src[t] = sum_j d_rows[tok_rows[t, j]]: the combine kernel with unit weights
            out.append(None if g is None else _combine_raw(g, plan.tok_rows, None, None, g.dtype, plan.T, plan.K))
        return (None, *out)


def permute(plan: DispatchPlan, *srcs: torch.Tensor):
    """Gather rows of every `src` ([T, ...]) into expert-major order -> [cap, ...] (tail rows zero)."""
    return _Permute.apply(plan, *srcs)


# ----------------------------------------------------------------------------------------------------
# (3) combine
# ----------------------------------------------------------------------------------------------------
class _Combine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, sparse_w, base, plan, out_dtype):
        _cuda(rows, sparse_w, base)
        ctx.plan = plan
        ctx.has_base = base is not None
        ctx.need_w = ctx.needs_input_grad[1]
        ctx.save_for_backward(rows if ctx.need_w else None)
        ctx.rows_meta = (rows.dtype, tuple(rows.shape))
        return _combine_raw(rows, plan.tok_rows, plan.row_w, base, out_dtype, plan.T, plan.K)

    @staticmethod
    def backward(ctx, dY):
        plan = ctx.plan
        (rows,) = ctx.saved_tensors
        lib = L.lib()
        dY = dY.contiguous()
        rdt, rshape = ctx.rows_meta
        d_rows = torch.empty(rshape, dtype=rdt, device=dY.device)
        d_sparse = torch.empty(plan.T, plan.E, dtype=torch.float32, device=dY.device) if ctx.need_w else None
        D = dY[0].numel()
        L.check(lib.hdmoe_combine_rows_bwd(_p(rows), _DT[rdt], _p(dY), _dt(dY), _p(plan.row_src),
                                           _p(plan.row_expert), _p(plan.row_w), _p(plan.n_rows_dev), plan.cap,
                                           plan.T, plan.E, D, _p(d_rows), _p(d_sparse), _st()), "combine_rows_bwd")
        return d_rows, d_sparse, (dY if ctx.has_base else None), None, None


def combine(rows: torch.Tensor, sparse_w: torch.Tensor, plan: DispatchPlan, base: Optional[torch.Tensor] = None,
            out_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """out[t] = (base[t]) + sum over dispatched experts (ascending) of sparse_w[t, e] * rows[row(t, e)]."""
    return _Combine.apply(rows, sparse_w, base, plan, out_dtype or rows.dtype)


# ----------------------------------------------------------------------------------------------------
# (4) EDM preconditioning / Heun step
# ----------------------------------------------------------------------------------------------------
def _sigma_vec(sigma: torch.Tensor, B: int) -> torch.Tensor:
    s = sigma.detach().to(torch.float32).reshape(-1).contiguous()
    if s.numel() not in (1, B):
        raise RuntimeError(f"sigma has {s.numel()} elements, expected 1 or batch size {B}")
    return s


class _PrecondIn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, sigma, sigma_data, out_dtype):
        _cuda(x, sigma)
        x = _f32c(x)
        B, per = x.shape[0], x[0].numel()
        s = _sigma_vec(sigma, B)
        out = torch.empty(x.shape, dtype=out_dtype, device=x.device)
        L.check(L.lib().hdmoe_edm_precond_in(_p(x), _p(s), s.numel(), float(sigma_data), _p(out), _dt(out), B, per,
                                             _st()), "edm_precond_in")
        ctx.save_for_backward(s)
        ctx.sd = float(sigma_data)
        return out

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        g = g.contiguous()
        B, per = g.shape[0], g[0].numel()
        dx = torch.empty(g.shape, dtype=torch.float32, device=g.device)
        L.check(L.lib().hdmoe_edm_precond_in_bwd(_p(g), _dt(g), _p(s), s.numel(), ctx.sd, _p(dx), B, per, _st()),
                "edm_precond_in_bwd")
        return dx, None, None, None


class _PrecondOut(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_in, F, sigma, sigma_data):
        _cuda(x_in, F, sigma)
        x_in, F = x_in.contiguous(), F.contiguous()
        B, per = x_in.shape[0], x_in[0].numel()
        s = _sigma_vec(sigma, B)
        D = torch.empty(x_in.shape, dtype=torch.float32, device=x_in.device)
        L.check(L.lib().hdmoe_edm_precond_out(_p(x_in), _dt(x_in), _p(F), _dt(F), _p(s), s.numel(), float(sigma_data),
                                              _p(D), B, per, _st()), "edm_precond_out")
        ctx.save_for_backward(s)
        ctx.meta = (float(sigma_data), x_in.dtype, F.dtype)
        return D

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        sd, xdt, fdt = ctx.meta
        g = _f32c(g)
        B, per = g.shape[0], g[0].numel()
        dF = torch.empty(g.shape, dtype=fdt, device=g.device)
        dX = torch.empty(g.shape, dtype=xdt, device=g.device)
        L.check(L.lib().hdmoe_edm_precond_out_bwd(_p(g), _p(s), s.numel(), sd, _p(dF), _DT[fdt], _p(dX), _DT[xdt], B,
                                                  per, _st()), "edm_precond_out_bwd")
        return dX, dF, None, None


def edm_precond_in(x, sigma, sigma_data: float, out_dtype=torch.float32):
    """x * c_in(sigma)  (models/model_config2.py:434,440)."""
    return _PrecondIn.apply(x, sigma, sigma_data, out_dtype)


def edm_precond_out(x_in, F, sigma, sigma_data: float):
    """c_skip * x_in + c_out * F  (models/model_config2.py:449, with quirk Q1)."""
    return _PrecondOut.apply(x_in, F, sigma, sigma_data)


def edm_heun_pre(x_cur, eps, noise_scale: float, t_hat: float, sigma_data: float, x_in_dtype=torch.float32):
    _cuda(x_cur, eps)
    x_cur = _f32c(x_cur)
    eps = _f32c(eps) if (eps is not None and noise_scale != 0.0) else None
    x_hat = torch.empty_like(x_cur)
    x_in = torch.empty(x_cur.shape, dtype=x_in_dtype, device=x_cur.device)
    L.check(L.lib().hdmoe_edm_heun_pre(_p(x_cur), _p(eps), float(noise_scale), float(t_hat), float(sigma_data),
                                       _p(x_hat), _p(x_in), _dt(x_in), x_cur.numel(), _st()), "edm_heun_pre")
    return x_hat, x_in


def edm_heun_euler(x_hat, x_in, F, F_guide, guidance: float, t_hat: float, t_next: float, sigma_data: float):
    _cuda(x_hat, x_in, F, F_guide)
    F = F.contiguous()
    F_guide = None if F_guide is None else F_guide.to(F.dtype).contiguous()
    d_cur, x_next = torch.empty_like(x_hat), torch.empty_like(x_hat)
    x_in_next = torch.empty_like(x_in) if (t_next > 0 and x_in is not None) else None
    L.check(L.lib().hdmoe_edm_heun_euler(_p(x_hat), _p(x_in), _dt(x_in) if x_in is not None else L.F32, _p(F), _p(F_guide), _dt(F), float(guidance),
                                         float(t_hat), float(t_next), float(sigma_data), _p(d_cur), _p(x_next),
                                         _p(x_in_next), x_hat.numel(), _st()), "edm_heun_euler")
    return d_cur, x_next, x_in_next


def edm_heun_correct(x_hat, x_next, x_in_next, F, F_guide, guidance: float, t_hat: float, t_next: float,
                     sigma_data: float, d_cur):
    _cuda(x_hat, x_next, x_in_next, F, F_guide, d_cur)
    F = F.contiguous()
    F_guide = None if F_guide is None else F_guide.to(F.dtype).contiguous()
    out = torch.empty_like(x_hat)
    L.check(L.lib().hdmoe_edm_heun_correct(_p(x_hat), _p(x_next), _p(x_in_next),
                                           _dt(x_in_next) if x_in_next is not None else L.F32, _p(F), _p(F_guide),
                                           _dt(F), float(guidance), float(t_hat), float(t_next), float(sigma_data),
                                           _p(d_cur), _p(out), x_hat.numel(), _st()), "edm_heun_correct")
    return out


# ----------------------------------------------------------------------------------------------------
# (5) W-PREP
# ----------------------------------------------------------------------------------------------------
_LAYOUTS = {"same": L.WLAYOUT_SAME, "taps": L.WLAYOUT_TAPS, "taps_t": L.WLAYOUT_TAPS_T}


class WeightPrep:
    """Multi-tensor weight preparation plan: ONE launch normalises / scales / casts many MP_Conv weights
    (include/hdmoe_b200.h §5).  entries: list of dicts with keys
        w: fp32 Parameter [rows, ...];  out: Tensor (layout `layout`);  out2 / layout2: optional second output;
        gain: float | 0-dim CUDA tensor;  layout: 'same' | 'taps' | 'taps_t';  cin_pad / cin_rows / cout_pad;
        active: optional int32 CUDA tensor (1 element): skip the in-place rewrite when it is 0."""

    def __init__(self, entries, device):
        self.n = len(entries)
        self.entries = entries
        self.descs = (L.WprepDesc * self.n)()
        self.dev_buf = torch.empty(self.n * C.sizeof(L.WprepDesc), dtype=torch.uint8, device=device)

    def _fill(self):
        for d, e in zip(self.descs, self.entries):
            w, out = e["w"], e["out"]
            assert w.dtype == torch.float32 and w.is_contiguous() and w.is_cuda
            rows = w.shape[0]
            fan_in = w[0].numel()
            taps = 1
            for s_ in w.shape[2:]:
                taps *= s_
            cin = fan_in // taps
            g = e.get("gain", 1.0)
            d.w, d.w_hat = w.data_ptr(), out.data_ptr()
            out2 = e.get("out2")
            d.w_hat2 = out2.data_ptr() if out2 is not None else None
            if torch.is_tensor(g):
                d.gain_ptr, d.gain = g.data_ptr(), 0.0
            else:
                d.gain_ptr, d.gain = None, float(g)
            act = e.get("active")
            d.active = act.data_ptr() if act is not None else None
            d.rows, d.fan_in, d.cin, d.taps = rows, fan_in, cin, taps
            d.cin_pad = int(e.get("cin_pad", cin))
            d.cin_rows = int(e.get("cin_rows", cin))
            d.cout_pad = int(e.get("cout_pad", rows))
            d.out_dtype = _dt(out)
            d.layout = _LAYOUTS[e.get("layout", "same")]
            d.layout2 = _LAYOUTS[e.get("layout2", "same")]

    def run(self, force: bool):
        self._fill()      # pointers may have moved (optimizer swaps, .to())
        L.check(L.lib().hdmoe_wprep_fwd(self.descs, _p(self.dev_buf), self.n, int(bool(force)), _st()), "wprep_fwd")

    def signature(self):
        """Pointer signature of the plan: equal signatures mean the uploaded table is still valid."""
        sig = []
        for e in self.entries:
            for k in ("w", "out", "out2", "active"):
                t = e.get(k)
                sig.append(t.data_ptr() if torch.is_tensor(t) else 0)
            g = e.get("gain", 1.0)
            sig.append(g.data_ptr() if torch.is_tensor(g) else float(g))
        return tuple(sig)

    def run_uploaded(self, force: bool):
        """Launch with the table already resident on the device (no host->device copy: CUDA-graph capturable)."""
        L.check(L.lib().hdmoe_wprep_fwd_resident(_p(self.dev_buf), self.n, self.total_rows, int(bool(force)), _st()),
                "wprep_fwd_resident")

    def upload(self):
        self._fill()
        tot = 0
        for d in self.descs:
            d.block_start = tot
            tot += d.rows
        self.total_rows = tot
        host = torch.frombuffer(bytearray(bytes(self.descs)), dtype=torch.uint8)
        self.dev_buf.copy_(host)


class WeightPrepBackward:
    """ONE launch turning accumulated d_w_hat buffers into master-weight gradients.  entries: dicts with
    w, d_w_hat (fp32, layout 'same' or 'taps'), d_w (fp32 out), gain (float | tensor), d_gain (optional 0-dim
    fp32 tensor, accumulated), cin_pad."""

    def __init__(self, entries, device):
        self.n = len(entries)
        self.entries = entries
        self.descs = (L.WprepBwdDesc * self.n)()
        self.dev_buf = torch.empty(self.n * C.sizeof(L.WprepBwdDesc), dtype=torch.uint8, device=device)

    def upload(self):
        self._fill()
        tot = 0
        for d in self.descs:
            d.block_start = tot
            tot += d.rows
        self.total_rows = tot
        host = torch.frombuffer(bytearray(bytes(self.descs)), dtype=torch.uint8)
        self.dev_buf.copy_(host)

    def run_uploaded(self):
        L.check(L.lib().hdmoe_wprep_bwd_multi_resident(_p(self.dev_buf), self.n, self.total_rows, _st()),
                "wprep_bwd_multi_resident")

    def signature(self):
        sig = []
        for e in self.entries:
            for k in ("w", "d_w_hat", "d_w", "d_gain"):
                t = e.get(k)
                sig.append(t.data_ptr() if torch.is_tensor(t) else 0)
            g = e.get("gain", 1.0)
            sig.append(g.data_ptr() if torch.is_tensor(g) else float(g))
        return tuple(sig)

    def run(self):
        self._fill()
        L.check(L.lib().hdmoe_wprep_bwd_multi(self.descs, _p(self.dev_buf), self.n, _st()), "wprep_bwd_multi")

    def _fill(self):
        for d, e in zip(self.descs, self.entries):
            w = e["w"]
            rows, fan_in = w.shape[0], w[0].numel()
            taps = 1
            for s_ in w.shape[2:]:
                taps *= s_
            g = e.get("gain", 1.0)
            d.w, d.d_w_hat, d.d_w = w.data_ptr(), e["d_w_hat"].data_ptr(), e["d_w"].data_ptr()
            if torch.is_tensor(g):
                d.gain_ptr, d.gain = g.data_ptr(), 0.0
            else:
                d.gain_ptr, d.gain = None, float(g)
            dg = e.get("d_gain")
            d.d_gain = dg.data_ptr() if dg is not None else None
            d.rows, d.fan_in, d.cin, d.taps = rows, fan_in, fan_in // taps, taps
            d.cin_pad = int(e.get("cin_pad", fan_in // taps))
            d.layout = _LAYOUTS[e.get("layout", "same")]


def wprep_bwd(w, d_w_hat, gain):
    """d_w (and d_gain when gain is a tensor) from d_w_hat through one normalisation."""
    _cuda(w, d_w_hat)
    w2 = w.detach().reshape(w.shape[0], -1).contiguous()
    g2 = _f32c(d_w_hat.reshape(w.shape[0], -1))
    d_w = torch.empty_like(w2)
    gt = gain if torch.is_tensor(gain) else None
    d_gain = torch.zeros((), dtype=torch.float32, device=w.device) if gt is not None else None
    L.check(L.lib().hdmoe_wprep_bwd(_p(w2), _p(g2), _p(gt), 0.0 if gt is not None else float(gain), w2.shape[0],
                                    w2.shape[1], _p(d_w), _p(d_gain), _st()), "wprep_bwd")
    return d_w.reshape(w.shape), d_gain


# ----------------------------------------------------------------------------------------------------
# (6) grouped implicit-GEMM convolution (tcgen05)
# ----------------------------------------------------------------------------------------------------
def gconv_raw(x, w_cat, cout: int, w_rows_total: int, row_expert, n_rows_dev, ksizes, wrows, scale=None,
              act: int = 0, residual=None, res_a: float = 0.0, res_b: float = 1.0):
    """y[cap,H,W,cout] = grouped 'same' conv of NHWC bf16 x with the tap-major prepared weights of each row's
    expert, fused epilogue out = res_a*residual + res_b*act(scale*conv)."""
    _cuda(x, w_cat)
    assert x.dtype == torch.bfloat16 and w_cat.dtype == torch.bfloat16 and x.is_contiguous() and w_cat.is_contiguous()
    cap, H, W, cin_pad = x.shape
    E = len(ksizes)
    y = torch.empty(cap, H, W, cout, dtype=torch.bfloat16, device=x.device)
    ks = (C.c_int32 * E)(*ksizes)
    wr = (C.c_int32 * E)(*wrows)
    if scale is not None:
        scale = _f32c(scale)
    if residual is not None:
        assert residual.dtype == torch.bfloat16 and residual.is_contiguous() and residual.shape == y.shape
    fn = L.lib().hdmoe_gconv2_fwd
    L.check(fn(_p(x), _p(w_cat), _p(y), cap, H, W, cin_pad, cout, w_rows_total, _p(row_expert), _p(n_rows_dev), E, ks, wr,
               _p(scale), int(act), _p(residual), float(res_a), float(res_b), _st()), "gconv_fwd")
    return y


# The grouped convolution has ONE product implementation (gconv2.cu: halo reuse).  The tap-group variant gconv3 passed
# the parity tests on B200 in round 2 and lost on every layer shape (profiles/r2_gconv3_vs_gconv2.md); it is archived in
# tools/legacy/.  set_gconv_impl / get_gconv_impl stay for callers that pinned the default.
def set_gconv_impl(v: int, experimental: bool = False) -> None:
    if v != 2:
        raise ValueError("gconv implementation 2 (halo reuse) is the only one in the library; see tools/legacy/")


def get_gconv_impl() -> int:
    return 2


def gconv_wgrad_raw(x, dy, dw, row_expert, n_rows_dev, ksizes, wrows):
    """dw[fp32, tap-major blocks, rows_total x cin_pad] += grouped convolution weight gradient (tcgen05)."""
    _cuda(x, dy, dw)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and dw.dtype == torch.float32
    assert x.is_contiguous() and dy.is_contiguous() and dw.is_contiguous()
    cap, H, W, cin_pad = x.shape
    cout = dy.shape[-1]
    E = len(ksizes)
    ks = (C.c_int32 * E)(*ksizes)
    wr = (C.c_int32 * E)(*wrows)
    L.check(L.lib().hdmoe_gconv_wgrad(_p(x), _p(dy), _p(dw), cap, H, W, cin_pad, cout, dw.shape[0], _p(row_expert),
                                      _p(n_rows_dev), E, ks, wr, _st()), "gconv_wgrad")


# ----------------------------------------------------------------------------------------------------
# (7) trunk attention, head_dim = 4
# ----------------------------------------------------------------------------------------------------
class _GN1Relu(torch.autograd.Function):
    """relu(group_norm(x, 1 group)) [+ mean over H, W] on channels-last fp32 x (csrc/gn_relu.cu)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, pool):
        _cuda(x, gamma, beta)
        B, Cn, H, W = x.shape
        if x.dtype != torch.float32 or not x.is_contiguous(memory_format=torch.channels_last):
            raise RuntimeError("gn1_relu: x must be channels-last float32")
        gamma, beta = _f32c(gamma), _f32c(beta)
        stats = torch.empty(B, 2, dtype=torch.float32, device=x.device)
        y = None if pool else torch.empty_like(x)
        pooled = torch.empty(B, Cn, dtype=torch.float32, device=x.device) if pool else None
        L.check(L.lib().hdmoe_gn1_relu_fwd(_p(x), _p(gamma), _p(beta), _p(y), _p(pooled), _p(stats), B, H * W, Cn,
                                           float(eps), _st()), "gn1_relu_fwd")
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.pool = pool
        return pooled if pool else y

    @staticmethod
    def backward(ctx, g):
        x, gamma, beta, stats = ctx.saved_tensors
        B, Cn, H, W = x.shape
        g = g.float()
        g = g.contiguous() if ctx.pool else g.contiguous(memory_format=torch.channels_last)
        dx = torch.empty_like(x)
        parts = torch.empty(2, B, Cn, dtype=torch.float32, device=x.device)
        L.check(L.lib().hdmoe_gn1_relu_bwd(_p(x), _p(gamma), _p(beta), _p(stats), None if ctx.pool else _p(g),
                                           _p(g) if ctx.pool else None, _p(dx), _p(parts[0]), _p(parts[1]), B, H * W, Cn,
                                           _st()), "gn1_relu_bwd")
        sums = parts.sum(dim=1)
        return dx, sums[0], sums[1], None, None


def gn1_relu(x, gamma, beta, eps: float = 1e-5, pool: bool = False):
    """ReLU(GroupNorm(1, C)(x)) for channels-last fp32 x [B, C, H, W]; pool=True returns the [B, C] spatial mean
    instead (the tail of Router.hard_route, models/model_components.py:92-103)."""
    return _GN1Relu.apply(x, gamma, beta, eps, pool)


class _GN1ReluNHWC(torch.autograd.Function):
    """relu(group_norm(x, 1 group)) [+ mean over H, W] on NHWC bf16 rows x [R, H, W, C]: the activations between the
    tcgen05 router-trunk convolutions (csrc/gn_relu.cu, bf16 instantiation; statistics, pooled output and the
    gamma / beta gradients in fp32).  gamma / beta are [G, C]: rows [g * R / G, (g + 1) * R / G) use table row g (the
    trunks of several routers run as one grouped launch)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, pool):
        _cuda(x, gamma, beta)
        if x.dtype != torch.bfloat16 or not x.is_contiguous() or x.ndim != 4:
            raise RuntimeError("gn1_relu_nhwc: x must be contiguous bfloat16 [R, H, W, C]")
        R, H, W, Cn = x.shape
        gamma, beta = _f32c(gamma).reshape(-1, Cn), _f32c(beta).reshape(-1, Cn)
        G = gamma.shape[0]
        if R % G != 0 or beta.shape[0] != G:
            raise RuntimeError("gn1_relu_nhwc: rows must split evenly over the gamma / beta groups")
        stats = torch.empty(R, 2, dtype=torch.float32, device=x.device)
        y = None if pool else torch.empty_like(x)
        pooled = torch.empty(R, Cn, dtype=torch.float32, device=x.device) if pool else None
        L.check(L.lib().hdmoe_gn1_relu_fwd_t(_p(x), L.BF16, _p(gamma), _p(beta), _p(y), _p(pooled), _p(stats), R, H * W, Cn,
                                             float(eps), R // G, _st()), "gn1_relu_fwd_t")
        ctx.save_for_backward(x, gamma, beta, stats)
        ctx.pool, ctx.G = pool, G
        return pooled if pool else y

    @staticmethod
    def backward(ctx, g):
        x, gamma, beta, stats = ctx.saved_tensors
        R, H, W, Cn = x.shape
        G = ctx.G
        g = g.float().contiguous() if ctx.pool else g.to(torch.bfloat16).contiguous()
        dx = torch.empty_like(x)
        parts = torch.empty(2, R, Cn, dtype=torch.float32, device=x.device)
        L.check(L.lib().hdmoe_gn1_relu_bwd_t(_p(x), L.BF16, _p(gamma), _p(beta), _p(stats), None if ctx.pool else _p(g),
                                             _p(g) if ctx.pool else None, _p(dx), _p(parts[0]), _p(parts[1]), R, H * W, Cn,
                                             R // G, _st()), "gn1_relu_bwd_t")
        sums = parts.view(2, G, R // G, Cn).sum(dim=2)
        return dx, sums[0], sums[1], None, None


def gn1_relu_nhwc(x, gamma, beta, eps: float = 1e-5, pool: bool = False):
    """ReLU(GroupNorm(1, C)(x)) for NHWC bf16 rows [R, H, W, C] with [G, C] (or [C]) affine tables; pool=True returns
    the fp32 [R, C] spatial mean."""
    return _GN1ReluNHWC.apply(x, gamma, beta, eps, pool)


def gn1_relu_supported(channels: int) -> bool:
    return channels % 4 == 0 and 1024 % (channels // 4) == 0


def _attn_split_p() -> int:
    """Strict fp32 (p and dS split into hi + lo) unless TF32 matmuls are enabled, like the library matmuls."""
    return 0 if torch.backends.cuda.matmul.allow_tf32 else 1


class _AttnD4(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, heads, scale):
        _cuda(q, k, v)
        q, k, v = _f32c(q), _f32c(k), _f32c(v)
        B, Sq, Cc = q.shape
        Sk = k.shape[1]
        assert Cc == heads * 4 and k.shape == (B, Sk, Cc) and v.shape == (B, Sk, Cc)
        o = torch.empty_like(q)
        lse = torch.empty(B, heads, Sq, dtype=torch.float32, device=q.device)
        split = _attn_split_p()
        L.check(L.lib().hdmoe_attn_d4_tc_fwd(_p(q), _p(k), _p(v), _p(o), _p(lse), B, Sq, Sk, heads, float(scale), split,
                                             _st()), "attn_d4_tc_fwd")
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.meta = (heads, float(scale), split)
        return o

    @staticmethod
    def backward(ctx, dO):
        q, k, v, o, lse = ctx.saved_tensors
        heads, scale, split = ctx.meta
        dO = _f32c(dO)
        B, Sq, _ = q.shape
        Sk = k.shape[1]
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        Dbuf = torch.empty_like(lse)
        L.check(L.lib().hdmoe_attn_d4_tc_bwd(_p(q), _p(k), _p(v), _p(o), _p(dO), _p(lse), _p(dq), _p(dk), _p(dv),
                                             _p(Dbuf), B, Sq, Sk, heads, scale, split, _st()), "attn_d4_tc_bwd")
        return dq, dk, dv, None, None


def attention_d4(q, k, v, heads: int, scale: float):
    """softmax(q k^T * scale) v per head for head_dim 4; q [B,Sq,heads*4], k/v [B,Sk,heads*4] fp32."""
    return _AttnD4.apply(q, k, v, heads, scale)


# ------------------------------------------------------------------------------------------------ thin projections
class _Linear32(torch.autograd.Function):
    """y = x W^T for the 32 -> 32 trunk projections over B*S rows (MP_Attention._proj,
    models/model_internals.py:364-372,407).  Forward and input gradient stay library GEMMs; the weight gradient
    dW = dY^T X (a [32, rows] x [rows, 32] product the library runs without split-K, ~120 us) is the streaming
    kernel csrc/lin_wgrad.cu (~15 us)."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return torch.nn.functional.linear(x, w)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = gy.matmul(w)
        if ctx.needs_input_grad[1]:
            g2 = gy.reshape(-1, 32).contiguous()
            x2 = x.reshape(-1, 32).contiguous()
            gw = torch.zeros(32, 32, dtype=torch.float32, device=x.device)
            L.check(L.lib().hdmoe_lin32_wgrad(_p(g2), _p(x2), _p(gw), g2.shape[0], _st()), "lin32_wgrad")
        return gx, gw


def linear32(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """F.linear(x, w) with the streaming weight-gradient kernel when it applies (fp32, 32 -> 32, many rows)."""
    if (x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and tuple(w.shape) == (32, 32)
            and x.shape[-1] == 32 and x.numel() >= 32 * 4096 and torch.is_grad_enabled() and w.requires_grad):
        return _Linear32.apply(x, w)
    return torch.nn.functional.linear(x, w)


# ----------------------------------------------------------------------------------------------------
# (9) HDMOEM glue on channels-last activations (csrc/trunk_glue.cu)
# ----------------------------------------------------------------------------------------------------
class _TrunkSwap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, v, w):
        _cuda(u, v, w)
        u, v, w = _f32c(u), _f32c(v), _f32c(w).reshape(-1)
        B = u.shape[0]
        per = u[0].numel()
        assert v.shape == u.shape and w.numel() == B
        q, c = torch.empty_like(u), torch.empty_like(u)
        L.check(L.lib().hdmoe_trunk_swap_fwd(_p(u), _p(v), _p(w), _p(q), _p(c), B, per, _st()), "trunk_swap_fwd")
        ctx.save_for_backward(u, v, w)
        return q, c

    @staticmethod
    def backward(ctx, dq, dc):
        u, v, w = ctx.saved_tensors
        B, per = u.shape[0], u[0].numel()
        dq = torch.zeros_like(u) if dq is None else _f32c(dq)
        dc = torch.zeros_like(u) if dc is None else _f32c(dc)
        du, dv = torch.empty_like(u), torch.empty_like(u)
        sl = L.lib().hdmoe_trunk_swap_slices()
        part = torch.empty(B, sl, dtype=torch.float32, device=u.device)
        L.check(L.lib().hdmoe_trunk_swap_bwd(_p(u), _p(v), _p(w), _p(dq), _p(dc), _p(du), _p(dv), _p(part), B, per, _st()),
                "trunk_swap_bwd")
        return du, dv, part.sum(dim=1)


def trunk_swap(u, v, w):
    """cfg1 soft query / context swap (models/model_config1.py:277-283): (w*v + (1-w)*u, w*u + (1-w)*v) with one
    weight per sample; u, v [B, ...] fp32 of any (equal) trailing shape."""
    return _TrunkSwap.apply(u, v, w)


class _TrunkGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, a, b, alpha, W1, W2, H, W, consts):
        _cuda(u, a, b, alpha, W1, W2)
        u, a, b = _f32c(u), _f32c(a), _f32c(b)
        ctx.in_shapes = (tuple(alpha.shape), tuple(W1.shape), tuple(W2.shape))
        W1, W2, alpha = _f32c(W1).reshape(W1.shape[0], -1), _f32c(W2).reshape(W2.shape[0], -1), _f32c(alpha).reshape(1)
        Cn = u.shape[-1]
        P = u.numel() // Cn
        HW = H * W
        if W1.shape != (Cn, 2 * Cn) or W2.shape != (2, Cn) or a.shape != u.shape or b.shape != u.shape:
            raise RuntimeError("trunk_gate: shapes must be u, a, b [P, C], W1 [C, 2C], W2 [2, C]")
        mix = torch.empty_like(u)
        g = torch.empty(P // HW, 2, H, W, dtype=torch.float32, device=u.device)
        L.check(L.lib().hdmoe_trunk_gate_fwd(_p(u), _p(a), _p(b), _p(alpha), _p(W1), _p(W2), _p(mix), _p(g), P, HW, Cn,
                                             *consts, _st()), "trunk_gate_fwd")
        ctx.save_for_backward(u, a, b, alpha, W1, W2)
        ctx.meta = (P, HW, Cn, consts)
        return mix, g

    @staticmethod
    def backward(ctx, d_mix, d_g):
        u, a, b, alpha, W1, W2 = ctx.saved_tensors
        P, HW, Cn, consts = ctx.meta
        d_mix = torch.zeros_like(u) if d_mix is None else _f32c(d_mix)
        d_g = None if d_g is None else _f32c(d_g)
        du, da, db = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)
        acc = torch.zeros(W1.numel() + W2.numel() + 1, dtype=torch.float32, device=u.device)
        dW1, dW2, d_alpha = acc[:W1.numel()], acc[W1.numel():W1.numel() + W2.numel()], acc[-1:]
        L.check(L.lib().hdmoe_trunk_gate_bwd(_p(u), _p(a), _p(b), _p(alpha), _p(W1), _p(W2), _p(d_mix), _p(d_g), _p(du), _p(da),
                                             _p(db), _p(dW1), _p(dW2), _p(d_alpha), P, HW, Cn, *consts, _st()),
                "trunk_gate_bwd")
        sa, s1, s2 = ctx.in_shapes
        return du, da, db, d_alpha.reshape(sa), dW1.reshape(s1), dW2.reshape(s2), None, None, None


def trunk_gate(u, a, b, alpha_txt, w1_hat, w2_hat, H: int, W: int, cat_t: float = 0.5, sum_t: float = 0.5):
    """Text blend + gated mix of HDMOEM.forward (models/model_config2.py:291-301) on channels-last activations.
    u (U-Net MoE output), a (cross_attn output), b (cross_attn_text output): [B, H*W, C] fp32; alpha_txt 0-dim;
    w1_hat / w2_hat: the PREPARED gate1 [C, 2C(,1,1)] / gate2 [2, C(,1,1)] weights.  Returns (mix [B, H*W, C] -- the input
    of output_proj --, out_gate [B, 2, H, W])."""
    import math
    Cn = u.shape[-1]
    c = math.sqrt((2 * Cn) / ((1 - cat_t) ** 2 + cat_t ** 2))
    c1, c2 = c * (1 - cat_t) / math.sqrt(Cn), c * cat_t / math.sqrt(Cn)
    n = math.sqrt((1 - sum_t) ** 2 + sum_t ** 2)
    consts = (float(c1), float(c2), float((1 - sum_t) / n), float(sum_t / n))
    return _TrunkGate.apply(u, a, b, alpha_txt, w1_hat, w2_hat, H, W, consts)


# ----------------------------------------------------------------------------------------------------
# (10) EDM_LOSS data term
# ----------------------------------------------------------------------------------------------------
class _SqErrRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d, x):
        _cuda(d, x)
        d, x = _f32c(d), _f32c(x)
        B, per = d.shape[0], d[0].numel()
        se = torch.empty(B, dtype=torch.float32, device=d.device)
        L.check(L.lib().hdmoe_sqerr_rows(_p(d), _p(x), _p(se), B, per, _st()), "sqerr_rows")
        ctx.save_for_backward(d, x)
        return se

    @staticmethod
    def backward(ctx, g):
        d, x = ctx.saved_tensors
        B, per = d.shape[0], d[0].numel()
        dd = torch.empty_like(d)
        L.check(L.lib().hdmoe_sqerr_rows_bwd(_p(d), _p(x), _p(_f32c(g)), _p(dd), B, per, _st()), "sqerr_rows_bwd")
        return dd, None


def sqerr_rows(d, x0):
    """se[b] = sum over the sample of (d - x0)^2 (fp32, deterministic); gradient flows to d only (x0 is data)."""
    return _SqErrRows.apply(d, x0)


def train_inputs(x0, eps, sigma, gen_a=None, gen_b=None):
    """x = x0 + eps * sigma[b] and the band masks of up to two mask generators in ONE launch (csrc/edm_step.cu).
    gen_* = (centers list, p_mean, p_std, bandwidth at this step, min_active) or None.  Returns (x, mask_a, mask_b)."""
    _cuda(x0, eps, sigma)
    x0, eps = _f32c(x0), _f32c(eps)
    sig = _f32c(sigma).reshape(-1)
    B, per = x0.shape[0], x0[0].numel()
    assert sig.numel() == B and eps.shape == x0.shape
    x = torch.empty_like(x0)
    descs, masks = [], []
    for g in (gen_a, gen_b):
        if g is None:
            descs.append(None)
            masks.append(None)
            continue
        centers, p_mean, p_std, bw, min_active = g
        d = L.MaskGenDesc()
        if len(centers) > L.MAX_MASK_EXPERTS:
            raise ValueError(f"train_inputs: at most {L.MAX_MASK_EXPERTS} experts per mask generator")
        for i, c in enumerate(centers):
            d.centers[i] = float(c)
        d.p_mean, d.p_std, d.bandwidth, d.n_experts, d.min_active = float(p_mean), float(p_std), float(bw), len(centers), int(min_active)
        descs.append(d)
        masks.append(torch.empty(B, len(centers), dtype=torch.float32, device=x0.device))
    ref = [None if d is None else C.byref(d) for d in descs]
    L.check(L.lib().hdmoe_train_inputs(_p(x0), _p(eps), _p(sig), _p(x), B, per, ref[0], _p(masks[0]), ref[1], _p(masks[1]),
                                       _st()), "train_inputs")
    return x, masks[0], masks[1]


# ----------------------------------------------------------------------------------------------------
# (11) branch scaling (csrc/trunk_glue.cu)
# ----------------------------------------------------------------------------------------------------
def analytic_scaling(time_vec, transition_point: float, softness: float):
    """[B, 2] = ((w + .01) * 2, (1 - w + .01) * 2), w = sigmoid((4 t - tp) / soft); models/model_config2.py:244-249.
    time_vec is data (log sigma / 4): no gradient."""
    _cuda(time_vec)
    t = _f32c(time_vec.detach()).reshape(-1)
    out = torch.empty(t.numel(), 2, dtype=torch.float32, device=t.device)
    L.check(L.lib().hdmoe_analytic_scaling(_p(t), float(transition_point), float(softness), _p(out), t.numel(), _st()),
            "analytic_scaling")
    return out


class _ScalePair(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, scaling, want_trunk):
        _cuda(feats, scaling)
        feats, scaling = _f32c(feats), _f32c(scaling)
        B, Cn, H, W = feats.shape
        assert scaling.shape == (B, 2)
        in_v, in_u = torch.empty_like(feats), torch.empty_like(feats)
        trunk = torch.empty(2 * B, H, W, Cn, dtype=torch.bfloat16, device=feats.device) if want_trunk else None
        L.check(L.lib().hdmoe_scale_pair_fwd(_p(feats), _p(scaling), _p(in_v), _p(in_u), _p(trunk), B, Cn, H * W, _st()),
                "scale_pair_fwd")
        ctx.save_for_backward(feats, scaling)
        ctx.want_trunk = want_trunk
        if want_trunk:
            return in_v, in_u, trunk
        ctx.mark_non_differentiable()
        return in_v, in_u, None

    @staticmethod
    def backward(ctx, g_v, g_u, g_t):
        feats, scaling = ctx.saved_tensors
        B, Cn, H, W = feats.shape
        g_v = None if g_v is None else _f32c(g_v)
        g_u = None if g_u is None else _f32c(g_u)
        g_t = None if (g_t is None or not ctx.want_trunk) else g_t.to(torch.bfloat16).contiguous()
        d_feats = torch.empty_like(feats)
        tiles = L.lib().hdmoe_scale_pair_tiles(H * W)
        part = torch.empty(B, tiles, 2, dtype=torch.float32, device=feats.device)
        L.check(L.lib().hdmoe_scale_pair_bwd(_p(feats), _p(scaling), _p(g_v), _p(g_u), _p(g_t), _p(d_feats), _p(part), B, Cn,
                                             H * W, _st()), "scale_pair_bwd")
        return d_feats, part.sum(dim=1), None


def scale_pair(feats, scaling, want_trunk: bool = False):
    """(in_vit, in_unet, trunk_in): scaling[b, 0] * feats, scaling[b, 1] * feats (fp32 [B, C, H, W]) and optionally the
    channels-last bf16 copy [2B, H, W, C] of both for the tcgen05 router trunk -- one pass over feats."""
    return _ScalePair.apply(feats, scaling, want_trunk)


# ----------------------------------------------------------------------------------------------------
# (12) Scaling_router (cfg1) as one kernel per direction (csrc/trunk_glue.cu)
# ----------------------------------------------------------------------------------------------------
class _ScalingRouter(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W1, g1, b1, W2, g2, b2, W3, noise, zeta, keep, eps):
        _cuda(x, W1, g1, b1, W2, g2, b2, W3, noise, keep)
        shapes = [tuple(t.shape) for t in (W1, g1, b1, W2, g2, b2, W3)]
        x, W1, g1, b1, W2, g2, b2, W3 = (_f32c(t) for t in (x, W1, g1, b1, W2, g2, b2, W3))
        noise, keep = _f32c(noise), _f32c(keep)
        B, D = x.shape
        if W1.numel() != 2 * D * D or W2.numel() != 8 * D * D or W3.numel() != 8 * D:
            raise RuntimeError("scaling_router: weights must be [2D, D], [4D, 2D], [2, 4D]")
        out = torch.empty(B, 2, dtype=torch.float32, device=x.device)
        L.check(L.lib().hdmoe_scaling_router_fwd(_p(x), _p(W1), _p(g1), _p(b1), _p(W2), _p(g2), _p(b2), _p(W3), _p(noise),
                                                 float(zeta), _p(keep), float(eps), _p(out), B, D, _st()),
                "scaling_router_fwd")
        ctx.save_for_backward(x, W1, g1, b1, W2, g2, b2, W3, noise, keep)
        ctx.meta = (float(zeta), float(eps), shapes)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, W1, g1, b1, W2, g2, b2, W3, noise, keep = ctx.saved_tensors
        zeta, eps, shapes = ctx.meta
        B, D = x.shape
        sizes = [W1.numel(), g1.numel(), b1.numel(), W2.numel(), g2.numel(), b2.numel(), W3.numel()]
        acc = torch.zeros(sum(sizes), dtype=torch.float32, device=x.device)
        parts = list(acc.split(sizes))
        dx = torch.empty_like(x)
        L.check(L.lib().hdmoe_scaling_router_bwd(_p(x), _p(W1), _p(g1), _p(b1), _p(W2), _p(g2), _p(b2), _p(W3), _p(noise), zeta,
                                                 _p(keep), eps, _p(_f32c(d_out)), _p(dx), _p(parts[0]), _p(parts[1]),
                                                 _p(parts[2]), _p(parts[3]), _p(parts[4]), _p(parts[5]), _p(parts[6]), B, D,
                                                 _st()), "scaling_router_bwd")
        grads = [p_.view(sh) for p_, sh in zip(parts, shapes)]
        return (dx, *grads, None, None, None, None)


def scaling_router(x, W1, g1, b1, W2, g2, b2, W3, noise=None, zeta: float = 0.0, keep=None, eps: float = 1e-5):
    """Scaling_router.forward on prepared weights (see include/hdmoe_b200.h); x [B, 64] -> [B, 2] = softmax * 2."""
    return _ScalingRouter.apply(x, W1, g1, b1, W2, g2, b2, W3, noise, zeta, keep, eps)
