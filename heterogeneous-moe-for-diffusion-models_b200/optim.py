"""Fused optimizer side of the train step: gradient-norm clipping + AdamW over every parameter
tensor in three launches (csrc/optim.cu), replacing `torch.nn.utils.clip_grad_norm_(params, max_norm)` followed by
`torch.optim.AdamW.step()` of the reference's loop (Utils/training.py:195-197).

`FusedAdamW` is a `torch.optim.Optimizer`: param groups (per-group lr / weight_decay / betas / eps as the reference's
get_optimizer builds them), `zero_grad`, `state_dict` / `load_state_dict` in torch's AdamW format (step, exp_avg,
exp_avg_sq), so checkpoints interchange.  The moments live in two flat buffers; the per-tensor descriptor table is
device resident and refreshed only when a pointer or a learning rate changed, so `step()` replays inside a CUDA graph
(during stream capture the table is staged through a pinned buffer that the capture keeps alive).

Per-tensor step counters as in torch (a parameter without a gradient is skipped and does not advance); the total norm
is available on the device (`last_grad_norm`) without a host synchronisation."""
import ctypes as C
from typing import Optional

import torch

from . import _lib as L


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None, clip_in_place: bool = True):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        b, e = self.param_groups[0]["betas"], self.param_groups[0]["eps"]
        for g in self.param_groups:
            if tuple(g["betas"]) != tuple(b) or g["eps"] != e:
                raise ValueError("FusedAdamW: betas / eps must be the same in every param group (lr and weight_decay may differ)")
        self.max_grad_norm = max_grad_norm
        self.clip_in_place = clip_in_place
        self._params = [p for g in self.param_groups for p in g["params"]]
        for p in self._params:
            if p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                raise RuntimeError("FusedAdamW: parameters must be contiguous float32 CUDA tensors (no CPU fallback)")
        self._built = False
        self._sig = None
        self._capture_tables = []

    # ------------------------------------------------------------------------------------------ state
    def _build(self):
        dev = self._params[0].device
        n = sum(p.numel() for p in self._params)
        self._m, self._v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        self._state = torch.zeros(3, device=dev)                 # calls, grad norm, clip coefficient
        self._steps = torch.zeros(len(self._params), device=dev)  # per-tensor update counts
        o = 0
        for i, p in enumerate(self._params):
            k = p.numel()
            st = self.state[p]
            st["exp_avg"], st["exp_avg_sq"] = self._m[o:o + k].view_as(p), self._v[o:o + k].view_as(p)
            st["step"] = self._steps[i]
            o += k
        chunk = L.lib().hdmoe_optim_chunk_elems()
        self._chunk_start, c = [], 0
        for p in self._params:
            self._chunk_start.append(c)
            c += max(1, (p.numel() + chunk - 1) // chunk)
        self._n_chunks = c
        self._partial = torch.empty(c, device=dev)
        self._descs = (L.OptTensorDesc * len(self._params))()
        self._nbytes = C.sizeof(L.OptTensorDesc) * len(self._params)
        self._table = torch.empty(self._nbytes, dtype=torch.uint8, device=dev)
        self._pinned = torch.empty(self._nbytes, dtype=torch.uint8).pin_memory()
        # staging buffers for tables uploaded DURING stream capture (one per capture; host allocation is not allowed
        # while capturing, so they exist up front)
        self._capture_pool = [torch.empty(self._nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
        self._uploaded = None
        self._built = True

    @property
    def last_grad_norm(self) -> torch.Tensor:
        """total gradient norm of the last step (0-dim device tensor, no synchronisation)"""
        return self._state[1]

    def _signature(self):
        sig = []
        for g in self.param_groups:
            for p in g["params"]:
                gr = p.grad
                if gr is not None and (gr.dtype != torch.float32 or not gr.is_contiguous()):
                    raise RuntimeError("FusedAdamW: gradients must be contiguous float32")
                sig.append((p.data_ptr(), 0 if gr is None else gr.data_ptr(), float(g["lr"]), float(g["weight_decay"])))
        return tuple(sig)

    def _upload(self, sig):
        moments = [(self.state[p]["exp_avg"].data_ptr(), self.state[p]["exp_avg_sq"].data_ptr()) for p in self._params]
        for d, p, (pp, gp, lr, wd), (mp, vp), cs in zip(self._descs, self._params, sig, moments, self._chunk_start):
            d.p, d.g, d.m, d.v = pp, (gp or None), mp, vp
            d.numel, d.lr, d.weight_decay, d.chunk_start, d.pad = p.numel(), lr, wd, cs, 0
        capturing = torch.cuda.is_current_stream_capturing()
        if capturing:
            # the captured copy node re-reads its source at every replay: give it a buffer nobody overwrites
            if not self._capture_pool:
                raise RuntimeError("FusedAdamW: more than 4 CUDA-graph captures of step() on one optimizer")
            pinned = self._capture_pool.pop()
            self._capture_tables.append(pinned)
        else:
            pinned = self._pinned
            if self._uploaded is not None:
                self._uploaded.synchronize()          # the previous asynchronous copy has read the staging buffer
        C.memmove(pinned.data_ptr(), C.addressof(self._descs), self._nbytes)
        self._table.copy_(pinned, non_blocking=True)
        if not capturing:
            self._uploaded = torch.cuda.Event()
            self._uploaded.record()

    # ------------------------------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self._built:
            self._build()
        sig = self._signature()
        if sig != self._sig or torch.cuda.is_current_stream_capturing():
            self._upload(sig)
            self._sig = sig
        g0 = self.param_groups[0]
        L.check(L.lib().hdmoe_adamw_step(C.c_void_p(self._table.data_ptr()), len(self._params), self._n_chunks,
                                         C.c_void_p(self._partial.data_ptr()), C.c_void_p(self._state.data_ptr()),
                                         C.c_void_p(self._steps.data_ptr()), float(self.max_grad_norm or 0.0), float(g0["betas"][0]), float(g0["betas"][1]),
                                         float(g0["eps"]), int(self.clip_in_place),
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "adamw_step")
        return loss

    def load_state_dict(self, state_dict):
        """torch AdamW format in, values copied into the flat moment buffers (the views stay pointer-stable)."""
        if not self._built:
            self._build()
        idx = {i: p for i, p in enumerate(self._params)}
        with torch.no_grad():
            step = 0.0
            for i, st in state_dict.get("state", {}).items():
                p = idx[int(i)]
                self.state[p]["exp_avg"].copy_(st["exp_avg"])
                self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
                self._steps[int(i)] = float(st.get("step", 0.0))
                step = max(step, float(st.get("step", 0.0)))
            self._state[0] = step
        for g_new, g in zip(state_dict.get("param_groups", []), self.param_groups):
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in g_new:
                    g[k] = g_new[k]
        self._sig = None
