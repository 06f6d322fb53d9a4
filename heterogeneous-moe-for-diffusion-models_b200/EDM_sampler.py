"""EDM Heun 2nd-order sampler, same call surface as the reference's Utils/EDM_sampler.py
(EDM_Sampler(model, Guide_net, ...).sample / .denoise).  The per-step elementwise glue and the EDM
preconditioning run as three fused kernels (csrc/edm_step.cu); the loop never synchronises the host."""
import math

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._denoiser import preconditioned_HDMOEM as _Native


class EDM_Sampler:
    def __init__(self, model: nn.Module, Guide_net: nn.Module, num_solve_steps: int = 32, sigma_min: float = 0.002,
                 sigma_max: float = 80, rho: int = 7, S_churn: float = 0.0, S_min: float = 0.0,
                 S_max: float = float("inf"), S_noise: float = 1.0, guidance: float = 1.0, dtype=torch.float32,
                 use_cuda_graph: bool = False):
        self.model, self.gnet = model, Guide_net
        self.num_steps = num_solve_steps
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho
        self.s_churn, self.s_min, self.s_max, self.s_noise = S_churn, S_min, S_max, S_noise
        self.guide = guidance
        self.dtype = dtype
        self.nfe = 0
        # B200 extra (not in the reference): record ONE denoiser evaluation as a CUDA graph and replay it for all
        # 2N-1 function evaluations (shapes are identical across steps; sigma and the input live in static buffers)
        self.use_cuda_graph = use_cuda_graph
        self._graphs = {}

    # -- reference-compatible helper (Utils/EDM_sampler.py:34-70) --------------------------------------
    def _call(self, net, x, sigma, text_emb, transition_mean, softness, **fast):
        bs = x.shape[0]
        ones = torch.ones((bs, self.model.num_experts), device=x.device)
        self.nfe += 1
        return net(x=x, sigma=sigma, text_emb=text_emb, Unet_router_mask=ones, Vit_router_mask=ones, zeta=0,
                   transition_point=transition_mean, softness=softness, **fast)["denoised"]

    def denoise(self, x, sigma, text_emb, transition_mean, softness, uncond_text_emb=None):
        """D(x; sigma) with classifier-free guidance folded in: ref.lerp(cond, guidance)."""
        if self.use_cuda_graph and isinstance(self.model, _Native) and isinstance(self.gnet, _Native) and x.is_cuda:
            # the evaluation sample() replays: preconditioned input -> graphed network(s) -> c_skip / c_out mix
            with torch.no_grad():
                sd = float(getattr(self.model, "sigma_data", 0.5))
                guided = self.guide != 1.0
                g_emb = uncond_text_emb if uncond_text_emb is not None else text_emb
                sig = torch.as_tensor(sigma, dtype=torch.float32, device=x.device).reshape(())
                x_in = ops.edm_precond_in(x.to(torch.float32), sig, sd)
                F, Fg = self._eval_graphed(x_in, sig, text_emb, g_emb, transition_mean, softness, guided)
                D_x = ops.edm_precond_out(x_in, F.to(torch.float32), sig, sd).to(self.dtype)
                if not guided:
                    return D_x
                return ops.edm_precond_out(x_in, Fg.to(torch.float32), sig, sd).to(self.dtype).lerp(D_x, self.guide)
        D_x = self._call(self.model, x, sigma, text_emb, transition_mean, softness).to(self.dtype)
        if self.guide == 1.0:
            return D_x
        emb = uncond_text_emb if uncond_text_emb is not None else text_emb
        ref = self._call(self.gnet, x, sigma, emb, transition_mean, softness).to(self.dtype)
        return ref.lerp(D_x, self.guide)

    def _eval_graphed(self, x_in, sigma_t, text_emb, g_emb, transition_mean, softness, guided):
        """One (or, guided, two) raw network evaluation(s) through a recorded CUDA graph.  Everything the capture
        bakes in is part of the cache key: shapes / dtype, the conditioning tensors (identity AND shape -- the entry
        keeps them alive so a recycled address cannot alias), the python scalars transition_mean / softness, and the
        train / eval state of both networks."""
        key = (tuple(x_in.shape), x_in.dtype, text_emb.data_ptr() if text_emb is not None else 0,
               tuple(text_emb.shape) if text_emb is not None else None,
               (g_emb.data_ptr(), tuple(g_emb.shape)) if (guided and g_emb is not None) else None, guided,
               float(transition_mean), float(softness), bool(self.model.training), bool(self.gnet.training),
               id(self.model), id(self.gnet))
        ent = self._graphs.get(key)
        if ent is None:
            st_x, st_s = x_in.clone(), sigma_t.detach().clone().reshape(())
            kw = dict(precomputed_x_in=st_x, raw_output=True)

            def run():
                F_ = self._call(self.model, st_x, st_s, text_emb, transition_mean, softness, **kw)
                Fg_ = self._call(self.gnet, st_x, st_s, g_emb, transition_mean, softness, **kw) if guided else None
                return F_, Fg_

            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    run()
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                outs = run()
            self.nfe -= 3 * (2 if guided else 1)        # warm-up / capture calls are not function evaluations
            ent = (graph, st_x, st_s, outs, (text_emb, g_emb))   # the conditioning tensors stay alive with the graph
            self._graphs[key] = ent
        graph, st_x, st_s, outs, _keep = ent
        st_x.copy_(x_in)
        st_s.copy_(sigma_t.reshape(()))
        graph.replay()
        self.nfe += 2 if guided else 1
        return outs

    def t_steps(self) -> torch.Tensor:
        """Karras rho-schedule with a trailing zero (Utils/EDM_sampler.py:82-87), fp32 on the host."""
        i = torch.arange(self.num_steps, dtype=self.dtype)
        t = (self.sigma_max ** (1 / self.rho) + i / (self.num_steps - 1)
             * (self.sigma_min ** (1 / self.rho) - self.sigma_max ** (1 / self.rho))) ** self.rho
        return torch.cat([t, torch.zeros_like(t[:1])])

    @torch.no_grad()
    def sample(self, noise: torch.Tensor, text_emb: torch.Tensor, transition_mean: float, softness: float,
               uncond_text_emb: torch.Tensor = None) -> torch.Tensor:
        """ref Utils/EDM_sampler.py:72-109"""
        if self.dtype != torch.float32:
            raise RuntimeError("hdmoe_b200.EDM_Sampler keeps the ODE state in float32 (as the reference default)")
        native = isinstance(self.model, _Native) and isinstance(self.gnet, _Native)
        sd = float(getattr(self.model, "sigma_data", 0.5))
        guided = self.guide != 1.0
        g_emb = uncond_text_emb if uncond_text_emb is not None else text_emb
        t_host = self.t_steps()
        t_dev = t_host.to(noise.device)
        f32 = np.float32
        x_next = noise.to(self.dtype) * t_dev[0]

        def evaluate(x_in_or_x, sigma_t):
            """-> (F, F_guide): raw network outputs (native) or denoised estimates (foreign model)."""
            if native and self.use_cuda_graph:
                return self._eval_graphed(x_in_or_x, sigma_t, text_emb, g_emb, transition_mean, softness, guided)
            if native:
                kw = dict(precomputed_x_in=x_in_or_x, raw_output=True)
                F = self._call(self.model, x_in_or_x, sigma_t, text_emb, transition_mean, softness, **kw)
                Fg = self._call(self.gnet, x_in_or_x, sigma_t, g_emb, transition_mean, softness, **kw) if guided else None
            else:
                F = self._call(self.model, x_in_or_x, sigma_t, text_emb, transition_mean, softness).to(self.dtype)
                Fg = (self._call(self.gnet, x_in_or_x, sigma_t, g_emb, transition_mean, softness).to(self.dtype)
                      if guided else None)
            return F, Fg

        for i in range(self.num_steps):
            t_cur, t_next = f32(t_host[i].item()), f32(t_host[i + 1].item())
            x_cur = x_next
            if self.s_churn > 0 and self.s_min <= t_cur <= self.s_max:
                gamma = min(self.s_churn / self.num_steps, math.sqrt(2) - 1)
            else:
                gamma = 0
            # fp32 scalar arithmetic in the reference's operation order (:98-99)
            t_hat = f32(t_cur + f32(f32(gamma) * t_cur)) if gamma else t_cur
            nscale = f32(np.sqrt(f32(f32(t_hat * t_hat) - f32(t_cur * t_cur))) * f32(self.s_noise)) if gamma else f32(0)
            eps = torch.randn_like(x_cur)          # always drawn (quirk Q14)
            sig_hat = t_dev[i] if not gamma else torch.tensor(float(t_hat), device=noise.device)
            x_hat, x_in = ops.edm_heun_pre(x_cur, eps, float(nscale), float(t_hat), sd)
            F, Fg = evaluate(x_in if native else x_hat, sig_hat)
            d_cur, x_next, x_in_next = ops.edm_heun_euler(x_hat, x_in if native else None, F, Fg, self.guide,
                                                          float(t_hat), float(t_next), sd)
            if i < self.num_steps - 1:
                F2, Fg2 = evaluate(x_in_next if native else x_next, t_dev[i + 1])
                x_next = ops.edm_heun_correct(x_hat, x_next, x_in_next if native else None, F2, Fg2, self.guide,
                                              float(t_hat), float(t_next), sd, d_cur)
        return x_next
