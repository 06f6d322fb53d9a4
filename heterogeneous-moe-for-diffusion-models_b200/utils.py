"""Loss-side router statistics and the host-side producers of the hot path's inputs, with the public
names of the reference's Utils/utils.py: EDM_LOSS, MaskGenerator, ZetaScheduler, sample_sigma_hybrid; plus
make_train_inputs, the loop body's noise add and both mask generators as one device launch."""
import math
from typing import Optional

import numpy as np
import torch
import torch.nn as nn


def sample_sigma_hybrid(batch_size, sigma_min=0.002, sigma_max=80.0, p_mean=-0.4, p_std=1.0, extreme_prob=0.2,
                        device="cuda", generator=None):
    """(1-p) log-normal + p log-uniform noise levels, clamped and shuffled; ref Utils/utils.py:26-61."""
    n_ln = int(batch_size * (1 - extreme_prob))
    n_u = batch_size - n_ln
    ln = (torch.randn([n_ln, 1, 1, 1], device=device, generator=generator) * p_std + p_mean).exp()
    lo, hi = math.log(sigma_min), math.log(sigma_max)
    un = (torch.rand([n_u, 1, 1, 1], device=device, generator=generator) * (hi - lo) + lo).exp()
    sigma = torch.cat([ln, un], dim=0).clamp(sigma_min, sigma_max)
    return sigma[torch.randperm(batch_size, device=device, generator=generator)]


def make_train_inputs(latent, sigma, unet_mask_gen=None, vit_mask_gen=None, step: int = 0, eps=None, generator=None):
    """The input producers of the reference's loop body (Utils/training.py:133-142) on the device in one launch:
    `images_noised = latent + randn_like(latent) * sigma` and both router masks `mask_gen(sigma, step)`.
    `eps` may be supplied (parity runs); otherwise it is drawn with torch's device generator, as the reference does.
    Returns (images_noised, Unet_mask, vit_mask); a mask is None when its generator is None."""
    from . import ops
    if eps is None:
        eps = torch.randn(latent.shape, device=latent.device, dtype=latent.dtype, generator=generator)

    def desc(g):
        if g is None:
            return None
        return (g._centers_host, g.p_mean, g.p_std, g.bandwidth_scheduler(step), g.min_active)
    return ops.train_inputs(latent, eps, sigma, desc(unet_mask_gen), desc(vit_mask_gen))


class EDM_LOSS(nn.Module):
    """ref Utils/utils.py:105-172.  `router_stats=(unet_stats, vit_stats)` (the partial sums the fused
    router kernel emitted: [sum_t probs | counts | sum_t z]) replaces ~16 small launches for the
    load-balance and z-loss terms; without it the same formulas run on the (B, E) tensors."""

    def __init__(self, num_experts: int, sigma_data: float = 0.5, Unet_bal: float = 0.0005, vit_bal: float = 0.0005,
                 z_bal: float = 0.0001, prior_bal: float = 0.001, transition_sigma: float = 1.0, sharpness: float = 2.0):
        super().__init__()
        self.num_experts = num_experts
        self.sigma_data = sigma_data
        self.Unet_lambda, self.vit_lambda, self.z_bal, self.prior_bal = Unet_bal, vit_bal, z_bal, prior_bal

    def forward(self, sigma_vec, x, sigma, out_model, router_stats=None):
        D = out_model["denoised"]
        E = self.num_experts
        if router_stats is None:
            # the fused router kernel attaches its partial sums to the tensors it returned (model_components.Router)
            su = getattr(out_model["Unet_router_loss"], "_hdmoe_stats", None)
            sv = getattr(out_model["vit_router_loss"], "_hdmoe_stats", None)
            if su is not None and sv is not None:
                router_stats = (su, sv)
        if D.is_cuda and D.dtype == torch.float32 and x.dtype == torch.float32 and D[0].numel() % 4 == 0 and D.shape == x.shape:
            # every image-dependent term is a function of the per-sample squared error: ONE reduction kernel forward, one
            # scaling kernel backward (csrc/edm_step.cu), the rest on [B]-sized vectors
            from . import ops
            B, per = D.shape[0], D[0].numel()
            se = ops.sqerr_rows(D, x)
            mse = se.sum() / (B * per)
            if out_model["log_var"] is None:
                pure = mse
            else:
                lv = out_model["log_var"].clamp(min=-10, max=10).reshape(-1)
                lv = lv.expand(B) if lv.numel() == 1 else lv
                pure = torch.mean(se / per * torch.exp(-lv) + lv)
            pure = pure.clamp(max=50)
            return self._finish(pure, mse, out_model, router_stats)
        err = (D - x) ** 2
        if out_model["log_var"] is None:
            pure = torch.mean(err)
        else:
            lv = out_model["log_var"].clamp(min=-10, max=10)
            pure = torch.mean(err / lv.exp() + lv)
        pure = pure.clamp(max=50)
        return self._finish(pure, torch.mean(err), out_model, router_stats)

    def _finish(self, pure, mse, out_model, router_stats):
        E = self.num_experts
        if router_stats is not None:
            su, sv = router_stats
            B = out_model["Unet_router_loss"].shape[0]
            lb_u = E * torch.sum((su[:E] / B) ** 2)
            lb_v = E * torch.sum((sv[:E] / B) ** 2)
            z_u, z_v = su[2 * E] / B, sv[2 * E] / B
        else:
            lb_u = self.load_balance(out_model["Unet_router_loss"], E)
            lb_v = self.load_balance(out_model["vit_router_loss"], E)
            z_u, z_v = self.z_loss(out_model["Unet_raw"]), self.z_loss(out_model["vit_raw"])
        bal = (self.Unet_lambda * lb_u + self.vit_lambda * lb_v).clamp(max=50)
        zl = (self.z_bal * z_u + self.z_bal * z_v).clamp(max=50)
        return {"loss": (pure + zl + bal).clamp(max=50), "denoising": mse, "balance": bal, "z_loss": zl,
                "entropy": 0.0, "pure_loss": pure}

    @staticmethod
    def load_balance(gate_probs, num_experts):
        return num_experts * torch.sum(gate_probs.mean(dim=0) ** 2)

    @staticmethod
    def entropy_loss(probs):
        return -torch.mean(torch.sum(probs * torch.log(probs + 1e-8), dim=-1))

    @staticmethod
    def z_loss(logits):
        z = torch.logsumexp(logits.clamp(min=-50, max=50), dim=-1) ** 2
        return torch.mean(z.clamp(max=100))


class ZetaScheduler:
    """Exploration-noise schedule; ref Utils/utils.py:175-225."""

    def __init__(self, total_steps: int, max_zeta: float, min_zeta: float = 0.0, strategy: str = "cos",
                 alpha: float = 4.0, warmup_ratio: float = 0.05):
        self.total_steps, self.max_zeta, self.min_zeta = total_steps, max_zeta, min_zeta
        self.strategy, self.alpha = strategy, alpha
        self.warmup_steps = int(total_steps * warmup_ratio)

    def get_zeta(self, step: int) -> float:
        if step < self.warmup_steps:
            return self.max_zeta
        if step >= self.total_steps:
            return self.min_zeta
        cur, tot = step - self.warmup_steps, self.total_steps - self.warmup_steps
        span = self.max_zeta - self.min_zeta
        if self.strategy == "cos":
            return float(self.min_zeta + span * 0.5 * (1 + np.cos(np.pi * cur / tot)))
        if self.strategy == "exp":
            term = max(min(-self.alpha * (cur - (self.max_zeta / tot)), 10), -10)
            return float(max(min(span * np.exp(term) + self.min_zeta, self.max_zeta), self.min_zeta))
        raise ValueError(f"Unknown strategy: {self.strategy}")


class MaskGenerator(nn.Module):
    """Rank-spaced band mask over the log-normal percentile of sigma; ref Utils/utils.py:228-330."""

    def __init__(self, expert_attributes: list, p_mean: float = -0.4, p_std: float = 1.0, bandwidth: float = 0.3,
                 max_bandwidth: float = 0.9, min_active: int = 1, total_steps: int = 5000, step_size: float = 0.1,
                 noise_range: tuple = (0.0, 1.0), strat_band: str = "step"):
        super().__init__()
        self.num_intervals = len(expert_attributes)
        self.strat_band, self.total_steps, self.max_bw, self.step_size = strat_band, total_steps, max_bandwidth, step_size
        self.p_mean, self.p_std, self.bandwidth, self.min_active = p_mean, p_std, bandwidth, min_active
        attrs = torch.tensor(expert_attributes, dtype=torch.float32)
        order = torch.sort(attrs, stable=True).indices
        centers = torch.zeros_like(attrs)
        centers[order] = torch.linspace(noise_range[0], noise_range[1], steps=len(attrs))
        self.register_buffer("expert_centers", centers)
        self._centers_host = centers.tolist()          # for the fused device producer (no device read per step)

    @torch.no_grad()
    def __call__(self, sigma: torch.Tensor, step: int) -> torch.Tensor:
        ls = torch.log(sigma.flatten())
        pct = (0.5 * (1 + torch.erf((ls - self.p_mean) / (self.p_std * np.sqrt(2))))).clamp(0, 1)
        dist = torch.abs(pct.view(-1, 1) - self.expert_centers.to(sigma.device).view(1, -1))
        mask = (dist <= self.bandwidth_scheduler(step)).float()
        top = torch.topk(-dist, k=self.min_active, dim=-1).indices
        mask.scatter_(1, top, 1.0)
        return mask

    def bandwidth_scheduler(self, step: int) -> float:
        if step >= self.total_steps:
            return self.max_bw
        if self.strat_band == "linear":
            return self.bandwidth + (self.max_bw - self.bandwidth) * (step / float(self.total_steps))
        if self.strat_band == "step":
            cur = int(step / (self.total_steps * self.step_size))
            return self.bandwidth + (self.max_bw - self.bandwidth) * min(cur / int(1.0 / self.step_size), 1.0)
        raise ValueError(self.strat_band)
