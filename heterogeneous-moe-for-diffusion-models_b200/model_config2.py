"""model_config2 surface (ref models/model_config2.py): analytic sigma-sigmoid path scaling; the only
configuration the reference's sampler and training loop can drive (quirk Q22)."""
from ._denoiser import HDMOEM as _Base
from ._denoiser import preconditioned_HDMOEM as _PBase
from ._denoiser import router_to_unet_experts  # noqa: F401  (same helper name as the reference)


class HDMOEM(_Base):
    _variant = 2

    def forward(self, x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta, transition_point, softness,
                noise=None):
        """-> (out, Unet_gate_probs, Unet_raw, Vit_gate_probs, vit_raw, scaling_factors, out_gate);
        ref models/model_config2.py:206-303."""
        return self._forward(x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta,
                             transition_point=transition_point, softness=softness, noise=noise)


class preconditioned_HDMOEM(_PBase):
    _net_cls = HDMOEM

    def forward(self, x, sigma, text_emb, Unet_router_mask, Vit_router_mask, zeta, transition_point, softness,
                return_log_var: bool = False, noise=None, **fast):
        """-> dict(denoised, Unet_router_loss, Unet_raw, vit_router_loss, vit_raw, scaling_net_out, out_gate,
        log_var); ref models/model_config2.py:389-468."""
        return self._forward(x, sigma, text_emb, Unet_router_mask, Vit_router_mask, zeta,
                             return_log_var=return_log_var, transition_point=transition_point, softness=softness,
                             noise=noise, **fast)
