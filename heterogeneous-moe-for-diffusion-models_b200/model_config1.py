"""model_config1 surface (ref models/model_config1.py): learned Scaling_router gains and the soft
query/context swap; forward has no transition_point/softness (quirk Q22)."""
from ._denoiser import HDMOEM as _Base
from ._denoiser import preconditioned_HDMOEM as _PBase
from ._denoiser import router_to_unet_experts  # noqa: F401


class HDMOEM(_Base):
    _variant = 1

    def forward(self, x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta, alpha_routing: float = 10,
                noise=None):
        """ref models/model_config1.py:210-309"""
        return self._forward(x, time_vec, text_emb, Unet_router_mask, Vit_router_mask, zeta,
                             alpha_routing=alpha_routing, noise=noise)


class preconditioned_HDMOEM(_PBase):
    _net_cls = HDMOEM

    def forward(self, x, sigma, text_emb, Unet_router_mask, Vit_router_mask, zeta, return_log_var: bool = False,
                noise=None, **fast):
        """ref models/model_config1.py:391-468"""
        return self._forward(x, sigma, text_emb, Unet_router_mask, Vit_router_mask, zeta,
                             return_log_var=return_log_var, noise=noise, **fast)
