"""hdmoe_b200 -- B200-native heterogeneous-MoE hot path of the EDM denoiser.

Python surface = the reference's module interface for this path:
    model_internals, model_components, model_config1, model_config2, EDM_sampler, utils, training (checkpoint I/O)
underneath, hand-written sm_100a kernels (csrc/, C ABI in include/hdmoe_b200.h) reached through ops.py.
Importing the package does not need a GPU; calling any op does, and a missing libhdmoe_b200.so raises.
"""
from . import _lib, ops  # noqa: F401
from . import model_internals, model_components, model_config1, model_config2, EDM_sampler, utils  # noqa: F401
from . import expert_parallel, training  # noqa: F401
from ._denoiser import (disable_expert_parallel, enable_expert_parallel, get_expert_dtype,  # noqa: F401
                        set_expert_dtype, set_grouped_experts, set_branch_streams, set_model_options)
from .EDM_sampler import EDM_Sampler  # noqa: F401
from .ops import set_gconv_impl  # noqa: F401
from .router_trunk import set_router_tcgen05_trunk  # noqa: F401

__all__ = ["ops", "model_internals", "model_components", "model_config1", "model_config2", "EDM_sampler", "utils",
           "EDM_Sampler", "set_expert_dtype", "get_expert_dtype", "set_grouped_experts", "set_branch_streams",
           "set_model_options"]
