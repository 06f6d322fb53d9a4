"""pytest configuration: the `gpu` marker, repo-root imports and golden-fixture helpers."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    """Returns {key: torch tensor | python scalar} of tests/golden/<name>.npz."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = {}
    for k in z.files:
        v = z[k]
        if v.dtype.kind in "US":
            out[k] = str(v)
        elif v.ndim == 0 and k.startswith(("meta.", "in.zeta")) or k.endswith((".zeta", ".k")):
            out[k] = v.item()
        else:
            out[k] = torch.from_numpy(np.array(v))
    return out


def golden_weights(g):
    return load_golden(g["meta.weights_file"])


TINY = dict(IN_in_channels=4, IN_img_resolution=8, internal_channels=8, time_emb_dim=16, text_emb_dim=24,
            num_experts=4, top_k=2, Fourier_bandwidth=1.0, VIT_num_blocks=1, VIT_patch_sizes=[2, 4, 4, 8],
            VIT_num_groups=2, VIT_num_heads=2, VIT_emb_size=8, Unet_num_blocks=1, Unet_channel_mult=[1, 2],
            Unet_kernel_sizes=[(3, 3), (3, 3), (5, 5), (5, 5)], Unet_model_channels=8, Unet_channel_mult_emb=2,
            Unet_label_balance=0.5, Unet_concat_balance=0.5, sigma_data=0.5, log_var_channels=8)

# Utils/configs.py:3-35 of the reference (the shipped model hyper-parameters)
FULL = dict(IN_in_channels=4, IN_img_resolution=32, internal_channels=32, time_emb_dim=64, text_emb_dim=768,
            num_experts=4, top_k=1, Fourier_bandwidth=1.0, VIT_num_blocks=4, VIT_patch_sizes=[4, 8, 8, 16],
            VIT_num_groups=4, VIT_num_heads=8, VIT_emb_size=32, Unet_num_blocks=2, Unet_channel_mult=[1, 2],
            Unet_kernel_sizes=[(3, 3), (3, 3), (5, 5), (5, 5)], Unet_model_channels=32, Unet_channel_mult_emb=2,
            Unet_label_balance=0.5, Unet_concat_balance=0.5, sigma_data=0.5, log_var_channels=32)


def rel_l2(a, b):
    a = a.detach().double().flatten()
    b = b.detach().double().flatten()
    return float(((a - b).norm() / b.norm().clamp_min(1e-30)).detach())
