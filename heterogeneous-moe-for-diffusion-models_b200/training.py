"""Checkpoint I/O either side of the hot path, in the reference's file format.

`save_checkpoint` keeps the signature and the dictionary layout of the reference's Utils/training.py:242-271
(`step`, `model_state_dict`, `optimizer_state_dict`, `mse`, `config`, written with `torch.save`), so a file written
here loads into the unmodified reference (`modeli.load_state_dict(ckeck['model_state_dict'])`, Utils/training.py:
301-304) and a file written by the reference loads here: the `state_dict` keys and shapes are the reference's
(tests/test_boundary_cpu.py), and `hdmoe_b200.optim.FusedAdamW` reads and writes torch's AdamW state format.

The VAE / CLIP wrappers of Utils/VAE_CLIP.py need `diffusers` and pretrained weights that are not available offline;
they are out of scope."""
import os
from typing import Any, Dict, Optional

import torch


def _to_cpu(obj):
    if torch.is_tensor(obj):
        return obj.detach().to("cpu")
    if isinstance(obj, dict):
        return {k: _to_cpu(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_to_cpu(v) for v in obj)
    return obj


def checkpoint_dir(configs: Dict[str, Any]) -> str:
    """Directory rule of the reference (Utils/training.py:249-254)."""
    if "save_dir" in configs:
        return configs["save_dir"]
    if "model_configs" in configs and "save_dir" in configs["model_configs"]:
        return configs["model_configs"]["save_dir"]
    return "./checkpoints"


def save_checkpoint(model, optimizer, step, mse_score, configs, filename) -> str:
    """ref Utils/training.py:242-271.  Returns the path written.  Tensors are moved to the CPU first so the file does
    not pin a device index (the reference loads with `map_location`, either works)."""
    save_path = checkpoint_dir(configs)
    os.makedirs(save_path, exist_ok=True)
    full_path = os.path.join(save_path, filename)
    net = model.module if hasattr(model, "module") else model
    checkpoint = {
        "step": step,
        "model_state_dict": _to_cpu(net.state_dict()),
        "optimizer_state_dict": _to_cpu(optimizer.state_dict()),
        "mse": mse_score,
        "config": configs,
    }
    torch.save(checkpoint, str(full_path))
    return full_path


def load_checkpoint(path: str, model, optimizer=None, map_location: Optional[Any] = None, strict: bool = True) -> Dict[str, Any]:
    """Inverse of `save_checkpoint` (the reference inlines it: Utils/training.py:301-304).  Restores the model (and the
    optimizer when given) in place and returns the remaining fields (`step`, `mse`, `config`).

    `weights_only=False` because the reference stores its config dictionaries (plain Python objects) in the file."""
    ck = torch.load(path, map_location=map_location, weights_only=False)
    net = model.module if hasattr(model, "module") else model
    net.load_state_dict(ck["model_state_dict"], strict=strict)
    if optimizer is not None and ck.get("optimizer_state_dict") is not None:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    return {"step": ck.get("step"), "mse": ck.get("mse"), "config": ck.get("config")}
