"""CPU-side checks of the drop-in boundary: the C-ABI library exports every symbol the header declares,
the Python modules keep the reference's interface (constructor / forward signatures, state_dict keys, init
RNG order), the host-side producers match the reference fixtures, and nothing falls back to the CPU."""
import ctypes
import inspect
import os
import re

import pytest
import torch

import hdmoe_b200
from conftest import ROOT, TINY, golden_weights, load_golden
from hdmoe_b200 import _lib, ops
from hdmoe_b200 import model_components as mc
from hdmoe_b200 import model_config1 as c1
from hdmoe_b200 import model_config2 as c2
from hdmoe_b200.EDM_sampler import EDM_Sampler
from hdmoe_b200.utils import EDM_LOSS, MaskGenerator, ZetaScheduler


def _header_symbols():
    names = set()
    for h in os.listdir(os.path.join(ROOT, "include")):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(hdmoe_[a-z0-9_]+)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libhdmoe_b200.so not built (run `make`)"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"
        assert s in _lib.PROTOTYPES, f"{s} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(syms)
    assert _lib.lib().hdmoe_version() >= 100


def test_no_cpu_fallback():
    w = torch.rand(8, 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.dispatch_plan(w)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.router_gate(torch.rand(8, 16), None, torch.rand(4, 16), 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.edm_precond_in(torch.rand(2, 4, 4, 4), torch.ones(2), 0.5)


def test_argument_errors_are_reported():
    lib = _lib.lib()
    rc = lib.hdmoe_router_gate_fwd(None, None, None, None, 0.0, None, None, 4, 16, 99, 1, None, None, None, None,
                                   None, None, None, None)
    assert rc == -1 and b"E <=" in lib.hdmoe_last_error()
    with pytest.raises(RuntimeError, match="code -1"):
        _lib.check(rc, "router_gate_fwd")


# reference signatures (models/model_components.py:79-85,18-22,281-293,589-604; model_config2.py:331-355;
# Utils/EDM_sampler.py:7-20,34-35,72-79)
REF_SIGS = {
    (mc.Router, "__init__"): ["in_channels", "time_dim", "top_k", "num_experts", "dropout"],
    (mc.Router, "forward"): ["x", "time_emb", "mask", "zeta"],
    (mc.Scaling_router, "__init__"): ["emb_dim", "num_experts", "dropout"],
    (mc.Scaling_router, "forward"): ["x", "zeta"],
    (mc.Unet_expert, "__init__"): ["img_resolution", "img_channels", "time_emb_dim", "text_emb_dim", "channel_mult",
                                   "model_channels", "channel_mult_emb", "num_blocks", "kernel_size", "label_balance",
                                   "concat_balance"],
    (mc.Unet_expert, "forward"): ["x", "time_emb", "text_emb"],
    (mc.Vit_expert, "__init__"): ["num_heads", "num_groups", "in_channels", "seq_ln", "emb_dim", "num_blocks",
                                  "patch_size", "time_dim", "text_dim", "res_balance", "attn_balance", "emb_balance",
                                  "gain_s", "gain_t"],
    (mc.Vit_expert, "forward"): ["x", "time_emb", "text_emb"],
    (c2.preconditioned_HDMOEM, "forward"): ["x", "sigma", "text_emb", "Unet_router_mask", "Vit_router_mask", "zeta",
                                            "transition_point", "softness", "return_log_var"],
    (c1.preconditioned_HDMOEM, "forward"): ["x", "sigma", "text_emb", "Unet_router_mask", "Vit_router_mask", "zeta",
                                            "return_log_var"],
    (c2.HDMOEM, "forward"): ["x", "time_vec", "text_emb", "Unet_router_mask", "Vit_router_mask", "zeta",
                             "transition_point", "softness"],
    (EDM_Sampler, "__init__"): ["model", "Guide_net", "num_solve_steps", "sigma_min", "sigma_max", "rho", "S_churn",
                                "S_min", "S_max", "S_noise", "guidance", "dtype"],
    (EDM_Sampler, "denoise"): ["x", "sigma", "text_emb", "transition_mean", "softness", "uncond_text_emb"],
    (EDM_Sampler, "sample"): ["noise", "text_emb", "transition_mean", "softness", "uncond_text_emb"],
}


@pytest.mark.parametrize("key", list(REF_SIGS), ids=lambda k: f"{k[0].__name__}.{k[1]}")
def test_reference_signatures(key):
    cls, meth = key
    params = [p for p in inspect.signature(getattr(cls, meth)).parameters if p != "self"]
    ref = REF_SIGS[key]
    assert params[:len(ref)] == ref     # extra trailing keyword-only conveniences (noise=...) are allowed
    assert c2.router_to_unet_experts.__name__ == "router_to_unet_experts"
    assert list(inspect.signature(c2.router_to_unet_experts).parameters)[:5] == ["x", "experts", "out_router",
                                                                                 "time_emb", "text_emb"]


@pytest.mark.parametrize("variant,wfile", [(2, "weights_cfg2_seed0"), (1, "weights_cfg1_seed0")])
def test_state_dict_layout_and_init_order_match_reference(variant, wfile):
    """Same seed -> same parameters as the reference constructor (so checkpoints and seeds interchange)."""
    ref = load_golden(wfile)
    torch.manual_seed(0)
    model = (c2 if variant == 2 else c1).preconditioned_HDMOEM(**TINY)
    sd = model.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    n_checked = 0
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(ref[k].shape), k
        if float(v.abs().max()) != 0:          # zero-initialised entries were re-drawn in the fixture
            assert torch.equal(v, ref[k]), k
            n_checked += 1
    assert n_checked > 100
    model.load_state_dict(ref)
    assert model.num_experts == 4 and hasattr(model.net, "Unet_router") and hasattr(model.net, "VIT_experts")


def test_host_producers_match_reference():
    g = load_golden("producers")
    sigma = g["mask.sigma"]
    for tag, attrs, rng in (("unet", [3, 3, 5, 5], (0.0, 0.6)), ("vit", [4, 8, 8, 16], (0.4, 1.0))):
        mg = MaskGenerator(expert_attributes=attrs, p_mean=-1.2, p_std=1.6, bandwidth=0.3, max_bandwidth=0.8,
                           min_active=1, total_steps=5000, step_size=0.1, noise_range=rng, strat_band="step")
        assert torch.equal(mg.expert_centers, g[f"mask.{tag}.centers"])
        for step in (0, 700, 2600, 6000):
            assert torch.equal(mg(sigma, step), g[f"mask.{tag}.{step}"])
    for strat in ("cos", "exp"):
        zs = ZetaScheduler(total_steps=900, max_zeta=2, min_zeta=0.01, strategy=strat, alpha=4.0, warmup_ratio=0.05)
        for i, s in enumerate(g["zeta.steps"].tolist()):
            assert abs(zs.get_zeta(s) - float(g[f"zeta.{strat}"][i])) < 1e-12
    smp = EDM_Sampler(None, None, num_solve_steps=18)
    assert torch.equal(smp.t_steps()[:-1], g["sampler.t_steps18"])
    assert abs(float(EDM_LOSS.load_balance(torch.full((16, 4), 0.25), 4)) - 1.0) < 1e-6


def test_direct_gradient_handoff_keeps_accumulation_semantics():
    """prepared.deliver_grads / unalias_grads (host logic of the clone-free gradient hand-off): a parameter without
    .grad receives the persistent view itself; one that already holds a gradient goes through autograd; a .grad that
    still aliases the persistent buffer is cloned before the buffer is overwritten."""
    import torch
    from hdmoe_b200 import prepared
    buf = torch.zeros(6)
    p1, p2 = torch.nn.Parameter(torch.ones(2, 2)), torch.nn.Parameter(torch.ones(2))
    views = [buf[0:4].view(2, 2), buf[4:6]]
    buf.copy_(torch.arange(6.0))
    p2.grad = torch.full((2,), 10.0)
    with torch.no_grad():
        out = prepared.deliver_grads([p1, p2], views)
    assert out[0] is None and p1.grad.data_ptr() == buf.data_ptr()          # handed off, no copy
    assert out[1] is views[1] and float(p2.grad[0]) == 10.0                  # existing gradient: autograd adds
    with torch.enable_grad():                                                # double backward: never hand off
        assert prepared.deliver_grads([p1, p2], views) == views
    prepared.unalias_grads([p1, p2], buf)                                    # next backward is about to overwrite buf
    assert p1.grad.data_ptr() != buf.data_ptr() and torch.equal(p1.grad, torch.tensor([[0.0, 1.0], [2.0, 3.0]]))
    buf.zero_()
    assert float(p1.grad.sum()) == 6.0


def test_fused_vit_weight_layout_matches_kernel_order():
    """Host logic of vit_fused.FusedVitExperts: the ten MP_Conv matrices of every (expert, block) lie contiguously in
    the prepared-weight buffer in the order csrc/vit_block.cu reads them (linear1, q, k, v, out, q_time, k_time,
    v_time, linear2, linear3 = 19 456 floats), and the aux offsets follow 256 + 8 S^2 per block.  No kernel is run."""
    import torch
    from hdmoe_b200 import model_components as mc, vit_fused
    patches = [4, 8, 8, 16]
    experts = torch.nn.ModuleList([mc.Vit_expert(num_heads=8, num_groups=4, in_channels=32, seq_ln=(32 // p) ** 2,
                                                 emb_dim=32, num_blocks=4, patch_size=p, time_dim=64, text_dim=768)
                                   for p in patches])
    runner = vit_fused.FusedVitExperts(experts)
    runner.group._build(torch.device("cpu"))
    runner._build_meta()                                   # asserts the layout internally
    meta = runner.meta
    assert meta.E == 4 and meta.nb == 4 and list(meta.tokens) == [64, 16, 16, 4]
    o = 0
    for e, ex in enumerate(experts):
        for b in range(4):
            assert meta.a_off[b][e] == o and o % 4 == 0
            o += 256 + 8 * ex.seq_ln ** 2
    assert runner._aux().numel() == o
    for b in range(4):
        for e in range(4):
            assert meta.w_off[b][e] % 4 == 0
            if b:
                assert meta.w_off[b][e] - meta.w_off[b - 1][e] == vit_fused._WBLOCK
    # the CPU never takes the fused path (no CPU fallback of the kernels; the composite torch path runs instead)
    assert not vit_fused.fusable(experts, torch.zeros(2, 32, 32, 32))


def test_checkpoint_file_format_round_trip_and_reference_load(tmp_path):
    """save_checkpoint writes the reference's dictionary (Utils/training.py:242-271): it reloads here (model + AdamW
    moments) and, where the unmodified reference is importable (the build container), into the reference's own model
    with strict key checking, as Utils/training.py:301-304 does."""
    import sys
    from hdmoe_b200.training import checkpoint_dir, load_checkpoint, save_checkpoint
    torch.manual_seed(0)
    model = c2.preconditioned_HDMOEM(**TINY)
    opt = torch.optim.AdamW(model.parameters(), lr=5e-4)
    for p in model.parameters():
        p.grad = torch.randn_like(p) * 1e-2
    opt.step()
    cfg = {"model_configs": {"save_dir": str(tmp_path / "ck")}, "note": "unit test"}
    assert checkpoint_dir(cfg) == str(tmp_path / "ck") and checkpoint_dir({}) == "./checkpoints"
    path = save_checkpoint(model, opt, step=7, mse_score=0.25, configs=cfg, filename="m.pt")
    raw = torch.load(path, weights_only=False)
    assert sorted(raw) == ["config", "model_state_dict", "mse", "optimizer_state_dict", "step"]
    torch.manual_seed(1)
    model2 = c2.preconditioned_HDMOEM(**TINY)
    opt2 = torch.optim.AdamW(model2.parameters(), lr=1e-3)
    meta = load_checkpoint(path, model2, opt2)
    assert meta["step"] == 7 and meta["mse"] == 0.25 and meta["config"]["note"] == "unit test"
    for (k, a), b in zip(model.state_dict().items(), model2.state_dict().values()):
        assert torch.equal(a, b), k
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys() and all(torch.equal(s1[i]["exp_avg_sq"], s2[i]["exp_avg_sq"]) for i in s1)
    assert opt2.param_groups[0]["lr"] == 5e-4
    if not os.path.isdir("/root/reference/models"):
        return
    sys.path.insert(0, "/root/reference")
    try:
        from models.model_config2 import preconditioned_HDMOEM as RefModel
    finally:
        sys.path.remove("/root/reference")
    ref = RefModel(**TINY)
    ref.load_state_dict(raw["model_state_dict"])           # strict: every key and shape is the reference's
    torch.save({"model_state_dict": ref.state_dict()}, tmp_path / "ref.pt")
    model2.load_state_dict(torch.load(tmp_path / "ref.pt")["model_state_dict"])


def test_model_options_pin_and_restore():
    """set_model_options pins switches on ONE model; the pins apply inside that model's forward only and the process
    defaults come back afterwards (two models with different settings can coexist)."""
    from hdmoe_b200 import _denoiser as D
    torch.manual_seed(0)
    m = c2.preconditioned_HDMOEM(**TINY)
    hdmoe_b200.set_model_options(m, expert_dtype=torch.bfloat16, grouped_experts=False)
    assert m.net._hdmoe_options == {"expert_dtype": torch.bfloat16, "grouped_experts": False}
    before = (D.get_expert_dtype(), D._GROUPED[0])
    with D._model_options(m.net):
        assert D.get_expert_dtype() == torch.bfloat16 and D._GROUPED[0] is False
    assert (D.get_expert_dtype(), D._GROUPED[0]) == before
    hdmoe_b200.set_model_options(m, grouped_experts=None)
    assert m.net._hdmoe_options == {"expert_dtype": torch.bfloat16}
    with pytest.raises(ValueError):
        hdmoe_b200.set_model_options(m, no_such_switch=True)
    with pytest.raises(ValueError):
        hdmoe_b200.set_model_options(m, expert_dtype=torch.float16)
    torch.manual_seed(0)
    assert list(m.state_dict().keys()) == list(c2.preconditioned_HDMOEM(**TINY).state_dict().keys())   # pins are not state
