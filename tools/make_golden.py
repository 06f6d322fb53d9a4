"""Generate golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python tools/make_golden.py

Writes tests/golden/*.npz.  The fixtures pin the oracle (oracle/hdmoe_oracle.py); the GPU box
has no /root/reference, so tests read only the committed .npz files.

What is recorded, per case: the full state_dict (tiny hyper-parameters so the file stays small),
every input, every entry of the output dict, router tuples, the dispatch order implied by the
reference's boolean-mask loop, gradients of a few parameters, and the post-forward weights
(train-mode MP_Conv mutates its parameter, quirk Q6).  Train-mode randomness: dropout is set to
0 and the exploration noise drawn by torch.randn_like is recorded in call order.
"""
import importlib.util
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
sys.path.insert(0, REF)

from models import model_components as mc  # noqa: E402
from models import model_config1 as c1  # noqa: E402
from models import model_config2 as c2  # noqa: E402
from models import model_internals as mi  # noqa: E402
from Utils.EDM_sampler import EDM_Sampler  # noqa: E402
from Utils import utils as U  # noqa: E402

TINY = dict(IN_in_channels=4, IN_img_resolution=8, internal_channels=8, time_emb_dim=16, text_emb_dim=24,
            num_experts=4, top_k=2, Fourier_bandwidth=1.0, VIT_num_blocks=1, VIT_patch_sizes=[2, 4, 4, 8],
            VIT_num_groups=2, VIT_num_heads=2, VIT_emb_size=8, Unet_num_blocks=1, Unet_channel_mult=[1, 2],
            Unet_kernel_sizes=[(3, 3), (3, 3), (5, 5), (5, 5)], Unet_model_channels=8, Unet_channel_mult_emb=2,
            Unet_label_balance=0.5, Unet_concat_balance=0.5, sigma_data=0.5, log_var_channels=8)


def randomize_zero_init(model, gen):
    """out_gain, alpha_txt, pos_emb, rel_pos_bias and norm biases are zero at init (quirk Q9)."""
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.abs().max() == 0:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)


def set_dropout_zero(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
        if isinstance(m, mc.Unet_block):
            m.dropout = 0


class NoiseTap:
    """Records every torch.randn_like draw (in call order) while active."""

    def __init__(self):
        self.draws = []
        self._orig = torch.randn_like

    def __enter__(self):
        def tapped(t, *a, **k):
            r = self._orig(t, *a, **k)
            self.draws.append(r.clone())
            return r
        torch.randn_like = tapped
        return self

    def __exit__(self, *a):
        torch.randn_like = self._orig


AFTER_KEYS = ("net.input_proj.weights", "net.Unet_experts.0.encoders.8x8_conv.weights",
              "net.Unet_experts.3.decoders.4x4_in0.conv_res1.weights", "net.VIT_experts.1.diffit.0.linear2.weights",
              "net.VIT_experts.0.unpatch_proj.weights", "net.cross_attn.q_proj.weights", "net.gate2.weights",
              "log_var_linear.weights", "net.out_fourier1.weights")
_saved = set()


def save_weights(fname, sd):
    if fname in _saved:
        return
    _saved.add(fname)
    np.savez_compressed(os.path.join(OUT, fname + ".npz"), **npify({k: v.clone() for k, v in sd.items()}))


def npify(d):
    out = {}
    for k, v in d.items():
        if v is None:
            continue
        if torch.is_tensor(v):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    return out


def ref_dispatch_order(out_router):
    """Order in which the reference's loop (models/model_config2.py:25-37) visits rows."""
    src, exp = [], []
    for i in range(out_router.shape[1]):
        m = out_router[:, i] > 0
        idx = torch.nonzero(m).flatten()
        src.append(idx)
        exp.append(torch.full_like(idx, i))
    return torch.cat(src).to(torch.int32), torch.cat(exp).to(torch.int32)


def model_case(name, variant, top_k, train, B=6, seed=0):
    cfg = dict(TINY, top_k=top_k)
    mod = c2 if variant == 2 else c1
    torch.manual_seed(seed)
    model = mod.preconditioned_HDMOEM(**cfg)
    gen = torch.Generator().manual_seed(seed + 100)
    randomize_zero_init(model, gen)
    set_dropout_zero(model)
    model.train(train)
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    res = cfg["IN_img_resolution"]
    x0 = torch.randn(B, 4, res, res, generator=gen) * 0.5
    sigma = torch.exp(torch.randn(B, 1, 1, 1, generator=gen) * 1.6 - 1.2).clamp(0.002, 80.0)
    x = (x0 + sigma * torch.randn(x0.shape, generator=gen)).requires_grad_(True)
    text = torch.randn(B, 5, cfg["text_emb_dim"], generator=gen)
    um = (torch.rand(B, 4, generator=gen) > 0.3).float()
    vm = (torch.rand(B, 4, generator=gen) > 0.3).float()
    um[:, 0] = 1.0  # keep at least one live expert per row (all-masked rows are a separate case)
    vm[:, 3] = 1.0
    zeta = 0.7
    kw = dict(x=x, sigma=sigma, text_emb=text, Unet_router_mask=um, Vit_router_mask=vm, zeta=zeta,
              return_log_var=True)
    if variant == 2:
        kw.update(transition_point=-1.2, softness=1.6)
    cap = {}
    hooks = []
    for rn in ("Unet_router", "vit_router"):
        hooks.append(getattr(model.net, rn).register_forward_hook(
            lambda m, i, o, rn=rn: cap.__setitem__(rn, [t.detach().clone() for t in o])))
    with NoiseTap() as tap:
        out = model(**kw)
    for h in hooks:
        h.remove()
    crit = U.EDM_LOSS(num_experts=4, sigma_data=0.5, Unet_bal=0.05, vit_bal=0.1, z_bal=0.005, prior_bal=0.0)
    loss = crit(sigma_vec=sigma, x=x0, sigma=sigma, out_model=out)
    loss["loss"].backward()
    rec = {"in.x": x, "in.x0": x0, "in.sigma": sigma, "in.text": text, "in.unet_mask": um, "in.vit_mask": vm,
           "in.zeta": zeta, "meta.variant": variant, "meta.top_k": top_k, "meta.train": int(train)}
    for k, v in out.items():
        rec["out." + k] = v
    for k, v in loss.items():
        if torch.is_tensor(v):
            rec["loss." + k] = v
    for i, d in enumerate(tap.draws):
        rec[f"noise.{i}"] = d
    rec["meta.n_noise"] = len(tap.draws)
    for rn in ("Unet_router", "vit_router"):
        sw, gp, lg = cap[rn]
        rec[f"router.{rn}.sparse"] = sw
        rec[f"router.{rn}.probs"] = gp
        rec[f"router.{rn}.logits"] = lg
        s, e = ref_dispatch_order(sw)
        rec[f"router.{rn}.src_row"] = s
        rec[f"router.{rn}.expert_of_row"] = e
    rec["grad.x"] = x.grad
    named = dict(model.named_parameters())
    for gk in ("net.input_proj.weights", "net.Unet_router.linear.weights", "net.vit_router.time_linear.weights",
               "net.Unet_experts.0.out_gain", "net.Unet_experts.2.encoders.8x8_conv.weights",
               "net.VIT_experts.1.diffit.0.linear2.weights", "net.cross_attn.q_proj.weights", "net.alpha_txt",
               "log_var_linear.weights", "net.output_proj.weights"):
        g = named[gk].grad
        rec["grad." + gk] = g if g is not None else torch.zeros_like(named[gk])
    save_weights(f"weights_cfg{variant}_seed{seed}", sd0)
    rec["meta.weights_file"] = f"weights_cfg{variant}_seed{seed}"
    if train:
        for k, v in model.state_dict().items():
            if k.endswith(".weights") and (k in AFTER_KEYS or "router" in k):
                rec["sd_after." + k] = v
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npify(rec))
    print(name, "rows", {rn: cap[rn][0].gt(0).sum(0).tolist() for rn in cap}, "loss", float(loss["loss"]))


def sampler_case(name, guidance, S_churn, seed=3):
    cfg = dict(TINY, top_k=1)
    torch.manual_seed(seed)
    model = c2.preconditioned_HDMOEM(**cfg)
    gen = torch.Generator().manual_seed(seed + 100)
    randomize_zero_init(model, gen)
    model.eval()
    B = 3
    noise = torch.randn(B, 4, 8, 8, generator=gen)
    text = torch.randn(B, 5, 24, generator=gen)
    uncond = torch.zeros_like(text)
    smp = EDM_Sampler(model, model, num_solve_steps=5, guidance=guidance, S_churn=S_churn, S_noise=1.003)
    torch.manual_seed(11)
    with NoiseTap() as tap:
        out = smp.sample(noise, text, -1.2, 1.6, uncond_text_emb=uncond)
    rec = {"in.noise": noise, "in.text": text, "in.uncond": uncond, "out.x": out, "meta.guidance": guidance,
           "meta.S_churn": S_churn, "meta.num_steps": 5, "meta.S_noise": 1.003}
    for i, d in enumerate(tap.draws):
        rec[f"noise.{i}"] = d
    rec["meta.n_noise"] = len(tap.draws)
    save_weights(f"weights_cfg2_seed{seed}", model.state_dict())
    rec["meta.weights_file"] = f"weights_cfg2_seed{seed}"
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npify(rec))
    print(name, "max|x|", float(out.abs().max()))


def router_tail_case():
    """Router tail on its own at E up to 64, k in {1,2}, incl. all-masked rows and a single live expert."""
    rec = {}
    gen = torch.Generator().manual_seed(5)
    for tag, (B, E, k) in {"a": (64, 4, 1), "b": (96, 8, 2), "c": (50, 64, 2), "d": (33, 16, 1)}.items():
        torch.manual_seed(7)
        r = mc.Router(in_channels=32, time_dim=64, top_k=k, num_experts=E, dropout=0.0)
        r.train()
        pooled = torch.randn(B, 128, generator=gen).abs()
        te = torch.randn(B, 64, generator=gen)
        mask = (torch.rand(B, E, generator=gen) > 0.4).float()
        mask[0] = 0.0                      # all-masked row -> NaN probs (quirk Q3)
        mask[1] = 0.0
        mask[1, E - 1] = 1.0               # one live expert: with k=2 weights are [1, 0]
        mask[2] = 1.0
        zeta = 0.5
        # drive the tail directly: replicate Router.forward from line 143 on with the module's own layers
        x = pooled
        cond = r.time_linear(mi.mp_silu(te))
        gamma, beta = cond.chunk(2, dim=1)
        x = x * (1 + gamma) + beta
        x = r.linear(x)
        nz = torch.randn(x.shape, generator=gen)
        x = x + nz * zeta
        x = x.masked_fill(mask == 0, float("-inf"))
        gp = torch.softmax(x, dim=-1)
        tv, ti = torch.topk(x, k, dim=-1)
        gw = torch.softmax(tv, dim=-1)
        sp = torch.zeros_like(x).scatter(-1, ti, gw)
        rec.update({f"{tag}.pooled": pooled, f"{tag}.time_emb": te, f"{tag}.mask": mask, f"{tag}.noise": nz,
                    f"{tag}.zeta": zeta, f"{tag}.k": k, f"{tag}.w_time": r.time_linear.weights,
                    f"{tag}.w_lin": r.linear.weights, f"{tag}.logits": x, f"{tag}.probs": gp,
                    f"{tag}.sparse": sp, f"{tag}.topk_idx": ti.to(torch.int32),
                    f"{tag}.load_balance": U.EDM_LOSS.load_balance(torch.nan_to_num(gp), E),
                    f"{tag}.z_loss": U.EDM_LOSS.z_loss(x)})
    np.savez_compressed(os.path.join(OUT, "router_tail.npz"), **npify(rec))
    print("router_tail ok")


def moe_identity_case():
    """router_to_unet_experts with scaling 'experts' (expert e multiplies by e+1): pins order + combine."""
    class Scale(torch.nn.Module):
        def __init__(self, s):
            super().__init__()
            self.s = s

        def forward(self, x, time_emb, text_emb):
            return x * self.s + time_emb.mean(dim=1).view(-1, 1, 1, 1) + text_emb.mean(dim=1).view(-1, 1, 1, 1)

    rec = {}
    gen = torch.Generator().manual_seed(9)
    for tag, (T, E, k) in {"a": (17, 4, 1), "b": (40, 8, 2), "c": (5, 4, 2)}.items():
        x = torch.randn(T, 3, 4, 4, generator=gen)
        te = torch.randn(T, 6, generator=gen)
        txt = torch.randn(T, 7, 10, generator=gen)
        logits = torch.randn(T, E, generator=gen)
        if tag == "c":
            logits[:, 1] = float("-inf")
            logits[0] = float("-inf")          # all-masked token contributes nothing
        tv, ti = torch.topk(logits, k, dim=-1)
        w = torch.zeros_like(logits).scatter(-1, ti, torch.softmax(tv, dim=-1))
        experts = torch.nn.ModuleList([Scale(float(e + 1)) for e in range(E)])
        out = c2.router_to_unet_experts(x, experts, w, te, txt)
        s, e = ref_dispatch_order(w)
        rec.update({f"{tag}.x": x, f"{tag}.time": te, f"{tag}.text": txt, f"{tag}.w": w, f"{tag}.out": out,
                    f"{tag}.src_row": s, f"{tag}.expert_of_row": e})
    np.savez_compressed(os.path.join(OUT, "moe_identity.npz"), **npify(rec))
    print("moe_identity ok")


def producers_case():
    rec = {}
    gen = torch.Generator().manual_seed(2)
    sigma = torch.exp(torch.randn(32, 1, 1, 1, generator=gen) * 1.6 - 1.2).clamp(0.002, 80)
    for tag, attrs, rng in (("unet", [3, 3, 5, 5], (0.0, 0.6)), ("vit", [4, 8, 8, 16], (0.4, 1.0))):
        mg = U.MaskGenerator(expert_attributes=attrs, p_mean=-1.2, p_std=1.6, bandwidth=0.3, max_bandwidth=0.8,
                             min_active=1, total_steps=5000, step_size=0.1, noise_range=rng, strat_band="step")
        rec[f"mask.{tag}.centers"] = mg.expert_centers
        for step in (0, 700, 2600, 6000):
            rec[f"mask.{tag}.{step}"] = mg(sigma, step)
    rec["mask.sigma"] = sigma
    zs = U.ZetaScheduler(total_steps=900, max_zeta=2, min_zeta=0.01, strategy="cos", alpha=4.0, warmup_ratio=0.05)
    rec["zeta.steps"] = np.array([0, 10, 44, 45, 46, 300, 899, 900, 1200])
    rec["zeta.cos"] = np.array([float(zs.get_zeta(int(s))) for s in rec["zeta.steps"]])
    ze = U.ZetaScheduler(total_steps=900, max_zeta=2, min_zeta=0.01, strategy="exp", alpha=4.0, warmup_ratio=0.05)
    rec["zeta.exp"] = np.array([float(ze.get_zeta(int(s))) for s in rec["zeta.steps"]])
    # sampler with the reference tests' mock denoiser (tests/test_utilities/test_sampler.py:6-23)
    class Mock(torch.nn.Module):
        num_experts = 4

        def __init__(self, c):
            super().__init__()
            self.c = c

        def forward(self, x, sigma, **kw):
            return {"denoised": x * 0.9 if self.c is None else torch.full_like(x, self.c)}
    noise = torch.randn(2, 4, 8, 8, generator=gen)
    smp = EDM_Sampler(Mock(None), Mock(None), num_solve_steps=18)
    rec["sampler.noise"] = noise
    rec["sampler.mock09"] = smp.sample(noise, None, -1.2, 1.6)
    smp = EDM_Sampler(Mock(1.0), Mock(0.0), num_solve_steps=6, guidance=3.0)
    rec["sampler.cfg3_denoise"] = smp.denoise(noise, torch.tensor(1.0), None, -1.2, 1.6)
    rec["sampler.cfg3"] = smp.sample(noise, None, -1.2, 1.6)
    st = torch.arange(18, dtype=torch.float32)
    rec["sampler.t_steps18"] = (80 ** (1 / 7) + st / 17 * (0.002 ** (1 / 7) - 80 ** (1 / 7))) ** 7
    np.savez_compressed(os.path.join(OUT, "producers.npz"), **npify(rec))
    print("producers ok")


def primitives_case():
    rec = {}
    gen = torch.Generator().manual_seed(4)
    for tag, (cin, cout, k, shape) in {"lin": (24, 16, (), (5, 24)), "c1": (8, 6, (1, 1), (2, 8, 6, 6)),
                                       "c3": (8, 6, (3, 3), (2, 8, 6, 6)), "c5": (5, 7, (5, 5), (2, 5, 7, 9)),
                                       "c2": (4, 4, (2, 2), (1, 4, 5, 5))}.items():
        torch.manual_seed(1)
        m = mi.MP_Conv(cin, cout, k)
        m.eval()
        x = torch.randn(*shape, generator=gen)
        rec[f"mpconv.{tag}.w"] = m.weights.detach().clone()
        rec[f"mpconv.{tag}.x"] = x
        rec[f"mpconv.{tag}.y"] = m(x, gain=0.7)
        m.train()
        y = m(x, gain=0.7)
        rec[f"mpconv.{tag}.y_train"] = y
        rec[f"mpconv.{tag}.w_after"] = m.weights.detach().clone()
    x = torch.randn(2, 3, 8, 8, generator=gen)
    rec["resample.x"] = x
    rec["resample.down"] = mi.resample(x, mode="down")
    rec["resample.up"] = mi.resample(x, mode="up")
    a, b = torch.randn(4, 6, generator=gen), torch.randn(4, 10, generator=gen)
    rec["mp.a"], rec["mp.b"] = a, b
    rec["mp.cat"] = mi.mp_cat(a, b, dim=1, t=0.3)
    rec["mp.sum"] = mi.mp_sum(a, a * 2 + 1, t=0.3)
    rec["mp.silu"] = mi.mp_silu(a)
    rec["mp.normalize"] = mi.normalize(torch.randn(3, 5, 4, 4, generator=gen) + 1, dim=[1])
    np.savez_compressed(os.path.join(OUT, "primitives.npz"), **npify(rec))
    print("primitives ok")


FULL = dict(IN_in_channels=4, IN_img_resolution=32, internal_channels=32, time_emb_dim=64, text_emb_dim=768,
            num_experts=4, top_k=1, Fourier_bandwidth=1.0, VIT_num_blocks=4, VIT_patch_sizes=[4, 8, 8, 16],
            VIT_num_groups=4, VIT_num_heads=8, VIT_emb_size=32, Unet_num_blocks=2, Unet_channel_mult=[1, 2],
            Unet_kernel_sizes=[(3, 3), (3, 3), (5, 5), (5, 5)], Unet_model_channels=32, Unet_channel_mult_emb=2,
            Unet_label_balance=0.5, Unet_concat_balance=0.5, sigma_data=0.5, log_var_channels=32)
FULL_GRAD_KEYS = ("net.input_proj.weights", "net.Unet_router.linear.weights", "net.Unet_router.hard_route.0.weights",
                  "net.vit_router.time_linear.weights", "net.Unet_experts.0.out_gain", "net.Unet_experts.3.out_gain",
                  "net.Unet_experts.1.encoders.32x32_conv.weights",
                  "net.Unet_experts.0.encoders.16x16_block0.conv_res1.weights",
                  "net.Unet_experts.3.decoders.32x32_block2.conv_skip.weights",
                  "net.Unet_experts.2.map_noise.weights", "net.VIT_experts.0.diffit.2.linear2.weights",
                  "net.VIT_experts.2.diffit.0.TMSA.rel_pos_bias", "net.VIT_experts.0.patch.weight",
                  "net.cross_attn.q_proj.weights", "net.cross_attn_text.k_proj.weights", "net.alpha_txt",
                  "net.gate1.weights", "log_var_linear.weights", "net.output_proj.weights")


def full_case(name, variant, B=4, seed=0):
    """One train-mode step of the SHIPPED hyper-parameters (Utils/configs.py:3-35) on the unmodified reference.  Weights are
    NOT stored (9 M parameters): they are the constructor's draw under torch.manual_seed(seed) + the zero-init re-draw,
    which the drop-in modules reproduce (RNG order is part of the boundary); per-tensor norms are stored to verify that.
    Outputs only: the output dict, loss terms, grad.x, the L2 norm of every parameter gradient, a few full gradients and
    a few post-step weights (quirk Q6)."""
    mod = c2 if variant == 2 else c1
    torch.manual_seed(seed)
    model = mod.preconditioned_HDMOEM(**FULL)
    gen = torch.Generator().manual_seed(seed + 100)
    randomize_zero_init(model, gen)
    set_dropout_zero(model)
    model.train()
    names = [n for n, _ in model.named_parameters()]
    w_norms = torch.stack([p.detach().double().norm() for _, p in model.named_parameters()])
    gen = torch.Generator().manual_seed(seed + 200)
    x0 = torch.randn(B, 4, 32, 32, generator=gen) * 0.5
    sigma = torch.exp(torch.randn(B, 1, 1, 1, generator=gen) * 1.6 - 1.2).clamp(0.002, 80.0)
    x = (x0 + sigma * torch.randn(x0.shape, generator=gen)).requires_grad_(True)
    text = torch.randn(B, 8, 768, generator=gen)
    um = torch.ones(B, 4)
    vm = torch.ones(B, 4)
    um[0, 1] = 0.0
    vm[1, 2] = 0.0
    zeta = 0.7
    kw = dict(x=x, sigma=sigma, text_emb=text, Unet_router_mask=um, Vit_router_mask=vm, zeta=zeta, return_log_var=True)
    if variant == 2:
        kw.update(transition_point=-1.2, softness=1.6)
    with NoiseTap() as tap:
        out = model(**kw)
    crit = U.EDM_LOSS(num_experts=4, sigma_data=0.5, Unet_bal=0.05, vit_bal=0.1, z_bal=0.005, prior_bal=0.0)
    loss = crit(sigma_vec=sigma, x=x0, sigma=sigma, out_model=out)
    loss["loss"].backward()
    rec = {"in.x": x, "in.x0": x0, "in.sigma": sigma, "in.text": text, "in.unet_mask": um, "in.vit_mask": vm,
           "in.zeta": zeta, "meta.variant": variant, "meta.top_k": 1, "meta.train": 1, "meta.seed": seed,
           "meta.w_norms": w_norms}
    for k, v in out.items():
        rec["out." + k] = v
    for k, v in loss.items():
        if torch.is_tensor(v):
            rec["loss." + k] = v
    for i, d in enumerate(tap.draws):
        rec[f"noise.{i}"] = d
    rec["meta.n_noise"] = len(tap.draws)
    rec["grad.x"] = x.grad
    named = dict(model.named_parameters())
    rec["gradnorm.all"] = torch.stack([(named[n].grad.double().norm() if named[n].grad is not None
                                        else torch.zeros((), dtype=torch.float64)) for n in names])
    for gk in FULL_GRAD_KEYS:
        g = named[gk].grad
        rec["grad." + gk] = g if g is not None else torch.zeros_like(named[gk])
    for k in ("net.input_proj.weights", "net.Unet_router.hard_route.0.weights", "net.gate2.weights"):
        rec["sd_after." + k] = model.state_dict()[k]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npify(rec))
    print(name, "loss", float(loss["loss"]), "routing", out["Unet_raw"].argmax(1).tolist(), out["vit_raw"].argmax(1).tolist())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(4)
    model_case("cfg2_train_k2", 2, 2, True)
    model_case("cfg2_eval_k1", 2, 1, False)
    model_case("cfg1_train_k1", 1, 1, True)
    model_case("cfg1_eval_k2", 1, 2, False)
    sampler_case("sampler_cfg2_g1", 1.0, 0.0)
    sampler_case("sampler_cfg2_g2_churn", 2.0, 4.0)
    router_tail_case()
    moe_identity_case()
    producers_case()
    primitives_case()
    full_case("full_cfg1_train", 1)
    full_case("full_cfg2_train", 2)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
