"""End-to-end GPU parity at the configuration that is BENCHMARKED (bench.py): TF32 trunk matmuls / convolutions, bf16
grouped tcgen05 U-Net experts, fused ViT blocks, train mode -- against the fp32 CPU oracle and against fixtures produced
by the unmodified reference at the shipped hyper-parameters.

Bars (north_star): routing (top-k indices, expert assignment) identical on every row whose oracle margin exceeds the
stated epsilon; activations, gradients and sampled latents rel-L2 <= 1e-2 at bf16; fp32 paths <= 1e-4 end to end.
Measured values are written to gpurun_out/parity_e2e.json (when that directory exists) for the docs."""
import contextlib
import json
import os

import pytest
import torch

from conftest import FULL, ROOT, load_golden, rel_l2
from oracle import hdmoe_oracle as O

pytestmark = pytest.mark.gpu

TOLBF = 1e-2
MARGIN = 2e-2          # oracle top-1 / top-2 logit margin below which a routing decision is not claimed under bf16 / TF32


def _record(name, **vals):
    d = os.path.join(ROOT, "gpurun_out")
    if not os.path.isdir(d):
        return
    p = os.path.join(d, "parity_e2e.json")
    cur = json.load(open(p)) if os.path.exists(p) else {}
    cur[name] = {k: (float(v) if not isinstance(v, (list, str, int)) else v) for k, v in vals.items()}
    json.dump(cur, open(p, "w"), indent=1)


@contextlib.contextmanager
def bench_numerics():
    """exactly what bench.py sets: TF32 on, bf16 expert path"""
    import hdmoe_b200
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    hdmoe_b200.set_expert_dtype(torch.bfloat16)
    try:
        yield
    finally:
        hdmoe_b200.set_expert_dtype(torch.float32)
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


@contextlib.contextmanager
def strict_fp32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _no_dropout(model):
    for mod in model.modules():        # parity convention: dropout off, exploration noise supplied
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
        if hasattr(mod, "dropout") and not isinstance(mod, torch.nn.Dropout):
            mod.dropout = 0
    return model


def _model(variant, seed=0, res=32):
    from hdmoe_b200 import model_config1, model_config2
    torch.manual_seed(seed)
    model = (model_config2 if variant == 2 else model_config1).preconditioned_HDMOEM(**dict(FULL, IN_img_resolution=res))
    gen = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.abs().max() == 0:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)
    return _no_dropout(model)


def _margin(logits):
    v = torch.sort(torch.nan_to_num(logits, neginf=-1e30), dim=-1, descending=True).values
    return v[:, 0] - v[:, 1]


# ------------------------------------------------------------------------------------------------ reference fixture
@pytest.mark.parametrize("variant", [1, 2])
def test_full_config_fixture_fp32(variant):
    """fp32 modules + kernels on the GPU vs the fixture the UNMODIFIED reference produced at the shipped hyper-parameters
    (tools/make_golden.py full_case): outputs, routing, loss terms, grad.x, gradient norms of all 500+ parameters."""
    from hdmoe_b200.utils import EDM_LOSS
    g = load_golden(f"full_cfg{variant}_train")
    with strict_fp32():
        model = _model(variant, int(g["meta.seed"])).cuda().train()
        names = [n for n, _ in model.named_parameters()]
        w_norms = torch.stack([p.detach().double().norm().cpu() for _, p in model.named_parameters()])
        assert torch.allclose(w_norms, g["meta.w_norms"], rtol=1e-10, atol=0)
        d = [g[f"noise.{i}"].cuda() for i in range(g["meta.n_noise"])]
        noise = {"scaling": d[0], "vit": d[1], "unet": d[2]} if variant == 1 else {"vit": d[0], "unet": d[1]}
        x = g["in.x"].cuda().requires_grad_(True)
        kw = dict(x=x, sigma=g["in.sigma"].cuda(), text_emb=g["in.text"].cuda(), Unet_router_mask=g["in.unet_mask"].cuda(),
                  Vit_router_mask=g["in.vit_mask"].cuda(), zeta=g["in.zeta"], return_log_var=True, noise=noise)
        if variant == 2:
            kw.update(transition_point=-1.2, softness=1.6)
        out = model(**kw)
        errs = {}
        for key in ("denoised", "Unet_router_loss", "vit_router_loss", "scaling_net_out", "out_gate", "log_var"):
            errs[key] = rel_l2(out[key].cpu(), g["out." + key])
            assert errs[key] < (3e-4 if key == "out_gate" else 1e-4), (key, errs[key])
        for key in ("Unet_raw", "vit_raw"):
            assert torch.equal(out[key].cpu().argmax(1), g["out." + key].argmax(1)), key
        crit = EDM_LOSS(num_experts=4, sigma_data=0.5, Unet_bal=0.05, vit_bal=0.1, z_bal=0.005, prior_bal=0.0)
        loss = crit(g["in.sigma"].cuda(), g["in.x0"].cuda(), g["in.sigma"].cuda(), out)
        for key in ("loss", "denoising", "balance", "z_loss", "pure_loss"):
            assert abs(float(loss[key]) - float(g["loss." + key])) < 5e-5 * max(1.0, abs(float(g["loss." + key]))), key
        loss["loss"].backward()
        errs["grad.x"] = rel_l2(x.grad.cpu(), g["grad.x"])
        assert errs["grad.x"] < 5e-4
        named = dict(model.named_parameters())
        gn = torch.stack([(named[n].grad.double().norm().cpu() if named[n].grad is not None
                           else torch.zeros((), dtype=torch.float64)) for n in names])
        ref_gn = g["gradnorm.all"]
        big = ref_gn > 1e-4 * ref_gn.max()
        errs["gradnorm_max_rel"] = float(((gn - ref_gn).abs() / ref_gn.clamp_min(1e-30))[big].max())
        assert errs["gradnorm_max_rel"] < 5e-3
        for k_, v in g.items():
            if k_.startswith("grad.") and k_ != "grad.x":
                got = named[k_[5:]].grad
                got = torch.zeros_like(v) if got is None else got.cpu()
                if float(v.abs().max()) == 0:
                    assert float(got.abs().max()) < 1e-8, k_
                else:
                    # single tensors deep inside an expert (|g| ~ 1e-5, three routed samples) carry the library
                    # kernels' fp32 re-association noise of ~60 stacked layers: 2e-3 measured; the norm of EVERY
                    # parameter gradient is held to 5e-3 above and grad.x to 5e-4
                    assert rel_l2(got, v) < 5e-3, (k_, rel_l2(got, v))
            if k_.startswith("sd_after."):
                assert rel_l2(named[k_[9:]].detach().cpu(), v) < 1e-6, k_
    _record(f"full_fixture_fp32_cfg{variant}", **errs)


# ------------------------------------------------------------------------------------------------ benchmarked numerics
def _oracle_sd(model):
    return {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point and not k.endswith(("freqs", "phases")))
            for k, v in model.state_dict().items()}


@pytest.mark.parametrize("variant", [1, 2])
def test_train_step_bench_numerics_vs_fp32_oracle(variant):
    """One train step at bench.py's numeric configuration (allow_tf32, bf16 grouped tcgen05 experts, fused ViT kernels,
    band masks of MaskGenerator at step 0, B = 64, dropout 0, exploration noise supplied) against the fp32 oracle:
    routing identical, denoised / loss / grad.x / the concatenated gradient of ALL parameters within 1e-2."""
    from hdmoe_b200.utils import EDM_LOSS, MaskGenerator
    B, zeta = 64, 0.5
    model = _model(variant)
    gen = torch.Generator().manual_seed(4321)
    x0 = torch.randn(B, 4, 32, 32, generator=gen) * 0.5
    sigma = torch.exp(torch.randn(B, 1, 1, 1, generator=gen) * 1.6 - 1.2).clamp(0.002, 80.0)
    x = x0 + sigma * torch.randn(x0.shape, generator=gen)
    text = torch.randn(B, 77, 768, generator=gen)
    mk = dict(p_mean=-1.2, p_std=1.6, bandwidth=0.3, max_bandwidth=0.8, min_active=1, total_steps=5000, step_size=0.1,
              strat_band="step")
    um = MaskGenerator([3, 3, 5, 5], noise_range=(0.0, 0.6), **mk)(sigma, 0)
    vm = MaskGenerator([4, 8, 8, 16], noise_range=(0.4, 1.0), **mk)(sigma, 0)
    noise = {"scaling": torch.randn(B, 2, generator=gen), "vit": torch.randn(B, 4, generator=gen),
             "unet": torch.randn(B, 4, generator=gen)}

    def oracle(grad):
        sd = _oracle_sd(model)
        xr = x.clone().requires_grad_(grad)
        with (contextlib.nullcontext() if grad else torch.no_grad()):
            with O.training_mode():
                out = O.preconditioned(sd, FULL, xr, sigma, text, um, vm, zeta, -1.2, 1.6, return_log_var=True,
                                       noise=noise, variant=variant)
        return sd, xr, out

    # keep every routing decision away from a tie: rows whose oracle margin is small get their winner's noise raised
    # (the decision itself is tested bit-exactly in test_gpu_parity.py; here a flipped row would turn the gradient
    # comparison into a comparison of different experts)
    _, _, out0 = oracle(False)
    for key, nk in (("Unet_raw", "unet"), ("vit_raw", "vit")):
        lg = out0[key]
        small = _margin(lg) < 10 * MARGIN
        noise[nk][small, lg.argmax(1)[small]] += 1.0
    sd, xr, ref = oracle(True)
    assert float(_margin(ref["Unet_raw"]).min()) > MARGIN and float(_margin(ref["vit_raw"]).min()) > MARGIN
    loss_ref = O.edm_loss(x0, ref, 4, 0.05, 0.1, 0.005)["loss"]
    loss_ref.backward()

    with bench_numerics():
        model.cuda().train()
        crit = EDM_LOSS(num_experts=4, sigma_data=0.5, Unet_bal=0.05, vit_bal=0.1, z_bal=0.005, prior_bal=0.0)
        xd = x.cuda().requires_grad_(True)
        kw = dict(transition_point=-1.2, softness=1.6) if variant == 2 else {}
        out = model(x=xd, sigma=sigma.cuda(), text_emb=text.cuda(), Unet_router_mask=um.cuda(), Vit_router_mask=vm.cuda(),
                    zeta=zeta, return_log_var=True, noise={k: v.cuda() for k, v in noise.items()}, **kw)
        loss = crit(sigma.cuda(), x0.cuda(), sigma.cuda(), out)["loss"]
        loss.backward()
        torch.cuda.synchronize()
    for rn, key in (("Unet_router", "Unet_raw"), ("vit_router", "vit_raw")):
        got = getattr(model.net, rn).last["topk_idx"].cpu().long().flatten()
        assert torch.equal(got, ref[key].argmax(1)), f"{rn}: routing differs from the oracle"
        fin = torch.isfinite(ref[key])
        assert torch.equal(torch.isfinite(out[key].cpu()), fin)
        assert rel_l2(out[key].cpu()[fin], ref[key][fin]) < TOLBF
        logit_err = max(locals().get("logit_err", 0.0), float((out[key].cpu()[fin] - ref[key][fin]).abs().max()))
    e = {"logits_max_abs": logit_err, "denoised": rel_l2(out["denoised"].cpu(), ref["denoised"]),
         "out_gate": rel_l2(out["out_gate"].cpu(), ref["out_gate"]),
         "loss_abs": abs(float(loss) - float(loss_ref)), "grad_x": rel_l2(xd.grad.cpu(), xr.grad)}
    num = den = 0.0
    worst = ("", 0.0)
    per_group = {}
    for n, p in model.named_parameters():
        g_ref = sd[n].grad
        g_ref = torch.zeros_like(sd[n]) if g_ref is None else g_ref
        g_got = torch.zeros_like(g_ref) if p.grad is None else p.grad.detach().float().cpu()
        dn, dd = float(((g_got - g_ref).double() ** 2).sum()), float((g_ref.double() ** 2).sum())
        num += dn
        den += dd
        grp = n.split(".")[1] if n.startswith("net.") else n.split(".")[0]
        a = per_group.setdefault(grp, [0.0, 0.0])
        a[0] += dn
        a[1] += dd
    e["grad_params"] = (num / den) ** 0.5
    e["grad_by_module"] = str({k: round((v[0] / v[1]) ** 0.5, 5) if v[1] > 0 else 0.0 for k, v in per_group.items()})
    _record(f"train_step_bench_numerics_cfg{variant}", **e)
    assert e["denoised"] < TOLBF, e
    assert e["loss_abs"] < TOLBF * max(1.0, abs(float(loss_ref))), e
    assert e["grad_x"] < TOLBF, e
    assert e["grad_params"] < TOLBF, e


# ------------------------------------------------------------------------------------------------ sampler
SAMPLER_MARGIN = 5e-3   # eval-mode logits (zeta = 0) of the random-init routers sit close together; the measured logit
                        # error of the bf16 / TF32 path is ~1e-3 (recorded by the train-step test)


@pytest.mark.parametrize("guidance,B,steps", [(1.0, 64, 6), (2.0, 32, 4)])
def test_sampler_bf16_graph_teacher_forced_vs_oracle(guidance, B, steps):
    """EDM Heun sampler in the benchmarked mode (bf16 experts, TF32 trunk, one CUDA graph per denoiser evaluation) against
    the fp32 oracle sampler.  (i) Teacher-forced: at every NFE both sides evaluate the denoiser on the ORACLE's state;
    each network output (conditional and, with guidance, unconditional) over the rows whose routing margin exceeds
    SAMPLER_MARGIN: rel-L2 <= 1e-2 for sigma < 1 and <= 3e-2 for sigma >= 1.  Why two bars: at large sigma D(x) is the raw
    network output (c_skip -> 0) and the U-Net branch input is scaled by s_unet = 2(1 - w + 0.01) ~ 0.1, so inside the
    U-Net experts the signal rides on the response to the constant ones channel and every bf16 activation rounding is
    amplified by ~1 / s_unet; the fp32 GPU path itself differs from the fp32 oracle by 1.7e-3 there (library summation
    order, tools/dbg_numerics.py), bf16 by 0.8 - 2.5e-2.  The guided combination ref.lerp(cond, g) = g*cond + (1-g)*ref
    amplifies both errors by up to |g| + |1-g|.  (ii) Free-running: the final latents of our own trajectory over the
    rows that never were inside the margin, rel-L2 <= 1e-2 (measured 2e-3 / 4e-3)."""
    from hdmoe_b200 import EDM_Sampler
    model = _model(2, seed=1)
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    gen = torch.Generator().manual_seed(77)
    noise = torch.randn(B, 4, 32, 32, generator=gen)
    text = torch.randn(B, 77, 768, generator=gen)
    uncond = torch.zeros_like(text)
    trace = []
    ones = torch.ones(B, 4)
    guided = guidance != 1.0

    def ofn(xs, s):
        with torch.no_grad():
            o = O.preconditioned(sd, FULL, xs, s, text, ones, ones, 0.0, -1.2, 1.6, variant=2)
            ok = (_margin(o["Unet_raw"]) > SAMPLER_MARGIN) & (_margin(o["vit_raw"]) > SAMPLER_MARGIN)
            d_c, d_u, d = o["denoised"], None, o["denoised"]
            if guided:
                o2 = O.preconditioned(sd, FULL, xs, s, uncond, ones, ones, 0.0, -1.2, 1.6, variant=2)
                ok &= (_margin(o2["Unet_raw"]) > SAMPLER_MARGIN) & (_margin(o2["vit_raw"]) > SAMPLER_MARGIN)
                d_u = o2["denoised"]
                d = O.cfg_denoise(d_c, d_u, guidance)
        trace.append((xs.clone(), s.clone(), d_c.clone(), None if d_u is None else d_u.clone(), d.clone(), ok))
        return d

    ref = O.edm_sample(ofn, noise, num_steps=steps)
    assert len(trace) == 2 * steps - 1
    with bench_numerics():
        model.cuda().eval()
        smp = EDM_Sampler(model, model, num_solve_steps=steps, guidance=guidance, use_cuda_graph=True)
        plain = EDM_Sampler(model, model, num_solve_steps=steps, guidance=1.0, use_cuda_graph=True)
        tx, un = text.cuda(), uncond.cuda()
        e_c, e_u, e_g, kept, bars = [], [], [], [], []
        for xs, s, d_c, d_u, d, ok in trace:
            kept.append(float(ok.float().mean()))
            bars.append(3e-2 if float(s) >= 1.0 else TOLBF)
            xd, sdv = xs.cuda(), s.cuda()
            e_c.append(rel_l2(plain.denoise(xd, sdv, tx, -1.2, 1.6).float().cpu()[ok], d_c[ok]))
            if guided:
                e_u.append(rel_l2(plain.denoise(xd, sdv, un, -1.2, 1.6).float().cpu()[ok], d_u[ok]))
                e_g.append(rel_l2(smp.denoise(xd, sdv, tx, -1.2, 1.6, uncond_text_emb=un).float().cpu()[ok], d[ok]))
        out = smp.sample(noise.cuda(), tx, -1.2, 1.6, uncond_text_emb=un).cpu()
    ok_all = torch.stack([t[5] for t in trace]).all(0)
    final = rel_l2(out[ok_all], ref[ok_all])
    _record(f"sampler_bf16_graph_g{guidance}", per_nfe_cond=[round(v, 5) for v in e_c],
            per_nfe_uncond=[round(v, 5) for v in e_u], per_nfe_guided=[round(v, 5) for v in e_g], rows_kept=kept,
            final_latents=final, rows_final=float(ok_all.float().mean()))
    assert min(kept) > 0.5, kept
    amp = abs(guidance) + abs(1 - guidance)
    assert all(e < b for e, b in zip(e_c, bars)), (e_c, bars)
    if guided:
        assert all(e < b for e, b in zip(e_u, bars)), (e_u, bars)
        assert all(e < b * amp for e, b in zip(e_g, bars)), (e_g, bars)
    assert torch.isfinite(out).all()
    assert ok_all.float().mean() > 0.3 and final < TOLBF, final


def test_sampler_graph_cache_keyed_on_python_scalars():
    """A second sample() on the same sampler with another transition_mean / softness must not replay the first call's
    captured constants (they are baked into the recorded kernels)."""
    from hdmoe_b200 import EDM_Sampler
    model = _model(2, seed=2)
    gen = torch.Generator().manual_seed(5)
    noise = torch.randn(4, 4, 32, 32, generator=gen).cuda()
    text = torch.randn(4, 77, 768, generator=gen).cuda()
    with bench_numerics():
        model.cuda().eval()
        g = EDM_Sampler(model, model, num_solve_steps=3, use_cuda_graph=True)
        e = EDM_Sampler(model, model, num_solve_steps=3)
        a1, a2 = g.sample(noise, text, -1.2, 1.6), g.sample(noise, text, 0.7, 0.4)
        b1, b2 = e.sample(noise, text, -1.2, 1.6), e.sample(noise, text, 0.7, 0.4)
        a3 = g.sample(noise, text, -1.2, 1.6)
    assert rel_l2(a1, b1) < 1e-5 and rel_l2(a2, b2) < 1e-5 and torch.equal(a1, a3)
    assert rel_l2(b1, b2) > 1e-3          # the two settings really differ
    assert len(g._graphs) == 2


# ------------------------------------------------------------------------------------------------ autograd contracts
def test_eval_mode_autograd_reaches_expert_parameters():
    """Eval-mode forward under autograd: the grouped bf16 path propagates to the input and to the U-Net expert parameters
    like the per-expert path (the reference's composite does); only the forced weight norm and dropout depend on
    .training."""
    import hdmoe_b200
    model = _model(2, seed=3)
    B = 8
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(B, 4, 32, 32, generator=gen)
    sigma = torch.exp(torch.randn(B, 1, 1, 1, generator=gen) * 1.6 - 1.2).clamp(0.002, 80.0)
    text = torch.randn(B, 77, 768, generator=gen)
    ones = torch.ones(B, 4).cuda()
    res = {}
    with bench_numerics():
        model.cuda().eval()
        w_before = model.net.Unet_experts[0].out_conv.weights.detach().clone()
        try:
            for mode in (True, False):
                hdmoe_b200.set_grouped_experts(mode)
                model.zero_grad(set_to_none=True)
                xin = x.cuda().requires_grad_(True)
                out = model(x=xin, sigma=sigma.cuda(), text_emb=text.cuda(), Unet_router_mask=ones, Vit_router_mask=ones,
                            zeta=0, transition_point=-1.2, softness=1.6)["denoised"]
                out.square().mean().backward()
                res[mode] = (xin.grad.float().cpu(), {n: p.grad.detach().float().cpu() for n, p in model.named_parameters()
                                                      if "Unet_experts" in n and p.grad is not None})
        finally:
            hdmoe_b200.set_grouped_experts(True)
        assert torch.equal(w_before, model.net.Unet_experts[0].out_conv.weights.detach())     # eval: no forced weight norm
    gx_g, gp_g = res[True]
    gx_l, gp_l = res[False]
    assert len(gp_g) >= len(gp_l) > 100
    assert rel_l2(gx_g, gx_l) < 2e-2
    a = torch.cat([gp_g[n].flatten() for n in gp_l])
    b = torch.cat([gp_l[n].flatten() for n in gp_l])
    assert float(b.norm()) > 0 and rel_l2(a, b) < 2e-2


def test_second_forward_before_backward_is_rejected():
    """The grouped / prepared-weight paths hold the buffers of ONE forward: a backward that belongs to an older forward
    raises instead of silently using the newer forward's plan and a wiped gradient accumulator."""
    model = _model(2, seed=4)
    B = 4
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(B, 4, 32, 32, generator=gen).cuda()
    sigma = torch.full((B, 1, 1, 1), 0.7).cuda()
    text = torch.randn(B, 77, 768, generator=gen).cuda()
    ones = torch.ones(B, 4).cuda()
    with bench_numerics():
        model.cuda().train()
        kw = dict(sigma=sigma, text_emb=text, Unet_router_mask=ones, Vit_router_mask=ones, zeta=0.1, transition_point=-1.2,
                  softness=1.6)
        l1 = model(x=x, **kw)["denoised"].square().mean()
        l2 = model(x=x * 0.5, **kw)["denoised"].square().mean()
        with pytest.raises(RuntimeError, match="one backward per forward"):
            l1.backward()
        model.zero_grad(set_to_none=True)
        l3 = model(x=x, **kw)["denoised"].square().mean()
        l3.backward()                                         # the normal order still works afterwards
        assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    del l2


def test_two_models_with_different_pinned_options_coexist():
    """An fp32 model and a bf16 (grouped tcgen05) model in one process, options pinned per model
    (hdmoe_b200.set_model_options): each gives exactly what the process-wide switch gives when it runs alone, in either
    call order, and the process default is untouched."""
    import hdmoe_b200
    m32, m16 = _model(2, seed=3).cuda().eval(), _model(2, seed=3).cuda().eval()
    gen = torch.Generator().manual_seed(5)
    B = 8
    x = (torch.randn(B, 4, 32, 32, generator=gen) * 0.5).cuda()
    sigma = torch.exp(torch.randn(B, 1, 1, 1, generator=gen) * 1.6 - 1.2).clamp(0.002, 80.0).cuda()
    text = torch.randn(B, 77, 768, generator=gen).cuda()
    ones = torch.ones(B, 4).cuda()

    def run(m):
        with torch.no_grad():
            return m(x=x, sigma=sigma, text_emb=text, Unet_router_mask=ones, Vit_router_mask=ones, zeta=0,
                     transition_point=-1.2, softness=1.6)["denoised"].float()

    with strict_fp32():
        assert hdmoe_b200.get_expert_dtype() == torch.float32
        want32 = run(m32)
        hdmoe_b200.set_expert_dtype(torch.bfloat16)
        want16 = run(m16)
        hdmoe_b200.set_expert_dtype(torch.float32)
        assert not torch.equal(want32, want16)
        hdmoe_b200.set_model_options(m16, expert_dtype=torch.bfloat16)
        hdmoe_b200.set_model_options(m32, expert_dtype=torch.float32)
        for order in ((m16, m32), (m32, m16)):
            got = {id(m): run(m) for m in order}
            assert torch.equal(got[id(m32)], want32) and torch.equal(got[id(m16)], want16)
        assert hdmoe_b200.get_expert_dtype() == torch.float32
