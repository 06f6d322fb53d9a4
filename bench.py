#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native heterogeneous-MoE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload train|sample]

Workload at N=1 (BASELINE.json configs[1]): model_config1 training step, bf16 expert path, batch 256 per GPU,
4x32x32 VAE-shaped latents, CLIP-shaped random text embeddings (256,77,768), random-init weights (zero-init
parameters re-drawn from N(0,0.3^2)).  One step = forward(return_log_var=True) + EDM_LOSS + backward +
clip_grad_norm_(1.0) + AdamW over ALL parameters.  N>1: one process per GPU
(torchrun), data-parallel replicas of the step with an NCCL gradient all-reduce; weak scaling.

Prints ONE JSON line (see the repo brief for the contract): value = whole-job img/s with inputs resident in
HBM; e2e = the same metric through the public module API with pinned HOST inputs, H2D copies and the D2H loss
read inside the timed region; roofline = the dominant hand-written kernel timed live with CUDA events;
cpu_baseline = the unmodified reference (oracle/_ref; the oracle port if that is absent) timed on this box's host cores
on a bounded sample.  `--impl reference` times only that CPU arm, on all host threads, and prints the same line shape.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FULL = dict(IN_in_channels=4, IN_img_resolution=32, internal_channels=32, time_emb_dim=64, text_emb_dim=768,
            num_experts=4, top_k=1, Fourier_bandwidth=1.0, VIT_num_blocks=4, VIT_patch_sizes=[4, 8, 8, 16],
            VIT_num_groups=4, VIT_num_heads=8, VIT_emb_size=32, Unet_num_blocks=2, Unet_channel_mult=[1, 2],
            Unet_kernel_sizes=[(3, 3), (3, 3), (5, 5), (5, 5)], Unet_model_channels=32, Unet_channel_mult_emb=2,
            Unet_label_balance=0.5, Unet_concat_balance=0.5, sigma_data=0.5, log_var_channels=32)
LOSS = dict(num_experts=4, sigma_data=0.5, Unet_bal=0.05, vit_bal=0.1, z_bal=0.005, prior_bal=0.0)
MASK = dict(p_mean=-1.2, p_std=1.6, bandwidth=0.3, max_bandwidth=0.8, min_active=1, total_steps=5000, step_size=0.1,
            strat_band="step")
P_MEAN, P_STD = -1.2, 1.6
SEED = 1234


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


def synth_batch(B, res, rank, device, pinned=False):
    """Synthetic step inputs on the HOST (seed 1234 + rank): latents, sigma, noised latents, text, masks."""
    from hdmoe_b200.utils import MaskGenerator, sample_sigma_hybrid
    gen = torch.Generator().manual_seed(SEED + rank)
    x0 = torch.randn(B, 4, res, res, generator=gen) * 0.5
    sigma = sample_sigma_hybrid(B, 0.002, 80.0, p_mean=P_MEAN, p_std=P_STD, extreme_prob=0.5, device="cpu",
                                generator=gen)
    x = x0 + sigma * torch.randn(x0.shape, generator=gen)
    text = torch.randn(B, 77, 768, generator=gen)
    um = MaskGenerator([3, 3, 5, 5], noise_range=(0.0, 0.6), **MASK)(sigma, 0)
    vm = MaskGenerator([4, 8, 8, 16], noise_range=(0.4, 1.0), **MASK)(sigma, 0)
    batch = dict(x0=x0, sigma=sigma, x=x, text=text, um=um, vm=vm)
    if pinned:
        batch = {k: v.pin_memory() for k, v in batch.items()}
    return batch


def build_model(variant, device, seed=0):
    import hdmoe_b200
    mod = hdmoe_b200.model_config1 if variant == 1 else hdmoe_b200.model_config2
    torch.manual_seed(seed)
    model = mod.preconditioned_HDMOEM(**FULL)
    gen = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for p in model.parameters():
            if float(p.abs().max()) == 0:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)
    return model.to(device)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(v, world, device):
    if world == 1:
        return v
    import torch.distributed as dist
    t = torch.tensor([v], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class L2Flusher:
    def __init__(self, device):
        self.buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=device)   # > 126 MB L2

    def __call__(self):
        self.buf.add_(1)


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
class TrainRunner:
    """One training step of preconditioned_HDMOEM (variant 1 / 2) at `res` x `res`, batch B per GPU: forward
    (return_log_var) + EDM_LOSS + backward + gradient all-reduce (world > 1) + gradient-norm clip 1.0 + AdamW, recorded
    as CUDA graph(s) when possible.  parallelism: "dp" = replicas; "ep" = U-Net experts sharded over the ranks with the
    static-shape all-to-all layer (expert_parallel.py), trunk data-parallel."""

    def __init__(self, variant, res, B, rank, world, device, parallelism="dp", use_graph=True, warmup=3, pinned=False,
                 ep_transport="peer", ep_capacity="auto"):
        import hdmoe_b200
        from hdmoe_b200.optim import FusedAdamW
        from hdmoe_b200.utils import EDM_LOSS
        self.world, self.device, self.B = world, device, B
        self.ep = parallelism == "ep" and world > 1
        self.ep_degree = 1
        if self.ep:
            grp, self.ep_degree = ep_group(rank, world)
            cap = ep_capacity_for(self.ep_degree) if ep_capacity == "auto" else ep_capacity
            self.ep_capacity = cap
            hdmoe_b200.enable_expert_parallel([3, 3, 5, 5], group=grp, capacity_factor=cap, transport=ep_transport)
        else:
            hdmoe_b200.disable_expert_parallel()
        torch.manual_seed(0)
        mod = hdmoe_b200.model_config1 if variant == 1 else hdmoe_b200.model_config2
        model = mod.preconditioned_HDMOEM(**dict(FULL, IN_img_resolution=res))
        gen = torch.Generator().manual_seed(100)
        with torch.no_grad():
            for p in model.parameters():
                if float(p.abs().max()) == 0:
                    p.copy_(torch.randn(p.shape, generator=gen) * 0.3)
        self.model = model.to(device).train()
        crit = EDM_LOSS(**LOSS)
        params = self.params = list(self.model.parameters())
        # clip_grad_norm_(1.0) + AdamW(lr 5e-4) of the reference loop, as the three-launch multi-tensor kernel set
        opt = self.opt = FusedAdamW(params, lr=5e-4, max_grad_norm=1.0)
        flat_sizes = [p.numel() for p in params]
        self.host = synth_batch(B, res, rank, device, pinned=pinned)
        self.dev_batch = {k: v.to(device) for k, v in self.host.items()}
        kw = dict(transition_point=P_MEAN, softness=P_STD) if variant == 2 else {}

        def fwd_bwd(b):
            out = self.model(x=b["x"], sigma=b["sigma"], text_emb=b["text"], Unet_router_mask=b["um"],
                             Vit_router_mask=b["vm"], zeta=2.0, return_log_var=True, **kw)
            loss = crit(b["sigma"], b["x0"], b["sigma"], out)["loss"]
            opt.zero_grad(set_to_none=True)
            loss.backward()
            return loss

        def step(b):
            loss = fwd_bwd(b)
            if world > 1:
                import torch.distributed as dist
                grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
                flat = torch.cat([g.reshape(-1) for g in grads])
                dist.all_reduce(flat)
                flat.div_(world)
                for p, g in zip(params, flat.split(flat_sizes)):
                    p.grad = g.view_as(p)
            opt.step()                                # gradient-norm clip (max 1.0) fused into the optimizer launches
            return loss

        self.eager_step = step
        self.graphed, self.note = None, "eager"
        if use_graph:
            from hdmoe_b200.train_step import GraphedTrainStep
            # The capture (and its own warm-up iterations) must be the FIRST thing that touches autograd: AccumulateGrad
            # nodes created on the default stream by an earlier eager step would tie the legacy stream to the capture.
            if world == 1:
                self.graphed = GraphedTrainStep(step, self.dev_batch, warmup=max(3, warmup)).capture()
                self.note = "cuda_graph (whole step: fwd+loss+bwd+clip+AdamW)"
            else:
                # N > 1: two graphs around ONE eager NCCL launch.  Graph A = forward + loss + backward (+ the expert-
                # parallel all-to-alls in "ep" mode, captured) + flatten of all gradients into a static buffer; the
                # all-reduce of that buffer is the only eager launch; graph B = mean, clip and AdamW on views of it.
                import torch.distributed as dist
                flat = torch.zeros(sum(flat_sizes), device=device)

                def fwd_bwd_flat(b):
                    loss = fwd_bwd(b)
                    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
                    torch.cat([g.reshape(-1) for g in grads], out=flat)
                    return loss

                def update(_b):
                    flat.div_(world)
                    for p, g in zip(params, flat.split(flat_sizes)):
                        p.grad = g.view_as(p)
                    opt.step()
                    return flat

                inner = GraphedTrainStep(fwd_bwd_flat, self.dev_batch, warmup=max(3, warmup)).capture()
                upd = GraphedTrainStep(update, {}, warmup=1).capture()

                class _Outer:
                    static = inner.static
                    prefetch, commit = inner.prefetch, inner.commit
                    enqueue_loss_read, result = inner.enqueue_loss_read, inner.result

                    def __call__(self, batch=None):
                        loss = inner(batch)
                        dist.all_reduce(flat)
                        upd(None)
                        return loss

                self.graphed = _Outer()
                self.note = ("cuda_graph A (fwd+loss+bwd%s+flatten) + eager NCCL all-reduce + cuda_graph B (clip+AdamW)"
                             % (" incl. expert-parallel all-to-alls" if self.ep else ""))
        else:
            for _ in range(warmup):
                step(self.dev_batch)

    def run(self):
        return self.graphed(None) if self.graphed is not None else self.eager_step(self.dev_batch)

    def time_steps(self, steps, flush=None):
        """device time of `steps` steps (CUDA events per step, L2 flushed between them outside the events), ms total"""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s_, e_ in ev:
            if flush is not None:
                flush()
            s_.record()
            self.run()
            e_.record()
        torch.cuda.synchronize()
        return sum(s_.elapsed_time(e_) for s_, e_ in ev)


def ep_capacity_for(degree):
    """Rows a rank's experts may receive, in units of T*k (None = exact worst case degree*T*k); an overflow is detected on
    the device and the measurement is repeated at the worst case.  Every fixed-shape kernel of the expert path runs over
    the capacity, so it is sized for the imbalance a skewed router produces (the busiest expert drawing ~45 % of all rows;
    the measured figure is reported as recv_rows), not for the worst case."""
    return min(float(degree), max(1.5, 0.45 * degree))


_EP_GROUPS = {}


def ep_group(rank, world):
    """(process group, degree) of the expert-parallel exchange: the four U-Net experts shard over at most 4 ranks; with 8
    ranks the job is 2 expert-parallel groups of 4 (EP x DP = 4 x 2), the trunk and the gradient all-reduce stay
    over all ranks."""
    if world <= 4 or world % 4:
        return None, world
    import torch.distributed as dist
    if world not in _EP_GROUPS:
        _EP_GROUPS[world] = [dist.new_group(list(range(i, i + 4))) for i in range(0, world, 4)]   # every rank creates all
    return _EP_GROUPS[world][rank // 4], 4


def make_runner(args, variant, res, B, rank, world, device, parallelism, pinned=False):
    """TrainRunner; data-parallel steps are recorded as CUDA graphs.  The expert-parallel step has static shapes and no
    host synchronisation.  With the peer-memory transport (default; csrc/peer.cu) its exchange is plain kernels and the
    step records like the data-parallel one.  With --ep-transport nccl the all-to-alls are NCCL calls: replaying those
    from inside the graph deadlocked on this torch 2.11 / NCCL 2.28 stack (2 x B200, round-2 log), so that variant runs
    eagerly unless --ep-graph is given."""
    ep = parallelism == "ep" and world > 1
    transport = getattr(args, "ep_transport", "peer")
    use_graph = (not args.no_graph) and (not ep or transport == "peer" or args.ep_graph)
    r = TrainRunner(variant, res, B, rank, world, device, parallelism, use_graph=use_graph, warmup=args.warmup, pinned=pinned,
                    ep_transport=transport, ep_capacity=getattr(args, "ep_capacity", "auto"))
    if ep:
        r.note += " [expert-parallel degree %d, exchange: %s, capacity %s x T*k rows]" % (
            r.ep_degree, "peer-memory pull kernels + device barrier" if transport == "peer" else "NCCL all_to_all_single",
            r.ep_capacity)
    return r


def run_ours(args):
    import hdmoe_b200
    from hdmoe_b200 import _lib
    rank, world, local = dist_setup(args.gpus)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    _lib.lib()     # fail loudly if the CUDA library is missing
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    hdmoe_b200.set_expert_dtype(torch.bfloat16)
    B = args.batch
    flush = L2Flusher(device)
    graphed, graph_note = None, "eager"
    try:
        runner = make_runner(args, 1, 32, B, rank, world, device, args.parallelism, pinned=True)
    except Exception:                            # noqa: BLE001
        if world > 1 or args.no_graph:
            raise
        sys.stderr.write("bench.py: CUDA-graph capture failed, restarting in eager mode\n")
        sys.stderr.flush()
        os.execv(sys.executable, [sys.executable] + sys.argv + ["--no-graph"])
    graphed, graph_note = runner.graphed, runner.note
    host, dev_batch, step = runner.host, runner.dev_batch, runner.eager_step
    barrier(world)
    run_step = (lambda b: graphed(b)) if graphed is not None else step
    static_in = graphed.static if graphed is not None else dev_batch
    for _ in range(2):
        run_step(None if graphed is not None else dev_batch)
    barrier(world)

    # ---- device-resident timing: EXACTLY K steps back to back between two CUDA events (barrier + synchronize on both
    # sides), the steady state of a training loop.  No L2 flush is needed between steps: one step streams several GB of
    # activations and 3 x 36 MB of parameter / optimizer state through the 126 MB L2, nothing survives from one step to
    # the next.  (`ms_per_step_isolated` repeats the measurement with a 256 MiB L2 flush before every step and one event
    # pair per step -- it adds the graph-launch latency of a cold start, ~0.4 ms, that back-to-back steps hide.)
    clocks = ClockSampler(local)
    l0 = _lib.launch_count()
    barrier(world)
    s_all, e_all = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_all.record()
    for i in range(args.steps):
        if args.profile_step and i == 0:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()      # ncu --profile-from-start off captures exactly one timed step
        run_step(None if graphed is not None else dev_batch)
        if args.profile_step and i == 0:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
    e_all.record()
    barrier(world)
    launches = _lib.launch_count() - l0
    ms_total = max_over_ranks(s_all.elapsed_time(e_all), world, device)
    clk = clocks.stop()
    ms_step = ms_total / args.steps
    value = B * world / (ms_step / 1e3)
    if args.profile_step:
        # profiler pass (ncu --profile-from-start off): the captured step is all that is wanted; a number printed under a
        # profiler is never a bench value, so nothing else is measured
        if rank == 0:
            print(json.dumps({"profile_step": True, "ms_per_step_under_profiler": round(ms_step, 3),
                              "execution": graph_note, "gpu_launches": int(launches)}), flush=True)
        return None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 5))]
    for s, e in ev:
        flush()
        s.record()
        run_step(None if graphed is not None else dev_batch)
        e.record()
    barrier(world)
    ms_isolated = max_over_ranks(sum(s.elapsed_time(e) for s, e in ev), world, device) / len(ev)
    if graphed is not None:
        # graph replays do not pass through the C ABI: count the hand-written launches of ONE recorded step
        l1 = _lib.launch_count()
        step(dev_batch)
        torch.cuda.synchronize()
        launches = (_lib.launch_count() - l1) * args.steps

    # ---- end-to-end: pinned host inputs -> H2D -> step -> D2H loss, all inside the timed region
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    e2e_steps = max(3, args.steps)     # K steps as well; the first step's H2D copy (pipeline fill) cannot overlap anything
    if graphed is not None:
        graphed.prefetch(host)                       # untimed: allocates the staging set
        graphed.commit()
        torch.cuda.synchronize()
    barrier(world)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    loss_host = 0.0
    if graphed is not None:
        graphed.prefetch(host)                       # step 0's H2D copy (inside the timed region)
    pending = None
    for i in range(e2e_steps):
        if graphed is not None:
            # every step's inputs come from pinned host memory; step i+1's H2D copy overlaps step i's replay.  Every
            # step's loss is read on the host (asynchronous D2H into pinned memory): the read of step i completes while
            # step i+1 runs, as a training loop that logs its loss one step late does; the last one before the end event.
            graphed.commit()
            if i + 1 < e2e_steps:
                graphed.prefetch(host)
            graphed(None)                            # replay
            h = graphed.enqueue_loss_read()
            if pending is not None:
                loss_host = graphed.result(pending)
            pending = h
        else:
            b = {k: v.to(device, non_blocking=True) for k, v in host.items()}
            loss_host = float(step(b).item())        # D2H read of the step result
    if pending is not None:
        loss_host = graphed.result(pending)
    e.record()
    barrier(world)
    e2e_ms = max_over_ranks(s.elapsed_time(e), world, device) / e2e_steps
    e2e_value = B * world / (e2e_ms / 1e3)

    # ---- the other configurations of BASELINE.json at this N (all ranks take part): EDM sampler (configs[3], batch
    # 1024 split over the ranks) and model_config2 at 64x64 (configs[2], batch 64 per GPU), data-parallel and expert-parallel
    def make_line(roof, cpu, disp, ref_gpu, peaks_src, samp, cfg_c):
        return {"metric": "denoiser train img/s", "value": round(value, 2), "unit": "img/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 3),
                "ms_per_step_isolated": round(ms_isolated, 3),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic",
                "config": {"workload": "model_config1 train step (fwd+EDM_LOSS+bwd+clip+AdamW), batch 256/GPU, "
                                       "4x32x32 latent, text (B,77,768), bf16 expert path, fp32 trunk (TF32 matmul)",
                           "global_batch": B * world,
                           "parallelism": (f"ep{world} (U-Net experts) + dp{world} (trunk)" if args.parallelism == "ep" and world > 1
                                           else f"dp{world}"),
                           "l2": "no flush: K steps back to back, per-step working set (GBs of activations) >> 126 MB L2",
                           "execution": graph_note},
                "clocks": clk,
                "e2e": {"value": round(e2e_value, 2), "unit": "img/s", "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": 4, "ms_per_step": round(e2e_ms, 3), "last_loss": loss_host,
                        "d2h": "the loss of EVERY step is copied to pinned host memory and read there; the read of step i "
                               "completes while step i+1 runs (the last one inside the timed region)"},
                "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu, "peaks": peaks_src,
                "dispatch": disp, "sampler": samp, "config_c": cfg_c, "ref_cuda_eager": ref_gpu}

    samp = cfg_c = None
    if not args.no_sampler:
        del runner, graphed, run_step, step
        torch.cuda.empty_cache()
        # The extras exercise multi-rank code paths (expert-parallel exchange over peer memory, NCCL sub-groups) that must
        # never cost the headline measurement: if they do not finish in time every rank leaves, rank 0 first printing the
        # line with what was measured so far.
        progress = {"sampler": None, "config_c": None}

        def bail():
            if rank == 0:
                progress["timeout"] = "extras did not finish within %d s; partial" % EXTRAS_TIMEOUT_S
                print(json.dumps(make_line(None, None, None, None, "n/a", progress.get("sampler"),
                                           dict(progress.get("config_c") or {}, error=progress["timeout"]))), flush=True)
            os._exit(0)

        import threading
        wd = threading.Timer(EXTRAS_TIMEOUT_S, bail)
        wd.daemon = True
        wd.start()
        samp, cfg_c = scale_extras(args, rank, world, device, progress)
        wd.cancel()
    # every collective is done: release the other ranks before rank 0 measures the single-GPU extras (they must not
    # sit in an NCCL call for a minute while rank 0 runs the kernel tables and the CPU baseline)
    if world > 1:
        import torch.distributed as dist
        barrier(world)
        dist.destroy_process_group()
    line = None
    if rank == 0:
        peaks = load_peaks()
        roof = roofline_gconv(device, peaks, flush)
        solo = world == 1                            # the sweeps, the sampler and the CPU baseline are N = 1 extras
        disp = dispatch_sweep(device, peaks, flush, full=args.full_sweep) if solo or args.full_sweep else None
        cpu = cpu_baseline_train(sample_batch=8, iters=3) if solo else None
        ref_gpu = ref_cuda_eager(device) if solo and not args.no_sampler else None
        line = make_line(roof, cpu, disp, ref_gpu, peaks["src"], samp, cfg_c)
        print(json.dumps(line), flush=True)
    return line


def _time_us(fn, flush, iters, warm=3):
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / iters * 1e3


def dispatch_point(device, flush, T, E, k, D, iters=10, dtype=torch.bfloat16):
    """index build + permute + combine at one (tokens, experts, top-k, row width) point.  Algorithmic bytes per SURVEY
    §8d: PLAN reads the router's T*k (index, weight) pairs and writes R*12 (row_src, row_expert, row_w) + T*k*4
    (tok_rows); PERMUTE 2*R*D*s + R*4; COMBINE R*D*s + R*8 + T*D*s.  `dispatch_combine_GBs` counts all three kernels
    sets in bytes AND time (the plan is item (2) of north_star); `permute_combine_GBs` is the data movement alone."""
    from hdmoe_b200 import ops
    gen = torch.Generator(device="cpu").manual_seed(T + E + k)
    lg = torch.randn(T, E, generator=gen) + torch.log(1.0 / torch.arange(1, E + 1).float())     # Zipf-skewed load
    idx = lg.topk(k, dim=1).indices
    sp = torch.zeros(T, E).scatter_(1, idx, 1.0 / k).to(device)
    idx_d = idx.to(torch.int32).to(device)
    tw_d = torch.full((T, k), 1.0 / k, device=device)
    x = torch.randn(T, D, generator=gen).to(device=device, dtype=dtype)
    plan = ops.dispatch_plan_from_topk(idx_d, tw_d, E)
    rows = ops.permute(plan, x)[0]
    s = x.element_size()
    R = T * k
    t_plan = _time_us(lambda: ops.dispatch_plan_from_topk(idx_d, tw_d, E), flush, iters)
    t_plan_dense = _time_us(lambda: ops.dispatch_plan(sp, top_k=k), flush, max(3, iters // 3))
    t_perm = _time_us(lambda: ops.permute(plan, x), flush, iters)
    t_comb = _time_us(lambda: ops.combine(rows, sp, plan), flush, iters)
    b_plan = T * k * 8 + R * 12 + T * k * 4
    b_perm = 2 * R * D * s + R * 4
    b_comb = R * D * s + R * 8 + T * D * s
    return {"T": T, "E": E, "k": k, "row_bytes": D * s, "plan_us": round(t_plan, 1), "plan_dense_us": round(t_plan_dense, 1),
            "permute_us": round(t_perm, 1), "combine_us": round(t_comb, 1), "permute_GBs": round(b_perm / t_perm / 1e3, 1),
            "combine_GBs": round(b_comb / t_comb / 1e3, 1),
            "permute_combine_GBs": round((b_perm + b_comb) / (t_perm + t_comb) / 1e3, 1),
            "dispatch_combine_GBs": round((b_plan + b_perm + b_comb) / (t_plan + t_perm + t_comb) / 1e3, 1)}


def dispatch_sweep(device, peaks, flush, full=False):
    pts = [(1024, 4, 1, 32768), (1024, 8, 2, 32768), (65536, 16, 2, 512), (262144, 64, 2, 128), (1048576, 64, 1, 128),
           (1048576, 64, 2, 512)]
    if full:
        pts = [(T, E, k, D) for T in (4096, 16384, 65536, 262144, 1048576) for E in (4, 8, 16, 32, 64) for k in (1, 2)
               for D in (32, 128, 512) if T * k * D * 2 <= (4 << 30)] + [(256, 4, 1, 32768), (1024, 4, 1, 32768)]
    out = [dispatch_point(device, flush, *p) for p in pts]
    best = max(o["dispatch_combine_GBs"] for o in out)
    ref = out[-1] if full else out[0]            # the reference-faithful point (T = 1024 samples, 64 KiB bf16 rows)
    return {"unit": "GB/s", "peak": peaks["hbm"], "best_frac": round(best / peaks["hbm"], 4),
            "reference_point_frac": round(ref["dispatch_combine_GBs"] / peaks["hbm"], 4),
            "note": "dispatch_combine_GBs = (plan + permute + combine bytes) / (plan + permute + combine time); L2 flushed "
                    "(256 MiB write) before every timed launch", "points": out}


# Every MP_Conv of one Unet_expert that runs through the grouped tcgen05 kernels (SURVEY Appendix D, cfg1/cfg2 at 32x32):
# (Cin, Cout, H = W, k x k conv? (False = 1x1 skip), multiplicity per expert forward).  30 k x k convolutions + 7 skips.
UNET_LAYERS = [(64, 64, 16, True, 11), (32, 32, 32, True, 8), (64, 64, 32, True, 2), (64, 32, 32, True, 2),
               (128, 64, 16, True, 2), (32, 32, 16, True, 2), (96, 64, 16, True, 1), (96, 32, 32, True, 1),
               (33, 32, 32, True, 1), (128, 64, 16, False, 2), (64, 32, 32, False, 2), (32, 64, 16, False, 1),
               (96, 64, 16, False, 1), (96, 32, 32, False, 1)]
ROUTED = [36, 48, 75, 97]          # rows per expert at the bench batch (256 rows, the routing of the synthetic batch)
KSIZES = [3, 3, 5, 5]


def _layer_problem(device, cin, cout, H, kxk):
    """Operands of one grouped layer at the bench routing: NHWC bf16 rows, tap-major operands for forward / data gradient."""
    ks = KSIZES if kxk else [1, 1, 1, 1]
    R = sum(ROUTED)
    cin_pad = cin if cin % 32 == 0 else (cin + 63) // 64 * 64
    cin_rows = cin if cin % 32 == 0 else cin // 32 * 32
    row_e = torch.tensor(sum(([e] * c for e, c in enumerate(ROUTED)), []), dtype=torch.int32, device=device)
    n_rows = torch.tensor([R], dtype=torch.int32, device=device)
    x = torch.randn(R, H, H, cin_pad, device=device).to(torch.bfloat16)
    dy = torch.randn(R, H, H, cout, device=device).to(torch.bfloat16)
    wrow, wrow_t, tot, tot_t = [], [], 0, 0
    for k in ks:
        wrow.append(tot)
        wrow_t.append(tot_t)
        tot += k * k * cout
        tot_t += k * k * cin_rows
    w = (torch.randn(tot, cin_pad, device=device) / 30).to(torch.bfloat16)
    wt = (torch.randn(tot_t, cout, device=device) / 30).to(torch.bfloat16)
    dw = torch.zeros(tot, cin_pad, device=device)
    flops = sum(2.0 * c * H * H * cout * cin * k * k for c, k in zip(ROUTED, ks))
    flops_dg = sum(2.0 * c * H * H * cout * cin_rows * k * k for c, k in zip(ROUTED, ks))
    return dict(ks=ks, R=R, cin_pad=cin_pad, cin_rows=cin_rows, row_e=row_e, n_rows=n_rows, x=x, dy=dy, wrow=wrow,
                wrow_t=wrow_t, tot=tot, tot_t=tot_t, w=w, wt=wt, dw=dw, flops=flops, flops_dg=flops_dg)


def _cudnn_layer_us(pr, cin, cout, H, flush, iters):
    """The same layer through the library: one bf16 channels-last F.conv2d per expert on its contiguous row range
    (what set_grouped_experts(False) runs), forward and backward (data + weight gradient) timed separately."""
    import torch.nn.functional as F
    xs, ws, gs = [], [], []
    lo = 0
    for n, k in zip(ROUTED, pr["ks"]):
        xs.append(pr["x"][lo:lo + n, :, :, :cin].permute(0, 3, 1, 2))                 # NCHW view of NHWC memory
        ws.append((torch.randn(cout, cin, k, k, device=pr["x"].device) / 30).to(torch.bfloat16)
                  .contiguous(memory_format=torch.channels_last))
        gs.append(pr["dy"][lo:lo + n].permute(0, 3, 1, 2))
        lo += n

    def fwd():
        for x_, w_ in zip(xs, ws):
            F.conv2d(x_, w_, padding=(w_.shape[-1] - 1) // 2)

    def bwd():
        for x_, w_, g_ in zip(xs, ws, gs):
            p_ = (w_.shape[-1] - 1) // 2
            torch.ops.aten.convolution_backward(g_, x_, w_, None, (1, 1), (p_, p_), (1, 1), False, (0, 0), 1,
                                                (True, True, False))
    return _time_us(fwd, flush, iters), _time_us(bwd, flush, iters)


def gconv_shape_table(device, peaks, flush, iters=10, with_cudnn=True):
    """Per-shape and FLOP-weighted aggregate tensor-pipe fractions of the grouped convolution kernels over ONE train step
    of the U-Net experts: forward + data gradient (gconv) and weight gradient (gwgrad2) of every layer of UNET_LAYERS at
    the bench routing, each launch timed alone with CUDA events and a cold L2, against the measured burst bf16 peak."""
    from hdmoe_b200 import ops
    rows, agg = [], {"fwd": [0.0, 0.0], "dgrad": [0.0, 0.0], "wgrad": [0.0, 0.0], "cudnn_fwd": [0.0, 0.0],
                     "cudnn_bwd": [0.0, 0.0]}
    for cin, cout, H, kxk, mult in UNET_LAYERS:
        pr = _layer_problem(device, cin, cout, H, kxk)
        us_f = _time_us(lambda: ops.gconv_raw(pr["x"], pr["w"], cout, pr["tot"], pr["row_e"], pr["n_rows"], pr["ks"],
                                              pr["wrow"]), flush, iters)
        us_d = _time_us(lambda: ops.gconv_raw(pr["dy"], pr["wt"], pr["cin_rows"], pr["tot_t"], pr["row_e"], pr["n_rows"],
                                              pr["ks"], pr["wrow_t"]), flush, iters)
        us_w = _time_us(lambda: ops.gconv_wgrad_raw(pr["x"], pr["dy"], pr["dw"], pr["row_e"], pr["n_rows"], pr["ks"],
                                                    pr["wrow"]), flush, iters)
        row = {"cin": cin, "cout": cout, "hw": H, "k": "3,3,5,5" if kxk else "1x1", "mult": mult,
               "gflop": round(pr["flops"] / 1e9, 2), "fwd_us": round(us_f, 1), "dgrad_us": round(us_d, 1),
               "wgrad_us": round(us_w, 1), "fwd_frac": round(pr["flops"] / us_f / 1e6 / peaks["bf16"], 3),
               "dgrad_frac": round(pr["flops_dg"] / us_d / 1e6 / peaks["bf16"], 3),
               "wgrad_frac": round(pr["flops"] / us_w / 1e6 / peaks["bf16"], 3)}
        for key, fl, us in (("fwd", pr["flops"], us_f), ("dgrad", pr["flops_dg"], us_d), ("wgrad", pr["flops"], us_w)):
            agg[key][0] += fl * mult
            agg[key][1] += us * mult
        if with_cudnn and kxk:
            cf, cb = _cudnn_layer_us(pr, cin, cout, H, flush, max(3, iters // 2))
            row["cudnn_bf16_fwd_us"], row["cudnn_bf16_bwd_us"] = round(cf, 1), round(cb, 1)
            agg["cudnn_fwd"][0] += pr["flops"] * mult
            agg["cudnn_fwd"][1] += cf * mult
            agg["cudnn_bwd"][0] += (pr["flops"] + pr["flops_dg"]) * mult
            agg["cudnn_bwd"][1] += cb * mult
        rows.append(row)
        del pr
    def frac(fl_us):
        return round(fl_us[0] / fl_us[1] / 1e6 / peaks["bf16"], 4) if fl_us[1] else None
    conv = [agg["fwd"][0] + agg["dgrad"][0], agg["fwd"][1] + agg["dgrad"][1]]
    out = {"layers": rows, "routing": ROUTED,
           "aggregate": {"gconv_fwd_dgrad_frac": frac(conv), "gconv_fwd_frac": frac(agg["fwd"]),
                         "gconv_dgrad_frac": frac(agg["dgrad"]), "gwgrad_frac": frac(agg["wgrad"]),
                         "gconv_us_per_step": round(conv[1], 1), "gwgrad_us_per_step": round(agg["wgrad"][1], 1),
                         "gflop_per_step_fwd": round(agg["fwd"][0] / 1e9, 1),
                         "cudnn_bf16_fwd_frac": frac(agg["cudnn_fwd"]), "cudnn_bf16_bwd_frac": frac(agg["cudnn_bwd"]),
                         "basis": "FLOP-weighted over the 37 grouped launches of one U-Net-expert pass (x multiplicity), "
                                  "burst bf16 peak " + peaks["src"]}}
    return out


def roofline_gconv(device, peaks, flush, iters=10):
    """Dominant hand-written kernel of the train step by FLOPs: the tcgen05 grouped implicit-GEMM convolution.
    Headline shape: the most EXPENSIVE U-Net expert layer, decoders.32x32_up.conv_res1/2 (2 of the 30 k x k
    convolutions of an expert, 17 % of its FLOPs): 256 rows routed 36/48/75/97 to the 3x3, 3x3, 5x5, 5x5 experts,
    32x32, Cin = Cout = 64.  flops = sum_e 2*n_e*H*W*Cout*Cin*k_e^2.  `per_shape` / `aggregate` hold every
    distinct layer of Appendix D and the FLOP-weighted fraction over one step next to the library (cuDNN bf16) on the
    same layers."""
    from hdmoe_b200 import ops
    pr = _layer_problem(device, 64, 64, 32, True)
    us = _time_us(lambda: ops.gconv_raw(pr["x"], pr["w"], 64, pr["tot"], pr["row_e"], pr["n_rows"], pr["ks"], pr["wrow"]),
                  flush, iters)
    flops = pr["flops"]
    ach = flops / us / 1e6
    us_w = _time_us(lambda: ops.gconv_wgrad_raw(pr["x"], pr["dy"], pr["dw"], pr["row_e"], pr["n_rows"], pr["ks"],
                                                pr["wrow"]), flush, iters)
    name = "gconv2_fwd_kernel<64,64>"
    table = gconv_shape_table(device, peaks, flush, iters)
    return {"bound": "tensor", "kernel": name + " (256 rows 32x32, Cin=Cout=64, k=3,3,5,5 routed 36/48/75/97)",
            "achieved": round(ach, 1), "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": round(ach / peaks["bf16"], 4),
            "traffic": ncu_traffic(name), "us_per_launch": round(us, 1), "algorithmic_flops": flops,
            "peak_basis": "burst (kernel timed alone), " + peaks["src"],
            "wgrad": {"kernel": "gwgrad2_kernel<64,64> (same layer)", "us_per_launch": round(us_w, 1),
                      "achieved": round(flops / us_w / 1e6, 1), "frac": round(flops / us_w / 1e6 / peaks["bf16"], 4),
                      "traffic": ncu_traffic("gwgrad2_kernel<64,64>")},
            "aggregate": table["aggregate"], "per_shape": table["layers"]}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of this
    kernel at this shape (profiles/ncu_traffic.json, written from the .ncu-rep by tools/ncu_summary.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p)).get(kernel)


def sampler_throughput(device, rank, world, total_batch=1024, steps=18, guidance=1.0):
    """EDM Heun sampler, 18 steps / 35 NFE (70 with guidance), model_config2, CLIP-shaped text (BASELINE.json configs[3]):
    the batch of 1024 is split over the ranks (samples are independent: no collective), time = max over ranks."""
    import hdmoe_b200
    B = max(1, total_batch // world)
    model = build_model(2, device)
    model.eval()
    gen = torch.Generator().manual_seed(SEED + rank)
    noise = torch.randn(B, 4, 32, 32, generator=gen).to(device)
    text = torch.randn(B, 77, 768, generator=gen).to(device)
    uncond = torch.zeros_like(text) if guidance != 1.0 else None
    smp = hdmoe_b200.EDM_Sampler(model, model, num_solve_steps=steps, guidance=guidance, use_cuda_graph=True)
    smp.sample(noise, text, P_MEAN, P_STD, uncond_text_emb=uncond)      # warm-up: allocator, graph capture
    barrier(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    smp.nfe = 0
    a.record()
    out = smp.sample(noise, text, P_MEAN, P_STD, uncond_text_emb=uncond)
    b.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(a.elapsed_time(b), world, device)
    fin = bool(torch.isfinite(out).all())
    del smp, model
    torch.cuda.empty_cache()
    return {"metric": "EDM sample img/s", "value": round(B * world / (ms / 1e3), 1), "unit": "img/s", "batch": B * world,
            "batch_per_gpu": B, "nfe": (2 * steps - 1) * (2 if guidance != 1.0 else 1), "guidance": guidance,
            "ms": round(ms, 1), "finite": fin,
            "execution": "one CUDA graph per denoiser evaluation, fused Heun kernels between; ranks sample independently"}


def ep_exchange_breakdown(rank, world, device, B=64, res=64, top_k=1, iters=10):
    """What the expert-parallel layer puts on NVLink per train step at config C, timed alone: the equal-split
    all-to-alls of the image / time / text send buffers (dispatch), the image-row all-to-all back (combine) and their
    backward mirrors, at the layer's static buffer sizes (G segments of C = T*k rows, bf16).  Device time, max over
    ranks.  Bytes: `wire` = rows that actually leave the GPU for a balanced router ((G-1)/G of T*k rows), `buffer` =
    what the fixed-capacity exchange moves (every segment, live or zero-filled)."""
    import torch.distributed as dist
    C = B * top_k
    widths = {"image": 32 * res * res, "time": 64, "text": 768}
    bufs = {k: torch.randn(world * C, w, device=device).to(torch.bfloat16) for k, w in widths.items()}
    outs = {k: torch.empty_like(v) for k, v in bufs.items()}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        barrier(world)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return max_over_ranks(a.elapsed_time(b), world, device) / iters * 1e3

    def dispatch():
        for k in widths:
            dist.all_to_all_single(outs[k], bufs[k])

    def combine():
        dist.all_to_all_single(outs["image"], bufs["image"])

    us_d, us_c = timed(dispatch), timed(combine)
    cnt = torch.zeros(world * 4, dtype=torch.int64, device=device)
    us_counts = timed(lambda: dist.all_gather_into_tensor(cnt, cnt[:4].contiguous()))
    row = sum(widths.values()) * 2
    buf_bytes = world * C * (row + widths["image"] * 2) * 2            # dispatch + combine, forward + backward
    wire_bytes = int(C * (world - 1) / world * (row + widths["image"] * 2) * 2)
    per_step_us = 2 * (us_d + us_c) + us_counts
    return {"dispatch_a2a_us": round(us_d, 1), "combine_a2a_us": round(us_c, 1), "counts_allgather_us": round(us_counts, 1),
            "exchange_us_per_step": round(per_step_us, 1), "buffer_bytes_per_step": buf_bytes,
            "wire_bytes_per_step_balanced": wire_bytes,
            "buffer_GBs": round(buf_bytes / per_step_us / 1e3, 1),
            "note": "one expert-parallel MoE layer (U-Net experts) per step: forward dispatch + combine and their "
                    "backward mirrors, NCCL all_to_all_single timed alone (no overlap), bf16 payload"}


def config_c_throughput(args, rank, world, device, parallelism, B=64, steps=5):
    """BASELINE configs[2]: model_config2 at 4x64x64 (11.2 M parameters), bf16 expert path, batch 64 per GPU, whole train
    step; "dp" = data-parallel replicas, "ep" = U-Net experts sharded over the ranks (static-shape all-to-all layer)."""
    from hdmoe_b200 import expert_parallel as EP
    import hdmoe_b200
    try:
        r = make_runner(args, 2, 64, B, rank, world, device, parallelism)
        r.run()
        barrier(world)
        ms = max_over_ranks(r.time_steps(steps), world, device) / steps
        loss = r.run()
        fin = bool(torch.isfinite(loss).all())
        load = None
        if parallelism == "ep" and world > 1:
            from hdmoe_b200 import peer
            # every rank must take the same branch: reduce the device-side failure flags before acting on them
            if max_over_ranks(float(peer.any_failed()), world, device):
                raise RuntimeError("peer-memory barrier timed out on some rank")
            if max_over_ranks(float(EP.overflowed()), world, device):
                if getattr(args, "ep_capacity", "auto") is None:
                    raise RuntimeError("expert-parallel capacity overflow at the exact worst case (cannot happen)")
                del r
                hdmoe_b200.disable_expert_parallel()
                torch.cuda.empty_cache()
                import copy
                worst = copy.copy(args)
                worst.ep_capacity = None            # exact worst case: degree * T * k rows
                return config_c_throughput(worst, rank, world, device, parallelism, B, steps)
            st = EP.LAST_STATS
            if "recv_rows" in st:
                recv = float(st["recv_rows"])
                # with fewer experts than ranks in the group some ranks receive nothing; balanced = B * k rows per rank
                load = {"recv_rows_max": max_over_ranks(recv, world, device), "recv_rows_min": -max_over_ranks(-recv, world, device),
                        "rows_per_rank_balanced": B, "capacity_rows": int(st["capacity_rows"]),
                        "imbalance_max_over_mean": round(max_over_ranks(recv, world, device) / B, 3)}
        note = r.note
        EP.LAST_STATS.clear()                        # device tensors of the recorded step: do not outlive its graph
        del r
    except Exception as exc:                          # noqa: BLE001
        import traceback
        traceback.print_exc()
        hdmoe_b200.disable_expert_parallel()
        torch.cuda.empty_cache()
        return {"parallelism": parallelism, "error": str(exc)[:300]}
    hdmoe_b200.disable_expert_parallel()
    torch.cuda.empty_cache()
    return {"metric": "denoiser train img/s", "workload": "model_config2 4x64x64 train step, batch %d per GPU" % B,
            "parallelism": (f"ep{min(world, 4) if world % 4 == 0 or world < 4 else world} (U-Net experts) + dp{world} (trunk)" if parallelism == "ep" else f"dp{world}"),
            "value": round(B * world / (ms / 1e3), 1), "unit": "img/s", "ms_per_step": round(ms, 2), "execution": note,
            "finite": fin, **({"load": load} if load else {})}


EXTRAS_TIMEOUT_S = 420


def scale_extras(args, rank, world, device, progress=None):
    """(sampler, config_c) records of this N; every rank runs them (rank 0 reports).  `progress` receives the records as
    they complete (read by the watchdog of run_ours)."""
    progress = {} if progress is None else progress
    samp = {"g1": sampler_throughput(device, rank, world, guidance=1.0)}
    samp["g2"] = sampler_throughput(device, rank, world, guidance=2.0)
    samp.update({k: samp["g1"][k] for k in ("metric", "value", "unit", "batch", "nfe", "ms", "finite")})   # headline: g = 1
    progress["sampler"] = samp
    import copy
    cfg_c = {"dp": config_c_throughput(args, rank, world, device, "dp")}
    cfg_c.update({k: cfg_c["dp"].get(k) for k in ("metric", "value", "unit", "ms_per_step", "workload")})
    progress["config_c"] = cfg_c
    if world > 1:
        cfg_c["ep"] = config_c_throughput(args, rank, world, device, "ep")
        nccl_args = copy.copy(args)
        nccl_args.ep_transport, nccl_args.ep_graph = "nccl", False
        cfg_c["ep_nccl_eager"] = config_c_throughput(nccl_args, rank, world, device, "ep")
        # the expert-parallel step runs eagerly (make_runner): the like-for-like bar is the data-parallel step run eagerly
        # too; the difference between the two eager steps is what expert parallelism itself costs, and the exchange
        # timed alone says how much of that is NVLink time
        eager_args = copy.copy(args)
        eager_args.no_graph = True
        cfg_c["dp_eager"] = config_c_throughput(eager_args, rank, world, device, "dp")
        try:
            cfg_c["exchange"] = ep_exchange_breakdown(rank, world, device)
        except Exception as exc:                      # noqa: BLE001
            cfg_c["exchange"] = {"error": str(exc)[:200]}
        try:
            cfg_c["ep_over_dp"] = round(cfg_c["ep"]["value"] / cfg_c["dp"]["value"], 3)
            cfg_c["ep_nccl_eager_over_dp_eager"] = round(cfg_c["ep_nccl_eager"]["value"] / cfg_c["dp_eager"]["value"], 3)
            cfg_c["ep_extra_ms_vs_dp"] = round(cfg_c["ep"]["ms_per_step"] - cfg_c["dp"]["ms_per_step"], 2)
        except (KeyError, ZeroDivisionError):
            pass
    cfg_c.update({k: cfg_c["dp"].get(k) for k in ("metric", "value", "unit", "ms_per_step", "workload")})
    return samp, cfg_c


# ---------------------------------------------------------------------------------------------------------
# Reference arm: the UNMODIFIED reference (oracle/_ref, staged by oracle/make_ref.py) on the host cores; the committed
# oracle port only when _ref is absent.  Same synthetic batch, same init protocol, train mode (dropout on) in both arms.
# ---------------------------------------------------------------------------------------------------------
def _ref_available():
    from oracle import make_ref
    return make_ref.available()


def build_ref_model(variant, device, res=32, seed=0):
    """The reference's own preconditioned_HDMOEM with build_model's init protocol (seeded constructor, zero-init
    parameters re-drawn from N(0, 0.3^2))."""
    from oracle import make_ref
    P1, P2, _, _, _ = make_ref.import_ref()
    torch.manual_seed(seed)
    model = (P1 if variant == 1 else P2)(**dict(FULL, IN_img_resolution=res))
    gen = torch.Generator().manual_seed(seed + 100)
    with torch.no_grad():
        for p in model.parameters():
            if float(p.abs().max()) == 0:
                p.copy_(torch.randn(p.shape, generator=gen) * 0.3)
    return model.to(device)


def _ref_train_step_fn(model, b, variant=1):
    from oracle import make_ref
    crit = make_ref.import_ref()[2](**LOSS)
    params = list(model.parameters())
    opt = torch.optim.AdamW(params, lr=5e-4)
    kw = dict(transition_point=P_MEAN, softness=P_STD) if variant == 2 else {}

    def step():
        out = model(x=b["x"], sigma=b["sigma"], text_emb=b["text"], Unet_router_mask=b["um"], Vit_router_mask=b["vm"],
                    zeta=2.0, return_log_var=True, **kw)
        loss = crit(sigma_vec=b["sigma"], x=b["x0"], sigma=b["sigma"], out_model=out)["loss"]
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        return loss
    return step


def cpu_baseline_train(sample_batch=8, iters=3, warmup=1):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    b = synth_batch(sample_batch, 32, 0, "cpu")
    if _ref_available():
        model = build_ref_model(1, "cpu")
        model.train()
        step = _ref_train_step_fn(model, b)
        kind = "reference"
        what = "UNMODIFIED reference (oracle/_ref: models/model_config1.py + Utils/utils.py EDM_LOSS), train mode"
    else:
        from oracle import hdmoe_oracle as O
        model = build_model(1, "cpu")
        sd = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and not k.endswith(("freqs", "phases")))
              for k, v in model.state_dict().items()}
        params = [v for v in sd.values() if v.requires_grad]
        opt = torch.optim.AdamW(params, lr=5e-4)
        gen = torch.Generator().manual_seed(7)

        def step():
            noise = {"scaling": torch.randn(sample_batch, 2, generator=gen),
                     "vit": torch.randn(sample_batch, 4, generator=gen), "unet": torch.randn(sample_batch, 4, generator=gen)}
            with O.training_mode():
                out = O.preconditioned(sd, FULL, b["x"], b["sigma"], b["text"], b["um"], b["vm"], zeta=2.0,
                                       return_log_var=True, noise=noise, variant=1)
            loss = O.edm_loss(b["x0"], out, 4, LOSS["Unet_bal"], LOSS["vit_bal"], LOSS["z_bal"])["loss"]
            opt.zero_grad(set_to_none=True)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
        kind = "port"
        what = "oracle (CPU port of the reference; oracle/_ref absent), dropout off"

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(iters):
        step()
    dt = (time.perf_counter() - t0) / iters
    return {"value": round(sample_batch / dt, 3), "unit": "img/s", "cores": cores, "kind": kind,
            "sample": f"{what}: model_config1 train step (fwd+EDM_LOSS+bwd+clip+AdamW), batch {sample_batch}, fp32, "
                      f"{iters} steps after {warmup} warm-up", "s_per_step": round(dt, 3)}


def ref_cuda_eager(device, variant=1, B=256, res=32, iters=3, warmup=2):
    """Same-box GPU bar: the unmodified reference in eager CUDA fp32 (TF32 matmul / cuDNN on, as bench.py sets
    for the whole process) on this B200, same workload and inputs as the headline.  None when oracle/_ref is absent."""
    if not _ref_available():
        return None
    for batch in (B, B // 4):
        try:
            model = build_ref_model(variant, device, res)
            model.train()
            b = {k: v.to(device) for k, v in synth_batch(batch, res, 0, device).items()}
            step = _ref_train_step_fn(model, b, variant)
            for _ in range(warmup):
                step()
            torch.cuda.synchronize()
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                loss = step()
            c.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(c) / iters
            peak = torch.cuda.max_memory_allocated(device) / 2 ** 30
            del model, step, b
            torch.cuda.empty_cache()
            return {"metric": "denoiser train img/s", "value": round(batch / (ms / 1e3), 1), "unit": "img/s",
                    "ms_per_step": round(ms, 2), "batch": batch, "kind": "reference (oracle/_ref) eager CUDA fp32, TF32 on",
                    "workload": f"model_config{variant} train step {res}x{res}", "finite": bool(torch.isfinite(loss)),
                    "peak_mem_GiB": round(peak, 1)}
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            continue
        except Exception as exc:                        # noqa: BLE001
            return {"error": str(exc)[:200]}
    return {"error": "out of memory at every tried batch"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    iters = max(1, min(args.steps, 5))
    cpu = cpu_baseline_train(sample_batch=8, iters=iters, warmup=max(1, min(args.warmup, 2)))
    world = int(os.environ.get("WORLD_SIZE", 1))
    line = {"impl": "reference", "metric": "denoiser train img/s", "value": cpu["value"], "unit": "img/s",
            "n_gpus": world, "steps": iters, "warmup": max(1, min(args.warmup, 2)),
            "ms_per_step": round(cpu["s_per_step"] * 1e3, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "model_config1 train step (fwd+EDM_LOSS+bwd+clip+AdamW), bounded sample batch 8, "
                                   "4x32x32 latent, text (8,77,768), CPU fp32", "global_batch": 8,
                       "parallelism": "cpu"},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=256, help="per-GPU batch")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--profile-step", action="store_true", help="cudaProfilerStart/Stop around the first timed step")
    ap.add_argument("--full-sweep", action="store_true", help="full MoE dispatch/combine sweep (BASELINE configs[4])")
    ap.add_argument("--no-sampler", action="store_true", help="skip the EDM sampler throughput extra")
    ap.add_argument("--parallelism", default="dp", choices=["dp", "ep"], help="N>1: data-parallel replicas or expert-parallel U-Net experts")
    ap.add_argument("--no-graph", action="store_true", help="run the eager step instead of the whole-step CUDA graph")
    ap.add_argument("--ep-graph", action="store_true", help="record the expert-parallel step even with --ep-transport nccl")
    ap.add_argument("--ep-transport", default="peer", choices=["peer", "nccl"],
                    help="expert-parallel exchange: peer-memory kernels (graph-capturable) or NCCL all-to-all (eager)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (impl ours) needs a CUDA device: hdmoe_b200 has no CPU fallback")
        run_ours(args)


if __name__ == "__main__":
    main()
