#!/usr/bin/env python
"""Benchmark of the VAE training step (the north-star metric: train images/sec).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

A "step" is zero_grad -> Encoder -> reparameterise(+KL) -> Decoder -> mse + sum(kl) -> backward
(-> bucketed gradient all-reduce when N > 1) -> RMSprop(lr=1e-4) step, on one batch of synthetic
64x64 grayscale images at batch 256 per GPU (BASELINE.json configs[1], resolved in SURVEY.md section 0 to
the models/networks.py VAE).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic FLOPs per image per training step (fwd + dgrad + wgrad), SURVEY.md section 8d / BASELINE.md section 3
STEP_MFLOP = {(64, 1): 3935.6, (64, 3): 4027.3, (128, 1): 21802.6}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def _load_reference_modules():
    """The UNMODIFIED reference modules, vendored by oracle/build_ref.py into the git-ignored oracle/_ref/ (they travel to
    the GPU box like the built .so; /root/reference itself does not exist there).  None when the recipe has not run."""
    import importlib.util
    path = os.path.join(ROOT, "oracle", "_ref", "models", "networks.py")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("vaeplay_reference_networks", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _ReferenceCpuStep:
    """The reference's own CPU path for the VAE terms of train.py:43-78: its Encoder / Decoder / VaeGan.reparameterize
    modules (models/networks.py), F.mse_loss + sum(kl), backward, torch.optim.RMSprop(lr=1e-4) per sub-network (:136-140)."""

    def __init__(self, networks, img, cin):
        import torch
        from oracle import vae_numpy as vn
        L = int(math.log2(img // 8))
        P = vn.synth_vae_params(img, 128, cin, cin, 0)
        self.enc = networks.Encoder(channel_in=cin, z_size=128, iter_level=L)
        self.dec = networks.Decoder(z_size=128, size=self.enc.size, channel_out=cin, iter_level=L)
        self.enc.load_state_dict({k[8:]: torch.from_numpy(v) for k, v in P.items() if k.startswith("encoder.")}, strict=False)
        self.dec.load_state_dict({k[8:]: torch.from_numpy(v) for k, v in P.items() if k.startswith("decoder.")}, strict=False)
        self.enc.train(); self.dec.train()
        self.networks = networks
        self.opts = [torch.optim.RMSprop(self.enc.parameters(), lr=1e-4), torch.optim.RMSprop(self.dec.parameters(), lr=1e-4)]

    def step(self, x):
        import torch
        import torch.nn.functional as F
        self.enc.zero_grad(); self.dec.zero_grad()
        mus, log_variances = self.enc(x)
        z = self.networks.VaeGan.reparameterize(None, mus, log_variances)
        x_tilde = self.dec(z)
        kl = -0.5 * torch.sum(-log_variances.exp() - torch.pow(mus, 2) + log_variances + 1, 1)
        loss = F.mse_loss(x, x_tilde) + torch.sum(kl)
        loss.backward()
        for o in self.opts:
            o.step()
        return loss.detach()


def _cpu_stepper(img, cin):
    """(stepper with .step(x), kind, description): the unmodified reference when oracle/_ref exists, else the line-by-line port."""
    import torch
    from oracle import vae_numpy as vn
    ref = _load_reference_modules()
    if ref is not None:
        return _ReferenceCpuStep(ref, img, cin), "reference", "unmodified reference models/networks.py (oracle/_ref) Encoder+reparameterize+Decoder, F.mse_loss+sum(kl), RMSprop"
    from oracle.vae_torch import VaeTorchPort
    P = vn.synth_vae_params(img, 128, cin, cin, 0)
    return VaeTorchPort(P, torch.float32), "port", "oracle/vae_torch.py (torch-CPU restatement of models/networks.py + train.py step)"


def cpu_baseline_sample(img, cin, steps=5, warmup=2, batch=16):
    import torch
    from oracle import vae_numpy as vn
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    stepper, kind, what = _cpu_stepper(img, cin)
    x = torch.from_numpy(vn.synth_batch(batch, img, cin, 128, 0)[0])
    for _ in range(warmup):
        stepper.step(x)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        stepper.step(x)
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return {"value": round(batch / med, 2), "unit": "images/s", "cores": threads, "kind": kind,
            "sample": f"{what}, fp32, batch {batch}, median of {steps} steps after {warmup} warm-up, {med * 1e3:.1f} ms/step"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (torch CPU ops) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import vae_numpy as vn
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    sample_b = args.ref_batch
    stepper, kind, what = _cpu_stepper(args.img, args.cin)
    x = torch.from_numpy(vn.synth_batch(sample_b, args.img, args.cin, 128, 0)[0])
    for _ in range(args.warmup):
        stepper.step(x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        stepper.step(x)
    dt = time.perf_counter() - t0
    ips = sample_b * args.steps / dt
    line = {
        "impl": "reference", "metric": "train_images_per_sec", "value": round(ips, 2), "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, sample_b, note=f"CPU sample: each step is one train step on {sample_b} images"),
        "cpu_baseline": {"value": round(ips, 2), "unit": "images/s", "cores": threads, "kind": kind,
                         "sample": f"{what}, fp32, {sample_b} images per step, {args.steps} steps"},
        "e2e": {"value": round(ips, 2), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, batch, note=None):
    cfg = {"workload": f"models/networks.py VAE (Encoder+reparameterize+Decoder, KL+MSE) {args.img}x{args.img}x{args.cin}, z=128, "
                       f"batch {batch}/GPU, fwd+loss+bwd+RMSprop",
           "img_size": args.img, "channels": args.cin, "batch_per_gpu": batch, "z": 128,
           "optimizer": "RMSprop(lr=1e-4, alpha=.99, eps=1e-8) inside the timed step (train.py:136-140); ours: fused multi-tensor kernel"}
    if note:
        cfg["note"] = note
    return cfg


class StepRunner:
    """bench.py's view of vae_play_b200.engine.VaeTrainer: the trainer plus the timing loops."""

    def __init__(self, args, img, world, rank, dev):
        from vae_play_b200.engine import VaeTrainer
        self._t = VaeTrainer(args, img, world, rank, dev)

    def __getattr__(self, name):        # model, step, x_host, x_dev, static_x, graph, split, wire, launches_per_step, ...
        return getattr(self.__dict__["_t"], name)

    def barrier(self):
        import torch
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def _max_over_ranks(self, ms):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def time_resident(self, steps):
        """K steps with the input already resident in HBM; CUDA events, barrier + synchronize on both sides, max over ranks."""
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        for _ in range(steps):
            loss = self.step(self.x_dev)
        e1.record()
        self.barrier()
        return self._max_over_ranks(e0.elapsed_time(e1))

    def time_e2e(self, steps):
        """K steps through the public API with HOST inputs: every step's batch is copied from pinned host memory and every step's
        loss is read back to the host, all inside the timed region.  Both transfers are pipelined (vae_play_b200.host_io): the
        copy of batch i+1 runs on a copy stream under step i, and the loss of step i is read while step i+1 runs."""
        import torch
        from vae_play_b200.host_io import HostBatchPipeline, ScalarReadback
        if getattr(self, "_pipe", None) is None:
            self._pipe = HostBatchPipeline(tuple(self.x_host.shape), device=self.dev)
            self._reader = ScalarReadback(1, device=self.dev)
        pipe, reader = self._pipe, self._reader
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        e0.record()
        pipe.feed(self.x_host)
        for i in range(steps):
            xd = pipe.take()
            if i + 1 < steps:
                pipe.feed(self.x_host)             # H2D of the next batch, under this step
            loss = self.step(xd)                   # graph mode: one device-to-device copy into the graph's input, then the replays
            pipe.release()
            if reader.pending() == 2:
                self.loss_host = reader.pop()[0]   # D2H read of step i-2's loss: the host stays at most two steps ahead
            reader.push(loss)
        while reader.pending():
            self.loss_host = reader.pop()[0]
        e1.record()
        self.barrier()
        return self._max_over_ranks(e0.elapsed_time(e1))


def init_nccl(args, dev):
    from vae_play_b200.engine import init_nccl as _init
    _init(args.sm_reserve, dev)


def run_ours(args):
    import torch
    import torch.distributed as dist

    import vae_play_b200 as vp
    from vae_play_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")      # keep NCCL's version banner off stdout: the driver reads ONE JSON line
        init_nccl(args, dev)
    vp.set_precision(args.precision)
    B, img, cin = args.batch, args.img, args.cin
    if cin != 1:
        raise SystemExit("bench: the stock VaeGan is grayscale (models/networks.py:207-209); use --cin 1")
    peaks = load_peaks()
    run = StepRunner(args, img, world, rank, dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~0.3 s to produce its first sample: start before the warm-up
    # untimed warm-up: W steps as asked, plus a fixed number of extra steps (same on every rank: the step
    # contains collectives) so that the clock sampler is live and the SM clocks have ramped up
    for _ in range(max(args.warmup, 3) + args.extra_warmup):
        run.step(run.x_dev)
    run.barrier()
    if rank == 0:
        sampler.lines.clear()    # keep only samples taken during the timed regions

    # ---- timed regions: `repeats` regions of exactly K steps each; the MEDIAN region is reported ---------------------
    n0 = _lib.launch_count()
    regions = [run.time_resident(args.steps) for _ in range(args.repeats)]
    launches = (_lib.launch_count() - n0) // args.repeats if not run.graph else run.launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms = statistics.median(regions)
    ips = world * B * args.steps / (ms / 1e3)
    regions2 = [run.time_e2e(args.steps) for _ in range(max(1, min(args.repeats, 3)))]
    ms2 = statistics.median(regions2)
    e2e_ips = world * B * args.steps / (ms2 / 1e3)
    assert math.isfinite(run.loss_host), "non-finite loss"

    line = None
    if rank == 0:
        mflop = STEP_MFLOP.get((img, cin))
        step_tf = ips / world * mflop * 1e6 / 1e12 if mflop else None
        line = {
            "metric": "train_images_per_sec", "value": round(ips, 1), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": dict(workload_config(args, B), parallelism=f"dp{world}",
                           l2="no explicit flush: per-step working set (~1 GB of activations at batch 256) exceeds the 126 MB L2",
                           input="one synthetic batch, re-used every step (resident: already in HBM; e2e: copied from pinned host memory every step on a copy stream, double-buffered, and every step's loss read back one or two steps late -- vae_play_b200/host_io.py)",
                           timing=f"median of {args.repeats} timed regions of {args.steps} steps each (CUDA events, barrier + synchronize on both sides, max over ranks)",
                           cuda_graph=run.graph, pdl=os.environ.get("VP_PDL", "1") != "0",
                           streams=("weight gradients on a side stream next to the BatchNorm-backward passes" if run.async_wgrad else "single stream")
                                   + ("; optimiser of everything but the encoder convs on a third stream next to the encoder-conv backward (one graph per step)" if run.overlap_opt else ""),
                           allreduce=(None if world == 1 else
                                      f"bucketed NCCL all-reduce, {run.wire} on the wire (<= {args.sm_reserve} CTAs, high-priority stream); the decoder / fc / heads buckets "
                                      f"(94 % of the bytes) run next to the encoder-conv backward graph, whose persistent grids are capped at "
                                      f"#SMs - {args.sm_reserve}; per-bucket optimiser graphs start as each bucket completes"
                                      + ("; backward in three stages (the decoder buckets are exchanged next to the heads / encoder.fc backward)" if run.three_stage else "")
                                      if run.split else f"bucketed NCCL all-reduce ({run.wire} on the wire) between backward and optimiser")),
            "clocks": clocks,
            "region_ms_per_step": [round(r / args.steps, 4) for r in regions],
            "e2e": {"value": round(e2e_ips, 1), "unit": "images/s", "h2d_bytes_per_step": run.x_host.numel() * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms2 / args.steps, 4)},
            "gpu_launches": int(launches),
            "step_tflops_per_gpu": round(step_tf, 2) if step_tf else None,
            "step_frac_of_sustained_bf16": round(step_tf / peaks["tf_sustained"], 4) if step_tf else None,
            "step_frac_of_burst_bf16": round(step_tf / peaks["tf_burst"], 4) if step_tf else None,
            "peaks": peaks["src"],
            "loss": run.loss_host,
        }
        # ---- kernel probes (N = 1 semantics: rank 0 only, after the timed regions) ---------------------------------
        line["roofline"] = kernel_probe(run.model, B, dev, peaks)
        if not args.no_extras:
            line["roofline_hbm"] = hbm_probe(run.model, B, dev, peaks)
    if world == 1 and not args.no_extras:
        extra = {}
        # ---- the other image size the north-star names, same batch / steps, measured the same way ----------------------
        other = 128 if img == 64 else 64
        del run
        torch.cuda.empty_cache()
        run2 = StepRunner(args, other, world, rank, dev)
        s2 = ClockSampler(local)
        s2.start()
        for _ in range(max(args.warmup, 3) + 10):
            run2.step(run2.x_dev)
        run2.barrier()
        s2.lines.clear()
        reg = [run2.time_resident(args.steps) for _ in range(args.repeats)]
        c2 = s2.stop()
        m2 = statistics.median(reg)
        reg_e = [run2.time_e2e(args.steps) for _ in range(2)]
        ips2 = B * args.steps / (m2 / 1e3)
        mf2 = STEP_MFLOP.get((other, cin))
        tf2 = ips2 * mf2 * 1e6 / 1e12
        extra[f"img{other}"] = {"value": round(ips2, 1), "unit": "images/s", "ms_per_step": round(m2 / args.steps, 4),
                                "e2e_value": round(B * args.steps / (statistics.median(reg_e) / 1e3), 1),
                                "step_tflops": round(tf2, 2), "step_frac_of_sustained_bf16": round(tf2 / peaks["tf_sustained"], 4),
                                "step_frac_of_burst_bf16": round(tf2 / peaks["tf_burst"], 4), "clocks": c2,
                                "workload": f"models/networks.py VAE {other}x{other}x{cin}, z=128, batch {B}, fwd+loss+bwd+RMSprop"}
        del run2
        torch.cuda.empty_cache()
        # ---- the GPU-library bar: the same step through torch.nn.functional (cuDNN / cuBLAS), context only ---------------
        extra["gpu_library_baseline"] = gpu_library_baseline(img, cin, B, dev, args.steps)
        # ---- config 4: the full train.py VAE-GAN step at 128x128 -------------------------------------------------------------
        try:
            extra["vaegan128"] = vaegan_bench(args, dev, peaks)
        except Exception as e:       # an extra must not take the headline line down with it
            extra["vaegan128"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
        # ---- config 5: one train_Style_GAN.py iteration at 256x256 ---------------------------------------------------------------
        try:
            extra["style256"] = style_bench(args, dev)
        except Exception as e:
            extra["style256"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        torch.cuda.empty_cache()
        line["extra"] = extra
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_sample(img, cin)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def vaegan_flops_per_image(img, z=128):
    """Forward MACs x 2 of one train.py step per input image: encoder, decoder on z and z_p, discriminator on the 3 images
    (x, x~, x_p) in REC mode (conv stack up to the last block's convolution) and in GAN mode (full), DirectDecoder."""
    L = int(math.log2(img // 8))
    conv = lambda hw_out, cin, cout, k=5: hw_out * hw_out * cin * cout * k * k
    enc, c, hw = 0, 1, img
    chans = []
    for i in range(L):
        co = 64 if i == 0 else c * 2
        hw //= 2
        enc += conv(hw, c, co)
        c = co
        chans.append(co)
    enc += 64 * c * 1024 + 1024 * 2 * z
    dec, hw, cd = z * 64 * c, 8, c
    for i in range(L):
        co = cd if i == 0 else cd // 2
        dec += conv(hw, cd, co)            # transposed conv: MAC = in-pixels * cin * cout * k*k
        hw *= 2
        cd = co
    dec += conv(img, cd, 1)
    disc_rec, cdi, hw = conv(img, 1, 32), 32, img
    for i in range(L):
        hw //= 2
        disc_rec += conv(hw, cdi, cdi * 2)
        cdi *= 2
    disc_gan = disc_rec + 64 * cdi * 512 + 512
    aux = z * 512 + 512 * 256 + 256 * 128 + 128 * 64 + 2 * 64 * 32 + 32 * 3
    fwd_mac = enc + 2 * dec + 3 * (disc_rec + disc_gan) + aux
    return 2.0 * fwd_mac


def vaegan_bench(args, dev, peaks, img=128, B=64):
    """Config 4 of BASELINE.json: the FULL train.py step (VaeGan forward with the decoder on z and z_p and the discriminator on
    3B images in REC and GAN mode, VaeGan.loss, the summed-loss backward that equals the reference's five accumulating backward
    calls, four RMSprop optimisers), CUDA-graph replay, timed like the main workload."""
    import torch

    import vae_play_b200.functional as VF
    from vae_play_b200 import _lib
    from vae_play_b200 import train_steps as TS
    from vae_play_b200.models.networks import VaeGan
    from vae_play_b200.optim import FusedRMSprop
    torch.manual_seed(0)
    VF.set_async_wgrad(not args.no_async_wgrad)       # re-used weights (decoder x2, discriminator x2): later uses are added on the side stream too
    net = VaeGan(img, 128).to(dev).train()
    groups = [net.encoder, net.decoder, net.discriminator, net.param_encoder]             # train.py:136-140
    opts = [FusedRMSprop(list(m.parameters()), lr=1e-4, zero_grads=True) for m in groups]
    VF.persistent_grads(list(net.parameters()))
    x = torch.rand(B, 1, img, img, device=dev)
    x_host = x.cpu().pin_memory()
    targets = torch.rand(B, 3, device=dev)
    off_dev = torch.zeros(1, dtype=torch.int64, device=dev)
    inc = VF.philox_policy(B * 128, torch.cuda.get_device_properties(dev).multi_processor_count)[1]
    simt0 = _lib.simt_bf16_count()

    def fwd_bwd():
        b = x.size(0)
        x_tilde, disc_class, disc_layer, mus, lv, params = net(x, rng=(0, off_dev))
        VF.philox_advance(off_dev, 2 * inc)
        nle, kl, mse, bo, bp, bs, l1 = VaeGan.loss(x, x_tilde, disc_layer[:b], disc_layer[b:-b], disc_layer[-b:], disc_class[:b],
                                                    disc_class[b:-b], disc_class[-b:], mus, lv, targets, params)
        parts = dict(recon=VF.mse_loss(x, x_tilde), kl=kl, mse=mse, bce_o=bo, bce_p=bp, bce_s=bs, l1=l1)
        return TS.vaegan_backward(None, parts, fused=True)

    def eager():
        for o in opts:
            o.zero_grad(set_to_none=True)
        loss = fwd_bwd()
        for o in opts:
            o.step()
        return loss
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            eager()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    VF.invalidate_caches()
    for o in opts:
        o.zero_grad(set_to_none=True)
    ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    l0 = _lib.launch_count()
    with torch.cuda.graph(ga):
        static_loss = fwd_bwd()
    with torch.cuda.graph(gb, pool=ga.pool()):
        for o in opts:
            o.step()
    launches = _lib.launch_count() - l0

    def step():
        ga.replay()
        gb.replay()
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    if os.environ.get("VP_PROFILE_VAEGAN"):       # diagnostic: per-kernel device time of this step (VP_PDL=0 for clean durations) -> stderr
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                step()
            torch.cuda.synchronize()
        rows = sorted(((e.device_time_total / 3, e.count / 3, e.key) for e in prof.key_averages() if e.device_time_total), reverse=True)
        tot = sum(r[0] for r in rows)
        print(f"[vaegan{img} profile] {tot:.0f} us of kernels per step in {sum(r[1] for r in rows):.0f} launches", file=sys.stderr)
        for t, c, k in rows[:40]:
            print(f"[vaegan{img} profile] {t:9.1f} us {100 * t / tot:5.1f}%  n={c:5.1f}  {k.replace('void ', '').replace('vp::', '').replace('(anonymous namespace)::', '')[:120]}", file=sys.stderr)
    regs = []
    for _ in range(args.repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        regs.append(e0.elapsed_time(e1) / args.steps)
    ms = statistics.median(regs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        x.copy_(x_host, non_blocking=True)
        step()
        lh = static_loss.item()
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    gf = vaegan_flops_per_image(img) * 3 / 1e9
    ips = B / ms * 1e3
    out = {"workload": f"train.py full step: VaeGan({img},128) fwd (decoder x2, discriminator on 3B images in REC+GAN mode) + VaeGan.loss + summed-loss "
                       f"backward (== the five accumulating backward calls) + 4 x RMSprop, batch {B}",
           "value": round(ips, 1), "unit": "images/s", "ms_per_step": round(ms, 4), "e2e_value": round(B / ms_e2e * 1e3, 1),
           "gpu_launches_per_step": int(launches), "bf16_contractions_on_cuda_cores": int(_lib.simt_bf16_count() - simt0),
           "approx_gflop_per_image": round(gf, 1), "approx_step_tflops": round(ips * gf / 1e3, 1),
           "approx_frac_of_sustained_bf16": round(ips * gf / 1e3 / peaks["tf_sustained"], 4), "loss": lh,
           "note": "FLOPs = 3 x forward MACs x 2 (an upper bound: first-layer data gradients are not computed)"}
    VF.set_grad_sinks({})
    return out


def style_bench(args, dev, img=256, B=8, Z=64, ncls=3):
    """Config 5 of BASELINE.json: one train_random_gan iteration of train_Style_GAN.py (StyleEncoder + Generator with the
    label-gated dual convolutions + two-headed Discriminator, three Adam optimisers, three backward passes) at 256x256 RGB,
    launched from Python (no graph: the step interleaves optimiser steps with backward passes through retained graphs)."""
    import torch

    import vae_play_b200.functional as VF
    from vae_play_b200 import _lib
    from vae_play_b200 import train_steps as TS
    from vae_play_b200.models import network_Style_GAN as S
    from vae_play_b200.optim import FusedAdam
    torch.manual_seed(0)
    VF.set_async_wgrad(False)
    VF.set_grad_sinks({})
    G, E, D = S.Generator(img, Z).to(dev).train(), S.StyleEncoder(Z, img).to(dev).train(), S.Discriminator(img, ncls).to(dev).train()
    g_opt, e_opt, d_opt = (FusedAdam(list(m.parameters()), lr=1e-4, capturable=True) for m in (G, E, D))
    xt, xc = torch.rand(B, 3, img, img, device=dev), torch.rand(B, 3, img, img, device=dev)
    y = torch.randint(0, ncls, (B,), device=dev)
    eps, sz = torch.randn(B, Z, device=dev), torch.randn(B, Z, device=dev)
    simt0 = _lib.simt_bf16_count()

    def eager():
        return TS.style_gan_step(G, E, D, g_opt, e_opt, d_opt, xt, xc, y, eps, sz)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            eager()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    # the whole iteration -- three forward / backward groups with the three Adam steps in between -- as ONE CUDA graph: launched
    # from Python it is host-bound (2 300 launches ~ 57 ms per step at any image size)
    graphed = False
    step = eager
    try:
        VF.invalidate_caches()
        for o in (g_opt, e_opt, d_opt):
            o.zero_grad(set_to_none=True)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            static_losses = eager()

        def step():            # noqa: F811
            g.replay()
            return static_losses
        step()
        torch.cuda.synchronize()
        graphed = True
    except Exception as e:     # fall back to Python launches, say so in the line
        graph_error = f"{type(e).__name__}: {e}"[:200]
        step = eager
        torch.cuda.synchronize()
    if os.environ.get("VP_PROFILE_STYLE"):       # diagnostic: per-kernel device time of this iteration -> stderr
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
        rows = sorted(((e.device_time_total / 2, e.count / 2, e.key) for e in prof.key_averages() if e.device_time_total), reverse=True)
        tot = sum(r[0] for r in rows)
        print(f"[style{img} profile] {tot:.0f} us of kernels per step in {sum(r[1] for r in rows):.0f} launches", file=sys.stderr)
        for t, c, k in rows[:45]:
            print(f"[style{img} profile] {t:9.1f} us {100 * t / tot:5.1f}%  n={c:6.1f}  {k.replace('void ', '').replace('vp::', '').replace('(anonymous namespace)::', '')[:120]}", file=sys.stderr)
    l0 = _lib.launch_count()
    steps = max(3, min(args.steps, 10))
    regs = []
    for _ in range(max(1, min(args.repeats, 3))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            losses = step()
        e1.record()
        torch.cuda.synchronize()
        regs.append(e0.elapsed_time(e1) / steps)
    launches = (_lib.launch_count() - l0) / (steps * len(regs))
    if graphed:
        c0 = _lib.launch_count()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        launches = _lib.launch_count() - c0        # graph replays do not pass through the library's counter: one eager iteration does
    ms = statistics.median(regs)
    nparams = sum(p.numel() for m in (G, E, D) for p in m.parameters())
    return {"workload": f"train_Style_GAN.py train_random_gan iteration: StyleEncoder + Generator({img}) + Discriminator, 3 x Adam, {img}x{img}x3, batch {B}, "
                        + ("one CUDA graph per iteration" if graphed else f"launched from Python (graph capture failed: {graph_error})"),
            "value": round(B / ms * 1e3, 1), "unit": "images/s", "ms_per_step": round(ms, 3), "gpu_launches_per_step": int(launches),
            "bf16_contractions_on_cuda_cores": int(_lib.simt_bf16_count() - simt0), "parameters": int(nparams),
            "losses": {k: round(float(v), 5) for k, v in losses.items()}}


def gpu_library_baseline(img, cin, B, dev, steps):
    """CONTEXT, never on the product path: the reference's step expressed as torch.nn.functional calls (oracle/vae_torch.py, the
    line-by-line port of models/networks.py + train.py) executed by cuDNN / cuBLAS on this GPU, (a) fp32 storage with TF32 tensor
    cores and (b) bf16 autocast, channels_last input, torch.optim.RMSprop -- what stock PyTorch does for the same step."""
    import torch
    from oracle import vae_numpy as vn
    from oracle.vae_torch import VaeTorchPort
    out = {"what": "oracle/vae_torch.py (torch.nn.functional port of the reference step) on cuda via cuDNN/cuBLAS, same batch; context only"}
    P = vn.synth_vae_params(img, 128, cin, cin, 0)
    x = torch.rand(B, cin, img, img, device=dev)
    for name in ("tf32", "bf16_autocast"):
        try:
            torch.backends.cudnn.allow_tf32 = True
            torch.backends.cuda.matmul.allow_tf32 = True
            torch.backends.cudnn.benchmark = True
            port = VaeTorchPort({k: torch.as_tensor(v).to(dev) for k, v in P.items()}, torch.float32)
            xin = x.contiguous(memory_format=torch.channels_last)

            def one():
                if name == "tf32":
                    port.step(xin)
                else:
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        for k in port.train_keys:
                            port.P[k].grad = None
                        loss, _ = port.forward_loss(xin)
                    loss.backward()
                    port.opt.step()
            for _ in range(5):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": round(B / ms * 1e3, 1), "unit": "images/s", "ms_per_step": round(ms, 4)}
            del port
        except Exception as e:       # a baseline that cannot run must not take the bench line down with it
            out[name] = {"error": f"{type(e).__name__}: {e}"[:200]}
    torch.backends.cudnn.benchmark = False
    torch.cuda.empty_cache()
    return out


def hbm_probe(model, B, dev, peaks, iters=20):
    """The HBM-bound kernels of the step, each timed alone (CUDA events around a graph of `iters` launches): algorithmic bytes
    (what the pass must read + write once), us, GB/s and the fraction of the measured copy bandwidth.
    Inputs are larger than the 126 MB L2 for the big layers; for the small ones the number is an L2-assisted upper bound."""
    import ctypes as C

    import torch

    import vae_play_b200.functional as VF
    from vae_play_b200 import _lib
    from vae_play_b200.optim import FusedRMSprop
    dt = VF.act_dtype()
    es = 2 if dt == torch.bfloat16 else 4
    rows_out = []

    def add(name, nbytes, fn):
        us = _graph_time_us(fn, iters)
        gbs = nbytes / us / 1e3
        rows_out.append({"kernel": name, "bytes": int(nbytes), "us": round(us, 1), "gbs": round(gbs, 1), "frac": round(gbs / peaks["hbm_gbs"], 3)})

    ptr, stream = VF._ptr, VF._stream
    img = 8 * 2 ** len(list(model.encoder.conv))
    # BatchNorm streaming passes on the largest activation (last DecoderBlock output) and on all six conv layers together
    shapes = []
    hh = img
    for blk in model.encoder.conv:
        hh //= 2
        shapes.append((B * hh * hh, blk._layer.cout))
    hh = 8
    for blk in list(model.decoder.conv)[:-1]:
        hh *= 2
        shapes.append((B * hh * hh, blk._layer.cout))
    bufs = []
    for rows, c in shapes:
        y = torch.randn(rows, c, device=dev).to(dt)
        da = torch.randn(rows, c, device=dev).to(dt)
        out = torch.empty_like(y)
        st = torch.rand(4, c, device=dev) + 0.5
        sums = torch.zeros(2 * c, dtype=torch.float64, device=dev)
        bufs.append((rows, c, y, da, out, st, sums))

    def apply_all(sel):
        for rows, c, y, da, out, st, sums in sel:
            _lib.call("vp_norm_apply_act", ptr(y), ptr(st[2]), ptr(st[3]), ptr(out), VF._code(dt), 1, rows, c, 1, 0.0, stream())

    def reduce_all(sel):
        for rows, c, y, da, out, st, sums in sel:
            _lib.call("vp_norm_bwd_reduce", ptr(y), ptr(da), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), ptr(sums), None, VF._code(dt), 1, rows, c, 1, 0.0, stream())

    def bapply_all(sel):
        for rows, c, y, da, out, st, sums in sel:
            _lib.call("vp_norm_bwd_apply", ptr(y), ptr(da), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), ptr(sums), ptr(out), None, None,
                      VF._code(dt), 1, rows, c, 1, 0.0, stream())
    elems_all = sum(r * c for r, c, *_ in bufs)
    big = [max(bufs, key=lambda b: b[0] * b[1])]
    elems_big = big[0][0] * big[0][1]
    add("norm_stream apply (largest layer)", 2 * es * elems_big, lambda: apply_all(big))
    add("norm_stream bwd-reduce (largest layer)", 2 * es * elems_big, lambda: reduce_all(big))
    add("norm_stream bwd-apply (largest layer)", 3 * es * elems_big, lambda: bapply_all(big))
    add("norm_stream apply (all 6 BatchNorm2d layers)", 2 * es * elems_all, lambda: apply_all(bufs))
    add("norm_stream bwd-reduce (all 6)", 2 * es * elems_all, lambda: reduce_all(bufs))
    add("norm_stream bwd-apply (all 6)", 3 * es * elems_all, lambda: bapply_all(bufs))
    del bufs, big
    # optimiser: p, g, sq read; p, sq, g(zero), bf16 copy written
    params = [p for p in list(model.encoder.parameters()) + list(model.decoder.parameters())]
    ps = [torch.nn.Parameter(p.detach().clone()) for p in params]
    for p in ps:
        p.grad = torch.randn_like(p) * 1e-3
    opt = FusedRMSprop(ps, lr=1e-6, zero_grads=True)
    n = sum(p.numel() for p in ps)
    n_sh = sum(p.numel() for p in params if VF._SHADOWS.get(p.data_ptr()) is not None)
    add("rmsprop (fp32 masters + bf16 copies + gradient clearing)", 24 * n + 2 * n_sh, lambda: opt.step())
    del opt, ps
    # thin layers
    first, last = model.encoder.conv[0], list(model.decoder.conv)[-1][0]
    lf, ll = first._layer, model.decoder._out_layer
    x1 = torch.randn(B, img, img, 1, device=dev).to(dt)
    y1 = lf.fwd(x1, first.conv.weight.detach(), None)
    dy1 = torch.randn_like(y1)
    add("thin first-layer fwd (1->64)", x1.numel() * es + y1.numel() * es, lambda: lf.fwd(x1, first.conv.weight.detach(), None))
    add("thin first-layer wgrad", x1.numel() * es + y1.numel() * es, lambda: lf.wgrad(x1, dy1, first.conv.weight.detach()))
    xl = torch.randn(B, img, img, ll.cin, device=dev).to(dt)
    yl = ll.fwd(xl, last.weight.detach(), last.bias.detach(), "sigmoid")
    dyl = torch.randn_like(yl)
    add("thin last-layer fwd (64->1, +bias+sigmoid)", xl.numel() * es + yl.numel() * es, lambda: ll.fwd(xl, last.weight.detach(), last.bias.detach(), "sigmoid"))
    add("thin last-layer dgrad", xl.numel() * es + yl.numel() * es, lambda: ll.dgrad(dyl, last.weight.detach(), tuple(xl.shape)))
    add("thin last-layer wgrad", xl.numel() * es + yl.numel() * es, lambda: ll.wgrad(xl, dyl, last.weight.detach()))
    del x1, y1, dy1, xl, yl, dyl
    # reconstruction loss (x fp32, x~ fp32) and the 8x8 map transposes next to the fc layers
    xa, xb = torch.rand(B, 1, img, img, device=dev), torch.rand(B, 1, img, img, device=dev).requires_grad_(True)
    add("recon mse fwd", 8 * xa.numel(), lambda: VF.mse_loss(xa, xb))
    cmax = list(model.encoder.conv)[-1]._layer.cout
    t = torch.randn(B, 64, cmax, device=dev).to(dt)
    add("transpose [B,64,C]->[B,C,64]", 2 * es * t.numel(), lambda: VF._TransposeBT.apply(t, 64, cmax))
    return {"peak_gbs": peaks["hbm_gbs"], "peak_kind": f"{peaks['src']} copy bandwidth", "kernels": rows_out}


def _graph_time_us(fn, iters):
    """Device time per call: `iters` calls captured into one CUDA graph, replayed once warm and once timed with CUDA
    events on the replaying stream (no host launch overhead inside the timed region)."""
    import torch
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def kernel_probe(model, B, dev, peaks, iters=20):
    """Every contraction of the step timed alone (CUDA events around a graph of `iters` launches) and the dominant one --
    the last DecoderBlock's transposed conv forward, the single longest launch of the step -- reported as `roofline`."""
    import torch

    import vae_play_b200.functional as VF
    dt = VF.act_dtype()
    rows = []

    def add(name, layer, weight, hin):
        x = torch.randn(B, hin, hin, layer.cin, device=dev).to(dt)
        w = weight.detach()
        y = layer.fwd(x, w, None)
        dy = torch.randn_like(y)
        mac = B * (hin * hin if layer.kind != "conv" else y.shape[1] * y.shape[2]) * layer.cin * layer.cout * layer.k * layer.k
        gf = 2.0 * mac / 1e9
        t_f = _graph_time_us(lambda: layer.fwd(x, w, None), iters)
        t_d = _graph_time_us(lambda: layer.dgrad(dy, w, tuple(x.shape)), iters) if layer.cin > 1 else None
        t_w = _graph_time_us(lambda: layer.wgrad(x, dy, w), iters)
        tf = lambda t: round(gf / t * 1e3, 1) if t else None
        rows.append({"layer": name, "gflop": round(gf, 2), "fwd_us": round(t_f, 1), "fwd_tflops": tf(t_f),
                     "dgrad_us": round(t_d, 1) if t_d else None, "dgrad_tflops": tf(t_d), "wgrad_us": round(t_w, 1), "wgrad_tflops": tf(t_w)})
        return gf, t_f

    img = 8 * 2 ** len(list(model.encoder.conv))
    hh = img
    for i, blk in enumerate(model.encoder.conv):
        add(f"encoder.conv.{i}", blk._layer, blk.conv.weight, hh)
        hh //= 2
    add("encoder.fc", model.encoder._fc_layer, model.encoder.fc[0].weight, 1)
    add("decoder.fc", model.decoder._fc_layer, model.decoder.fc[0].weight, 1)
    hh = 8
    dom = None
    blocks = list(model.decoder.conv)
    for i, blk in enumerate(blocks[:-1]):
        gf, t_f = add(f"decoder.conv.{i}", blk._layer, blk.conv.weight, hh)
        dom = (f"decoder.conv.{i} ConvTranspose2d 5x5 s2 forward ({blk._layer.cin}->{blk._layer.cout} ch, {hh}x{hh}->{2*hh}x{2*hh}, one launch)", gf, t_f)
        hh *= 2
    add(f"decoder.conv.{len(blocks)-1}", model.decoder._out_layer, blocks[-1][0].weight, hh)
    name, gf, t_f = dom
    tflops = gf / t_f * 1e3
    total_us = sum((r["fwd_us"] or 0) + (r["dgrad_us"] or 0) + (r["wgrad_us"] or 0) for r in rows)
    total_gf = sum(r["gflop"] * (3 if r["dgrad_us"] else 2) for r in rows)
    return {"bound": "tensor", "kernel": name, "achieved": round(tflops, 2), "peak": peaks["tf_burst"], "unit": "TFLOP/s",
            "frac": round(tflops / peaks["tf_burst"], 4), "traffic": DOMINANT_TRAFFIC_BYTES.get(img),
            "us_per_launch": round(t_f, 2), "flop_per_launch": gf * 1e9, "peak_kind": f"{peaks['src']} burst",
            "all_contractions": {"sum_us": round(total_us, 1), "gflop": round(total_gf, 1), "tflops": round(total_gf / total_us * 1e3, 1),
                                 "frac_of_burst": round(total_gf / total_us * 1e3 / peaks["tf_burst"], 4), "layers": rows}}


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the ncu --set full capture
# summarised in profiles/ (per image size); None until captured
DOMINANT_TRAFFIC_BYTES = {64: 222577920}     # profiles/r02_ncu_ct3_fwd.summary.txt: 132.07 MB read + 90.51 MB written


def add_arguments(ap):
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--img", type=int, default=64)
    ap.add_argument("--cin", type=int, default=1)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--ref-batch", type=int, default=16)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of --steps steps each; the median region is reported")
    ap.add_argument("--no-extras", action="store_true", help="skip the other image size, the HBM-kernel table and the GPU-library baseline")
    ap.add_argument("--extra-warmup", type=int, default=60, help="additional untimed steps so clocks ramp up and the sampler is live")
    ap.add_argument("--torch-optim", action="store_true", help="use torch.optim.RMSprop instead of the fused multi-tensor kernel")
    ap.add_argument("--no-bucket-pipeline", action="store_true", help="data parallel: one optimiser launch after all all-reduces")
    ap.add_argument("--no-split-backward", action="store_true",
                    help="data parallel: do NOT cut the backward graph at the encoder conv stack (all-reduce fully exposed between backward and optimiser)")
    ap.add_argument("--grad-wire", default="bf16", choices=["bf16", "fp32"],
                    help="data parallel: gradient buckets cross NVLink in bf16 (half the bytes; the optimiser reads the reduced bf16 values) or fp32")
    ap.add_argument("--sm-reserve", type=int, default=32, help="data parallel: SMs left to NCCL while the all-reduce overlaps the encoder-conv backward")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--no-overlap-opt", action="store_true", help="one GPU: optimiser as its own graph after the backward (default: its larger part runs next to the encoder-conv backward)")
    ap.add_argument("--three-stage-backward", default="auto", choices=["auto", "on", "off"],
                    help="data parallel: cut the backward once more between decoder and sample so that the decoder's gradient exchange starts a stage earlier (auto: from 8 ranks on)")
    ap.add_argument("--no-async-wgrad", action="store_true", help="weight gradients on the main stream (default: a side stream, overlapping the BatchNorm-backward passes)")


def main():
    ap = argparse.ArgumentParser()
    add_arguments(ap)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
