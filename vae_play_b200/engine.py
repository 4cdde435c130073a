"""The VAE training step as one object: model, fused optimiser(s), persistent gradient slots, CUDA graph(s), the side streams of
the backward pass and -- with ``world > 1`` -- the bucketed gradient exchange.  ``VaeTrainer(...).step(x)`` is what
``bench.py`` times and what a training loop would call once per batch (reference: the loop body of train.py:38-78 for the
stock VAE path -- Encoder, reparameterize, Decoder, F.mse_loss + KL, backward, RMSprop).

One GPU: ONE CUDA graph per step with three streams ("Streams").  Data parallel: two backward graphs with
the exchange stream between them.  ``torch.distributed`` must be initialised by the caller
(``init_nccl`` below sets NCCL up with a bounded CTA count on a high-priority stream).
"""
from __future__ import annotations

import types

import torch


def init_nccl(sm_reserve: int, dev):
    """NCCL's kernels get a bounded number of CTAs (the SMs the persistent grids leave free while the exchange overlaps backward)
    and a high-priority stream (their CTAs are placed first when SMs free up)."""
    import torch.distributed as dist
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        opts.config.max_ctas = sm_reserve
        opts.config.min_ctas = min(sm_reserve, 4)
        dist.init_process_group("nccl", device_id=dev, pg_options=opts)
    except Exception:
        dist.init_process_group("nccl", device_id=dev)


class VaeTrainer:
    """One model + optimiser + (CUDA-graph) step closure for a given image size; ``step(x)`` runs one full training step and
    returns the (device) loss.  Options: see ``default_options`` (same names as bench.py's flags)."""

    @staticmethod
    def default_options():
        return dict(batch=256, cin=1, grad_wire="bf16", no_async_wgrad=False, no_bucket_pipeline=False, no_graph=False, no_overlap_opt=False,
                    no_split_backward=False, sm_reserve=32, torch_optim=False, three_stage_backward="auto")


    def __init__(self, opts=None, img=64, world=1, rank=0, dev=None, **overrides):
        import torch
        import torch.distributed as dist

        import vae_play_b200.functional as VF
        from vae_play_b200 import _lib
        from vae_play_b200.models.networks import VaeGan
        from vae_play_b200.parallel import GradBuckets
        args = types.SimpleNamespace(**{**self.default_options(), **(vars(opts) if opts is not None else {}), **overrides})
        dev = dev if dev is not None else torch.device("cuda", torch.cuda.current_device())
        self.args, self.img, self.world, self.rank, self.dev = args, img, world, rank, dev
        B, cin = args.batch, args.cin
        self.B = B
        torch.manual_seed(0)
        model = VaeGan(img, 128).to(dev).train()
        self.model = model
        params = list(model.encoder.parameters()) + list(model.decoder.parameters())
        use_graph = not args.no_graph
        if args.torch_optim:
            opt = torch.optim.RMSprop(params, lr=1e-4, capturable=use_graph)
        else:
            from vae_play_b200.optim import FusedRMSprop          # same update rule, one multi-tensor kernel
            # the kernel also refreshes the bf16 operand copies of the weights and clears each gradient after use, so the
            # next step's weight-gradient kernels accumulate into known-zero persistent slots (no cast pass, no memsets)
            opt = FusedRMSprop(params, lr=1e-4, zero_grads=True)
        wire_bf16 = world > 1 and args.grad_wire == "bf16" and not args.torch_optim
        # the decoder's gradients get buckets of their own: their exchange starts one backward stage earlier (see step())
        enc_last = list(model.encoder.parameters())[-1]
        # "auto": from 8 ranks on.  Measured: up to 4 ranks the decoder bucket's all-reduce (47-74 us) still fits under the encoder-conv
        # backward behind the big bucket and the extra graph costs ~11 us (N = 2: 1.820 vs 1.809 ms); at 8 ranks it does not fit
        # (110 us behind 214 us against ~240 us of stage 2: profiles/r02_dp8_timeline.txt) and would move under stage 1b
        three_stage = (world >= 8) if args.three_stage_backward in ("auto", None) else args.three_stage_backward in (True, "on")
        buckets = (GradBuckets(params, world, overlap=not use_graph, wire_dtype=torch.bfloat16 if wire_bf16 else None,
                               breaks=[enc_last] if three_stage else ())
                   if world > 1 else None)
        if buckets is None and not args.torch_optim:
            VF.persistent_grads(params)
        if wire_bf16:
            # bf16 on the wire: pack() converts + clears the fp32 buckets, the optimiser reads the reduced bf16 values directly
            views = {}
            for bi in range(len(buckets.buckets)):
                views.update(buckets.wire_views(bi))
            opt = FusedRMSprop(params, lr=1e-4, zero_grads=False, wire=views)
        # data parallel: one optimiser per gradient bucket, so that the update of bucket i runs while NCCL reduces bucket i+1
        bucket_opts = None
        if buckets is not None and not args.torch_optim and not args.no_bucket_pipeline:
            from vae_play_b200.optim import FusedRMSprop
            bucket_opts = [FusedRMSprop(b["params"], lr=1e-4, zero_grads=not wire_bf16, wire=buckets.wire_views(bi))
                           for bi, b in enumerate(buckets.buckets)]
        self.wire = "bf16" if wire_bf16 else "fp32"
        torch.manual_seed(1234 + rank)
        self.x_host = torch.rand(B, cin, img, img).pin_memory()
        self.x_dev = self.x_host.to(dev)
        off_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        _, eps_inc = VF.philox_policy(B * 128, torch.cuda.get_device_properties(dev).multi_processor_count)

        # weight gradients on a side stream: the tensor-bound wgrad kernel of block L runs next to the bandwidth-bound
        # BatchNorm-backward passes of block L-1 (functional.set_async_wgrad); joined at the end of every backward (stage)
        VF.set_async_wgrad(not args.no_async_wgrad and not args.torch_optim)

        def fwd_bwd(x):
            # disjoint, reproducible Philox streams per rank: seed = rank; the offset lives on the device and
            # advances by what Tensor.normal_() on B*z elements would consume, so CUDA-graph replays draw fresh eps
            xt, mulv, kl = model.vae_forward(x, rng=(rank, 0, off_dev))
            VF.philox_advance(off_dev, eps_inc)
            loss = VF.vae_loss(x, xt, kl, mse_scale=1.0 / world)
            loss.backward()
            VF.join_async()
            return loss

        # Data parallel, graph mode: the backward is cut at the output of the encoder's conv stack.  Stage 1 (decoder, heads,
        # encoder.fc: 94 % of the gradient bytes) and stage 2 (the encoder convs) are separate graphs, and the all-reduce of
        # the stage-1 buckets runs on NCCL's stream while stage 2 executes.
        enc_conv_params = [p for blk in model.encoder.conv for p in blk.parameters()]
        enc_conv_ids = {id(p) for p in enc_conv_params}
        stage1_params = [p for p in params if id(p) not in enc_conv_ids]
        cut = {}

        def fwd_bwd_stage1(x):
            taps = []
            xt, mulv, kl = model.vae_forward(x, rng=(rank, 0, off_dev), taps=taps)
            VF.philox_advance(off_dev, eps_inc)
            loss = VF.vae_loss(x, xt, kl, mse_scale=1.0 / world)
            a3 = taps[0]          # behind VF.grad_cut: naming it in `inputs` executes only that identity node
            a3.register_hook(lambda g: cut.__setitem__("g", g))
            loss.backward(inputs=stage1_params + [a3], retain_graph=True)
            VF.join_async()
            a3.grad = None        # `inputs` also accumulated it into .grad
            cut["a"] = taps[1]    # stage 2 starts one identity node further in: no retained .grad to clone or add into
            return loss

        def bwd_stage2():
            cut["a"].backward(cut["g"], inputs=enc_conv_params)
            VF.join_async()

        # Three-stage variant (data parallel, ``three_stage_backward``: "auto" = from 8 ranks on): stage 1 is cut once more between
        # the decoder and the sample z.
        #   1a: forward, loss, decoder backward          -> the decoder buckets are complete: their exchange starts
        #   1b: sample / KL / heads / encoder.fc backward  -> encoder.fc.0.weight (73 % of the bytes) is complete
        #   2 : encoder-conv backward
        # The three backward calls are gradient-equivalent to one (test_three_stage_backward_matches_single).
        dec_params = list(model.decoder.parameters())
        dec_ids = {id(p) for p in dec_params}
        stage1b_params = [p for p in stage1_params if id(p) not in dec_ids]
        ones_kl = torch.ones(B, dtype=torch.float32, device=dev)          # d loss / d kl_b (functional._VaeLoss: loss = s * mse + sum_b kl_b)

        def fwd_bwd_stage1a(x):
            taps = []
            xt, mulv, kl = model.vae_forward(x, rng=(rank, 0, off_dev), taps=taps)
            VF.philox_advance(off_dev, eps_inc)
            loss = VF.vae_loss(x, xt, kl, mse_scale=1.0 / world)
            z_out = taps[2]
            z_out.register_hook(lambda g: cut.__setitem__("gz", g))
            loss.backward(inputs=dec_params + [z_out], retain_graph=True)
            VF.join_async()
            z_out.grad = None
            cut["z"], cut["kl"], cut["a_out"], cut["a_in"] = taps[3], kl, taps[0], taps[1]
            return loss

        def bwd_stage1b():
            a3 = cut["a_out"]
            a3.register_hook(lambda g: cut.__setitem__("g", g))
            torch.autograd.backward([cut["z"], cut["kl"]], [cut["gz"], ones_kl], inputs=stage1b_params + [a3], retain_graph=True)
            VF.join_async()
            a3.grad = None
            cut["a"] = cut["a_in"]

        def eager_step(x):
            opt.zero_grad(set_to_none=True)
            loss = fwd_bwd(x)
            if buckets is not None:
                buckets.allreduce()
            if bucket_opts is not None:
                for o in bucket_opts:
                    o.step()
            else:
                opt.step()
            return loss

        # One GPU: the optimiser is split in two and its larger part (everything but the encoder convs: 97 % of the bytes) is
        # launched on its own stream as soon as those gradients are final -- a bandwidth-bound kernel next to the tensor-bound
        # encoder-conv backward.  The whole step is then ONE CUDA graph.
        overlap_opt = use_graph and buckets is None and not args.torch_optim and not args.no_overlap_opt
        if overlap_opt:
            from vae_play_b200.optim import FusedRMSprop
            opt1 = FusedRMSprop(stage1_params, lr=1e-4, zero_grads=True)
            opt2 = FusedRMSprop(enc_conv_params, lr=1e-4, zero_grads=True)
            opt_stream = torch.cuda.Stream()

            def eager_step(x):            # noqa: F811 -- same step, the optimiser in two parts
                main = torch.cuda.current_stream()
                opt1.zero_grad(set_to_none=True)                          # host-side only: the slots are handed out afresh
                opt2.zero_grad(set_to_none=True)
                loss = fwd_bwd_stage1(x)                                  # ends with join_async(): every stage-1 gradient is final
                opt_stream.wait_stream(main)
                with torch.cuda.stream(opt_stream):
                    opt1.step()
                bwd_stage2()
                opt2.step()
                main.wait_stream(opt_stream)
                return loss

        graph_a = graph_b = graph_a1 = graph_a2 = None
        early_a, early_b = [], []
        graph_bs = []
        early = []
        extra_launches = 0
        xstream = torch.cuda.Stream() if buckets is not None else None
        # data parallel: on by default -- the backward graph is cut after encoder.fc's weight gradient (94 % of the gradient
        # bytes are complete there) and the encoder-conv backward that follows is captured with its persistent grids capped at
        # (#SMs - sm_reserve), so that NCCL's CTAs find free SMs and the all-reduce really runs next to it
        split_backward = use_graph and buckets is not None and not args.no_split_backward
        static_x = self.x_dev.clone()
        self.static_x = static_x
        launches_per_replay = launches_opt = 0
        if use_graph:
            # two CUDA graphs per step: A = forward + loss + backward (gradients land in the flat buckets),
            # B = optimiser; the bucketed NCCL all-reduce runs between them, outside the captured regions
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    eager_step(static_x)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            VF.invalidate_caches()
            opt.zero_grad(set_to_none=True)
            graph_a, graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            if overlap_opt:
                with torch.cuda.graph(graph_a):
                    static_loss = eager_step(static_x)
                graph_b = None
            elif split_backward:
                by_size = lambda ids: sorted(ids, key=lambda i: -buckets.buckets[i]["buf"].numel())
                early_a = by_size(buckets.buckets_within(dec_params))
                early = by_size(buckets.buckets_within(stage1_params))
                early_b = [i for i in early if i not in early_a]
                graph_a1, graph_a2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                late = [i for i in range(len(buckets.buckets)) if i not in early]
                if not three_stage:
                    early_a, early_b, graph_a1 = [], early, None
                with torch.cuda.graph(graph_a):
                    static_loss = fwd_bwd_stage1a(static_x) if three_stage else fwd_bwd_stage1(static_x)
                l1 = _lib.launch_count()
                buckets.pack(early)              # per step this runs on the exchange stream, next to stages 1b / 2 (see step())
                extra_launches = _lib.launch_count() - l1
                nsm = torch.cuda.get_device_properties(dev).multi_processor_count
                _lib.call("vp_set_sm_limit", max(nsm - args.sm_reserve, nsm // 2))
                try:
                    if three_stage:
                        with torch.cuda.graph(graph_a1, pool=graph_a.pool()):
                            bwd_stage1b()
                    with torch.cuda.graph(graph_a2, pool=graph_a.pool()):
                        bwd_stage2()
                        buckets.pack(late)
                finally:
                    _lib.call("vp_set_sm_limit", 0)
            else:
                with torch.cuda.graph(graph_a):
                    static_loss = fwd_bwd(static_x)
                    if buckets is not None:
                        buckets.pack()
            launches_per_replay = _lib.launch_count() - l0
            if buckets is not None:
                buckets.allreduce_subset(range(len(buckets.buckets)), pre_packed=True)
                buckets.allreduce(check_missing=False)
            l0 = _lib.launch_count()
            if overlap_opt:
                pass
            elif bucket_opts is not None:
                graph_bs = []
                for o in bucket_opts:
                    gb_ = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gb_, pool=graph_a.pool()):
                        o.step()
                    graph_bs.append(gb_)
            else:
                with torch.cuda.graph(graph_b, pool=graph_a.pool()):
                    opt.step()
            launches_opt = _lib.launch_count() - l0
        # Everything the captured graphs address must outlive them.  Most of it stays reachable through the `step` closure, but the
        # capture-only closures die with __init__, and a freed block of the regular allocator is handed to the next small
        # allocation (found the hard way: `ones_kl`, 1 KB, referenced only by bwd_stage1b, was overwritten by the first
        # torch.tensor() after construction and every gradient upstream of the sample exploded)
        self._keep_alive = (off_dev, ones_kl, cut, static_x)
        self.graph = graph_a is not None
        self.overlap_opt = bool(overlap_opt)
        self.async_wgrad = not args.no_async_wgrad and not args.torch_optim
        self.split = graph_a2 is not None
        self.three_stage = graph_a1 is not None
        self.launches_per_step = (launches_per_replay + launches_opt + extra_launches) if self.graph else None

        def step(x):
            if graph_a is None:
                return eager_step(x)
            if x.data_ptr() != static_x.data_ptr():
                static_x.copy_(x, non_blocking=True)
            graph_a.replay()
            if graph_b is None:
                return static_loss
            main = torch.cuda.current_stream()
            if graph_a2 is not None:
                # exchange stream: bf16 packing of the stage-1 buckets, then their all-reduce -- all of it next to stage 2
                if early_a:
                    xstream.wait_stream(main)
                    with torch.cuda.stream(xstream):
                        for bi in early_a:                     # the decoder's gradients: next to stage 1b
                            buckets.pack([bi])
                            buckets.allreduce_subset([bi], pre_packed=True)
                if graph_a1 is not None:
                    graph_a1.replay()
                xstream.wait_stream(main)
                with torch.cuda.stream(xstream):
                    for bi in early_b:                     # largest first (encoder.fc's weight: 73 % of the bytes): next to stage 2
                        buckets.pack([bi])
                        buckets.allreduce_subset([bi], pre_packed=True)
                graph_a2.replay()
            if bucket_opts is not None:
                buckets.allreduce_subset(range(len(buckets.buckets)), pre_packed=True)      # the rest, queued on NCCL's stream in order
                with torch.cuda.stream(xstream):
                    for bi in early_a + early_b:                            # their updates too run next to stage 2 / the late exchange:
                        buckets.wait_bucket(bi)                             # nothing left in the step reads those weights
                        graph_bs[bi].replay()
                for bi, gb_ in enumerate(graph_bs):
                    if bi not in early:
                        buckets.wait_bucket(bi)
                        gb_.replay()
                main.wait_stream(xstream)
                return static_loss
            if buckets is not None:
                buckets.allreduce_subset(range(len(buckets.buckets)), pre_packed=True)
                buckets.allreduce(check_missing=False)
            graph_b.replay()
            return static_loss

        self.step = step

