"""Data-parallel gradient exchange for the VAE step: one process per GPU, batch sharded per rank,
parameters replicated, ONE sum all-reduce of the gradients per step in a few large buckets.

The reference is single-GPU; batch sharding is the path's only natural
partition (section 8e).  Buckets are flat fp32 buffers filled in reverse layer order; the wgrad
kernels write parameter gradients straight into their bucket slot (``functional.set_grad_sinks``), a
post-accumulate hook marks the slot ready, and as soon as a bucket is complete its NCCL all-reduce
is issued asynchronously so that it overlaps the rest of backward.  BatchNorm statistics stay
per-rank (DDP semantics), so a rank reproduces the single-GPU reference at its own batch.

Loss scaling: with L = mean-type mse + sum-type KL, the single-process equivalent of W ranks is
L = (1/W) sum_r mse_r + sum_r kl_r, i.e. each rank scales its mse term by 1/W and gradients are
summed -- no post-scaling pass.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import List

import torch
import torch.distributed as dist


class GradBuckets:
    def __init__(self, params: List[torch.nn.Parameter], world_size: int, bucket_mb: float = 32.0, group=None,
                 overlap: bool = True, wire_dtype=None, breaks=()):
        """overlap=True: each bucket's all-reduce is issued from the gradient hook as soon as the bucket is
        complete (overlaps the rest of backward).  overlap=False: hooks only place gradients into the buckets and
        ``allreduce()`` issues all collectives afterwards -- used when forward+backward is replayed as a CUDA
        graph (the collectives stay outside the captured region)."""
        self.params = [p for p in params if p.requires_grad]
        self.world = world_size
        self.group = group
        self.overlap = overlap
        # wire_dtype=torch.bfloat16: the buckets are exchanged in bf16 (half the NVLink bytes).  pack() converts a bucket
        # (and clears its fp32 slots for the next step) in one pass; the optimiser then reads the reduced bf16 values directly
        # (FusedRMSprop(wire=...)).  None: fp32 on the wire, bit-identical to summing the per-rank gradients in fp32.
        if wire_dtype not in (None, torch.float32, torch.bfloat16):
            raise ValueError("wire_dtype must be None, torch.float32 or torch.bfloat16")
        self.wire_dtype = wire_dtype if wire_dtype == torch.bfloat16 else None
        cap = int(bucket_mb * (1 << 20) / 4)
        # reverse registration order ~ the order gradients become ready in backward
        order = list(reversed(self.params))
        self.buckets = []          # list of dict(buf, params, ready, handle)
        cur, cur_n = [], 0
        # breaks: parameters that must START a bucket (in reverse registration order): lets a caller that cuts its backward into
        # stages keep each stage's gradients in buckets of their own (vae_play_b200.engine)
        break_ids = {id(p) for p in breaks}
        for p in order:
            if cur and (cur_n + self._padded(p) > cap or id(p) in break_ids):
                self._close(cur)
                cur, cur_n = [], 0
            cur.append(p)
            cur_n += self._padded(p)
        if cur:
            self._close(cur)
        self.slot = {}
        for bi, b in enumerate(self.buckets):
            off = 0
            for p in b["params"]:
                # same strides as the parameter (conv weights live in channels-last order): the slot IS the gradient
                view = torch.as_strided(b["buf"], p.shape, p.stride(), storage_offset=off)
                self.slot[id(p)] = (bi, view, off)
                off += self._padded(p)
        self._sync = True
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self._install_sinks()

    @contextmanager
    def no_sync(self):
        """Backward calls inside this context only accumulate into the buckets (no readiness counting, no collective), like
        ``DistributedDataParallel.no_sync``.  The reference step runs five ``backward(retain_graph=True)`` calls
        (train.py:68-73): wrap the first four, run the last one outside, then call ``allreduce()``.  With ``overlap=True``
        exactly ONE backward per step may run outside ``no_sync`` -- a gradient arriving for a bucket whose all-reduce is
        already in flight raises instead of silently diverging the ranks."""
        old, self._sync = self._sync, False
        try:
            yield self
        finally:
            self._sync = old

    @staticmethod
    def _padded(p):
        """slot size in elements: every slot starts 16-byte aligned (vector stores / reductions of the wgrad kernels)"""
        return (p.numel() + 3) & ~3

    def _close(self, plist):
        n = sum(self._padded(p) for p in plist)
        dev = plist[0].device
        self.buckets.append({"buf": torch.zeros(n, dtype=torch.float32, device=dev), "params": list(plist),
                             "ready": 0, "handle": None, "packed": False,
                             "wire": torch.zeros(n, dtype=torch.bfloat16, device=dev) if self.wire_dtype is not None else None})

    def _install_sinks(self):
        try:
            from . import functional as VF
            VF.set_grad_sinks({p.data_ptr(): (self.buckets[self.slot[id(p)][0]]["buf"], self.slot[id(p)][2], p) for p in self.params})
        except Exception:  # pragma: no cover - CPU-only host-logic tests
            pass

    # ---- bf16 wire format ---------------------------------------------------------------------------------------------
    def pack(self, bucket_ids=None):
        """fp32 bucket -> bf16 wire buffer, clearing the fp32 slots in the same pass (vp_pack_grads_bf16), stream-ordered on
        the current stream.  No-op with an fp32 wire.  Called by the hooks (eager) or by the caller inside its captured graph."""
        if self.wire_dtype is None:
            return
        self._join_side_stream()
        import ctypes as C
        from . import _lib
        from . import functional as VF
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for bi in (range(len(self.buckets)) if bucket_ids is None else bucket_ids):
            b = self.buckets[bi]
            _lib.call("vp_pack_grads_bf16", C.c_void_p(b["buf"].data_ptr()), C.c_void_p(b["wire"].data_ptr()), b["buf"].numel(), 1, stream)
            b["packed"] = True
            VF.sinks_zeroed(b["params"])

    def wire_views(self, bucket_id):
        """param -> bf16 view of its reduced gradient inside the wire buffer (what FusedRMSprop(wire=...) reads), or None."""
        if self.wire_dtype is None:
            return None
        b = self.buckets[bucket_id]
        return {id(p): torch.as_strided(b["wire"], p.shape, p.stride(), storage_offset=self.slot[id(p)][2]) for p in b["params"]}

    @staticmethod
    def _join_side_stream():
        if torch.cuda.is_available():
            from . import functional as VF           # weight gradients still running on functional's side stream
            VF.join_async()

    def _reduce(self, b):
        self._join_side_stream()
        if self.wire_dtype is not None and not b["packed"]:
            self.pack([next(i for i, x in enumerate(self.buckets) if x is b)])
        buf = b["wire"] if self.wire_dtype is not None else b["buf"]
        b["packed"] = False
        return dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    # ---- hook: called once per parameter per backward, after .grad has been accumulated -----------
    def _on_grad(self, p):
        bi, view, _ = self.slot[id(p)]
        g = p.grad
        if g.data_ptr() != view.data_ptr():
            # gradient was produced elsewhere (e.g. accumulated over several backward calls): copy in
            view.copy_(g)
            p.grad = view
        b = self.buckets[bi]
        if b["handle"] is not None:
            raise RuntimeError("GradBuckets: a gradient arrived for a bucket whose all-reduce is already in flight; run a single "
                               "backward per step, or wrap the accumulating backward calls in no_sync()")
        if not self._sync:
            return
        b["ready"] += 1
        if b["ready"] == len(b["params"]) and self.world > 1 and self.overlap:
            b["handle"] = self._reduce(b)

    def allreduce(self, check_missing: bool = True):
        """Finish the step's exchange: launch whatever is still pending, wait for everything."""
        for b in self.buckets:
            if self.world > 1 and b["handle"] is None:
                # some parameter of this bucket received no gradient this step: its slot must not carry stale data
                for p in b["params"]:
                    if check_missing and p.grad is None:
                        self.slot[id(p)][1].zero_()
                b["handle"] = self._reduce(b)
        for b in self.buckets:
            if b["handle"] is not None:
                b["handle"].wait()
                b["handle"] = None
            b["ready"] = 0

    def allreduce_subset(self, bucket_ids, pre_packed=False):
        """Issue (asynchronously, on NCCL's stream, ordered after the work already queued on the current stream) the
        all-reduce of the given buckets -- used when the caller knows from the structure of its step that they are
        complete (CUDA-graph replay, where the readiness hooks do not run).  ``allreduce()`` later waits for them."""
        if self.world <= 1:
            return
        for bi in bucket_ids:
            b = self.buckets[bi]
            if b["handle"] is None:
                if self.wire_dtype is not None and pre_packed:
                    b["packed"] = True          # the caller's captured graph ran pack() for this bucket
                b["handle"] = self._reduce(b)

    def wait_bucket(self, bi):
        """Make the current stream wait for bucket ``bi``'s all-reduce (issued by ``allreduce_subset``)."""
        b = self.buckets[bi]
        if b["handle"] is not None:
            b["handle"].wait()
            b["handle"] = None
        b["ready"] = 0

    def buckets_within(self, params):
        """Indices of the buckets all of whose parameters are in ``params``."""
        ids = {id(p) for p in params}
        return [i for i, b in enumerate(self.buckets) if all(id(p) in ids for p in b["params"])]

    def total_bytes(self):
        return sum(b["buf"].numel() * 4 for b in self.buckets)

    def remove(self):
        for h in self._hooks:
            h.remove()
        try:
            from . import functional as VF
            VF.set_grad_sinks({})
        except Exception:  # pragma: no cover
            pass
