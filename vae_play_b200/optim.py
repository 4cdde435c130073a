"""Fused optimisers on the fp32 master parameters.

``FusedRMSprop`` has the semantics and hyper-parameters of ``torch.optim.RMSprop`` as the reference uses it
(train.py:136-140: lr 1e-4, alpha 0.99, eps 1e-8, no momentum, not centered) but updates every parameter of a
group with ONE multi-tensor kernel launch instead of five foreach launches.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


class FusedRMSprop(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-2, alpha=0.99, eps=1e-8, weight_decay=0.0, zero_grads=False, wire=None):
        """zero_grads=True: ``step()`` also clears every gradient it has consumed (the tensors stay attached to the
        parameters).  Use with ``functional.persistent_grads`` / ``parallel.GradBuckets`` and do NOT call ``zero_grad()``:
        the next backward then accumulates into known-zero memory without any memset.
        wire: {id(param): bf16 tensor} -- the gradient VALUES are read from these buffers (the bf16 data-parallel wire format,
        ``GradBuckets.wire_views``) instead of ``param.grad``; ``param.grad`` must still exist (it names the slot)."""
        if lr < 0 or eps < 0 or alpha < 0 or weight_decay < 0:
            raise ValueError("invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, alpha=alpha, eps=eps, weight_decay=weight_decay))
        self._tables = {}
        self.zero_grads = bool(zero_grads)
        self.wire = wire

    def _table(self, gi, group):
        from . import functional as VF
        plist = [p for p in group["params"] if p.grad is not None]
        shadows = [VF._SHADOWS.get(p.data_ptr()) for p in plist]
        shadows = [s if s is not None and s[1].stride() == p.stride() else None for s, p in zip(shadows, plist)]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), s[1].data_ptr() if s else 0) for p, s in zip(plist, shadows))
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1]
        for p in plist:
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            if not p.is_cuda or p.dtype != torch.float32 or not dense or p.grad.stride() != p.stride():
                raise _lib.VaePlayError("FusedRMSprop needs dense fp32 CUDA parameters whose gradients share their strides")
            st = self.state[p]
            if "square_avg" not in st:
                st["step"] = 0
                st["square_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        n = len(plist)
        arr = lambda vals: (C.c_void_p * n)(*vals)
        wire = None
        if self.wire is not None:
            for p in plist:
                w = self.wire.get(id(p))
                if w is None or w.dtype != torch.bfloat16 or w.stride() != p.stride():
                    raise _lib.VaePlayError("FusedRMSprop(wire=...): every parameter needs a bf16 wire view with its own strides")
            wire = arr([self.wire[id(p)].data_ptr() for p in plist])
        tab = (arr([p.data_ptr() for p in plist]), arr([p.grad.data_ptr() for p in plist]),
               arr([self.state[p]["square_avg"].data_ptr() for p in plist]), (C.c_int64 * n)(*[p.numel() for p in plist]), n, plist,
               arr([s[1].data_ptr() if s else None for s in shadows]), shadows, wire)
        self._tables[gi] = (key, tab)
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        from . import functional as VF
        VF.join_async()          # weight gradients still running on the side stream (functional.set_async_wgrad)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for gi, group in enumerate(self.param_groups):
            pa, ga, sa, na, n, plist, sha, shadows, wire = self._table(gi, group)
            if n == 0:
                continue
            # the bf16 operand copies of the weights (functional.TapLayer._shadow) are refreshed by the same kernel
            _lib.call("vp_rmsprop_step_wire", pa, ga, sa, sha, wire, na, n, float(group["lr"]), float(group["alpha"]), float(group["eps"]),
                      float(group["weight_decay"]), int(self.zero_grads), stream)
            # the parameters were modified by a kernel torch does not know about: bump their version counters so that
            # everything keyed on tensor._version (the packed-weight caches, autograd's saved-tensor checks) sees it
            torch._C._autograd._unsafe_set_version_counter(plist, [p._version + 1 for p in plist])
            for p, sh in zip(plist, shadows):
                if sh is not None:
                    sh[0].shadow_refreshed(p, sh[1])
            if self.zero_grads:
                from . import functional as VF
                VF.sinks_zeroed(plist)
        return loss


class FusedAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` (no amsgrad) as the reference uses it (train_BE.py:131, train_Style_GAN.py:318-321) with one
    multi-tensor kernel launch per parameter group; refreshes the bf16 operand copies and (``zero_grads=True``) clears the
    consumed gradients like ``FusedRMSprop``.  ``capturable=True`` keeps the step count on the device (CUDA-graph replay)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, zero_grads=False, capturable=False):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.zero_grads, self.capturable = bool(zero_grads), bool(capturable)
        self._tables, self._steps, self._step_dev = {}, {}, {}

    def _table(self, gi, group):
        from . import functional as VF
        plist = [p for p in group["params"] if p.grad is not None]
        shadows = [VF._SHADOWS.get(p.data_ptr()) for p in plist]
        shadows = [s if s is not None and s[1].stride() == p.stride() else None for s, p in zip(shadows, plist)]
        key = tuple((p.data_ptr(), p.grad.data_ptr(), s[1].data_ptr() if s else 0) for p, s in zip(plist, shadows))
        hit = self._tables.get(gi)
        if hit is not None and hit[0] == key:
            return hit[1]
        for p in plist:
            dense = p.is_contiguous() or (p.dim() == 4 and p.is_contiguous(memory_format=torch.channels_last))
            if not p.is_cuda or p.dtype != torch.float32 or not dense or p.grad.stride() != p.stride():
                raise _lib.VaePlayError("FusedAdam needs dense fp32 CUDA parameters whose gradients share their strides")
            st = self.state[p]
            if "exp_avg" not in st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        n = len(plist)
        arr = lambda vals: (C.c_void_p * n)(*vals)
        tab = (arr([p.data_ptr() for p in plist]), arr([p.grad.data_ptr() for p in plist]), arr([self.state[p]["exp_avg"].data_ptr() for p in plist]),
               arr([self.state[p]["exp_avg_sq"].data_ptr() for p in plist]), (C.c_int64 * n)(*[p.numel() for p in plist]), n, plist,
               arr([s[1].data_ptr() if s else None for s in shadows]), shadows)
        self._tables[gi] = (key, tab)
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        from . import functional as VF
        VF.join_async()          # weight gradients still running on the side stream (functional.set_async_wgrad)
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        for gi, group in enumerate(self.param_groups):
            pa, ga, ma, va, na, n, plist, sha, shadows = self._table(gi, group)
            if n == 0:
                continue
            self._steps[gi] = self._steps.get(gi, 0) + 1
            step_dev = None
            if self.capturable:
                if gi not in self._step_dev:
                    self._step_dev[gi] = torch.zeros(1, dtype=torch.int64, device=plist[0].device)
                _lib.call("vp_philox_advance", C.c_void_p(self._step_dev[gi].data_ptr()), 1, stream)
                step_dev = C.c_void_p(self._step_dev[gi].data_ptr())
            b1, b2 = group["betas"]
            _lib.call("vp_adam_step", pa, ga, ma, va, sha, na, n, float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                      float(group["weight_decay"]), self._steps[gi], step_dev, int(self.zero_grads), stream)
            torch._C._autograd._unsafe_set_version_counter(plist, [p._version + 1 for p in plist])
            for p, sh in zip(plist, shadows):
                if sh is not None:
                    sh[0].shadow_refreshed(p, sh[1])
            if self.zero_grads:
                from . import functional as VF
                VF.sinks_zeroed(plist)
        return loss
