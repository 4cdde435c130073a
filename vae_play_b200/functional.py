"""Autograd functions of the VAE hot path, each a thin host wrapper over libvaeplay_b200 (C ABI).

Activations flow between layers as channels-last tensors ``[N,H,W,C]`` in the activation dtype of the
current precision mode (bf16: tensor-core mode, fp32: check mode); NCHW fp32 exists only at the graph
edges (``to_channels_last`` / ``from_channels_last``).  torch is used for memory, streams and autograd
bookkeeping only -- every arithmetic step is one of our CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import ACT, BF16, F32, VpConvGeom

_STATE = {"precision": "bf16", "engine": _lib.ENGINE_AUTO}
_EPOCH = [0]       # bumped by invalidate_caches(): forces every packed-weight cache to miss once
_TRACE = None      # when a list: every fused layer appends its activated output (tests / diagnostics only)
_GRAD_SINKS = {}   # param.data_ptr() -> (bucket view, param): where wgrad writes directly (vae_play_b200.parallel)
_SHADOWS = {}      # weight.data_ptr() -> (TapLayer, bf16 copy): lets a fused optimiser refresh the copy while it updates the master


_SINK_HOOKS = []   # post-accumulate hooks that re-arm a slot once autograd has consumed the tensor handed out for it


def set_grad_sinks(sinks_by_param_id, params=None):
    """Register persistent slots for parameter gradients: param.data_ptr() -> (flat fp32 buffer, element offset, param).
    The weight-gradient kernels write straight into them (flat data-parallel buckets, ``persistent_grads``)."""
    for h in _SINK_HOOKS:
        h.remove()
    _SINK_HOOKS.clear()
    _GRAD_SINKS.clear()
    for k, v in sinks_by_param_id.items():
        # [flat, offset, param, slot is known to hold zeros, slot handed out and not yet accumulated by autograd]
        entry = [v[0], int(v[1]), v[2], False, False]
        _GRAD_SINKS[k] = entry
        if hasattr(v[2], "register_post_accumulate_grad_hook"):
            _SINK_HOOKS.append(v[2].register_post_accumulate_grad_hook(lambda p, e=entry: e.__setitem__(4, False)))


def persistent_grads(params):
    """Give every parameter a slot in ONE flat fp32 buffer that serves as its ``.grad`` step after step.  Together with an
    optimiser that clears each gradient once it has consumed it (``FusedRMSprop(zero_grads=True)``) the weight-gradient
    kernels accumulate into known-zero memory: no per-tensor memset, no allocation.  Returns the flat buffer."""
    params = [p for p in params if p.requires_grad]
    pad = lambda p: (p.numel() + 3) & ~3          # every slot starts 16-byte aligned
    flat = torch.zeros(sum(pad(p) for p in params), dtype=torch.float32, device=params[0].device)
    sinks, off = {}, 0
    for p in params:
        sinks[p.data_ptr()] = (flat, off, p)
        off += pad(p)
    set_grad_sinks(sinks)
    for e in _GRAD_SINKS.values():
        e[3] = True
    return flat


def sinks_zeroed(params):
    """An optimiser reports that it has cleared the gradients of ``params`` (those living in registered slots)."""
    for p in params:
        e = _GRAD_SINKS.get(p.data_ptr())
        if e is not None and p.grad is not None and p.grad.data_ptr() == e[0].data_ptr() + 4 * e[1]:
            e[3] = True


def _grad_target(weight):
    """Where a weight gradient is written: (tensor, holds_zeros).  A registered slot is used when it can become the
    parameter's .grad as is: no gradient accumulated yet this step AND the slot not already handed to an earlier use of
    the same weight in this backward (a weight used twice -- the decoder on z and z_p, the discriminator in REC and GAN
    mode, reference train.py:43-73 -- gets a fresh buffer for the later uses; autograd sums them into the slot).  The
    view is created afresh and referenced by nobody else, so that autograd's AccumulateGrad adopts it instead of cloning."""
    hit = _GRAD_SINKS.get(weight.data_ptr())
    if hit is not None:
        flat, off, param, zeroed, handed = hit
        if param.grad is None and not handed and param.shape == weight.shape and param.stride() == weight.stride():
            hit[3] = False
            hit[4] = True
            return torch.as_strided(flat, weight.shape, weight.stride(), storage_offset=off), zeroed
    return torch.empty_like(weight, dtype=torch.float32), False


def invalidate_caches():
    """Force re-packing of all weight panels on next use (call before CUDA-graph capture so that the
    pack kernels are part of the captured step)."""
    _EPOCH[0] += 1


def set_precision(mode: str):
    """'bf16' (tcgen05 tensor cores, bf16 activations) or 'fp32' (CUDA-core check mode)."""
    if mode not in ("bf16", "fp32"):
        raise ValueError(mode)
    _STATE["precision"] = mode


def get_precision() -> str:
    return _STATE["precision"]


def set_epilogue_stats(on: bool):
    """BatchNorm statistics from the GEMM epilogue (default) or from a separate pass over the stored output (tests)."""
    _STATE["epilogue_stats"] = bool(on)


def set_fuse_bn_backward(on: bool, thin: bool = False):
    """BatchNorm backward: first pass in the epilogue of the consumer's data-gradient kernel (default) or as its own pass (tests).
    ``thin``: also in the output layer's thin data-gradient kernel (vp_thin_conv_dgrad_bnred).  Off by default: that kernel runs at
    ~50 % of HBM bandwidth and the extra 134 MB read of y costs it more (+55 us at batch 256, 64x64) than the separate reduce pass
    over the last decoder block (-15 us net loss measured: 1.913 vs 1.874 ms per step)."""
    _STATE["fuse_bn_bwd"] = bool(on)
    _STATE["fuse_bn_bwd_thin"] = bool(on) and bool(thin)


_ASYNC = {"on": False, "stream": None, "pending": [], "forked": False, "home": None, "slots": set()}


def set_async_wgrad(on: bool):
    """Weight gradients of the fused layers on a side stream, concurrent with the rest of the backward pass (the tensor-bound
    weight-gradient kernel of block L overlaps the bandwidth-bound BatchNorm-backward passes of block L-1).  Only gradients that
    go straight into a persistent slot (``persistent_grads`` / data-parallel buckets) take the side stream: nothing on the main
    stream touches them until the optimiser.  A weight used several times per backward (the VAE-GAN step) gets its later
    contributions added into the slot on the side stream as well (``_async_fork``).  The caller MUST call ``join_async()`` after ``backward()`` -- inside the same
    CUDA-graph capture when capturing; the fused optimisers and ``GradBuckets`` also join before they read gradients."""
    _ASYNC["on"] = bool(on)
    if not on:
        join_async()


def _async_fork(weight):
    """(event on the current stream, again) if this weight's gradient may be computed on the side stream, else None.
    ``again``: the weight's slot was already handed to an earlier use in this backward whose kernel runs on the side stream
    (the decoder on z and z_p, the discriminator in REC and GAN mode): this use is computed into a scratch buffer on the side
    stream and ADDED into the slot there -- in stream order behind the first use -- and autograd is told nothing (None)."""
    if not _ASYNC["on"]:
        return None
    ptr = weight.data_ptr()
    hit = _GRAD_SINKS.get(ptr)
    if hit is None or hit[2].shape != weight.shape or hit[2].stride() != weight.stride():
        return None
    again = ptr in _ASYNC["slots"]
    if not again and (hit[2].grad is not None or hit[4]):
        return None            # the slot is taken by a gradient that lives on the main stream: autograd will add there
    if _ASYNC["stream"] is None or _ASYNC["stream"].device != weight.device:
        _ASYNC["stream"] = torch.cuda.Stream(weight.device)
    ev = torch.cuda.Event()
    _ASYNC["home"] = torch.cuda.current_stream(weight.device)
    ev.record(_ASYNC["home"])
    _ASYNC["forked"] = True
    _ASYNC["slots"].add(ptr)
    return ev, again


def _wgrad_maybe_async(layer, x, dy, weight, fork):
    if fork is None:
        return layer.wgrad(x, dy, weight)
    ev, again = fork
    side = _ASYNC["stream"]
    with torch.cuda.stream(side):
        side.wait_event(ev)
        dw = layer.wgrad(x, dy, weight)          # first use: written into the slot; later uses: a scratch buffer (the slot is taken)
        if again:
            flat, off = _GRAD_SINKS[weight.data_ptr()][0], _GRAD_SINKS[weight.data_ptr()][1]
            if dw.data_ptr() == flat.data_ptr() + 4 * off:
                raise _lib.VaePlayError("async weight gradient: a re-used weight was handed its slot twice")
            _lib.call("vp_axpy", 1.0, _ptr(dw), C.c_void_p(flat.data_ptr() + 4 * off), dw.numel(), _stream())
            _ASYNC["pending"].append((x, dy, dw))
            return None
    _ASYNC["pending"].append((x, dy))            # keep the operands allocated until join_async()
    return dw


def join_async():
    """The current stream waits for every weight gradient launched on the side stream.  When the current stream is the one the
    backward ran on, their operands may be freed too (a third stream -- an optimiser overlapping the rest of the backward --
    only waits: the allocator hands freed blocks straight back to the stream that allocated them)."""
    cur = None
    if _ASYNC["forked"]:
        ev = torch.cuda.Event()
        ev.record(_ASYNC["stream"])
        cur = torch.cuda.current_stream(_ASYNC["stream"].device)
        cur.wait_event(ev)
    if cur is None or _ASYNC.get("home") is None or cur == _ASYNC["home"]:
        _ASYNC["forked"] = False
        _ASYNC["pending"].clear()
        _ASYNC["slots"].clear()


def set_pad_route(on: bool):
    """bf16 layers whose channel counts are not multiples of 64: zero-pad onto the tcgen05 kernels (default) or fall back to the
    packed route (CUDA-core kernel unless the shape happens to be tensor-core eligible) -- tests / A-B measurements."""
    _STATE["pad_route"] = bool(on)


def set_engine(name: str):
    """'auto' | 'simt' | 'tc' -- engine used for the contractions (tests force one or the other)."""
    _STATE["engine"] = {"auto": _lib.ENGINE_AUTO, "simt": _lib.ENGINE_SIMT, "tc": _lib.ENGINE_TC}[name]


def act_dtype() -> torch.dtype:
    return torch.bfloat16 if _STATE["precision"] == "bf16" else torch.float32


def _code(dt: torch.dtype) -> int:
    if dt == torch.float32:
        return F32
    if dt == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {dt}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    _lib.ensure_workspace(torch.device("cuda", torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise _lib.VaePlayError(f"{what}: tensor is on {t.device}; vae_play_b200 has no CPU path")
    if not t.is_contiguous():
        raise _lib.VaePlayError(f"{what}: tensor must be contiguous")


# ------------------------------------------------------------------------------------------------
# layer geometry + weight packing recipes
# ------------------------------------------------------------------------------------------------
@dataclass
class Pack:
    taps: int
    n: int
    k: int
    sn: int
    sk: int
    st: int


def weights_channels_last(module):
    """Keep the conv weights of ``module`` in torch.channels_last memory format (same shapes, same ``state_dict``):
    physically [co][kh][kw][ci] (nn.Conv2d) / [ci][kh][kw][co] (nn.ConvTranspose2d), which the TMA reads in place as a
    tcgen05 operand for forward AND data gradient, and which the weight-gradient kernel writes in place -- no packed
    copies, no unpack pass.  Only layers the in-place kernels take (both channel counts multiples of 64) are converted."""
    for m in module.modules():
        if isinstance(m, (torch.nn.Conv2d, torch.nn.ConvTranspose2d)) and m.in_channels % 64 == 0 and m.out_channels % 64 == 0:
            m.weight.data = m.weight.data.contiguous(memory_format=torch.channels_last)
    return module


class TapLayer:
    """One contraction layer: how its forward / dgrad / wgrad map onto the kernel library.

    kind: 'conv' (nn.Conv2d), 'convT' (nn.ConvTranspose2d), 'linear' (nn.Linear on [B,in]).
    Three routes, chosen per call:
      * in place  (vp_conv_*_cl):  bf16, both channel counts multiples of 64, weight dense in channels-last order
      * thin      (vp_thin_conv_*): bf16, a single channel on one side
      * padded    (vp_pad_* + vp_conv_*_cl): bf16, any other channel counts -- activations and weight are zero-padded to the next
                  multiples of 64 and the in-place tcgen05 kernels run on the padded geometry (exact: zero channels add zeros)
      * packed    (vp_conv_*):      the fp32 check mode and a forced CUDA-core engine (tap-major panels built by vp_pack_weight)
    """

    def __init__(self, kind, cin, cout, k=1, stride=1, pad=0, out_pad=0):
        if kind not in ("conv", "convT", "linear"):
            raise ValueError(kind)
        self.kind, self.cin, self.cout = kind, cin, cout
        self.k, self.stride, self.pad, self.out_pad = k, stride, pad, out_pad
        self._cache = {}

    # ---- geometry -------------------------------------------------------------------------------
    def out_shape(self, n, h, w):
        if self.kind == "conv":
            return (n, (h + 2 * self.pad - self.k) // self.stride + 1, (w + 2 * self.pad - self.k) // self.stride + 1, self.cout)
        if self.kind == "convT":
            return (n, (h - 1) * self.stride - 2 * self.pad + self.k + self.out_pad,
                    (w - 1) * self.stride - 2 * self.pad + self.k + self.out_pad, self.cout)
        return (n, 1, 1, self.cout)

    def _geom(self, n, hi, wi, ci, ho, wo, co, k, stride, pad, transposed):
        return VpConvGeom(n, hi, wi, ci, ho, wo, co, k, k, stride, pad, transposed)

    def _layer_geom(self, n, h, w, ho, wo):
        return self._geom(n, h, w, self.cin, ho, wo, self.cout, self.k, self.stride, self.pad, int(self.kind == "convT"))

    # ---- packed route ---------------------------------------------------------------------------------
    def _recipe(self, which, weight) -> Pack:
        """Pack recipe from the weight's actual strides (torch-contiguous or channels-last)."""
        T = self.k * self.k
        if weight.dim() == 2:
            s_out, s_in, st = weight.stride(0), weight.stride(1), 1
        else:
            if weight.stride(2) != weight.shape[3] * weight.stride(3) and weight.shape[2] > 1:
                raise _lib.VaePlayError("conv weight with non-uniform tap strides")
            s_out, s_in, st = weight.stride(0), weight.stride(1), weight.stride(3)
        if self.kind == "convT":
            s_out, s_in = s_in, s_out       # weight is [cin][cout][kh][kw]
        if (which == "dgrad") != (self.kind == "convT" and which == "wgrad"):
            # [T][cin][cout]: dgrad panel; also the wgrad layout of a transposed conv
            return Pack(T, self.cin, self.cout, s_in, s_out, st)
        return Pack(T, self.cout, self.cin, s_out, s_in, st)

    def _packed(self, weight, which, dtype):
        key = (which, dtype, weight.data_ptr(), weight._version, _EPOCH[0], weight.stride())
        hit = self._cache.get(which)
        if hit is not None and hit[0] == key:
            return hit[1]
        p = self._recipe(which, weight)
        wp = torch.empty(p.taps * p.n * p.k, dtype=dtype, device=weight.device)
        _lib.call("vp_pack_weight", _ptr(weight), _ptr(wp), _code(dtype), p.taps, p.n, p.k, p.sn, p.sk, p.st, _stream())
        self._cache[which] = (key, wp)
        return wp

    # ---- in-place route ---------------------------------------------------------------------------------
    def _cl(self, dt, weight):
        if dt != torch.bfloat16 or _STATE["engine"] == _lib.ENGINE_SIMT or self.cin % 64 or self.cout % 64:
            return False
        if weight.dim() == 4:
            return weight.is_contiguous(memory_format=torch.channels_last)
        return weight.is_contiguous()

    def _shadow(self, weight):
        """bf16 copy of the master weight, element for element (same strides)."""
        hit = self._cache.get("shadow")
        key = (weight.data_ptr(), weight._version, _EPOCH[0])
        if hit is not None and (hit[0] == key or (len(hit) > 2 and hit[0][:2] == key[:2])):
            return hit[1]      # up to date (a copy refreshed by the optimiser stays valid across invalidate_caches())
        sh = hit[1] if hit is not None and hit[1].shape == weight.shape and hit[1].stride() == weight.stride() else \
            torch.empty_like(weight, dtype=torch.bfloat16)
        _lib.call("vp_cast", _ptr(weight), F32, _ptr(sh), BF16, weight.numel(), _stream())
        self._cache["shadow"] = (key, sh)
        _SHADOWS[weight.data_ptr()] = (self, sh)
        return sh

    def shadow_refreshed(self, weight, shadow):
        """Called by an optimiser that has written the bf16 copy itself."""
        self._cache["shadow"] = ((weight.data_ptr(), weight._version, _EPOCH[0]), shadow, "optimiser")

    # ---- padded route: channel counts that are not multiples of 64 on the 64-multiple tcgen05 kernels ------------------------
    def _padded(self, dt, weight):
        if dt != torch.bfloat16 or _STATE["engine"] == _lib.ENGINE_SIMT or _STATE.get("pad_route", True) is False:
            return False
        if weight.dim() == 4 and weight.shape[2] > 1 and weight.stride(2) != weight.shape[3] * weight.stride(3):
            return False
        return True

    @staticmethod
    def _up64(c):
        return (c + 63) & ~63

    def _pad_dims(self, weight):
        """(d0, d1, taps, s0, s1, st, d0p, d1p) of the weight: axes 0 / 1 are (co, ci) for conv / linear, (ci, co) for convT."""
        d0, d1 = weight.shape[0], weight.shape[1]
        taps = self.k * self.k if weight.dim() == 4 else 1
        st = weight.stride(3) if weight.dim() == 4 else 1
        return d0, d1, taps, weight.stride(0), weight.stride(1), st, self._up64(d0), self._up64(d1)

    def _padded_weight(self, weight):
        key = (weight.data_ptr(), weight._version, _EPOCH[0], weight.stride())
        hit = self._cache.get("padw")
        if hit is not None and hit[0] == key:
            return hit[1]
        d0, d1, taps, s0, s1, st, d0p, d1p = self._pad_dims(weight)
        wp = hit[1] if hit is not None and hit[1].numel() == d0p * taps * d1p else torch.empty(d0p * taps * d1p, dtype=torch.bfloat16, device=weight.device)
        _lib.call("vp_pad_weight_cl", _ptr(weight), _ptr(wp), d0, d1, taps, s0, s1, st, d0p, d1p, _stream())
        self._cache["padw"] = (key, wp)
        return wp

    # padded copies of layer inputs made by a training-mode forward, for the weight gradient of the same call.  The ORIGINAL
    # tensor is held too (autograd saves it anyway): while it lives its address cannot be handed to another tensor, so
    # (address, version, shape) identifies it.  At most 4 pending entries per layer (REC + GAN passes, z + z_p).
    def _remember_padded(self, x, xp):
        pend = self._cache.setdefault("padded_inputs", [])
        pend.append((x.data_ptr(), x._version, tuple(x.shape), x, xp))
        if len(pend) > 4:
            pend.pop(0)

    def _recall_padded(self, x, cp):
        pend = self._cache.get("padded_inputs")
        if pend:
            for i in range(len(pend) - 1, -1, -1):
                e = pend[i]
                if e[0] == x.data_ptr() and e[1] == x._version and e[2] == tuple(x.shape) and e[4].shape[-1] == cp:
                    pend.pop(i)
                    return e[4]
        return None

    def _padded_rows(self, weight, cop):
        """fp32 weight [co][...] with zero rows appended up to cop (cached per weight version)."""
        key = ("rows", cop)
        hit = self._cache.get(key)
        if hit is not None and hit[0] == weight.data_ptr() and hit[1] == weight._version and hit[3] == _EPOCH[0]:
            return hit[2]
        per = weight.numel() // weight.shape[0]
        wq = hit[2] if hit is not None and hit[2].numel() == cop * per else torch.empty(cop * per, dtype=torch.float32, device=weight.device)
        _lib.call("vp_pad_channels", _ptr(weight), weight.numel(), _ptr(wq), cop * per, 1, F32, _stream())     # one "row" of co*per values, zero tail
        self._cache[key] = (weight.data_ptr(), weight._version, wq, _EPOCH[0])
        return wq

    @staticmethod
    def _pad_act(x, cp):
        c = x.shape[-1]
        if c == cp:
            return x
        out = torch.empty(x.shape[:-1] + (cp,), dtype=x.dtype, device=x.device)
        _lib.call("vp_pad_channels", _ptr(x), c, _ptr(out), cp, x.numel() // c, _code(x.dtype), _stream())
        return out

    @staticmethod
    def _slice_act(xp, c):
        cp = xp.shape[-1]
        if c == cp:
            return xp
        out = torch.empty(xp.shape[:-1] + (c,), dtype=xp.dtype, device=xp.device)
        _lib.call("vp_copy_channels", _ptr(xp), cp, 0, _ptr(out), c, 0, c, xp.numel() // cp, _code(xp.dtype), 0, _stream())
        return out

    def _padded_geom(self, n, h, w, ho, wo):
        return self._geom(n, h, w, self._up64(self.cin), ho, wo, self._up64(self.cout), self.k, self.stride, self.pad, int(self.kind == "convT"))

    # ---- thin layers (a single channel on one side): tcgen05 kernels that read the fp32 master weight directly ----
    def _thin(self, which, dt, weight):
        if self.kind != "conv" or dt != torch.bfloat16 or _STATE["engine"] == _lib.ENGINE_SIMT or not weight.is_contiguous():
            return False
        T, ci, co, s = self.k * self.k, self.cin, self.cout, self.stride
        thin_in = ci == 1 and self.k <= 8 and s <= 3
        if which == "fwd":
            return (thin_in and co % 32 == 0 and co <= 128) or (s == 1 and self.k <= 5 and co * T <= 32 and ci % 64 == 0 and ci <= 256)
        thin_out = co == 1 and s == 1 and self.k <= 8
        if which == "dgrad":
            return thin_out and ci % 32 == 0 and ci <= 128
        return (thin_in and co % 64 == 0) or (thin_out and ci % 64 == 0)

    # ---- the three contractions --------------------------------------------------------------------
    def fwd(self, x, weight, bias, act="none", slope=0.0, out_dtype=None):
        n, h, w, _ = x.shape
        dt = x.dtype
        out_dtype = out_dtype or dt
        weight = weight.detach()
        shp = self.out_shape(n, h, w)
        y = torch.empty(shp, dtype=out_dtype, device=x.device)
        g = self._layer_geom(n, h, w, shp[1], shp[2])
        if self._thin("fwd", dt, weight):
            _lib.call("vp_thin_conv_fwd", C.byref(g), _ptr(x), _ptr(weight), _ptr(bias), _ptr(y), _code(out_dtype),
                      ACT[act], float(slope), _stream())
        elif self._cl(dt, weight):
            _lib.call("vp_conv_fwd_cl", C.byref(g), _ptr(x), _ptr(self._shadow(weight)), _ptr(bias), _ptr(y), _code(out_dtype),
                      ACT[act], float(slope), _stream())
        elif self._padded(dt, weight):
            cip, cop = self._up64(self.cin), self._up64(self.cout)
            gp = self._padded_geom(n, h, w, shp[1], shp[2])
            bp = None
            if bias is not None:
                bp = torch.empty(cop, dtype=torch.float32, device=x.device)
                _lib.call("vp_pad_channels", _ptr(bias.detach()), self.cout, _ptr(bp), cop, 1, F32, _stream())     # zero-padded bias, one kernel
            yp = y if cop == self.cout else torch.empty(shp[:3] + (cop,), dtype=out_dtype, device=x.device)
            # the padded temporaries stay referenced until the call has been issued: a tensor dropped right after its pointer was
            # taken would hand its block to the NEXT allocation (the padded weight on a cache miss) before the kernel is queued
            xp, wp = self._pad_act(x, cip), self._padded_weight(weight)
            if xp is not x and torch.is_grad_enabled() and weight.requires_grad:
                self._remember_padded(x, xp)          # the weight gradient needs the same padded copy: keep it instead of padding again
            _lib.call("vp_conv_fwd_cl", C.byref(gp), _ptr(xp), _ptr(wp), _ptr(bp), _ptr(yp), _code(out_dtype), ACT[act], float(slope), _stream())
            if yp is not y:
                _lib.call("vp_copy_channels", _ptr(yp), cop, 0, _ptr(y), self.cout, 0, self.cout, y.numel() // self.cout, _code(out_dtype), 0, _stream())
        else:
            wp = self._packed(weight, "fwd", dt)
            _lib.call("vp_conv_fwd", C.byref(g), _ptr(x), _ptr(wp), _ptr(bias), _ptr(y), _code(dt), _code(out_dtype),
                      ACT[act], float(slope), _STATE["engine"], _stream())
        return y

    def stats_in_epilogue(self, dt, weight):
        """True when the forward GEMM can also deliver the BatchNorm statistics of its output (vp_*_fwd_*_stats)."""
        weight = weight.detach()
        if self.cout % 64 or self.cout > 512 or _STATE.get("epilogue_stats", True) is False:
            return False
        if self._thin("fwd", dt, weight):
            return self.cin == 1 and self.cout <= 128
        return self._cl(dt, weight)

    def fwd_stats(self, x, weight):
        """y (bf16, no bias / activation) and the per-CTA partial sums of its batch statistics: (y, parts, nparts)."""
        n, h, w, _ = x.shape
        weight = weight.detach()
        shp = self.out_shape(n, h, w)
        y = torch.empty(shp, dtype=x.dtype, device=x.device)
        g = self._layer_geom(n, h, w, shp[1], shp[2])
        cap = 2 * _num_sms(x.device)
        parts = torch.empty(cap * 2 * self.cout, dtype=torch.float32, device=x.device)
        nparts = C.c_int(0)
        if self._thin("fwd", x.dtype, weight):
            _lib.call("vp_thin_conv_fwd_stats", C.byref(g), _ptr(x), _ptr(weight), _ptr(y), _ptr(parts), cap, C.byref(nparts), _stream())
        else:
            _lib.call("vp_conv_fwd_cl_stats", C.byref(g), _ptr(x), _ptr(self._shadow(weight)), _ptr(y), _ptr(parts), cap,
                      C.byref(nparts), _stream())
        return y, parts, nparts.value

    def dgrad(self, dy, weight, x_shape, out_dtype=None):
        n, h, w, _ = x_shape
        dt = dy.dtype
        out_dtype = out_dtype or dt
        weight = weight.detach()
        dx = torch.empty(x_shape, dtype=out_dtype, device=dy.device)
        g = self._layer_geom(n, h, w, dy.shape[1], dy.shape[2])
        if self._thin("dgrad", dt, weight):
            _lib.call("vp_thin_conv_dgrad", C.byref(g), _ptr(dy), _ptr(weight), _ptr(dx), _code(out_dtype), _stream())
        elif self._cl(dt, weight):
            _lib.call("vp_conv_dgrad_cl", C.byref(g), _ptr(dy), _ptr(self._shadow(weight)), _ptr(dx), _code(out_dtype), _stream())
        elif (self.kind == "conv" and self.cin == 1 and self.stride == 1 and self.k <= 5 and dt == torch.bfloat16 and weight.is_contiguous()
              and _STATE["engine"] != _lib.ENGINE_SIMT and self._up64(self.cout) <= 256):
            # gradient w.r.t. a single-channel input (the discriminator's first layer on x_tilde): the thin-output forward kernel on dy
            # with flipped taps; a 32-channel dy is zero-padded to 64 (its weight rows beyond cout are never read: K blocks of 64
            # channels, the padding channels of dy are zero and meet whatever the weight tile holds -- so pad the weight view too)
            cop = self._up64(self.cout)
            gp = self._geom(n, h, w, 1, dy.shape[1], dy.shape[2], cop, self.k, self.stride, self.pad, 0)
            dyp = self._pad_act(dy, cop)
            wq = weight if cop == self.cout else self._padded_rows(weight, cop)
            _lib.call("vp_thin_conv_dgrad_in1", C.byref(gp), _ptr(dyp), _ptr(wq), _ptr(dx), _code(out_dtype), _stream())
        elif self._padded(dt, weight):
            cip, cop = self._up64(self.cin), self._up64(self.cout)
            gp = self._padded_geom(n, h, w, dy.shape[1], dy.shape[2])
            dxp = dx if cip == self.cin else torch.empty(tuple(x_shape[:3]) + (cip,), dtype=out_dtype, device=dy.device)
            dyp, wp = self._pad_act(dy, cop), self._padded_weight(weight)
            _lib.call("vp_conv_dgrad_cl", C.byref(gp), _ptr(dyp), _ptr(wp), _ptr(dxp), _code(out_dtype), _stream())
            if dxp is not dx:
                _lib.call("vp_copy_channels", _ptr(dxp), cip, 0, _ptr(dx), self.cin, 0, self.cin, dx.numel() // self.cin, _code(out_dtype), 0, _stream())
        else:
            wp = self._packed(weight, "dgrad", dt)
            _lib.call("vp_conv_dgrad", C.byref(g), _ptr(dy), _ptr(wp), _ptr(dx), _code(dt), _code(out_dtype),
                      _STATE["engine"], _stream())
        return dx

    def dgrad_bnred(self, dy, weight, x_shape, prev):
        """Data gradient fused with the first pass of the PRODUCER block's BatchNorm backward (``prev``: that block's holder, see
        ``fused_layer``).  Returns (dx, parts, nparts) or None when no kernel with that epilogue serves the shape."""
        weight = weight.detach()
        thin = self._thin("dgrad", dy.dtype, weight) and self.cin == 64
        if thin and not _STATE.get("fuse_bn_bwd_thin", False):
            return None
        if dy.dtype != torch.bfloat16 or not (thin or self._cl(dy.dtype, weight)) or _STATE.get("fuse_bn_bwd", True) is False:
            return None
        y = prev["y"]
        if tuple(y.shape) != tuple(x_shape) or y.dtype != torch.bfloat16 or not y.is_contiguous():
            return None
        n, h, w, _ = x_shape
        dx = torch.empty(x_shape, dtype=torch.bfloat16, device=dy.device)
        g = self._layer_geom(n, h, w, dy.shape[1], dy.shape[2])
        cap = 2 * _num_sms(dy.device)
        parts = torch.empty(cap * 2 * self.cin, dtype=torch.float32, device=dy.device)
        nparts = C.c_int(0)
        name = "vp_thin_conv_dgrad_bnred" if thin else "vp_conv_dgrad_cl_bnred"
        rc = getattr(_lib.load(), name)(C.byref(g), _ptr(dy), _ptr(weight if thin else self._shadow(weight)), _ptr(dx), _ptr(y),
                                        _ptr(prev["scale"]), _ptr(prev["shift"]), _ptr(prev["mean"]), _ptr(parts), cap, C.byref(nparts),
                                        _stream())
        if rc == -3:          # VP_EUNSUPPORTED: nothing was launched
            return None
        _lib.check(rc, name)
        return dx, parts, nparts.value

    def wgrad(self, x, dy, weight):
        n, h, w, _ = x.shape
        dt = x.dtype
        g = self._layer_geom(n, h, w, dy.shape[1], dy.shape[2])
        thin, cl = self._thin("wgrad", dt, weight), self._cl(dt, weight)
        if thin or cl:
            dw, zeroed = _grad_target(weight)       # same strides as the weight: the kernel writes the gradient in place
            _lib.call("vp_thin_conv_wgrad" if thin else "vp_conv_wgrad_cl", C.byref(g), _ptr(x), _ptr(dy), _ptr(dw), int(zeroed), _stream())
            return dw
        if (self._padded(dt, weight) and self.kind == "conv" and self.cin == 1 and self.k <= 8 and self.stride <= 3 and weight.is_contiguous()
                and self.cout % 64 != 0):
            # single-channel input, output channels not a multiple of 64 (the discriminator's 1 -> 32 first layer): pad only the
            # wide side and use the thin weight-gradient kernel; the gradient rows of the padding channels are dropped
            cop = self._up64(self.cout)
            gp = self._geom(n, h, w, 1, dy.shape[1], dy.shape[2], cop, self.k, self.stride, self.pad, 0)
            dwp = torch.empty((cop,) + tuple(weight.shape[1:]), dtype=torch.float32, device=x.device)
            dyp = self._pad_act(dy, cop)
            _lib.call("vp_thin_conv_wgrad", C.byref(gp), _ptr(x), _ptr(dyp), _ptr(dwp), 0, _stream())
            dw, _ = _grad_target(weight)
            _lib.call("vp_cast", _ptr(dwp), F32, _ptr(dw), F32, weight.numel(), _stream())
            return dw
        if self._padded(dt, weight):
            cip, cop = self._up64(self.cin), self._up64(self.cout)
            gp = self._padded_geom(n, h, w, dy.shape[1], dy.shape[2])
            d0, d1, taps, s0, s1, st, d0p, d1p = self._pad_dims(weight)
            dwp = torch.empty(d0p * taps * d1p, dtype=torch.float32, device=x.device)
            xp = self._recall_padded(x, cip)
            xp, dyp = (xp if xp is not None else self._pad_act(x, cip)), self._pad_act(dy, cop)
            _lib.call("vp_conv_wgrad_cl", C.byref(gp), _ptr(xp), _ptr(dyp), _ptr(dwp), 0, _stream())
            dw, _ = _grad_target(weight)
            _lib.call("vp_unpad_wgrad_cl", _ptr(dwp), _ptr(dw), d0, d1, taps, s0, s1, st, d1p, _stream())
            return dw
        p = self._recipe("wgrad", weight)
        dwp = torch.empty(p.taps * p.n * p.k, dtype=torch.float32, device=x.device)
        _lib.call("vp_conv_wgrad", C.byref(g), _ptr(x), _ptr(dy), _ptr(dwp), _code(dt), _STATE["engine"], _stream())
        dw, _ = _grad_target(weight)
        _lib.call("vp_unpack_wgrad", _ptr(dwp), _ptr(dw), p.taps, p.n, p.k, p.sn, p.sk, p.st, _stream())
        return dw


@dataclass
class NormCfg:
    kind: Optional[str]  # 'batch' | 'instance' | None
    eps: float = 1e-5
    momentum: float = 0.1


class _FusedLayerFn(torch.autograd.Function):
    """contraction (+bias) -> [BatchNorm | InstanceNorm] -> activation, with the matching backward.

    Mirrors models/blocks.py:31-34 (Conv2d block), models/networks.py:27-30 (EncoderBlock),
    :42-46 (DecoderBlock) and the Linear->BatchNorm1d->ReLU stacks at :65-67,88-90.
    Returns (activated output, pre-norm contraction output).
    """

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, layer: TapLayer, norm: NormCfg, act, slope, training, bn_module,
                out_dtype, holder=None, prev_bn=None):
        _require_cuda(x, "fused layer input")
        ctx.set_materialize_grads(False)
        ctx.holder, ctx.prev_bn = holder, prev_bn
        dt = x.dtype
        out_dtype = out_dtype or dt
        ctx.layer, ctx.norm, ctx.act, ctx.slope = layer, norm, act, slope
        ctx.x_shape = tuple(x.shape)
        ctx.has_bias = bias is not None
        ctx.few_rows = False
        if norm.kind is None:
            a = layer.fwd(x, weight, bias, act, slope, out_dtype)
            ctx.save_for_backward(x, weight, a)
            ctx.out_dtype = out_dtype
            return a, None
        fuse_stats = norm.kind == "batch" and training and bias is None and layer.stats_in_epilogue(dt, weight)
        if fuse_stats:
            y, parts, nparts = layer.fwd_stats(x, weight)
        else:
            y = layer.fwd(x, weight, bias, "none", 0.0, dt)
        n, h, w, c = y.shape
        if norm.kind == "batch":
            groups, rpg, cc = 1, n * h * w, c
        else:
            groups, rpg, cc = n, h * w, c
        dev = x.device
        stats = torch.empty(4, groups * cc, dtype=torch.float32, device=dev)  # mean, invstd, scale, shift
        mean, invstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
        g_, b_ = gamma, beta
        # few rows, many channels (the BatchNorm1d behind the fc layers): statistics + finalize + apply in ONE launch
        few_rows = norm.kind == "batch" and training and not fuse_stats and rpg <= 8192 and cc % 4 == 0 and out_dtype == dt
        ctx.few_rows = few_rows
        if few_rows:
            rm = rv = nbt = None
            if bn_module is not None and bn_module.track_running_stats:
                rm, rv, nbt = bn_module.running_mean, bn_module.running_var, bn_module.num_batches_tracked
            a = torch.empty_like(y)
            _lib.call("vp_bn_rows_fwd", _ptr(y), _ptr(g_), _ptr(b_), _ptr(rm), _ptr(rv), float(norm.momentum), float(norm.eps), _ptr(a),
                      _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift), _code(dt), rpg, cc, ACT[act], float(slope), _ptr(nbt), _stream())
            ctx.save_for_backward(x, weight, y, stats)
            ctx.dims = (groups, rpg, cc, c)
            ctx.train_stats = True
            ctx.has_affine = gamma is not None
            return a, y
        if training or norm.kind == "instance":
            rm = rv = nbt = None
            if norm.kind == "batch" and bn_module is not None and bn_module.track_running_stats:
                # running statistics AND the num_batches_tracked counter are updated by the finalize kernel
                rm, rv, nbt = bn_module.running_mean, bn_module.running_var, bn_module.num_batches_tracked
            if fuse_stats:
                _lib.call("vp_norm_finalize_parts", _ptr(parts), nparts, _ptr(g_), _ptr(b_), _ptr(rm), _ptr(rv), float(norm.momentum),
                          float(norm.eps), _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift), rpg, cc, _ptr(nbt), _stream())
            else:
                sums = torch.empty(2 * groups * cc, dtype=torch.float64, device=dev)
                _lib.call("vp_norm_stats", _ptr(y), _ptr(sums), _code(dt), groups, rpg, cc, _stream())
                _lib.call("vp_norm_finalize", _ptr(sums), _ptr(g_), _ptr(b_), _ptr(rm), _ptr(rv), float(norm.momentum), float(norm.eps),
                          _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift), groups, rpg, cc, _ptr(nbt), _stream())
        else:
            # eval-mode BatchNorm: running statistics (not on the training hot path; tiny [C] vectors)
            rm, rv = bn_module.running_mean, bn_module.running_var
            invstd.copy_(torch.rsqrt(rv + norm.eps))
            mean.copy_(rm)
            scale.copy_(invstd * (g_ if g_ is not None else 1.0))
            shift.copy_((b_ if b_ is not None else 0.0) - rm * scale)
        a = torch.empty(y.shape, dtype=out_dtype, device=dev) if out_dtype != dt else torch.empty_like(y)
        if out_dtype != dt:
            raise _lib.VaePlayError("normalised layers keep the activation dtype")
        _lib.call("vp_norm_apply_act", _ptr(y), _ptr(scale), _ptr(shift), _ptr(a), _code(dt), groups, rpg, cc, ACT[act],
                  float(slope), _stream())
        ctx.save_for_backward(x, weight, y, stats)
        ctx.dims = (groups, rpg, cc, c)
        ctx.train_stats = bool(training or norm.kind == "instance")
        ctx.has_affine = gamma is not None
        if holder is not None and norm.kind == "batch" and training and act == "relu" and dt == torch.bfloat16:
            # what the NEXT layer's data-gradient epilogue needs to do the first pass of this block's BatchNorm backward
            holder.update(ok=True, y=y, scale=scale, shift=shift, mean=mean)
        return a, y

    @staticmethod
    def backward(ctx, da, dy_extra):
        layer, norm, act, slope = ctx.layer, ctx.norm, ctx.act, ctx.slope
        dev = ctx.saved_tensors[0].device
        dgamma = dbeta = dbias = None
        if norm.kind is None:
            x, weight, a = ctx.saved_tensors
            if da is None:
                return (None,) * 14
            da = da.contiguous()
            dt = x.dtype
            n, h, w, c = a.shape
            rows = n * h * w
            sums = None
            if a.dtype != dt:
                # fp32 output of a bf16 layer (mu/logvar heads, the discriminator's sigmoid score): the activation
                # backward runs in fp32 on the fp32 output, then the gradient is brought to the activation dtype
                d32 = da if da.dtype == torch.float32 else da.float()
                if act not in (None, "none"):
                    dy32 = torch.empty_like(a)
                    sums = torch.empty(2 * c, dtype=torch.float64, device=dev)
                    _lib.call("vp_norm_bwd_reduce", _ptr(a), _ptr(d32), None, None, None, None, _ptr(sums), _ptr(dy32), F32,
                              1, rows, c, ACT[act] | 16, float(slope), _stream())
                    d32 = dy32
                d_in = torch.empty(a.shape, dtype=dt, device=dev)
                _lib.call("vp_cast", _ptr(d32), F32, _ptr(d_in), _code(dt), d32.numel(), _stream())
                dy = d_in
            elif act in (None, "none"):
                dy = da
            else:
                dy = torch.empty_like(a)
                # 16 / 32 channels (the discriminator's first layer at full resolution): an elementwise pass does not care where a
                # row ends -- present f = 64 / c consecutive pixels as ONE 64-channel row, which the streaming kernel serves, and
                # fold the f partial column sums afterwards
                f = 64 // c if (dt == torch.bfloat16 and c < 64 and 64 % c == 0 and rows % (64 // c) == 0 and rows >= 4096) else 1
                sums = torch.empty(2 * c * f, dtype=torch.float64, device=dev)
                _lib.call("vp_norm_bwd_reduce", _ptr(a), _ptr(da), None, None, None, None, _ptr(sums), _ptr(dy), _code(dt),
                          1, rows // f, c * f, ACT[act] | 16, float(slope), _stream())
                if f > 1:
                    sums = sums[:c * f].reshape(f, c).sum(0)
            if ctx.has_bias and sums is not None:
                dbias = sums[:c].float()        # the activation backward already reduced d over the rows
            elif ctx.has_bias:
                dbias = torch.empty(c, dtype=torch.float32, device=dev)
                scratch = torch.empty(2 * c, dtype=torch.float64, device=dev)
                _lib.call("vp_colsum", _ptr(dy), _ptr(dbias), _ptr(scratch), _code(dt), rows, c, _stream())
        else:
            x, weight, y, stats = ctx.saved_tensors
            mean, invstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
            groups, rpg, cc, c = ctx.dims
            dt = y.dtype
            if da is None and dy_extra is None:
                return (None,) * 14
            if da is None:
                # only the pre-norm output was used (Discriminator 'REC' mode, networks.py:180-185)
                dy = dy_extra.contiguous()
                fork = _async_fork(weight)
                dx = layer.dgrad(dy, weight, ctx.x_shape) if ctx.needs_input_grad[0] else None
                dw = _wgrad_maybe_async(layer, x, dy, weight, fork)
                return dx, dw, None, None, None, None, None, None, None, None, None, None, None, None
            da = da.contiguous()
            dy = torch.empty_like(y)
            if ctx.few_rows:
                if ctx.has_affine:
                    dgamma = torch.empty(cc, dtype=torch.float32, device=dev)
                    dbeta = torch.empty(cc, dtype=torch.float32, device=dev)
                _lib.call("vp_bn_rows_bwd", _ptr(y), _ptr(da), _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift), _ptr(dy), _ptr(dgamma),
                          _ptr(dbeta), _code(dt), rpg, cc, ACT[act], float(slope), _stream())
            elif ctx.train_stats:
                sums = torch.empty(2 * groups * cc, dtype=torch.float64, device=dev)
                pre = ctx.holder.pop("pre", None) if ctx.holder else None
                if pre is not None and pre[0] == da.data_ptr() and pre[1] == da._version and tuple(da.shape) == tuple(y.shape):
                    # the consumer's data-gradient epilogue already reduced d = da * relu'(.) over this very tensor
                    _lib.call("vp_norm_bwd_finish_parts", _ptr(pre[2]), pre[3], _ptr(invstd), _ptr(sums), cc, _stream())
                else:
                    _lib.call("vp_norm_bwd_reduce", _ptr(y), _ptr(da), _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift),
                              _ptr(sums), None, _code(dt), groups, rpg, cc, ACT[act], float(slope), _stream())
                if ctx.has_affine:
                    dgamma = torch.empty(cc, dtype=torch.float32, device=dev)
                    dbeta = torch.empty(cc, dtype=torch.float32, device=dev)
                _lib.call("vp_norm_bwd_apply", _ptr(y), _ptr(da), _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift),
                          _ptr(sums), _ptr(dy), _ptr(dgamma), _ptr(dbeta), _code(dt), groups, rpg, cc, ACT[act],
                          float(slope), _stream())
            else:
                raise _lib.VaePlayError("backward through eval-mode BatchNorm is not part of the training path")
            if dy_extra is not None:
                if dt != torch.float32:
                    raise _lib.VaePlayError("gradient into the pre-norm output is supported in fp32 mode only")
                extra = dy_extra.contiguous()
                _lib.call("vp_axpy", 1.0, _ptr(extra), _ptr(dy), dy.numel(), _stream())
            if ctx.has_bias:
                # a bias directly in front of Batch/InstanceNorm (StyleUp's ConvTranspose2d, network_Style_GAN.py:49-50) is
                # removed by the mean subtraction: its gradient is identically zero (the reference computes round-off noise)
                dbias = torch.zeros(cc if norm.kind == "batch" else c, dtype=torch.float32, device=dev)
        fork = _async_fork(weight)          # dy is complete at this point of the stream: a concurrent weight gradient may start here
        dx = None
        if ctx.needs_input_grad[0]:
            prev = ctx.prev_bn
            fused = layer.dgrad_bnred(dy, weight, ctx.x_shape, prev) if prev is not None and prev.get("ok") else None
            if fused is not None:
                dx, parts, nparts = fused
                prev["pre"] = (dx.data_ptr(), dx._version, parts, nparts)
            else:
                dx = layer.dgrad(dy, weight, ctx.x_shape)
        # launched AFTER the data gradient (the critical path keeps the SMs first); on the side stream it runs next to the
        # bandwidth-bound BatchNorm-backward passes of the block below
        dw = _wgrad_maybe_async(layer, x, dy, weight, fork)
        return dx, dw, dbias, dgamma, dbeta, None, None, None, None, None, None, None, None, None


def fused_layer(x, weight, bias, gamma, beta, layer, norm, act, slope, training, bn_module=None, out_dtype=None):
    # producer -> consumer link for the fused BatchNorm backward: the activated output of a conv-BatchNorm-ReLU block carries a
    # holder; the layer that consumes it hands it to its own backward, whose data-gradient epilogue does the reduction pass
    holder = {} if norm.kind == "batch" else None
    out = _FusedLayerFn.apply(x, weight, bias, gamma, beta, layer, norm, act, slope, training, bn_module, out_dtype, holder,
                              getattr(x, "_vp_bn", None))
    if holder:
        out[0]._vp_bn = holder
    if _TRACE is not None:
        _TRACE.append(out[0].detach())
    return out


def trace_activations(enable: bool):
    """Start (returns the list that will be filled) or stop recording each fused layer's activated output."""
    global _TRACE
    _TRACE = [] if enable else None
    return _TRACE


# ------------------------------------------------------------------------------------------------
# graph edges
# ------------------------------------------------------------------------------------------------
class _GradCut(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g


def grad_cut(x):
    """Identity whose autograd node does nothing: the place to cut a backward pass in two.  ``loss.backward(inputs=[..., t])``
    with a non-leaf ``t`` EXECUTES t's grad_fn (torch marks the nodes of all ``inputs`` as needed); cutting at the output of a
    conv block itself would therefore run that block's whole backward in the first stage and again in the second."""
    return _GradCut.apply(x)


class _ToCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        _require_cuda(x, "to_channels_last")
        n, c, h, w = x.shape
        y = torch.empty((n, h, w, c), dtype=dtype, device=x.device)
        if c == 1:
            _lib.call("vp_cast", _ptr(x), F32, _ptr(y), _code(dtype), x.numel(), _stream())
        else:
            _lib.call("vp_nchw_to_nhwc", _ptr(x), _ptr(y), _code(dtype), n, c, h, w, _stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        return _FromCL.apply(dy), None


class _FromCL(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        _require_cuda(a, "from_channels_last")
        n, h, w, c = a.shape
        ctx.dtype = a.dtype
        y = torch.empty((n, c, h, w), dtype=torch.float32, device=a.device)
        if c == 1:
            _lib.call("vp_cast", _ptr(a), _code(a.dtype), _ptr(y), F32, a.numel(), _stream())
        else:
            _lib.call("vp_nhwc_to_nchw", _ptr(a), _ptr(y), _code(a.dtype), n, c, h, w, _stream())
        return y

    @staticmethod
    def backward(ctx, dy):
        return _ToCL.apply(dy.contiguous(), ctx.dtype)


def to_channels_last(x_nchw: torch.Tensor) -> torch.Tensor:
    """NCHW fp32 -> NHWC activation dtype (graph entry)."""
    if x_nchw.dtype != torch.float32:
        x_nchw = x_nchw.float()
    return _ToCL.apply(x_nchw.contiguous(), act_dtype())


def from_channels_last(a: torch.Tensor) -> torch.Tensor:
    """NHWC activation dtype -> NCHW fp32 (graph exit)."""
    return _FromCL.apply(a)


class _TransposeBT(torch.autograd.Function):
    """[B, R, C] -> [B, C, R] in the same dtype (vp_transpose_bt); its own inverse with (R, C) swapped."""

    @staticmethod
    def forward(ctx, a, rows, cols):
        _require_cuda(a, "transpose")
        b = a.numel() // (rows * cols)
        out = torch.empty(a.numel(), dtype=a.dtype, device=a.device)
        _lib.call("vp_transpose_bt", _ptr(a), _ptr(out), _code(a.dtype), b, rows, cols, _stream())
        ctx.rc = (rows, cols)
        ctx.in_shape = a.shape
        return out

    @staticmethod
    def backward(ctx, d):
        rows, cols = ctx.rc
        return _TransposeBT.apply(d.contiguous(), cols, rows).reshape(ctx.in_shape), None, None


def hwc_to_chw_flat(a: torch.Tensor) -> torch.Tensor:
    """channels-last map [B,H,W,C] -> the NCHW-flatten vector ``ten.view(len(ten), -1)`` of the reference
    (models/networks.py:74-75), as [B,1,1,C*H*W]."""
    n, h, w, c = a.shape
    return _TransposeBT.apply(a, h * w, c).reshape(n, 1, 1, c * h * w)


def chw_flat_to_hwc(v: torch.Tensor, c: int, h: int, w: int) -> torch.Tensor:
    """[B,1,1,C*H*W] in NCHW-flatten order (``ten.view(len(ten), -1, 8, 8)``, models/networks.py:110) -> [B,H,W,C]."""
    n = v.shape[0]
    return _TransposeBT.apply(v, c, h * w).reshape(n, h, w, c)


# ------------------------------------------------------------------------------------------------
# reparameterisation + KL
# ------------------------------------------------------------------------------------------------
def _num_sms(dev) -> int:
    return torch.cuda.get_device_properties(dev).multi_processor_count


def philox_policy(n: int, num_sms: int):
    """(grid, offset increment) of ATen's calc_execution_policy for an n-element normal_()."""
    grid = min((n + 255) // 256, num_sms * 8)
    return grid, ((n - 1) // (256 * grid * 4) + 1) * 4


def _take_generator_state(dev, n):
    """Consume what Tensor.normal_() on n elements would consume from torch's default CUDA generator."""
    gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed(), gen.get_offset()
    _, inc = philox_policy(n, _num_sms(dev))
    gen.set_offset(offset + inc)
    return seed, offset


class _ReparamKL(torch.autograd.Function):
    """z = eps*exp(0.5*logvar) + mu and kl_b = -0.5*sum_j(1 + lv - mu^2 - e^lv) in one kernel.

    models/networks.py:228-231 (reparameterize) and :270 (KL).  eps is drawn in-kernel from the same
    Philox stream position ``logvar.new(size).normal_()`` would use, or taken from ``eps`` if given.
    """

    @staticmethod
    def forward(ctx, mu, logvar, eps, z_dtype, rng):
        if not mu.is_cuda:
            raise _lib.VaePlayError(f"reparameterize mu: tensor is on {mu.device}; vae_play_b200 has no CPU path")
        ctx.set_materialize_grads(False)
        ctx.packed = logvar is None
        if ctx.packed:  # mu is the fused head output [B, 2Z] = (mu | logvar)
            packed = mu
            zdim = packed.shape[1] // 2
            mu, logvar = packed[:, :zdim], packed[:, zdim:]
        rows, zdim = mu.shape
        if mu.stride(1) != 1 or logvar.stride(1) != 1 or mu.stride(0) != logvar.stride(0):
            mu, logvar = mu.contiguous(), logvar.contiguous()
        ld = mu.stride(0)
        dev = mu.device
        z = torch.empty((rows, zdim), dtype=z_dtype, device=dev)
        kl = torch.empty(rows, dtype=torch.float32, device=dev)
        if eps is None:
            eps_saved = torch.empty((rows, zdim), dtype=torch.float32, device=dev)
            if rng is None:
                seed, offset = _take_generator_state(dev, rows * zdim)
                off_dev = None
            else:
                seed, offset, off_dev = rng
            _lib.call("vp_reparam_kl_fwd", _ptr(mu), _ptr(logvar), ld, None, seed, offset, _ptr(off_dev), _num_sms(dev),
                      _ptr(z), _code(z_dtype), _ptr(eps_saved), _ptr(kl), rows, zdim, _stream())
        else:
            eps_saved = eps.contiguous().float()
            _lib.call("vp_reparam_kl_fwd", _ptr(mu), _ptr(logvar), ld, _ptr(eps_saved), 0, 0, None, _num_sms(dev),
                      _ptr(z), _code(z_dtype), None, _ptr(kl), rows, zdim, _stream())
        ctx.save_for_backward(mu, logvar, eps_saved)
        ctx.ld = ld
        return z, kl

    @staticmethod
    def backward(ctx, dz, dkl):
        mu, logvar, eps = ctx.saved_tensors
        rows, zdim = eps.shape
        dev = eps.device
        dzc = dz.contiguous() if dz is not None else None
        dklc = dkl.contiguous().float() if dkl is not None else None
        if ctx.packed:
            dml = torch.empty((rows, 2 * zdim), dtype=torch.float32, device=dev)
            d_mu, d_lv, ldo = dml[:, :zdim], dml[:, zdim:], 2 * zdim
        else:
            dml = torch.empty((2, rows, zdim), dtype=torch.float32, device=dev)
            d_mu, d_lv, ldo = dml[0], dml[1], zdim
        _lib.call("vp_reparam_kl_bwd", _ptr(mu), _ptr(logvar), ctx.ld, _ptr(eps), _ptr(dzc),
                  _code(dzc.dtype) if dzc is not None else F32, _ptr(dklc), _ptr(d_mu), _ptr(d_lv), F32, ldo, rows, zdim,
                  _stream())
        if ctx.packed:
            return dml, None, None, None, None
        return d_mu, d_lv, None, None, None


def reparam_kl(mu, logvar=None, eps=None, z_dtype=torch.float32, rng=None):
    """Returns (z [B,Z], kl [B]).  ``logvar=None``: ``mu`` is the fused head output [B,2Z] = (mu | logvar).
    ``rng=(seed, offset, offset_dev_tensor)`` pins the Philox position (CUDA-graph replay)."""
    return _ReparamKL.apply(mu, logvar, eps, z_dtype, rng)


def philox_advance(offset_dev: torch.Tensor, inc: int):
    """offset_dev (int64[1] on the device) += inc, stream-ordered (advances the Philox position inside a CUDA graph)."""
    _lib.call("vp_philox_advance", _ptr(offset_dev), int(inc), _stream())


def philox_normal(shape, device, rng=None):
    """Drop-in for ``torch.randn(shape, device='cuda')`` / ``Tensor.normal_()`` (same generator stream)."""
    n = int(math.prod(shape))
    out = torch.empty(shape, dtype=torch.float32, device=device)
    dev = out.device
    if rng is None:
        seed, offset = _take_generator_state(dev, n)
        off_dev = None
    else:
        seed, offset, off_dev = rng
    _lib.call("vp_philox_normal", _ptr(out), n, seed, offset, _ptr(off_dev), _num_sms(dev), _stream())
    return out


# ------------------------------------------------------------------------------------------------
# reconstruction losses
# ------------------------------------------------------------------------------------------------
_SCRATCH = {}


def _loss_scratch(dev):
    key = (dev.type, dev.index)
    if key not in _SCRATCH:
        _SCRATCH[key] = (torch.zeros(1, dtype=torch.float64, device=dev), torch.zeros(1, dtype=torch.int32, device=dev))
    return _SCRATCH[key]


class _ReconLoss(torch.autograd.Function):
    """mean((x_tilde-x)^2) (F.mse_loss, train.py:62) or mean(|x_tilde-x|) (F.l1_loss, train_Style_GAN.py:220)."""

    @staticmethod
    def forward(ctx, x, xt, kind):
        _require_cuda(xt, "recon loss")
        x = x.contiguous().float()
        xt = xt.contiguous()
        acc, counter = _loss_scratch(x.device)
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        _lib.call("vp_recon_loss_fwd", _ptr(x), _ptr(xt), x.numel(), kind, _ptr(acc), _ptr(counter), _ptr(loss), _stream())
        ctx.save_for_backward(x, xt)
        ctx.kind = kind
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        x, xt = ctx.saved_tensors
        dxt = torch.empty_like(xt)
        g = g.contiguous().float()
        _lib.call("vp_recon_loss_bwd", _ptr(x), _ptr(xt), x.numel(), ctx.kind, _ptr(g), _ptr(dxt), _stream())
        return None, dxt, None


def mse_loss(x, x_tilde):
    return _ReconLoss.apply(x, x_tilde, 0)


def l1_loss(x, x_tilde):
    return _ReconLoss.apply(x, x_tilde, 1)


class _BceDice(torch.autograd.Function):
    """bce_weight*BCEWithLogits(mean) + dice(sigmoid(logits)) (train_BE.py:58-59, tools/ops.py:12-19)."""

    @staticmethod
    def forward(ctx, logits, target, bce_weight):
        _require_cuda(logits, "bce_dice")
        logits = logits.contiguous().float()
        target = target.contiguous().float()
        rows = logits.shape[0]
        per = logits.numel() // rows
        dev = logits.device
        acc = torch.empty(rows * 4, dtype=torch.float64, device=dev)
        _, counter = _loss_scratch(dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.call("vp_bce_dice_fwd", _ptr(logits), _ptr(target), rows, per, float(bce_weight), _ptr(acc), _ptr(counter),
                  _ptr(loss), _stream())
        ctx.save_for_backward(logits, target, acc)
        ctx.w = float(bce_weight)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        logits, target, acc = ctx.saved_tensors
        rows = logits.shape[0]
        per = logits.numel() // rows
        d = torch.empty_like(logits)
        g = g.contiguous().float()
        _lib.call("vp_bce_dice_bwd", _ptr(logits), _ptr(target), rows, per, ctx.w, _ptr(acc), _ptr(g), _ptr(d), _stream())
        return d, None, None


def bce_dice_loss(logits, target, bce_weight=0.5):
    return _BceDice.apply(logits, target, bce_weight)


class _VaeLoss(torch.autograd.Function):
    """loss = F.mse_loss(x, x_tilde) + sum_b kl_b  (train.py:62-63, VAE terms) in two kernels, fused backward."""

    @staticmethod
    def forward(ctx, x, xt, kl, mse_scale):
        _require_cuda(xt, "vae_loss")
        x = x.contiguous().float()
        xt = xt.contiguous()
        kl = kl.contiguous()
        acc, counter = _loss_scratch(x.device)
        loss = torch.empty(1, dtype=torch.float32, device=x.device)
        _lib.call("vp_recon_loss_fwd", _ptr(x), _ptr(xt), x.numel(), 0, _ptr(acc), _ptr(counter), _ptr(loss), _stream())
        if mse_scale != 1.0:
            _lib.call("vp_axpy", float(mse_scale) - 1.0, _ptr(loss), _ptr(loss), 1, _stream())   # loss *= mse_scale
        _lib.call("vp_sum_into", _ptr(kl), kl.numel(), 1.0, _ptr(loss), _stream())
        ctx.save_for_backward(x, xt)
        ctx.nkl = kl.numel()
        ctx.mse_scale = float(mse_scale)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        x, xt = ctx.saved_tensors
        g = g.contiguous().float()
        dxt = torch.empty_like(xt)
        dkl = torch.empty(ctx.nkl, dtype=torch.float32, device=xt.device)
        if ctx.mse_scale != 1.0:
            gs = torch.empty_like(g)
            _lib.call("vp_fill_from", _ptr(g), ctx.mse_scale, _ptr(gs), 1, _stream())
        else:
            gs = g
        _lib.call("vp_recon_loss_bwd", _ptr(x), _ptr(xt), x.numel(), 0, _ptr(gs), _ptr(dxt), _stream())
        _lib.call("vp_fill_from", _ptr(g), 1.0, _ptr(dkl), ctx.nkl, _stream())
        return None, dxt, dkl, None


def vae_loss(x, x_tilde, kl, mse_scale=1.0):
    """mse_scale * F.mse_loss(x, x_tilde) + kl.sum() as one autograd node (mse_scale = 1/world_size under
    data parallelism, see vae_play_b200.parallel)."""
    return _VaeLoss.apply(x, x_tilde, kl, mse_scale)


class DualLinear:
    """Two nn.Linear heads that share an input (Encoder.l_mu / l_var, models/networks.py:69-70,76-77)
    evaluated as ONE contraction with N = 2*out: output [B, 2*out] = (mu | logvar), fp32."""

    def __init__(self, cin, cout):
        self.cin, self.cout = cin, cout
        self.layer = TapLayer("linear", cin, 2 * cout)
        self._cache = {}

    def _packed(self, w1, w2, which, dtype):
        key = (dtype, w1.data_ptr(), w1._version, w2.data_ptr(), w2._version, _EPOCH[0])
        hit = self._cache.get(which)
        if hit is not None and hit[0] == key:
            return hit[1]
        z, k = self.cout, self.cin
        wp = torch.empty(2 * z * k, dtype=dtype, device=w1.device)
        if which == "fwd":      # [2z][k]
            _lib.call("vp_pack_weight", _ptr(w1), _ptr(wp[: z * k]), _code(dtype), 1, z, k, k, 1, 1, _stream())
            _lib.call("vp_pack_weight", _ptr(w2), _ptr(wp[z * k:]), _code(dtype), 1, z, k, k, 1, 1, _stream())
        else:                   # dgrad: [k][2z]; built through a [2z][k] fp32 staging copy
            stage = torch.empty((2 * z, k), dtype=torch.float32, device=w1.device)
            _lib.call("vp_cast", _ptr(w1), F32, _ptr(stage[:z]), F32, z * k, _stream())
            _lib.call("vp_cast", _ptr(w2), F32, _ptr(stage[z:]), F32, z * k, _stream())
            _lib.call("vp_pack_weight", _ptr(stage), _ptr(wp), _code(dtype), 1, k, 2 * z, 1, k, 1, _stream())
        self._cache[which] = (key, wp)
        return wp


class _DualLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, dual: DualLinear):
        _require_cuda(x, "dual linear")
        n = x.shape[0]
        dt = x.dtype
        z, k = dual.cout, dual.cin
        dev = x.device
        wp = dual._packed(w1.detach(), w2.detach(), "fwd", dt)
        bias = torch.empty(2 * z, dtype=torch.float32, device=dev)
        _lib.call("vp_cast", _ptr(b1), F32, _ptr(bias[:z]), F32, z, _stream())
        _lib.call("vp_cast", _ptr(b2), F32, _ptr(bias[z:]), F32, z, _stream())
        y = torch.empty((n, 2 * z), dtype=torch.float32, device=dev)
        g = VpConvGeom(n, 1, 1, k, 1, 1, 2 * z, 1, 1, 1, 0, 0)
        _lib.call("vp_conv_fwd", C.byref(g), _ptr(x), _ptr(wp), _ptr(bias), _ptr(y), _code(dt), F32, 0, 0.0,
                  _STATE["engine"], _stream())
        ctx.save_for_backward(x, w1, w2)
        ctx.dual = dual
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w1, w2 = ctx.saved_tensors
        dual = ctx.dual
        n = x.shape[0]
        dt = x.dtype
        z, k = dual.cout, dual.cin
        dev = x.device
        dy = dy.contiguous()
        if dt != torch.float32:
            d = torch.empty((n, 2 * z), dtype=dt, device=dev)
            _lib.call("vp_cast", _ptr(dy), F32, _ptr(d), _code(dt), dy.numel(), _stream())
        else:
            d = dy
        db = torch.empty(2 * z, dtype=torch.float32, device=dev)
        scratch = torch.empty(4 * z, dtype=torch.float64, device=dev)
        _lib.call("vp_colsum", _ptr(d), _ptr(db), _ptr(scratch), _code(dt), n, 2 * z, _stream())
        g = VpConvGeom(n, 1, 1, k, 1, 1, 2 * z, 1, 1, 1, 0, 0)
        dw = torch.empty((2 * z, k), dtype=torch.float32, device=dev)  # packed [1][2z][k] == torch layout
        _lib.call("vp_conv_wgrad", C.byref(g), _ptr(x), _ptr(d), _ptr(dw), _code(dt), _STATE["engine"], _stream())
        dx = None
        if ctx.needs_input_grad[0]:
            wpt = dual._packed(w1.detach(), w2.detach(), "dgrad", dt)
            dx = torch.empty_like(x)
            _lib.call("vp_conv_dgrad", C.byref(g), _ptr(d), _ptr(wpt), _ptr(dx), _code(dt), _code(dt), _STATE["engine"], _stream())
        return dx, dw[:z], db[:z], dw[z:], db[z:], None


def dual_linear(x, w1, b1, w2, b2, dual: DualLinear):
    return _DualLinearFn.apply(x, w1, b1, w2, b2, dual)
