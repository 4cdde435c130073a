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


def set_grad_sinks(sinks_by_param_id, params=None):
    """Register flat-bucket slots for parameter gradients (data-parallel training)."""
    _GRAD_SINKS.clear()
    _GRAD_SINKS.update(sinks_by_param_id)


def invalidate_caches():
    """Force re-packing of all weight panels on next use (call before CUDA-graph capture so that the
    pack kernels are part of the captured step)."""
    _EPOCH[0] += 1


def _grad_target(weight):
    hit = _GRAD_SINKS.get(weight.data_ptr())
    if hit is not None:
        view, param = hit
        if param.grad is None and view.shape == weight.shape:
            return view
    return torch.empty_like(weight, dtype=torch.float32)


def set_precision(mode: str):
    """'bf16' (tcgen05 tensor cores, bf16 activations) or 'fp32' (CUDA-core check mode)."""
    if mode not in ("bf16", "fp32"):
        raise ValueError(mode)
    _STATE["precision"] = mode


def get_precision() -> str:
    return _STATE["precision"]


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


class TapLayer:
    """One contraction layer: how its forward / dgrad / wgrad map onto vp_conv_{fwd,dgrad,wgrad}.

    kind: 'conv' (nn.Conv2d), 'convT' (nn.ConvTranspose2d), 'linear' (nn.Linear on [B,in]),
          'flatten_in'  (nn.Linear on the NCHW-flattened SxS map, models/networks.py:65,74-75),
          'flatten_out' (nn.Linear whose output is viewed as [B,C,S,S], models/networks.py:88,110).
    """

    def __init__(self, kind, cin, cout, k=1, stride=1, pad=0, out_pad=0, spatial=1):
        self.kind, self.cin, self.cout = kind, cin, cout
        self.k, self.stride, self.pad, self.out_pad, self.S = k, stride, pad, out_pad, spatial
        self._cache = {}
        T = k * k
        if kind == "conv":
            self.p_fwd = Pack(T, cout, cin, cin * T, T, 1)
            self.p_dgrad = Pack(T, cin, cout, T, cin * T, 1)
            self.p_wgrad = self.p_fwd
        elif kind == "convT":
            self.p_fwd = Pack(T, cout, cin, T, cout * T, 1)
            self.p_dgrad = Pack(T, cin, cout, cout * T, T, 1)
            self.p_wgrad = self.p_dgrad
        elif kind == "linear":
            self.p_fwd = Pack(1, cout, cin, cin, 1, 1)
            self.p_dgrad = Pack(1, cin, cout, 1, cin, 1)
            self.p_wgrad = self.p_fwd
        elif kind == "flatten_in":
            # weight [out, C*S*S] == conv weight [out, C, S, S]; cin = C
            T = spatial * spatial
            self.p_fwd = Pack(T, cout, cin, cin * T, T, 1)
            self.p_dgrad = Pack(T, cin, cout, T, cin * T, 1)     # [T][C][out] == plain linear with N' = t*C + c
            self.p_wgrad = self.p_fwd
        elif kind == "flatten_out":
            # weight [C*S*S, z]; cout = C, cin = z
            T = spatial * spatial
            self.p_fwd = Pack(T, cout, cin, T * cin, 1, cin)     # [T][C][z] == plain linear with N' = t*C + c
            self.p_dgrad = Pack(T, cin, cout, 1, T * cin, cin)   # [T][z][C]: 'conv' over the SxS map
            self.p_wgrad = self.p_dgrad
        else:
            raise ValueError(kind)

    # ---- geometry -------------------------------------------------------------------------------
    def out_shape(self, n, h, w):
        if self.kind == "conv":
            return (n, (h + 2 * self.pad - self.k) // self.stride + 1, (w + 2 * self.pad - self.k) // self.stride + 1, self.cout)
        if self.kind == "convT":
            return (n, (h - 1) * self.stride - 2 * self.pad + self.k + self.out_pad,
                    (w - 1) * self.stride - 2 * self.pad + self.k + self.out_pad, self.cout)
        if self.kind in ("linear", "flatten_in"):
            return (n, 1, 1, self.cout)
        return (n, self.S, self.S, self.cout)

    def _geom(self, n, hi, wi, ci, ho, wo, co, k, stride, pad, transposed):
        return VpConvGeom(n, hi, wi, ci, ho, wo, co, k, k, stride, pad, transposed)

    def _packed(self, weight, which, dtype):
        key = (which, dtype, weight.data_ptr(), weight._version, _EPOCH[0])
        hit = self._cache.get(which)
        if hit is not None and hit[0] == key:
            return hit[1]
        p: Pack = getattr(self, "p_" + which)
        wp = torch.empty(p.taps * p.n * p.k, dtype=dtype, device=weight.device)
        _lib.call("vp_pack_weight", _ptr(weight), _ptr(wp), _code(dtype), p.taps, p.n, p.k, p.sn, p.sk, p.st, _stream())
        self._cache[which] = (key, wp)
        return wp

    # ---- thin layers (<= 2 channels on one side): tcgen05 kernels that read the fp32 master weight directly ----
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
        if self._thin("fwd", dt, weight):
            shp = self.out_shape(n, h, w)
            y = torch.empty(shp, dtype=out_dtype, device=x.device)
            g = self._geom(n, h, w, self.cin, shp[1], shp[2], self.cout, self.k, self.stride, self.pad, 0)
            _lib.call("vp_thin_conv_fwd", C.byref(g), _ptr(x), _ptr(weight.detach()), _ptr(bias), _ptr(y), _code(out_dtype),
                      ACT[act], float(slope), _stream())
            return y
        wp = self._packed(weight.detach(), "fwd", dt)
        shp = self.out_shape(n, h, w)
        y = torch.empty(shp, dtype=out_dtype, device=x.device)
        if self.kind == "flatten_out":
            g = self._geom(n, 1, 1, self.cin, 1, 1, self.S * self.S * self.cout, 1, 1, 0, 0)
        elif self.kind == "flatten_in":
            g = self._geom(n, h, w, self.cin, 1, 1, self.cout, self.S, 1, 0, 0)
        else:
            g = self._geom(n, h, w, self.cin, shp[1], shp[2], self.cout, self.k, self.stride, self.pad, int(self.kind == "convT"))
        _lib.call("vp_conv_fwd", C.byref(g), _ptr(x), _ptr(wp), _ptr(bias), _ptr(y), _code(dt), _code(out_dtype),
                  ACT[act], float(slope), _STATE["engine"], _stream())
        return y

    def dgrad(self, dy, weight, x_shape, out_dtype=None):
        n, h, w, _ = x_shape
        dt = dy.dtype
        out_dtype = out_dtype or dt
        if self._thin("dgrad", dt, weight):
            dx = torch.empty(x_shape, dtype=out_dtype, device=dy.device)
            g = self._geom(n, h, w, self.cin, dy.shape[1], dy.shape[2], self.cout, self.k, self.stride, self.pad, 0)
            _lib.call("vp_thin_conv_dgrad", C.byref(g), _ptr(dy), _ptr(weight.detach()), _ptr(dx), _code(out_dtype), _stream())
            return dx
        wp = self._packed(weight.detach(), "dgrad", dt)
        dx = torch.empty(x_shape, dtype=out_dtype, device=dy.device)
        if self.kind == "flatten_in":
            # expansion: plain linear [n,out] -> [n, S*S*C] with the permuted packing
            g = self._geom(n, 1, 1, self.cout, 1, 1, self.S * self.S * self.cin, 1, 1, 0, 0)
            _lib.call("vp_conv_fwd", C.byref(g), _ptr(dy), _ptr(wp), None, _ptr(dx), _code(dt), _code(out_dtype), 0, 0.0,
                      _STATE["engine"], _stream())
        elif self.kind == "flatten_out":
            # contraction over the SxS map: 'conv' k=S p=0 with A = dy
            g = self._geom(n, self.S, self.S, self.cout, 1, 1, self.cin, self.S, 1, 0, 0)
            _lib.call("vp_conv_fwd", C.byref(g), _ptr(dy), _ptr(wp), None, _ptr(dx), _code(dt), _code(out_dtype), 0, 0.0,
                      _STATE["engine"], _stream())
        else:
            ho, wo = dy.shape[1], dy.shape[2]
            g = self._geom(n, h, w, self.cin, ho, wo, self.cout, self.k, self.stride, self.pad, int(self.kind == "convT"))
            _lib.call("vp_conv_dgrad", C.byref(g), _ptr(dy), _ptr(wp), _ptr(dx), _code(dt), _code(out_dtype),
                      _STATE["engine"], _stream())
        return dx

    def wgrad(self, x, dy, weight):
        n, h, w, _ = x.shape
        dt = x.dtype
        if self._thin("wgrad", dt, weight):
            dw = _grad_target(weight)
            g = self._geom(n, h, w, self.cin, dy.shape[1], dy.shape[2], self.cout, self.k, self.stride, self.pad, 0)
            _lib.call("vp_thin_conv_wgrad", C.byref(g), _ptr(x), _ptr(dy), _ptr(dw), _stream())
            return dw
        p: Pack = self.p_wgrad
        dwp = torch.empty(p.taps * p.n * p.k, dtype=torch.float32, device=x.device)
        if self.kind == "flatten_out":
            # dW[c*T+t][k] = sum_b dY[b,t,c] z[b,k]: conv-wgrad with the roles x := dY (SxS map), dy := z
            g = self._geom(n, self.S, self.S, self.cout, 1, 1, self.cin, self.S, 1, 0, 0)
            _lib.call("vp_conv_wgrad", C.byref(g), _ptr(dy), _ptr(x), _ptr(dwp), _code(dt), _STATE["engine"], _stream())
        elif self.kind == "flatten_in":
            g = self._geom(n, h, w, self.cin, 1, 1, self.cout, self.S, 1, 0, 0)
            _lib.call("vp_conv_wgrad", C.byref(g), _ptr(x), _ptr(dy), _ptr(dwp), _code(dt), _STATE["engine"], _stream())
        else:
            ho, wo = dy.shape[1], dy.shape[2]
            g = self._geom(n, h, w, self.cin, ho, wo, self.cout, self.k, self.stride, self.pad, int(self.kind == "convT"))
            _lib.call("vp_conv_wgrad", C.byref(g), _ptr(x), _ptr(dy), _ptr(dwp), _code(dt), _STATE["engine"], _stream())
        dw = _grad_target(weight)
        _lib.call("vp_unpack_wgrad", _ptr(dwp), _ptr(dw), p.taps, p.n, p.k, p.sn, p.sk, p.st, _stream())
        return dw


def _permute_vec(v, T, Cn, inverse=False):
    """feature order c*T+t (torch BatchNorm1d after Linear) <-> t*C+c (our channels-last view)."""
    out = torch.empty_like(v)
    if not inverse:
        _lib.call("vp_pack_weight", _ptr(v), _ptr(out), F32, T, Cn, 1, T, 0, 1, _stream())
    else:
        _lib.call("vp_unpack_wgrad", _ptr(v), _ptr(out), T, Cn, 1, T, 0, 1, _stream())
    return out


@dataclass
class NormCfg:
    kind: Optional[str]  # 'batch' | 'instance' | None
    eps: float = 1e-5
    momentum: float = 0.1
    perm_T: int = 0       # >0: BatchNorm1d over a flatten_out layer, features permuted with T = S*S


class _FusedLayerFn(torch.autograd.Function):
    """contraction (+bias) -> [BatchNorm | InstanceNorm] -> activation, with the matching backward.

    Mirrors models/blocks.py:31-34 (Conv2d block), models/networks.py:27-30 (EncoderBlock),
    :42-46 (DecoderBlock) and the Linear->BatchNorm1d->ReLU stacks at :65-67,88-90.
    Returns (activated output, pre-norm contraction output).
    """

    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, layer: TapLayer, norm: NormCfg, act, slope, training, bn_module,
                out_dtype):
        _require_cuda(x, "fused layer input")
        ctx.set_materialize_grads(False)
        dt = x.dtype
        out_dtype = out_dtype or dt
        ctx.layer, ctx.norm, ctx.act, ctx.slope = layer, norm, act, slope
        ctx.x_shape = tuple(x.shape)
        ctx.has_bias = bias is not None
        if norm.kind is None:
            a = layer.fwd(x, weight, bias, act, slope, out_dtype)
            ctx.save_for_backward(x, weight, a)
            ctx.out_dtype = out_dtype
            return a, None
        y = layer.fwd(x, weight, bias, "none", 0.0, dt)
        n, h, w, c = y.shape
        if norm.perm_T:
            feat = h * w * c  # BatchNorm1d over the flattened (permuted) feature vector
            groups, rpg, cc = 1, n, feat
        elif norm.kind == "batch":
            groups, rpg, cc = 1, n * h * w, c
        else:
            groups, rpg, cc = n, h * w, c
        dev = x.device
        stats = torch.empty(4, groups * cc, dtype=torch.float32, device=dev)  # mean, invstd, scale, shift
        mean, invstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
        g_, b_ = gamma, beta
        if norm.perm_T and gamma is not None:
            g_, b_ = _permute_vec(gamma.detach(), norm.perm_T, c), _permute_vec(beta.detach(), norm.perm_T, c)
        if training or norm.kind == "instance":
            sums = torch.empty(2 * groups * cc, dtype=torch.float64, device=dev)
            _lib.call("vp_norm_stats", _ptr(y), _ptr(sums), _code(dt), groups, rpg, cc, _stream())
            rm = rv = None
            if norm.kind == "batch" and bn_module is not None and bn_module.track_running_stats:
                rm, rv = bn_module.running_mean, bn_module.running_var
                if norm.perm_T:
                    rm_p, rv_p = _permute_vec(rm, norm.perm_T, c), _permute_vec(rv, norm.perm_T, c)
                else:
                    rm_p, rv_p = rm, rv
            _lib.call("vp_norm_finalize", _ptr(sums), _ptr(g_), _ptr(b_), _ptr(rm_p if rm is not None else None),
                      _ptr(rv_p if rm is not None else None), float(norm.momentum), float(norm.eps),
                      _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift), groups, rpg, cc, _stream())
            if rm is not None:
                if norm.perm_T:
                    rm.copy_(_permute_vec(rm_p, norm.perm_T, c, inverse=True))
                    rv.copy_(_permute_vec(rv_p, norm.perm_T, c, inverse=True))
                bn_module.num_batches_tracked += 1
        else:
            # eval-mode BatchNorm: running statistics (not on the training hot path; tiny [C] vectors)
            rm, rv = bn_module.running_mean, bn_module.running_var
            if norm.perm_T:
                rm, rv = _permute_vec(rm, norm.perm_T, c), _permute_vec(rv, norm.perm_T, c)
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
        return a, y

    @staticmethod
    def backward(ctx, da, dy_extra):
        layer, norm, act, slope = ctx.layer, ctx.norm, ctx.act, ctx.slope
        dev = da.device
        dgamma = dbeta = dbias = None
        if norm.kind is None:
            x, weight, a = ctx.saved_tensors
            if da is None:
                return (None,) * 12
            da = da.contiguous()
            dt = x.dtype
            n, h, w, c = a.shape
            rows = n * h * w
            if a.dtype != dt:
                # fp32 output of a bf16 layer (mu/logvar heads): bring the gradient to the activation dtype
                d_in = torch.empty(a.shape, dtype=dt, device=dev)
                _lib.call("vp_cast", _ptr(da), _code(da.dtype), _ptr(d_in), _code(dt), da.numel(), _stream())
                if act not in (None, "none"):
                    raise _lib.VaePlayError("fp32-output layers must have no activation")
                dy = d_in
            elif act in (None, "none"):
                dy = da
            else:
                dy = torch.empty_like(a)
                sums = torch.empty(2 * c, dtype=torch.float64, device=dev)
                _lib.call("vp_norm_bwd_reduce", _ptr(a), _ptr(da), None, None, None, None, _ptr(sums), _ptr(dy), _code(dt),
                          1, rows, c, ACT[act] | 16, float(slope), _stream())
            if ctx.has_bias:
                dbias = torch.empty(c, dtype=torch.float32, device=dev)
                scratch = torch.empty(2 * c, dtype=torch.float64, device=dev)
                _lib.call("vp_colsum", _ptr(dy), _ptr(dbias), _ptr(scratch), _code(dt), rows, c, _stream())
        else:
            x, weight, y, stats = ctx.saved_tensors
            mean, invstd, scale, shift = stats[0], stats[1], stats[2], stats[3]
            groups, rpg, cc, c = ctx.dims
            dt = y.dtype
            if da is None and dy_extra is None:
                return (None,) * 12
            if da is None:
                # only the pre-norm output was used (Discriminator 'REC' mode, networks.py:180-185)
                dy = dy_extra.contiguous()
                dw = layer.wgrad(x, dy, weight)
                dx = layer.dgrad(dy, weight, ctx.x_shape) if ctx.needs_input_grad[0] else None
                return dx, dw, None, None, None, None, None, None, None, None, None, None
            da = da.contiguous()
            dy = torch.empty_like(y)
            if ctx.train_stats:
                sums = torch.empty(2 * groups * cc, dtype=torch.float64, device=dev)
                _lib.call("vp_norm_bwd_reduce", _ptr(y), _ptr(da), _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift),
                          _ptr(sums), None, _code(dt), groups, rpg, cc, ACT[act], float(slope), _stream())
                if ctx.has_affine:
                    dgamma = torch.empty(cc, dtype=torch.float32, device=dev)
                    dbeta = torch.empty(cc, dtype=torch.float32, device=dev)
                _lib.call("vp_norm_bwd_apply", _ptr(y), _ptr(da), _ptr(mean), _ptr(invstd), _ptr(scale), _ptr(shift),
                          _ptr(sums), _ptr(dy), _ptr(dgamma), _ptr(dbeta), _code(dt), groups, rpg, cc, ACT[act],
                          float(slope), _stream())
                if norm.perm_T and dgamma is not None:
                    dgamma = _permute_vec(dgamma, norm.perm_T, c, inverse=True)
                    dbeta = _permute_vec(dbeta, norm.perm_T, c, inverse=True)
            else:
                raise _lib.VaePlayError("backward through eval-mode BatchNorm is not part of the training path")
            if dy_extra is not None:
                if dt != torch.float32:
                    raise _lib.VaePlayError("gradient into the pre-norm output is supported in fp32 mode only")
                _lib.call("vp_axpy", 1.0, _ptr(dy_extra.contiguous()), _ptr(dy), dy.numel(), _stream())
        dw = layer.wgrad(x, dy, weight)
        dx = layer.dgrad(dy, weight, ctx.x_shape) if ctx.needs_input_grad[0] else None
        return dx, dw, dbias, dgamma, dbeta, None, None, None, None, None, None, None


def fused_layer(x, weight, bias, gamma, beta, layer, norm, act, slope, training, bn_module=None, out_dtype=None):
    out = _FusedLayerFn.apply(x, weight, bias, gamma, beta, layer, norm, act, slope, training, bn_module, out_dtype)
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
        _require_cuda(mu, "reparameterize mu")
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
