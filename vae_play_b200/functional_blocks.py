"""Autograd wrappers of the block-level operators around the contractions (models/blocks.py, models/network_Style_GAN.py,
tools/ops.py of the reference): channel concat, AddCoords, bilinear x2, adaptive average pooling, the SCSE gate, the
label-gated blend of ``myConv2d``, row softmax / the attention core, dice on probabilities and the edge loss.

Same conventions as ``functional.py``: channels-last activations ``[N,H,W,C]`` in the activation dtype, every arithmetic
step is one kernel of libvaeplay_b200 (csrc/blocks_ops.cu) called through the C ABI, torch only carries memory, streams and
the autograd graph.  No CPU path.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import F32
from .functional import _code, _loss_scratch, _ptr, _require_cuda, _stream


def _cl(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------
# torch.cat([a, b], dim=1) on channels-last tensors
# ------------------------------------------------------------------------------------------------
class _CatChannels(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, "cat_channels")
        _require_cuda(b, "cat_channels")
        if a.shape[:3] != b.shape[:3] or a.dtype != b.dtype:
            raise _lib.VaePlayError(f"cat_channels: {tuple(a.shape)} vs {tuple(b.shape)}")
        n, h, w, ca = a.shape
        cb = b.shape[3]
        out = torch.empty((n, h, w, ca + cb), dtype=a.dtype, device=a.device)
        rows = n * h * w
        _lib.call("vp_copy_channels", _ptr(a), ca, 0, _ptr(out), ca + cb, 0, ca, rows, _code(a.dtype), 0, _stream())
        _lib.call("vp_copy_channels", _ptr(b), cb, 0, _ptr(out), ca + cb, ca, cb, rows, _code(a.dtype), 0, _stream())
        ctx.ca, ctx.cb = ca, cb
        return out

    @staticmethod
    def backward(ctx, d):
        d = _cl(d)
        n, h, w, c = d.shape
        rows = n * h * w
        da = db = None
        if ctx.needs_input_grad[0]:
            da = torch.empty((n, h, w, ctx.ca), dtype=d.dtype, device=d.device)
            _lib.call("vp_copy_channels", _ptr(d), c, 0, _ptr(da), ctx.ca, 0, ctx.ca, rows, _code(d.dtype), 0, _stream())
        if ctx.needs_input_grad[1]:
            db = torch.empty((n, h, w, ctx.cb), dtype=d.dtype, device=d.device)
            _lib.call("vp_copy_channels", _ptr(d), c, ctx.ca, _ptr(db), ctx.cb, 0, ctx.cb, rows, _code(d.dtype), 0, _stream())
        return da, db


def cat_channels(a, b):
    """``torch.cat([a, b], dim=1)`` of the reference (network_Style_GAN.py:62,140,222) on channels-last tensors."""
    return _CatChannels.apply(a, b)


# ------------------------------------------------------------------------------------------------
# AddCoords (blocks.py:97-112)
# ------------------------------------------------------------------------------------------------
class _AddCoords(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, normalize):
        _require_cuda(x, "add_coords")
        n, h, w, c = x.shape
        out = torch.empty((n, h, w, c + 2), dtype=x.dtype, device=x.device)
        _lib.call("vp_add_coords", _ptr(x), _ptr(out), _code(x.dtype), n, h, w, c, int(bool(normalize)), _stream())
        ctx.c = c
        return out

    @staticmethod
    def backward(ctx, d):
        d = _cl(d)
        n, h, w, c2 = d.shape
        dx = torch.empty((n, h, w, ctx.c), dtype=d.dtype, device=d.device)
        _lib.call("vp_copy_channels", _ptr(d), c2, 0, _ptr(dx), ctx.c, 0, ctx.c, n * h * w, _code(d.dtype), 0, _stream())
        return dx, None


def add_coords(x, normalize=False):
    return _AddCoords.apply(x, normalize)


# ------------------------------------------------------------------------------------------------
# F.interpolate(scale_factor=2, mode='bilinear')  (blocks.py:145)
# ------------------------------------------------------------------------------------------------
class _Up2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "upsample2x")
        n, h, w, c = x.shape
        y = torch.empty((n, 2 * h, 2 * w, c), dtype=x.dtype, device=x.device)
        _lib.call("vp_upsample2x_fwd", _ptr(x), _ptr(y), _code(x.dtype), n, h, w, c, _stream())
        return y

    @staticmethod
    def backward(ctx, d):
        d = _cl(d)
        n, h2, w2, c = d.shape
        dx = torch.empty((n, h2 // 2, w2 // 2, c), dtype=d.dtype, device=d.device)
        _lib.call("vp_upsample2x_bwd", _ptr(d), _ptr(dx), _code(d.dtype), n, h2 // 2, w2 // 2, c, _stream())
        return dx


def upsample2x(x):
    return _Up2.apply(x)


# ------------------------------------------------------------------------------------------------
# nn.AdaptiveAvgPool2d
# ------------------------------------------------------------------------------------------------
class _AvgPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, oh, ow):
        _require_cuda(x, "adaptive_avgpool")
        n, h, w, c = x.shape
        y = torch.empty((n, oh, ow, c), dtype=x.dtype, device=x.device)
        _lib.call("vp_avgpool_fwd", _ptr(x), _ptr(y), _code(x.dtype), n, h, w, c, oh, ow, _stream())
        ctx.hw = (h, w)
        return y

    @staticmethod
    def backward(ctx, d):
        d = _cl(d)
        n, oh, ow, c = d.shape
        h, w = ctx.hw
        dx = torch.empty((n, h, w, c), dtype=d.dtype, device=d.device)
        _lib.call("vp_avgpool_bwd", _ptr(d), _ptr(dx), _code(d.dtype), n, h, w, c, oh, ow, _stream())
        return dx, None, None


def adaptive_avgpool(x, oh=1, ow=1):
    return _AvgPool.apply(x, oh, ow)


# ------------------------------------------------------------------------------------------------
# SCSE gate (blocks.py:64-65): x * cSE(x) + x * sSE(x)
# ------------------------------------------------------------------------------------------------
class _ScseGate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cse, sse):
        _require_cuda(x, "scse_gate")
        n, h, w, c = x.shape
        cse, sse = _cl(cse), _cl(sse)
        if cse.numel() != n * c or sse.numel() != n * h * w or cse.dtype != x.dtype or sse.dtype != x.dtype:
            raise _lib.VaePlayError("scse_gate: cse must be [n,1,1,c] and sse [n,h,w,1] in the activation dtype")
        y = torch.empty_like(x)
        _lib.call("vp_scse_fwd", _ptr(x), _ptr(cse), _ptr(sse), _ptr(y), _code(x.dtype), n, h * w, c, _stream())
        ctx.save_for_backward(x, cse, sse)
        return y

    @staticmethod
    def backward(ctx, d):
        x, cse, sse = ctx.saved_tensors
        d = _cl(d)
        n, h, w, c = x.shape
        dx = torch.empty_like(x)
        dcse32 = torch.empty((n, c), dtype=torch.float32, device=x.device)
        dsse = torch.empty_like(sse)
        _lib.call("vp_scse_bwd", _ptr(x), _ptr(cse), _ptr(sse), _ptr(d), _ptr(dx), _ptr(dcse32), _ptr(dsse), _code(x.dtype), n, h * w, c,
                  _stream())
        if x.dtype == torch.float32:
            dcse = dcse32.reshape(cse.shape)
        else:
            dcse = torch.empty(cse.shape, dtype=x.dtype, device=x.device)
            _lib.call("vp_cast", _ptr(dcse32), F32, _ptr(dcse), _code(x.dtype), dcse32.numel(), _stream())
        return dx, dcse, dsse


def scse_gate(x, cse, sse):
    return _ScseGate.apply(x, cse, sse)


# ------------------------------------------------------------------------------------------------
# myConv2d blend (network_Style_GAN.py:78-79)
# ------------------------------------------------------------------------------------------------
class _Blend(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a1, a2, label):
        _require_cuda(a1, "blend")
        a2 = _cl(a2)
        n = a1.shape[0]
        lab = label.reshape(-1).to(torch.float32).contiguous()
        if lab.numel() != n or a1.shape != a2.shape or a1.dtype != a2.dtype:
            raise _lib.VaePlayError("blend: one label per sample, two tensors of the same shape and dtype")
        y = torch.empty_like(a1)
        _lib.call("vp_blend_fwd", _ptr(a1), _ptr(a2), _ptr(lab), _ptr(y), _code(a1.dtype), n, a1.numel() // n, _stream())
        ctx.save_for_backward(lab)
        return y

    @staticmethod
    def backward(ctx, d):
        (lab,) = ctx.saved_tensors
        d = _cl(d)
        n = d.shape[0]
        d1, d2 = torch.empty_like(d), torch.empty_like(d)
        _lib.call("vp_blend_bwd", _ptr(d), _ptr(lab), _ptr(d1), _ptr(d2), _code(d.dtype), n, d.numel() // n, _stream())
        return d1, d2, None


def blend(a1, a2, label):
    """``a1 * (1 - label) + a2 * label`` with one label per sample."""
    return _Blend.apply(a1, a2, label)


# ------------------------------------------------------------------------------------------------
# softmax over the last axis / the attention core (blocks.py:84-91)
# ------------------------------------------------------------------------------------------------
class _Softmax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "softmax_rows")
        cols = x.shape[-1]
        y = torch.empty_like(x)
        _lib.call("vp_softmax_fwd", _ptr(x), _ptr(y), _code(x.dtype), x.numel() // cols, cols, _stream())
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, d):
        (y,) = ctx.saved_tensors
        d = _cl(d)
        cols = y.shape[-1]
        dx = torch.empty_like(y)
        _lib.call("vp_softmax_bwd", _ptr(y), _ptr(d), _ptr(dx), _code(y.dtype), y.numel() // cols, cols, _stream())
        return dx


def softmax_rows(x):
    return _Softmax.apply(x)


def _bmm(a, b, out, batch, m, n, k, sa, sb, accumulate=0):
    _lib.call("vp_bmm", _ptr(a), _ptr(b), _ptr(out), _code(out.dtype), batch, m, n, k, sa[0], sa[1], sa[2], sb[0], sb[1], sb[2], accumulate,
              _stream())


class _AttentionCore(torch.autograd.Function):
    """energy = q k^T, attention = softmax_j, out[b,i,:] = sum_j attention[b,i,j] v[b,j,:]; q, k [B,P,Cq], v [B,P,C] (P = h*w)."""

    @staticmethod
    def forward(ctx, q, k, v):
        _require_cuda(q, "attention")
        k, v = _cl(k), _cl(v)
        b, p, cq = q.shape
        c = v.shape[2]
        dt, dev = q.dtype, q.device
        energy = torch.empty((b, p, p), dtype=dt, device=dev)
        _bmm(q, k, energy, b, p, p, cq, (p * cq, cq, 1), (p * cq, 1, cq))                      # op(B)[k, j] = k[b, j, k]
        attn = torch.empty_like(energy)
        _lib.call("vp_softmax_fwd", _ptr(energy), _ptr(attn), _code(dt), b * p, p, _stream())
        out = torch.empty((b, p, c), dtype=dt, device=dev)
        _bmm(attn, v, out, b, p, c, p, (p * p, p, 1), (p * c, c, 1))
        ctx.save_for_backward(q, k, v, attn)
        return out

    @staticmethod
    def backward(ctx, d):
        q, k, v, attn = ctx.saved_tensors
        d = _cl(d)
        b, p, cq = q.shape
        c = v.shape[2]
        dt, dev = q.dtype, q.device
        dattn = torch.empty_like(attn)
        _bmm(d, v, dattn, b, p, p, c, (p * c, c, 1), (p * c, 1, c))                              # dattn[i,j] = sum_c d[i,c] v[j,c]
        dv = torch.empty_like(v)
        _bmm(attn, d, dv, b, p, c, p, (p * p, 1, p), (p * c, c, 1))                              # dv[j,c] = sum_i attn[i,j] d[i,c]
        de = torch.empty_like(attn)
        _lib.call("vp_softmax_bwd", _ptr(attn), _ptr(dattn), _ptr(de), _code(dt), b * p, p, _stream())
        dq, dk = torch.empty_like(q), torch.empty_like(k)
        _bmm(de, k, dq, b, p, cq, p, (p * p, p, 1), (p * cq, cq, 1))                             # dq[i,c] = sum_j de[i,j] k[j,c]
        _bmm(de, q, dk, b, p, cq, p, (p * p, 1, p), (p * cq, cq, 1))                             # dk[j,c] = sum_i de[i,j] q[i,c]
        return dq, dk, dv


def attention_core(q, k, v):
    return _AttentionCore.apply(q, k, v)


class _ScaleAdd(torch.autograd.Function):
    """gamma * a + x with a learnable device scalar gamma (blocks.py:72,93)."""

    @staticmethod
    def forward(ctx, gamma, a, x):
        _require_cuda(a, "scale_add")
        x = _cl(x)
        g32 = gamma.detach().reshape(1).to(torch.float32).contiguous()
        y = torch.empty_like(a)
        _lib.call("vp_scale_add", _ptr(g32), _ptr(a), _ptr(x), _ptr(y), _code(a.dtype), a.numel(), _stream())
        ctx.save_for_backward(g32, a)
        return y

    @staticmethod
    def backward(ctx, d):
        g32, a = ctx.saved_tensors
        d = _cl(d)
        da = torch.empty_like(a)
        zero = torch.zeros_like(a)
        _lib.call("vp_scale_add", _ptr(g32), _ptr(d), _ptr(zero), _ptr(da), _code(a.dtype), a.numel(), _stream())
        acc = torch.empty(1, dtype=torch.float64, device=a.device)
        _lib.call("vp_dot", _ptr(d), _ptr(a), _ptr(acc), _code(a.dtype), a.numel(), _stream())
        return acc.to(torch.float32).reshape(1), da, d


def scale_add(gamma, a, x):
    return _ScaleAdd.apply(gamma, a, x)


# ------------------------------------------------------------------------------------------------
# dice on probabilities (tools/ops.py:12-19) and the edge loss (tools/ops.py:187-214)
# ------------------------------------------------------------------------------------------------
class _Dice(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, t, smooth):
        _require_cuda(p, "dice_loss")
        p = _cl(p).float()
        t = _cl(t).float()
        rows = p.shape[0]
        per = p.numel() // rows
        dev = p.device
        acc = torch.empty(rows * 3, dtype=torch.float64, device=dev)
        _, counter = _loss_scratch(dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        _lib.call("vp_dice_fwd", _ptr(p), _ptr(t), rows, per, float(smooth), _ptr(acc), _ptr(counter), _ptr(loss), _stream())
        ctx.save_for_backward(t, acc)
        ctx.smooth, ctx.shape = float(smooth), p.shape
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        t, acc = ctx.saved_tensors
        rows = t.shape[0]
        per = t.numel() // rows
        dp = torch.empty(ctx.shape, dtype=torch.float32, device=t.device)
        g = g.contiguous().float()
        _lib.call("vp_dice_bwd", _ptr(t), rows, per, ctx.smooth, _ptr(acc), _ptr(g), _ptr(dp), _stream())
        return dp, None, None


def dice_loss(probabilities, targets, smooth=1.0):
    """``compute_dice_loss`` / ``dice_loss`` of tools/ops.py:12-19,178-185 (inputs are probabilities, NOT logits)."""
    return _Dice.apply(probabilities, targets, smooth)


class _EdgeMap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _require_cuda(x, "edge_map")
        x = x.float().contiguous()
        if x.dim() != 4 or x.shape[1] != 1:
            raise _lib.VaePlayError("edge_map: single-channel NCHW maps [n,1,h,w]")
        n, _, h, w = x.shape
        e, sgn = torch.empty_like(x), torch.empty_like(x)
        _lib.call("vp_edge_fwd", _ptr(x), _ptr(e), _ptr(sgn), n, h, w, _stream())
        ctx.save_for_backward(sgn)
        return e

    @staticmethod
    def backward(ctx, de):
        (sgn,) = ctx.saved_tensors
        de = de.contiguous().float()
        n, _, h, w = sgn.shape
        dx = torch.empty_like(sgn)
        _lib.call("vp_edge_bwd", _ptr(de), _ptr(sgn), _ptr(dx), n, h, w, _stream())
        return dx


def edge_map(x):
    """``filter(x).abs()`` of edge_loss: the 3x3 Laplacian-like kernel / 8, zero padding (tools/ops.py:193-211)."""
    return _EdgeMap.apply(x)


def edge_loss(mask_probabilities, mask_targets):
    """``edge_loss`` of tools/ops.py:187-214: dice between the edge maps of prediction and target."""
    return dice_loss(edge_map(mask_probabilities), edge_map(mask_targets).detach())


# ------------------------------------------------------------------------------------------------
# stand-alone pointwise activations (nn.ReLU / nn.Tanh / .sigmoid() applied to a tensor, not fused behind a contraction)
# ------------------------------------------------------------------------------------------------
class _Pointwise(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act, slope):
        _require_cuda(x, "activation")
        c = x.shape[-1]
        rows = x.numel() // c
        y = torch.empty_like(x)
        _lib.call("vp_norm_apply_act", _ptr(x), None, None, _ptr(y), _code(x.dtype), 1, rows, c, _lib.ACT[act], float(slope), _stream())
        ctx.save_for_backward(y)
        ctx.act, ctx.slope = act, slope
        return y

    @staticmethod
    def backward(ctx, d):
        (y,) = ctx.saved_tensors
        d = _cl(d)
        c = y.shape[-1]
        rows = y.numel() // c
        dx = torch.empty_like(y)
        sums = torch.empty(2 * c, dtype=torch.float64, device=y.device)
        # derivative from the OUTPUT (relu: y > 0, tanh: 1 - y^2, sigmoid: y (1 - y)); the per-channel sums are a by-product
        _lib.call("vp_norm_bwd_reduce", _ptr(y), _ptr(d), None, None, None, None, _ptr(sums), _ptr(dx), _code(y.dtype), 1, rows, c,
                  _lib.ACT[ctx.act] | 16, float(ctx.slope), _stream())
        return dx, None, None


def relu(x):
    return _Pointwise.apply(x, "relu", 0.0)


def tanh(x):
    return _Pointwise.apply(x, "tanh", 0.0)


def sigmoid(x):
    return _Pointwise.apply(x, "sigmoid", 0.0)


# ------------------------------------------------------------------------------------------------
# the remaining terms of VaeGan.loss / the train.py step (models/networks.py:264-281, train.py:62-67)
# ------------------------------------------------------------------------------------------------
class _FeatureMse(torch.autograd.Function):
    """out[r] = 0.5 * sum_j (a[r,j] - b[r,j])^2"""

    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, "feature_mse")
        a, b = a.float().contiguous(), b.float().contiguous()
        rows = a.shape[0]
        cols = a.numel() // rows
        out = torch.empty(rows, dtype=torch.float32, device=a.device)
        _lib.call("vp_feature_mse_fwd", _ptr(a), _ptr(b), _ptr(out), rows, cols, _stream())
        ctx.save_for_backward(a, b)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        rows = a.shape[0]
        cols = a.numel() // rows
        g = g.contiguous().float()
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        _lib.call("vp_feature_mse_bwd", _ptr(a), _ptr(b), _ptr(g), _ptr(da), _ptr(db), rows, cols, _stream())
        return da, db


def feature_mse(a, b):
    """``torch.sum(0.5 * (a - b) ** 2, 1)`` (networks.py:273)."""
    return _FeatureMse.apply(a, b)


def half_sqdiff(a, b):
    """``0.5 * (a - b) ** 2`` element-wise on [B, P] (the nle term, networks.py:267)."""
    shp = a.shape
    return _FeatureMse.apply(a.reshape(-1, 1), b.reshape(-1, 1)).reshape(shp)


class _NegLog(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p, sign, offset):
        _require_cuda(p, "neglog")
        p = p.float().contiguous()
        out = torch.empty_like(p)
        _lib.call("vp_neglog_fwd", _ptr(p), _ptr(out), p.numel(), float(sign), float(offset), _stream())
        ctx.save_for_backward(p)
        ctx.so = (float(sign), float(offset))
        return out

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        g = g.contiguous().float()
        dp = torch.empty_like(p)
        _lib.call("vp_neglog_bwd", _ptr(p), _ptr(g), _ptr(dp), p.numel(), ctx.so[0], ctx.so[1], _stream())
        return dp, None, None


def neglog(p, sign=1.0, offset=1e-3):
    """``-log(sign * p + offset)``: -log(D + 1e-3) with (1, 1e-3), -log(1 - D + 1e-3) with (-1, 1 + 1e-3) (networks.py:276-278)."""
    return _NegLog.apply(p, sign, offset)


class _SmoothL1Sum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, scale):
        _require_cuda(a, "smooth_l1_sum")
        a, b = a.float().contiguous(), b.float().contiguous()
        out = torch.empty(1, dtype=torch.float32, device=a.device)
        _lib.call("vp_smooth_l1_sum_fwd", _ptr(a), _ptr(b), _ptr(out), a.numel(), float(scale), _stream())
        ctx.save_for_backward(a, b)
        ctx.scale = float(scale)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g = g.contiguous().float().reshape(1)
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        _lib.call("vp_smooth_l1_sum_bwd", _ptr(a), _ptr(b), _ptr(g), _ptr(da), _ptr(db), a.numel(), ctx.scale, _stream())
        return da, db, None


def smooth_l1_sum(a, b, scale=1.0):
    """``scale * F.smooth_l1_loss(a, b, reduction="sum")`` (networks.py:279 with scale = 1 / batch)."""
    return _SmoothL1Sum.apply(a, b, scale)


class _KlPerSample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu, logvar):
        if not mu.is_cuda:
            raise _lib.VaePlayError(f"kl_per_sample: tensor is on {mu.device}; vae_play_b200 has no CPU path")
        if mu.stride(-1) != 1 or logvar.stride(-1) != 1 or mu.stride(0) != logvar.stride(0):
            mu, logvar = mu.contiguous(), logvar.contiguous()
        rows, z = mu.shape
        kl = torch.empty(rows, dtype=torch.float32, device=mu.device)
        _lib.call("vp_kl_fwd", _ptr(mu), _ptr(logvar), mu.stride(0), _ptr(kl), rows, z, _stream())
        ctx.save_for_backward(mu, logvar)
        return kl

    @staticmethod
    def backward(ctx, dkl):
        mu, logvar = ctx.saved_tensors
        rows, z = mu.shape
        dkl = dkl.contiguous().float()
        d = torch.empty((2, rows, z), dtype=torch.float32, device=mu.device)
        # dmu = dkl * mu, dlogvar = dkl * 0.5 * (exp(logvar) - 1): the KL half of the fused reparameterisation backward
        _lib.call("vp_reparam_kl_bwd", _ptr(mu), _ptr(logvar), mu.stride(0), _ptr(mu), None, F32, _ptr(dkl), _ptr(d[0]), _ptr(d[1]), F32, z, rows, z,
                  _stream())
        return d[0], d[1]


def kl_per_sample(mu, logvar):
    """``-0.5 * torch.sum(-logvar.exp() - mu ** 2 + logvar + 1, 1)`` (networks.py:270)."""
    return _KlPerSample.apply(mu, logvar)


class _WeightedSums(torch.autograd.Function):
    """sum_i w_i * sum(t_i) as one scalar (the loss combinations of train.py:63-67); backward fills each t_i's gradient with w_i * g."""

    @staticmethod
    def forward(ctx, weights, *tensors):
        dev = tensors[0].device
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        flat = []
        for w, t in zip(weights, tensors):
            _require_cuda(t, "weighted_sums")
            t = t.float().contiguous()
            flat.append(t)
            _lib.call("vp_sum_into", _ptr(t), t.numel(), float(w), _ptr(out), _stream())
        ctx.weights = [float(w) for w in weights]
        ctx.shapes = [t.shape for t in tensors]
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().float().reshape(1)
        outs = []
        for i, (w, shp) in enumerate(zip(ctx.weights, ctx.shapes)):
            if not ctx.needs_input_grad[1 + i]:
                outs.append(None)
                continue
            d = torch.empty(shp, dtype=torch.float32, device=g.device)
            _lib.call("vp_fill_from", _ptr(g), w, _ptr(d), d.numel(), _stream())
            outs.append(d)
        return (None, *outs)


def weighted_sums(tensors, weights):
    return _WeightedSums.apply(list(weights), *tensors)


class _NllPick(torch.autograd.Function):
    """out[b] = -log(s[b, label[b]])"""

    @staticmethod
    def forward(ctx, s, label):
        _require_cuda(s, "nll_pick")
        s = s.float().contiguous()
        label = label.to(torch.int64).contiguous()
        rows, cols = s.shape
        out = torch.empty(rows, dtype=torch.float32, device=s.device)
        _lib.call("vp_nll_pick_fwd", _ptr(s), _ptr(label), _ptr(out), rows, cols, _stream())
        ctx.save_for_backward(s, label)
        return out

    @staticmethod
    def backward(ctx, g):
        s, label = ctx.saved_tensors
        rows, cols = s.shape
        ds = torch.empty_like(s)
        g = g.contiguous().float()
        _lib.call("vp_nll_pick_bwd", _ptr(s), _ptr(label), _ptr(g), _ptr(ds), rows, cols, _stream())
        return ds, None


def cross_entropy(inputs, labels):
    """``F.cross_entropy(inputs, labels)`` (mean over the batch): log-softmax of ``inputs`` -- which in train_Style_GAN.py:219 are
    already softmax probabilities: the reference's double softmax is preserved -- then the negative log-likelihood of the label."""
    s = softmax_rows(inputs.float().contiguous())
    nll = _NllPick.apply(s, labels)
    return weighted_sums([nll], [1.0 / nll.numel()])


def binary_cross_entropy_const(p, target_one: bool):
    """``F.binary_cross_entropy(p, ones)`` / ``(p, zeros)`` (mean): -mean(log p) / -mean(log(1 - p)), log clamped at -100."""
    t = neglog(p, 1.0, 0.0) if target_one else neglog(p, -1.0, 1.0)
    return weighted_sums([t], [1.0 / t.numel()])
