"""Drop-in mirror of the reference's ``models/blocks.py`` operator blocks on the B200 kernel library.

``Conv2d`` (conv -> [BatchNorm2d | InstanceNorm2d] -> [ReLU | LeakyReLU(0.02) | Tanh]) and ``Linear``
(linear -> [ReLU | LeakyReLU(0.2) | Tanh]) keep the reference's constructor signatures and
``state_dict`` keys (``conv.0.weight``, ``conv.1.running_mean``, ``fc.0.weight`` ...), reference
/root/reference/models/blocks.py:5-50.  ``SCSEBlock`` (:52-65), ``SelfAttentionBlock`` (:67-95), ``AddCoords`` (:97-112),
``Down`` (:114-127) and ``Up`` (:129-146) follow the same rule: reference attribute tree and keys, every arithmetic step a
kernel of libvaeplay_b200 (contractions: the tap-GEMM engines; the rest: csrc/blocks_ops.cu).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as VF
from .. import functional_blocks as VB
from ..functional import NormCfg, TapLayer


class Conv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, stride=1, bn=None, activate='relu'):
        super().__init__()
        mods = [nn.Conv2d(in_channel, out_channel, kernel_size, stride=stride, padding=(kernel_size - 1) // 2,
                          bias=(bn is None))]
        self._norm_idx = None
        if bn == "batch":
            mods.append(nn.BatchNorm2d(out_channel))
            self._norm_idx = 1
        elif bn == "instance":
            mods.append(nn.InstanceNorm2d(out_channel))
            self._norm_idx = 1
        self._act, self._slope = None, 0.0
        if activate == 'relu':
            mods.append(nn.ReLU())
            self._act = "relu"
        elif activate == 'lrelu':
            mods.append(nn.LeakyReLU(0.02))
            self._act, self._slope = "lrelu", 0.02
        elif activate == 'tanh':
            mods.append(nn.Tanh())
            self._act = "tanh"
        self.conv = nn.Sequential(*mods)
        self._bn = bn if bn in ("batch", "instance") else None
        self._layer = TapLayer("conv", in_channel, out_channel, k=kernel_size, stride=stride, pad=(kernel_size - 1) // 2)
        VF.weights_channels_last(self)

    def forward_cl(self, a):
        conv = self.conv[0]
        if self._bn == "batch":
            bn = self.conv[1]
            y, _ = VF.fused_layer(a, conv.weight, None, bn.weight, bn.bias, self._layer,
                                  NormCfg("batch", eps=bn.eps, momentum=bn.momentum), self._act, self._slope, self.training, bn)
        elif self._bn == "instance":
            inn = self.conv[1]
            y, _ = VF.fused_layer(a, conv.weight, None, None, None, self._layer, NormCfg("instance", eps=inn.eps),
                                  self._act, self._slope, self.training, None)
        else:
            y, _ = VF.fused_layer(a, conv.weight, conv.bias, None, None, self._layer, NormCfg(None), self._act, self._slope,
                                  self.training, None)
        return y

    def forward(self, input):
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(input)))


class Linear(nn.Module):
    def __init__(self, in_channel, out_channel, bias=True, activate='relu'):
        super().__init__()
        mods = [nn.Linear(in_channel, out_channel, bias=bias)]
        self._act, self._slope = None, 0.0
        if activate == 'relu':
            mods.append(nn.ReLU())
            self._act = "relu"
        elif activate == 'lrelu':
            mods.append(nn.LeakyReLU(0.2))
            self._act, self._slope = "lrelu", 0.2
        elif activate == 'tanh':
            mods.append(nn.Tanh())
            self._act = "tanh"
        self.fc = nn.Sequential(*mods)
        self._layer = TapLayer("linear", in_channel, out_channel)

    def forward_cl(self, a):
        lin = self.fc[0]
        y, _ = VF.fused_layer(a, lin.weight, lin.bias, None, None, self._layer, NormCfg(None), self._act, self._slope,
                              self.training, None)
        return y

    def forward(self, x):
        shp = x.shape
        a = VF.to_channels_last(x.reshape(-1, shp[-1], 1, 1))
        return VF.from_channels_last(self.forward_cl(a)).reshape(*shp[:-1], -1)


def _conv1x1(a, conv: nn.Conv2d, layer: TapLayer, act, training, out_dtype=None):
    y, _ = VF.fused_layer(a, conv.weight, conv.bias, None, None, layer, NormCfg(None), act, 0.0, training, None, out_dtype)
    return y


class SCSEBlock(nn.Module):
    """x * cSE(x) + x * sSE(x) -- reference blocks.py:52-65 (keys cSE.1/cSE.3/sSE.0 .weight/.bias)."""

    def __init__(self, in_channels, reduction=16):
        super().__init__()
        mid = in_channels // reduction
        self.cSE = nn.Sequential(
            nn.AdaptiveAvgPool2d(1),
            nn.Conv2d(in_channels, mid, 1),
            nn.ReLU(inplace=True),
            nn.Conv2d(mid, in_channels, 1),
            nn.Sigmoid(),
        )
        self.sSE = nn.Sequential(nn.Conv2d(in_channels, 1, 1), nn.Sigmoid())
        self._c1 = TapLayer("conv", in_channels, mid, k=1)
        self._c2 = TapLayer("conv", mid, in_channels, k=1)
        self._s = TapLayer("conv", in_channels, 1, k=1)

    def forward_cl(self, a):
        pooled = VB.adaptive_avgpool(a, 1, 1)                                        # [n,1,1,c]
        h = _conv1x1(pooled, self.cSE[1], self._c1, "relu", self.training)
        cse = _conv1x1(h, self.cSE[3], self._c2, "sigmoid", self.training)           # [n,1,1,c]
        sse = _conv1x1(a, self.sSE[0], self._s, "sigmoid", self.training)            # [n,h,w,1]
        return VB.scse_gate(a, cse, sse)

    def forward(self, x):
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(x)))


class SelfAttentionBlock(nn.Module):
    """gamma * attention(x) + x -- reference blocks.py:67-95 (q / k / v are blocks.Conv2d 1x1 with bias + ReLU, gamma init 0)."""

    def __init__(self, in_channel):
        super().__init__()
        self.q = Conv2d(in_channel, in_channel // 8, 1)
        self.k = Conv2d(in_channel, in_channel // 8, 1)
        self.v = Conv2d(in_channel, in_channel, 1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)

    def forward_cl(self, a):
        n, h, w, c = a.shape
        q = self.q.forward_cl(a).reshape(n, h * w, -1)          # channels-last: [b, pixel, channel] == proj_query of the reference
        k = self.k.forward_cl(a).reshape(n, h * w, -1)
        v = self.v.forward_cl(a).reshape(n, h * w, c)
        out = VB.attention_core(q, k, v).reshape(n, h, w, c)
        return VB.scale_add(self.gamma, out, a)

    def forward(self, x):
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(x)))


class AddCoords(nn.Module):
    """reference blocks.py:97-112: two extra channels holding the column / row index (optionally normalised to [-1, 1))."""

    def __init__(self, if_normalize=False):
        super().__init__()
        self.if_normalize = if_normalize

    def forward_cl(self, a):
        return VB.add_coords(a, self.if_normalize)

    def forward(self, x):
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(x)))


class Down(nn.Module):
    """reference blocks.py:114-127: [AddCoords ->] Conv2d(k, stride 2, bias, ReLU)."""

    def __init__(self, in_channel, out_channel, kernel_size, if_add_coord=False):
        super().__init__()
        self.if_add_coord = if_add_coord
        coord_channel = 2 if if_add_coord else 0
        self.conv = Conv2d(in_channel + coord_channel, out_channel, kernel_size, stride=2)
        if if_add_coord:
            self.add_coord = AddCoords()

    def forward_cl(self, a):
        if self.if_add_coord:
            a = self.add_coord.forward_cl(a)
        return self.conv.forward_cl(a)

    def forward(self, x):
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(x)))


class Up(nn.Module):
    """reference blocks.py:129-146: [AddCoords ->] 2 x (conv3x3 - BatchNorm - ReLU) -> bilinear x2 (align_corners=False)."""

    def __init__(self, in_channel, out_channel, if_add_coord=False):
        super().__init__()
        self.if_add_coord = if_add_coord
        coord_channel = 2 if if_add_coord else 0
        self.conv = nn.Sequential(
            Conv2d(in_channel + coord_channel, out_channel, 3, stride=1, bn="batch"),
            Conv2d(out_channel, out_channel, 3, stride=1, bn="batch")
        )
        if if_add_coord:
            self.add_coord = AddCoords()

    def forward_cl(self, a):
        if self.if_add_coord:
            a = self.add_coord.forward_cl(a)
        for blk in self.conv:
            a = blk.forward_cl(a)
        return VB.upsample2x(a)

    def forward(self, x):
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(x)))
