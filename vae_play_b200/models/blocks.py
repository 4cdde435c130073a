"""Drop-in mirror of the reference's ``models/blocks.py`` operator blocks on the B200 kernel library.

``Conv2d`` (conv -> [BatchNorm2d | InstanceNorm2d] -> [ReLU | LeakyReLU(0.02) | Tanh]) and ``Linear``
(linear -> [ReLU | LeakyReLU(0.2) | Tanh]) keep the reference's constructor signatures and
``state_dict`` keys (``conv.0.weight``, ``conv.1.running_mean``, ``fc.0.weight`` ...), reference
/root/reference/models/blocks.py:5-50.  SCSEBlock / SelfAttentionBlock / AddCoords / Down / Up are the
"next" rows of SURVEY.md section 8f and are not built yet.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as VF
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
