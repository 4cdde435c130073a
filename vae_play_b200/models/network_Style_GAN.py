"""Drop-in mirror of the reference's ``models/network_Style_GAN.py`` (config 5 of BASELINE.json) on the B200 kernel library.

Same class names, constructor signatures, attribute tree and ``state_dict`` keys as the reference
(/root/reference/models/network_Style_GAN.py:12-229): ``StyleEncoder`` (:12-41), ``StyleUp`` (:45-65), ``myConv2d`` (:72-79),
``Generator`` (:81-180), ``MLP`` (:182-199), ``Discriminator`` (:201-229).  Parameters live in the same ``nn`` shells the
reference builds (through the ``blocks`` mirrors), every arithmetic step is a kernel of libvaeplay_b200; activations travel
channels-last between layers and are NCHW fp32 only at the module boundaries the reference exposes.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .. import functional as VF
from .. import functional_blocks as VB
from ..functional import NormCfg, TapLayer
from .blocks import Conv2d, Linear, SCSEBlock

IMAGE_CHANNEL = 3


class StyleEncoder(nn.Module):
    def __init__(self, z_dim, image_size, max_channels=1024):
        super().__init__()
        in_dim = IMAGE_CHANNEL
        out_dim = 64
        convs = [Conv2d(in_dim, out_dim, 5, 1, activate=None)]
        n_level = int(np.log2(image_size)) - 2
        for _ in range(n_level):
            in_dim = out_dim
            out_dim = min(out_dim * 2, max_channels)
            convs.append(Conv2d(in_dim, out_dim, 3, stride=2, bn="instance"))
        convs.append(Conv2d(out_dim, out_dim, 3, stride=2))
        convs.append(Conv2d(out_dim, out_dim, 3, stride=2))
        self.convs = nn.Sequential(*convs)
        self.fc_mu = Linear(out_dim, z_dim, activate=None)
        self.fc_logvar = Linear(out_dim, z_dim, activate=None)

    def forward(self, x):
        a = VF.to_channels_last(x)
        for blk in self.convs:
            a = blk.forward_cl(a)
        n, h, w, c = a.shape
        if h * w != 1:
            # x.reshape(x.size(0), -1) of the reference flattens NCHW: bring the map to (c, y, x) order first
            a = VF.hwc_to_chw_flat(a)
        a = a.reshape(n, 1, 1, -1)
        mu = VF.from_channels_last(self.fc_mu.forward_cl(a)).reshape(n, -1)
        logvar = VF.from_channels_last(self.fc_logvar.forward_cl(a)).reshape(n, -1)
        return mu, logvar


class StyleUp(nn.Module):
    """ConvTranspose2d(k4 s2 p1, bias) -> InstanceNorm -> ReLU; cat with the skip; Conv2d 3x3 (+bias, ReLU); 2 x SCSE; ReLU."""

    def __init__(self, in_channel, out_channel):
        super().__init__()
        self.up_convs = nn.Sequential(
            nn.ConvTranspose2d(in_channel, out_channel, 4, 2, 1),
            nn.InstanceNorm2d(out_channel),
            nn.ReLU()
        )
        self.cat_convs = nn.Sequential(
            Conv2d(out_channel * 2, out_channel, 3),
            SCSEBlock(out_channel, reduction=4),
            SCSEBlock(out_channel, reduction=4),
            nn.ReLU()
        )
        self._up = TapLayer("convT", in_channel, out_channel, k=4, stride=2, pad=1, out_pad=0)
        VF.weights_channels_last(self.up_convs)

    def forward_cl(self, a, skip):
        ct = self.up_convs[0]
        a, _ = VF.fused_layer(a, ct.weight, ct.bias, None, None, self._up, NormCfg("instance", eps=self.up_convs[1].eps), "relu", 0.0,
                              self.training, None)
        a = VB.cat_channels(a, skip)
        a = self.cat_convs[0].forward_cl(a)
        a = self.cat_convs[1].forward_cl(a)
        a = self.cat_convs[2].forward_cl(a)
        return VB.relu(a)

    def forward(self, x, skip):
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(x), VF.to_channels_last(skip)))


class myConv2d(nn.Module):
    """conv_1(x) * (1 - label) + conv_2(x) * label (reference :72-79): two ``blocks.Conv2d`` gated per sample."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, bn=None, activate='relu'):
        super().__init__()
        self.conv_1 = Conv2d(in_channel, out_channel, kernel_size, stride, bn, activate)
        self.conv_2 = Conv2d(in_channel, out_channel, kernel_size, stride, bn, activate)

    def forward_cl(self, a, label):
        return VB.blend(self.conv_1.forward_cl(a), self.conv_2.forward_cl(a), label)

    def forward(self, x, label):
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(x), label))


class MLP(nn.Module):
    def __init__(self, nf_in, nf_out, num_blocks):
        super().__init__()
        model = []
        in_dim = nf_in
        out_dim = nf_in
        model.append(Linear(in_dim, out_dim, activate=None))
        ratio = int(2 ** (int(np.log2(nf_out / nf_in)) / (num_blocks - 1)))
        for _ in range(num_blocks - 2):
            in_dim = out_dim
            out_dim = min(in_dim * ratio, nf_out)
            model.append(Linear(in_dim, out_dim, activate=None))
        model.append(Linear(out_dim, nf_out, activate=None))
        self.model = nn.Sequential(*model)

    def forward_cl(self, a):
        for m in self.model:
            a = m.forward_cl(a)
        return a

    def forward(self, x):
        x = x.reshape(x.size(0), -1)
        a = VF.to_channels_last(x.reshape(x.size(0), -1, 1, 1))
        return VF.from_channels_last(self.forward_cl(a)).reshape(x.size(0), -1)


class Generator(nn.Module):
    def __init__(self, image_size, z_dim, max_channels=256):
        super().__init__()
        self.z_dim = z_dim
        self.image_size = image_size
        self.conv1 = myConv2d(IMAGE_CHANNEL + 1, 32, 3, 1, activate=None)
        self.conv2 = myConv2d(32, 32, 3, 1, activate=None)
        self.down1 = myConv2d(32, 64, 4, 2, bn="instance")
        self.down2 = myConv2d(64, 128, 4, 2, bn="instance")
        self.down3 = myConv2d(128, 256, 4, 2, bn="instance")
        self.down4 = myConv2d(256, 256, 4, 2, bn="instance")
        self.up1 = StyleUp(256, 256)
        self.up2 = StyleUp(256, 128)
        self.up3 = StyleUp(128, 64)
        self.skip1 = Conv2d(256, 256, 3, 1, bn="instance")
        self.skip2 = Conv2d(128, 128, 3, 1, bn="instance")
        self.skip3 = Conv2d(64, 64, 3, 1, bn="instance")
        self.final = nn.Sequential(
            nn.ConvTranspose2d(64, 32, 4, 2, 1),
            Conv2d(32, 32, 3, 1, bn=None),
            Conv2d(32, 32, 3, 1, bn=None),
            Conv2d(32, IMAGE_CHANNEL, 3, 1, bn=None, activate=None),
            nn.Tanh()
        )
        self._final_up = TapLayer("convT", 64, 32, k=4, stride=2, pad=1, out_pad=0)
        self.mlp = MLP(z_dim, image_size * image_size, 3)

    # ---- channels-last internals -----------------------------------------------------------------------------------------
    def encode_cl(self, x, style_code, labels):
        n = x.size(0)
        code = self.mlp.forward_cl(VF.to_channels_last(style_code.reshape(n, -1, 1, 1)))      # [n,1,1,S*S]
        code = code.reshape(n, self.image_size, self.image_size, 1)                          # == reshape(n, 1, S, S) in NCHW
        a = VB.cat_channels(VF.to_channels_last(x), code)
        labels = labels.reshape(n)
        c0 = self.conv2.forward_cl(self.conv1.forward_cl(a, labels), labels)
        d1 = self.down1.forward_cl(c0, labels)
        d2 = self.down2.forward_cl(d1, labels)
        d3 = self.down3.forward_cl(d2, labels)
        d4 = self.down4.forward_cl(d3, labels)
        return c0, d1, d2, d3, d4

    def decode_cl(self, c0, d1, d2, d3, d4):
        up1 = self.up1.forward_cl(d4, self.skip1.forward_cl(d3))
        up2 = self.up2.forward_cl(up1, self.skip2.forward_cl(d2))
        up3 = self.up3.forward_cl(up2, self.skip3.forward_cl(d1))
        ct = self.final[0]
        a, _ = VF.fused_layer(up3, ct.weight, ct.bias, None, None, self._final_up, NormCfg(None), "none", 0.0, self.training, None)
        a = self.final[1].forward_cl(a)
        a = self.final[2].forward_cl(a)
        a = self.final[3].forward_cl(a)
        return VB.tanh(a)

    # ---- reference API (NCHW fp32 in / out) ------------------------------------------------------------------------------
    def encode(self, x, style_code, labels):
        return tuple(VF.from_channels_last(t) for t in self.encode_cl(x, style_code, labels))

    def decode(self, c0, d1, d2, d3, d4, style_code):
        cl = [VF.to_channels_last(t) for t in (c0, d1, d2, d3, d4)]
        return VF.from_channels_last(self.decode_cl(*cl))

    def forward(self, x, style_code, labels):
        return VF.from_channels_last(self.decode_cl(*self.encode_cl(x, style_code, labels)))


class Discriminator(nn.Module):
    def __init__(self, image_size, num_of_classes, max_channels=256):
        super().__init__()
        in_dim = IMAGE_CHANNEL * 2
        out_dim = 64
        convs = [Conv2d(in_dim, out_dim, 5, 1)]
        n_level = int(np.log2(image_size)) - 2
        for _ in range(n_level):
            in_dim = out_dim
            out_dim = min(out_dim * 2, max_channels)
            convs.append(Conv2d(in_dim, out_dim, 3, stride=2, bn="instance"))
        self.convs = nn.Sequential(*convs)
        self.adv_convs = nn.Sequential(
            Conv2d(out_dim, out_dim, 3, stride=2, activate="lrelu"),
            Conv2d(out_dim, 1, 3, stride=2, activate=None)
        )
        self.aux_convs = nn.Sequential(
            Conv2d(out_dim, out_dim, 3, stride=2, activate="lrelu"),
            Conv2d(out_dim, num_of_classes, 3, stride=2, activate=None)
        )

    def forward(self, x, x_content, y):
        a = VB.cat_channels(VF.to_channels_last(x), VF.to_channels_last(x_content))
        for blk in self.convs:
            a = blk.forward_cl(a)
        n = a.shape[0]
        adv = a
        for blk in self.adv_convs:
            adv = blk.forward_cl(adv)
        aux = a
        for blk in self.aux_convs:
            aux = blk.forward_cl(aux)
        # the heads end on 1x1 maps at the reference's sizes; flatten in NCHW order otherwise
        adv_res = VB.sigmoid(VF.from_channels_last(adv).reshape(n, -1))
        aux_res = VB.softmax_rows(VF.from_channels_last(aux).reshape(n, -1).contiguous())
        return adv_res, aux_res
