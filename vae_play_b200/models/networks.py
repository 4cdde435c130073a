"""Drop-in mirror of the reference's ``models/networks.py`` (VAE-GAN) on the B200 kernel library.

Same class names, constructor signatures, attribute tree and ``state_dict`` keys as the reference
(/root/reference/models/networks.py:10-281), so ``train.py``-style code, ``init_parameters``-style
module walks, optimisers and checkpoints work unchanged -- but every forward/backward arithmetic step
runs in libvaeplay_b200 (hand-written sm_100a kernels) instead of ATen/cuDNN/cuBLAS.

Parameters live in real ``nn.Conv2d`` / ``nn.ConvTranspose2d`` / ``nn.Linear`` / ``nn.BatchNorm*``
shells (never called); activations travel between layers channels-last in the activation dtype.
"""
from __future__ import annotations

import math

import numpy
import torch
import torch.nn as nn

from .. import functional as VF
from ..functional import DualLinear, NormCfg, TapLayer


def _bn_cfg(bn: nn.modules.batchnorm._BatchNorm):
    return NormCfg("batch", eps=bn.eps, momentum=bn.momentum)


# encoder block (used in encoder and discriminator) -- reference networks.py:10-30
class EncoderBlock(nn.Module):
    def __init__(self, channel_in, channel_out):
        super().__init__()
        self.conv = nn.Conv2d(in_channels=channel_in, out_channels=channel_out, kernel_size=5, padding=2, stride=2, bias=False)
        self.bn = nn.BatchNorm2d(num_features=channel_out, momentum=0.9)
        self._layer = TapLayer("conv", channel_in, channel_out, k=5, stride=2, pad=2)
        VF.weights_channels_last(self)

    def forward_cl(self, a, out=False):
        act, pre = VF.fused_layer(a, self.conv.weight, None, self.bn.weight, self.bn.bias, self._layer, _bn_cfg(self.bn),
                                  "relu", 0.0, self.training, self.bn)
        return (act, pre) if out else act

    def forward(self, ten, out=False, t=False):
        a = VF.to_channels_last(ten)
        if out:
            act, pre = self.forward_cl(a, True)
            return VF.from_channels_last(act), VF.from_channels_last(pre)
        return VF.from_channels_last(self.forward_cl(a))


# decoder block -- reference networks.py:34-46
class DecoderBlock(nn.Module):
    def __init__(self, channel_in, channel_out):
        super().__init__()
        self.conv = nn.ConvTranspose2d(channel_in, channel_out, kernel_size=5, padding=2, stride=2, output_padding=1, bias=False)
        self.bn = nn.BatchNorm2d(channel_out, momentum=0.9)
        self._layer = TapLayer("convT", channel_in, channel_out, k=5, stride=2, pad=2, out_pad=1)
        VF.weights_channels_last(self)

    def forward_cl(self, a):
        act, _ = VF.fused_layer(a, self.conv.weight, None, self.bn.weight, self.bn.bias, self._layer, _bn_cfg(self.bn),
                                "relu", 0.0, self.training, self.bn)
        return act

    def forward(self, ten):
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(ten)))


# reference networks.py:49-81
class Encoder(nn.Module):
    def __init__(self, channel_in=3, z_size=128, iter_level=3):
        super().__init__()
        self.size = channel_in
        layers_list = []
        for i in range(iter_level):
            if i == 0:
                layers_list.append(EncoderBlock(channel_in=self.size, channel_out=64))
                self.size = 64
            else:
                layers_list.append(EncoderBlock(channel_in=self.size, channel_out=self.size * 2))
                self.size *= 2
        self.conv = nn.Sequential(*layers_list)
        self.fc = nn.Sequential(nn.Linear(in_features=8 * 8 * self.size, out_features=1024, bias=False),
                                nn.BatchNorm1d(num_features=1024, momentum=0.9),
                                nn.ReLU(True))
        self.l_mu = nn.Linear(in_features=1024, out_features=z_size)
        self.l_var = nn.Linear(in_features=1024, out_features=z_size)
        self._fc_layer = TapLayer("linear", 8 * 8 * self.size, 1024)
        self._heads = DualLinear(1024, z_size)
        self.z_size = z_size

    def forward_packed(self, ten, taps=None):
        """NCHW fp32 image -> fused head output [B, 2Z] = (mu | logvar), fp32.  ``taps`` (a list) receives the output of
        the conv stack twice (see below): the point where a two-stage backward can be cut (bench.py overlaps the gradient all-reduce of
        everything downstream of it with the backward of the conv stack)."""
        a = VF.to_channels_last(ten)
        for blk in self.conv:
            a = blk.forward_cl(a)
        if taps is not None:
            # two identity nodes: a two-stage backward names taps[0] in ``inputs`` of stage 1 (autograd then also retains its
            # .grad, and would clone / add into it again on every later pass through its node) and starts stage 2 from taps[1]
            inner = VF.grad_cut(a)
            a = VF.grad_cut(inner)
            taps.append(a)
            taps.append(inner)
        if a.shape[1] != 8 or a.shape[2] != 8:
            raise ValueError(f"Encoder expects an 8x8 map before fc, got {tuple(a.shape)} (img_size must be 8 * 2**iter_level)")
        h, _ = VF.fused_layer(VF.hwc_to_chw_flat(a), self.fc[0].weight, None, self.fc[1].weight, self.fc[1].bias, self._fc_layer,
                              _bn_cfg(self.fc[1]), "relu", 0.0, self.training, self.fc[1])
        return VF.dual_linear(h, self.l_mu.weight, self.l_mu.bias, self.l_var.weight, self.l_var.bias, self._heads)

    def forward(self, ten):
        mulv = self.forward_packed(ten)
        return mulv[:, : self.z_size], mulv[:, self.z_size:]


# reference networks.py:84-115
class Decoder(nn.Module):
    def __init__(self, z_size, size, channel_out=3, iter_level=3):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(in_features=z_size, out_features=8 * 8 * size, bias=False),
                                nn.BatchNorm1d(num_features=8 * 8 * size, momentum=0.9),
                                nn.ReLU(True))
        self.size = size
        self._fc_layer = TapLayer("linear", z_size, 8 * 8 * size)
        self._fc_size = size
        layers_list = [DecoderBlock(channel_in=self.size, channel_out=self.size)]
        for _ in range(iter_level - 1):
            layers_list.append(DecoderBlock(channel_in=self.size, channel_out=self.size // 2))
            self.size = self.size // 2
        layers_list.append(nn.Sequential(
            nn.Conv2d(in_channels=self.size, out_channels=channel_out, kernel_size=5, stride=1, padding=2),
            nn.Sigmoid()
        ))
        self.conv = nn.Sequential(*layers_list)
        self._out_layer = TapLayer("conv", self.size, channel_out, k=5, stride=1, pad=2)
        self._nonorm = NormCfg(None)

    def forward_cl(self, z_cl):
        """z as [B,1,1,Z] in the activation dtype -> x_tilde channels-last."""
        a, _ = VF.fused_layer(z_cl, self.fc[0].weight, None, self.fc[1].weight, self.fc[1].bias, self._fc_layer,
                              _bn_cfg(self.fc[1]), "relu", 0.0, self.training, self.fc[1])
        a = VF.chw_flat_to_hwc(a, self._fc_size, 8, 8)
        blocks = list(self.conv)
        for blk in blocks[:-1]:
            a = blk.forward_cl(a)
        last = blocks[-1][0]
        xt, _ = VF.fused_layer(a, last.weight, last.bias, None, None, self._out_layer, self._nonorm, "sigmoid", 0.0,
                               self.training, None)
        return xt

    def forward(self, ten):
        z = ten.reshape(len(ten), -1, 1, 1)
        return VF.from_channels_last(self.forward_cl(VF.to_channels_last(z)))


# reference networks.py:118-148: eight bias Linears without activation
class DirectDecoder(nn.Module):
    def __init__(self, z_size, num_of_param=3):
        super().__init__()
        self.head = nn.Sequential(
            nn.Linear(in_features=z_size, out_features=512),
            nn.Linear(in_features=512, out_features=256),
            nn.Linear(in_features=256, out_features=128),
            nn.Linear(in_features=128, out_features=64),
        )
        self.r_fc = nn.Sequential(nn.Linear(in_features=64, out_features=32), nn.Linear(in_features=32, out_features=1))
        self.xy_fc = nn.Sequential(nn.Linear(in_features=64, out_features=32), nn.Linear(in_features=32, out_features=2))
        self._layers = {}
        self._nonorm = NormCfg(None)

    def _lin(self, a, m: nn.Linear, out_dtype=None):
        key = id(m)
        if key not in self._layers:
            self._layers[key] = TapLayer("linear", m.in_features, m.out_features)
        y, _ = VF.fused_layer(a, m.weight, m.bias, None, None, self._layers[key], self._nonorm, "none", 0.0, self.training,
                              None, out_dtype)
        return y

    def forward(self, ten):
        a = VF.to_channels_last(ten.reshape(len(ten), -1, 1, 1))
        for m in self.head:
            a = self._lin(a, m)
        r = self._lin(self._lin(a, self.r_fc[0]), self.r_fc[1], torch.float32)
        xy = self._lin(self._lin(a, self.xy_fc[0]), self.xy_fc[1], torch.float32)
        return torch.cat([r.reshape(len(ten), -1), xy.reshape(len(ten), -1)], dim=-1)


# reference networks.py:151-198
class Discriminator(nn.Module):
    def __init__(self, channel_in=3, recon_level=3, iter_level=3):
        super().__init__()
        self.size = channel_in
        self.recon_levl = recon_level
        self.conv = nn.ModuleList()
        self.conv.append(nn.Sequential(
            nn.Conv2d(in_channels=self.size, out_channels=32, kernel_size=5, stride=1, padding=2),
            nn.ReLU(inplace=True)))
        self._first = TapLayer("conv", channel_in, 32, k=5, stride=1, pad=2)
        self.size = 32
        channel_out = self.size * 2
        for _ in range(iter_level):
            self.conv.append(EncoderBlock(channel_in=self.size, channel_out=channel_out))
            self.size = channel_out
            channel_out *= 2
        self.fc = nn.Sequential(
            nn.Linear(in_features=8 * 8 * self.size, out_features=512, bias=False),
            nn.BatchNorm1d(num_features=512, momentum=0.9),
            nn.ReLU(inplace=True),
            nn.Linear(in_features=512, out_features=1),
        )
        self._fc_layer = TapLayer("linear", 8 * 8 * self.size, 512)
        self._out_layer = TapLayer("linear", 512, 1)
        self._nonorm = NormCfg(None)

    def forward(self, ten_orig, ten_predicted, ten_sampled, mode='REC'):
        ten = torch.cat((ten_orig, ten_predicted, ten_sampled), 0)
        a = VF.to_channels_last(ten)
        for i, lay in enumerate(self.conv):
            if i == 0:
                c0 = lay[0]
                a, _ = VF.fused_layer(a, c0.weight, c0.bias, None, None, self._first, self._nonorm, "relu", 0.0,
                                      self.training, None)
            elif mode == "REC" and i == self.recon_levl:
                _, pre = lay.forward_cl(a, True)
                # pre-BatchNorm conv output, flattened in NCHW order like the reference (:183)
                return VF.from_channels_last(pre).reshape(len(ten), -1)
            else:
                a = lay.forward_cl(a)
        h, _ = VF.fused_layer(VF.hwc_to_chw_flat(a), self.fc[0].weight, None, self.fc[1].weight, self.fc[1].bias, self._fc_layer,
                              _bn_cfg(self.fc[1]), "relu", 0.0, self.training, self.fc[1])
        o, _ = VF.fused_layer(h, self.fc[3].weight, self.fc[3].bias, None, None, self._out_layer, self._nonorm, "sigmoid",
                              0.0, self.training, None, torch.float32)
        return o.reshape(len(ten), 1)


# reference networks.py:201-281
class VaeGan(nn.Module):
    def __init__(self, img_size, z_size=128, num_of_param=3):
        super().__init__()
        self.iter_level = int(math.log2(img_size // 8))
        self.z_size = z_size
        self.encoder = Encoder(channel_in=1, z_size=self.z_size, iter_level=self.iter_level)
        self.decoder = Decoder(z_size=self.z_size, size=self.encoder.size, channel_out=1, iter_level=self.iter_level)
        self.discriminator = Discriminator(channel_in=1, recon_level=self.iter_level, iter_level=self.iter_level)
        self.param_encoder = DirectDecoder(z_size, num_of_param=num_of_param)
        self.init_parameters()

    def init_parameters(self):
        # U(-s, s), s = 1/sqrt(fan)/sqrt(3), biases 0 -- reference networks.py:214-226
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d, nn.Linear)):
                if hasattr(m, "weight") and m.weight is not None and m.weight.requires_grad:
                    scale = 1.0 / numpy.sqrt(numpy.prod(m.weight.shape[1:]))
                    scale /= numpy.sqrt(3)
                    nn.init.uniform_(m.weight, -scale, scale)
                if hasattr(m, "bias") and m.bias is not None and m.bias.requires_grad:
                    nn.init.constant_(m.bias, 0.0)

    def reparameterize(self, mu, logvar, eps=None):
        """z = eps*exp(0.5*logvar) + mu with eps from the device generator's Philox stream (:228-231)."""
        z, _ = VF.reparam_kl(mu, logvar, eps=eps)
        return z

    def forward(self, x, gen_size=10, eps=None, z_p=None, rng=None):
        """Reference VaeGan.forward (networks.py:233-258): train -> (x_tilde, disc_class [3B,1], disc_layer [3B,F], mus,
        log_variances, params [B,3]); eval -> (x_tilde, params) or, with x None, decoded samples.  ``eps`` / ``z_p`` (tests)
        replace the two random draws; by default both come from the device generator's Philox stream exactly as the
        reference's ``normal_()`` / ``torch.randn(...).cuda()`` would consume it.  ``rng=(seed, offset_dev)`` pins both draws to
        a device-resident Philox offset instead (CUDA-graph replay: the caller advances it by ``2 * philox_policy(B*z)[1]`` per
        step, see bench.py)."""
        if self.training:
            mus, log_variances = self.encoder(x)
            if rng is not None and eps is None:
                inc = VF.philox_policy(len(x) * self.z_size, VF._num_sms(x.device))[1]
                z, _ = VF.reparam_kl(mus, log_variances, rng=(rng[0], 0, rng[1]))
                if z_p is None:
                    z_p = VF.philox_normal((len(x), self.z_size), x.device, rng=(rng[0], inc, rng[1]))
            else:
                z = self.reparameterize(mus, log_variances, eps)
            x_tilde = self.decoder(z)
            params = self.param_encoder(z)
            if z_p is None:
                z_p = VF.philox_normal((len(x), self.z_size), x.device)            # torch.randn(...).cuda(), :241
            z_p = z_p.detach().requires_grad_(True)
            x_p = self.decoder(z_p)
            disc_layer = self.discriminator(x, x_tilde, x_p, "REC")
            disc_class = self.discriminator(x, x_tilde, x_p, "GAN")
            return x_tilde, disc_class, disc_layer, mus, log_variances, params
        else:
            if x is None:
                dev = next(self.parameters()).device
                z_p = VF.philox_normal((gen_size, self.z_size), dev)
                return self.decoder(z_p)
            mus, log_variances = self.encoder(x)
            z = self.reparameterize(mus, log_variances, eps)
            x_tilde = self.decoder(z)
            params = self.param_encoder(z)
            return x_tilde, params

    # ---- fused hot path (the north-star step): encoder -> reparam+KL -> decoder ---------------------
    def vae_forward(self, x, eps=None, rng=None, taps=None):
        """Returns (x_tilde NCHW fp32, mu|logvar packed [B,2Z] fp32, kl [B]) without leaving channels-last."""
        mulv = self.encoder.forward_packed(x, taps)
        z, kl = VF.reparam_kl(mulv, None, eps=eps, z_dtype=VF.act_dtype(), rng=rng)
        if taps is not None:
            # a second place to cut the backward (taps[2] outer, taps[3] inner): between the decoder and the sample, so that a
            # data-parallel step can start exchanging the decoder's gradients before the heads / encoder.fc backward runs
            inner = VF.grad_cut(z)
            z = VF.grad_cut(inner)
            taps.append(z)
            taps.append(inner)
        xt = self.decoder.forward_cl(z.reshape(len(z), 1, 1, -1))
        return VF.from_channels_last(xt), mulv, kl

    @staticmethod
    def loss(x, x_tilde, disc_layer_original, disc_layer_predicted, disc_layer_sampled, disc_class_original,
             disc_class_predicted, disc_class_sampled, mus, variances, targets, params):
        """Same 7-tuple as the reference's ``VaeGan.loss`` (networks.py:264-281):
        (nle, kl, feature-mse, -log(D(x)+1e-3), -log(1-D(x~)+1e-3), -log(1-D(x_p)+1e-3), smooth-L1/B), every term one
        kernel of libvaeplay_b200 (csrc/gan_losses.cu) with its own backward -- no ATen arithmetic on the step."""
        from .. import functional_blocks as VB
        b = x.size(0)
        flat = lambda t: t.reshape(t.size(0), -1)
        nle = VB.half_sqdiff(flat(x), flat(x_tilde))
        kl = VB.kl_per_sample(mus, variances)
        feat = VB.feature_mse(disc_layer_original, disc_layer_predicted)
        d_real = VB.neglog(disc_class_original, 1.0, 1e-3)
        d_rec = VB.neglog(disc_class_predicted, -1.0, 1.0 + 1e-3)
        d_samp = VB.neglog(disc_class_sampled, -1.0, 1.0 + 1e-3)
        aux = VB.smooth_l1_sum(targets, params, 1.0 / b)
        return nle, kl, feat, d_real, d_rec, d_samp, aux
