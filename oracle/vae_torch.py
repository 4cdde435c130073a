"""torch-CPU port of the reference VAE step (TEST INFRASTRUCTURE / CPU BASELINE ONLY).

The reference's own CPU implementation of the hot path *is* ``torch.nn`` on the CPU backend
(oneDNN / MKL): models/networks.py builds nn.Conv2d / nn.ConvTranspose2d / nn.Linear /
nn.BatchNorm modules and train.py drives them with autograd and RMSprop.  /root/reference itself
cannot travel to the GPU box, so this file restates the same chain with ``torch.nn.functional``
calls on plain parameter tensors (state_dict naming), each line citing the reference line it
follows.  It is used (a) as the timed CPU baseline of ``bench.py`` (``cpu_baseline.kind = "port"``
and ``--impl reference``) and (b) as a mid-size fp64 checker.  It is pinned against the same
golden fixtures as the NumPy oracle (tests/test_oracle_golden.py::test_torch_port_*).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5
BN_MOMENTUM = 0.9  # models/networks.py:16,40,66,89


class VaeTorchPort:
    """Encoder -> reparameterize -> Decoder -> mse + sum(kl), parameters keyed like the reference state_dict."""

    def __init__(self, params: dict, dtype=torch.float32, lr=1e-4):
        self.P = {k: torch.as_tensor(v).to(dtype).clone() for k, v in params.items()}
        self.train_keys = [k for k in self.P if "running" not in k]
        for k in self.train_keys:
            self.P[k].requires_grad_(True)
        self.L = sum(1 for k in self.P if k.startswith("encoder.conv.") and k.endswith(".conv.weight"))
        # train.py:136-140: RMSprop(lr=1e-4) per sub-network; one optimiser over the same tensors is equivalent
        self.opt = torch.optim.RMSprop([self.P[k] for k in self.train_keys], lr=lr)

    def _bn(self, t, name):
        P = self.P
        return F.batch_norm(t, P[name + ".running_mean"], P[name + ".running_var"], P[name + ".weight"], P[name + ".bias"],
                            training=True, momentum=BN_MOMENTUM, eps=BN_EPS)

    def encoder(self, x):
        P, t = self.P, x
        for i in range(self.L):                                                        # networks.py:72-73, 27-30
            t = F.conv2d(t, P[f"encoder.conv.{i}.conv.weight"], None, stride=2, padding=2)   # :14
            t = F.relu(self._bn(t, f"encoder.conv.{i}.bn"))                                  # :16,28-29
        t = t.reshape(len(t), -1)                                                      # :74
        t = F.relu(self._bn(F.linear(t, P["encoder.fc.0.weight"]), "encoder.fc.1"))    # :65-67,75
        return (F.linear(t, P["encoder.l_mu.weight"], P["encoder.l_mu.bias"]),         # :69,76
                F.linear(t, P["encoder.l_var.weight"], P["encoder.l_var.bias"]))       # :70,77

    def decoder(self, z):
        P = self.P
        t = F.relu(self._bn(F.linear(z, P["decoder.fc.0.weight"]), "decoder.fc.1"))    # :88-90,109
        t = t.reshape(len(t), -1, 8, 8)                                                # :110
        for i in range(self.L):                                                        # :42-46
            t = F.conv_transpose2d(t, P[f"decoder.conv.{i}.conv.weight"], None, stride=2, padding=2, output_padding=1)  # :38
            t = F.relu(self._bn(t, f"decoder.conv.{i}.bn"))
        t = F.conv2d(t, P[f"decoder.conv.{self.L}.0.weight"], P[f"decoder.conv.{self.L}.0.bias"], stride=1, padding=2)   # :101
        return torch.sigmoid(t)                                                        # :102

    def forward_loss(self, x, eps=None):
        mu, logvar = self.encoder(x)
        std = logvar.mul(0.5).exp()                                                    # :229
        if eps is None:
            eps = torch.empty_like(std).normal_()                                      # :230
        z = eps * std + mu                                                             # :231
        xt = self.decoder(z)
        kl = -0.5 * torch.sum(-logvar.exp() - mu.pow(2) + logvar + 1, 1)               # :270
        loss = F.mse_loss(x, xt) + kl.sum()                                            # train.py:62-63 (VAE terms)
        return loss, dict(mu=mu, logvar=logvar, z=z, x_tilde=xt, kl=kl)

    def step(self, x, eps=None, optimize=True):
        """zero_grad -> forward -> loss -> backward (-> RMSprop step): train.py:68-78 restricted to the VAE terms."""
        for k in self.train_keys:
            self.P[k].grad = None
        loss, out = self.forward_loss(x, eps)
        loss.backward()
        if optimize:
            self.opt.step()
        return loss.detach(), out

    def grads(self):
        return {k: self.P[k].grad for k in self.train_keys}


def time_cpu_steps(img=64, cin=1, z=128, batch=16, steps=5, warmup=2, threads=None, seed=0):
    """Time the port on the host cores.  Returns (images_per_second, ms_per_step, threads)."""
    import os
    import time

    from . import vae_numpy as vn
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    P = vn.synth_vae_params(img, z, cin, cin, seed)
    port = VaeTorchPort(P, torch.float32)
    x_np, _ = vn.synth_batch(batch, img, cin, z, seed)
    x = torch.from_numpy(x_np)
    for _ in range(warmup):
        port.step(x)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        port.step(x)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    med = ts[len(ts) // 2]
    return batch / med, med * 1e3, threads
