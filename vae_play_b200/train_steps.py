"""The reference's training steps on the B200 kernel library.

``vaegan_step`` is train.py:43-78 of the reference (config 4 of BASELINE.json): VaeGan forward (encoder, decoder on z and on
z_p, discriminator on the 3B images in REC and in GAN mode), ``VaeGan.loss``, the five losses and their gradients, then one
RMSprop step per sub-network.  Two gradient modes with identical mathematics:

* ``fused=False``: the reference's five ``backward(retain_graph=True)`` calls, accumulated by autograd (:68-73);
* ``fused=True`` (default): ONE backward of the summed loss.  After ``zero_grad`` the five calls accumulate
  d(loss_recon + loss_encoder + loss_decoder + loss_discriminator + loss_aux)/d(theta), so a single backward of
      recon + aux + sum(kl) + (1 + lambda) sum(mse) + lambda (sum(bce_o) + sum(bce_p) + sum(bce_s))
  gives the same gradients at a fifth of the backward work -- and without the reference's cancellation of
  -(1 - lambda) and +1 times the discriminator gradient (lambda = 1e-6, train.py:21,66).
"""
from __future__ import annotations

import torch

from . import functional as VF
from . import functional_blocks as VB
from .models.networks import VaeGan

LAMBDA_MSE = 1e-6      # train.py:21


def vaegan_losses(net: VaeGan, x, targets, eps=None, z_p=None, lambda_mse=LAMBDA_MSE):
    """Forward + the five losses of train.py:43-67.  Returns (losses dict, parts dict)."""
    b = x.size(0)
    x_tilde, disc_class, disc_layer, mus, log_variances, params = net(x, eps=eps, z_p=z_p)
    dl_o, dl_p, dl_s = disc_layer[:b], disc_layer[b:-b], disc_layer[-b:]
    dc_o, dc_p, dc_s = disc_class[:b], disc_class[b:-b], disc_class[-b:]
    nle, kl, mse, bce_o, bce_p, bce_s, l1 = VaeGan.loss(x, x_tilde, dl_o, dl_p, dl_s, dc_o, dc_p, dc_s, mus, log_variances, targets, params)
    recon = VF.mse_loss(x, x_tilde)                                                     # F.mse_loss(imgs, x_tilde), :62
    lam = float(lambda_mse)
    parts = dict(recon=recon, kl=kl, mse=mse, bce_o=bce_o, bce_p=bce_p, bce_s=bce_s, l1=l1, x_tilde=x_tilde, nle=nle)
    losses = {
        "loss_recon": recon,
        "loss_encoder": VB.weighted_sums([kl, mse], [1.0, 1.0]),                        # :63
        "loss_discriminator": VB.weighted_sums([bce_o, bce_p, bce_s], [1.0, 1.0, 1.0]),  # :64
        "loss_decoder": VB.weighted_sums([mse, bce_o, bce_p, bce_s], [lam, -(1.0 - lam), -(1.0 - lam), -(1.0 - lam)]),   # :65
        "loss_aux": l1,                                                                # :66
    }
    return losses, parts


def vaegan_backward(losses, parts, fused=True, lambda_mse=LAMBDA_MSE):
    """Gradients of the step into ``param.grad`` (the caller has cleared them)."""
    if not fused:
        losses["loss_recon"].backward(retain_graph=True)
        losses["loss_encoder"].backward(retain_graph=True)
        losses["loss_decoder"].backward(retain_graph=True)
        losses["loss_discriminator"].backward(retain_graph=True)
        losses["loss_aux"].backward()
        return None
    lam = float(lambda_mse)
    total = VB.weighted_sums([parts["recon"], parts["l1"], parts["kl"], parts["mse"], parts["bce_o"], parts["bce_p"], parts["bce_s"]],
                             [1.0, 1.0, 1.0, 1.0 + lam, lam, lam, lam])
    total.backward()
    return total


def vaegan_step(net: VaeGan, optimizers, x, targets, fused=True, eps=None, z_p=None):
    """One full train.py step.  ``optimizers``: the four optimisers of train.py:136-140 (any iterable)."""
    for o in optimizers:
        o.zero_grad(set_to_none=True)
    losses, parts = vaegan_losses(net, x, targets, eps=eps, z_p=z_p)
    vaegan_backward(losses, parts, fused=fused)
    for o in optimizers:
        o.step()
    return losses
