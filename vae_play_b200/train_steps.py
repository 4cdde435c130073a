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
        VF.join_async()
        return None
    lam = float(lambda_mse)
    total = VB.weighted_sums([parts["recon"], parts["l1"], parts["kl"], parts["mse"], parts["bce_o"], parts["bce_p"], parts["bce_s"]],
                             [1.0, 1.0, 1.0, 1.0 + lam, lam, lam, lam])
    total.backward()
    VF.join_async()          # weight gradients on the side stream (functional.set_async_wgrad), if any
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


# ------------------------------------------------------------------------------------------------------------------------
# config 5: train_Style_GAN.py::train_random_gan (:162-281)
# ------------------------------------------------------------------------------------------------------------------------
def style_reparameterization(mu, logvar, eps):
    """train_Style_GAN.py:156-160: std = exp(logvar / 2); z = eps * std + mu with HOST-drawn eps (np.random.normal) -- the caller
    supplies it (``eps`` [B, z_dim] on the device)."""
    z, _ = VF.reparam_kl(mu.contiguous(), logvar.contiguous(), eps=eps)
    return z


def style_gan_step(G, E, D, g_opt, e_opt, d_opt, x_target, x_content, y_org, eps, sample_z):
    """One train_random_gan iteration, statement for statement (including its order: E steps before the latent loss is formed, G
    after it; the discriminator heads return probabilities that F.cross_entropy soft-maxes again).  Returns the seven logged losses."""
    b = x_target.size(0)
    e_opt.zero_grad()
    g_opt.zero_grad()
    mu, logvar = E(x_target)
    encode_z = style_reparameterization(mu, logvar, eps)
    x_rec = G(x_content, encode_z, y_org)
    d_rec_valid, d_rec_type = D(x_rec, x_content, y_org)
    g_rec_kl_loss = VB.weighted_sums([VB.kl_per_sample(mu, logvar)], [1.0])                     # 0.5 * sum(exp(lv) + mu^2 - lv - 1)
    g_rec_d_loss = VB.weighted_sums([VB.binary_cross_entropy_const(d_rec_valid, True), VB.cross_entropy(d_rec_type, y_org)], [1.0, 1.0])
    g_rec_pixel_loss = VF.l1_loss(x_target, x_rec)
    x_gen = G(x_content, sample_z, y_org)
    d_gen_valid, d_gen_type = D(x_gen, x_content, y_org)
    g_gen_d_loss = VB.weighted_sums([VB.binary_cross_entropy_const(d_gen_valid, True), VB.cross_entropy(d_gen_type, y_org)], [1.0, 1.0])
    g_loss = VB.weighted_sums([g_rec_pixel_loss, g_rec_d_loss, g_rec_kl_loss, g_gen_d_loss], [1.0, 1.0, 1.0, 1.0])
    g_loss.backward(retain_graph=True)
    e_opt.step()
    _mu, _ = E(x_gen)
    loss_latent = VB.weighted_sums([VF.l1_loss(sample_z, _mu)], [0.5])
    loss_latent.backward()
    g_opt.step()
    d_opt.zero_grad()
    d_real_valid, d_real_type = D(x_target, x_content, y_org)
    d_fake_valid, d_fake_type = D(x_rec.detach(), x_content, y_org)
    d_real_loss = VB.weighted_sums([VB.binary_cross_entropy_const(d_real_valid, True), VB.cross_entropy(d_real_type, y_org)], [1.0, 1.0])
    d_fake_loss = VB.weighted_sums([VB.binary_cross_entropy_const(d_fake_valid, False), VB.cross_entropy(d_fake_type, y_org)], [1.0, 1.0])
    d_adv_loss = VB.weighted_sums([d_real_loss, d_fake_loss], [0.5, 0.5])
    d_adv_loss.backward()
    d_opt.step()
    return {"g_rec_kl_loss": g_rec_kl_loss, "g_rec_d_loss": g_rec_d_loss, "g_rec_pixel_loss": g_rec_pixel_loss, "g_gen_d_loss": g_gen_d_loss,
            "loss_latent": loss_latent, "d_real_loss": d_real_loss, "d_fake_loss": d_fake_loss}
