"""Generate tests/golden/style_step32_b2.npz by running the UNMODIFIED reference Style_GAN modules through one
``train_random_gan`` iteration (TEST INFRASTRUCTURE ONLY).

    python oracle/gen_golden_style_step.py        # build container only (needs /root/reference)

Config 5 of BASELINE.json, reference train_Style_GAN.py:162-281 with models/network_Style_GAN.py, at image_size 32 / z_dim 16
/ batch 2 / 2 classes (the 256x256 production size has a 369 M-parameter MLP; the graph is the same).  The step body below is
the reference's, statement for statement (it lives in a script with CLI / dataset code around it and cannot be imported as a
function); the two host draws (np.random.normal) are replaced by preset tensors.  float64; Adam(lr=1e-4) as in :337-339.
Stored: the seven logged losses, and for E / G / D the parameters AFTER the step (digests) -- the step's observable result.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
IMG, Z, B, NCLS = 32, 16, 2, 2


def synth_step_inputs():
    rs = np.random.RandomState(5150)
    x_target = rs.uniform(-1, 1, size=(B, 3, IMG, IMG)).astype(np.float32)
    x_content = rs.uniform(-1, 1, size=(B, 3, IMG, IMG)).astype(np.float32)
    y = np.array([1, 0], dtype=np.int64)
    eps = rs.standard_normal(size=(B, Z)).astype(np.float32)
    sample_z = rs.standard_normal(size=(B, Z)).astype(np.float32)
    return x_target, x_content, y, eps, sample_z


def digest(a, nsamp=64):
    a = np.asarray(a, np.float64).ravel()
    idx = np.linspace(0, a.size - 1, num=min(nsamp, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()], a[idx]])


def run(dtype):
    import torch
    import torch.nn.functional as F
    from oracle.gen_golden import import_reference
    from oracle.gen_golden_blocks import synth_state
    import_reference()
    import models.network_Style_GAN as style
    G, E, D = style.Generator(IMG, Z), style.StyleEncoder(Z, IMG), style.Discriminator(IMG, NCLS)
    for i, m in enumerate((G, E, D)):
        m.load_state_dict(synth_state(m, 900 + i), strict=False)
        m.to(dtype).train()
    before = {n: {k: p.detach().clone() for k, p in m.named_parameters()} for n, m in (("G", G), ("E", E), ("D", D))}
    g_opt, e_opt, d_opt = (torch.optim.Adam(m.parameters(), lr=1e-4) for m in (G, E, D))
    xt, xc, y, eps, sz = synth_step_inputs()
    x_target, x_content = torch.from_numpy(xt).to(dtype), torch.from_numpy(xc).to(dtype)
    y_org, eps, samlpe_z = torch.from_numpy(y), torch.from_numpy(eps).to(dtype), torch.from_numpy(sz).to(dtype)
    b = B
    # ---- train_Style_GAN.py:213-262 --------------------------------------------------------------------------------------
    e_opt.zero_grad()
    g_opt.zero_grad()
    mu, logvar = E(x_target)
    std = torch.exp(logvar / 2)
    encode_z = eps * std + mu
    x_rec = G(x_content, encode_z, y_org)
    d_rec_valid, d_rec_type = D(x_rec, x_content, y_org)
    g_rec_kl_loss = 0.5 * torch.sum(torch.exp(logvar) + mu ** 2 - logvar - 1)
    g_rec_d_loss = F.binary_cross_entropy(d_rec_valid, torch.ones((b, 1), dtype=dtype)) + F.cross_entropy(d_rec_type, y_org)
    g_rec_pixel_loss = F.l1_loss(x_rec, x_target)
    g_rec_loss = g_rec_pixel_loss + g_rec_d_loss + g_rec_kl_loss
    x_gen = G(x_content, samlpe_z, y_org)
    d_gen_valid, d_gen_type = D(x_gen, x_content, y_org)
    g_gen_d_loss = F.binary_cross_entropy(d_gen_valid, torch.ones((b, 1), dtype=dtype)) + F.cross_entropy(d_gen_type, y_org)
    g_loss = g_rec_loss + g_gen_d_loss
    g_loss.backward(retain_graph=True)
    e_opt.step()
    _mu, _ = E(x_gen)
    loss_latent = F.l1_loss(_mu, samlpe_z) * 0.5
    loss_latent.backward()
    g_opt.step()
    d_opt.zero_grad()
    d_real_valid, d_real_type = D(x_target, x_content, y_org)
    d_fake_valid, d_fake_type = D(x_rec.detach(), x_content, y_org)
    d_real_loss = F.binary_cross_entropy(d_real_valid, torch.ones((b, 1), dtype=dtype)) + F.cross_entropy(d_real_type, y_org)
    d_fake_loss = F.binary_cross_entropy(d_fake_valid, torch.zeros((b, 1), dtype=dtype)) + F.cross_entropy(d_fake_type, y_org)
    d_adv_loss = (d_real_loss + d_fake_loss) * 0.5
    d_adv_loss.backward()
    d_opt.step()
    # -----------------------------------------------------------------------------------------------------------------------
    out = {"losses": np.array([float(v) for v in (g_rec_kl_loss, g_rec_d_loss, g_rec_pixel_loss, g_gen_d_loss, loss_latent, d_real_loss, d_fake_loss)]),
           "x_rec": digest(x_rec.detach().double().numpy())}
    for n, m in (("G", G), ("E", E), ("D", D)):
        for k, p in m.named_parameters():
            # Adam's first step moves every element by ~lr * sign(grad): store the UPDATE (after - before) / lr, i.e. g / (|g| + eps)
            out[f"upd/{n}/{k}"] = digest(((p.detach() - before[n][k]) / 1e-4).double().numpy())
            out[f"grad/{n}/{k}"] = digest(p.grad.double().numpy()) if p.grad is not None else np.zeros(3)
    return out


def main():
    import torch
    o64, o32 = run(torch.float64), run(torch.float32)
    res = dict(o64)
    keys = [k for k in o64 if k.startswith("grad/")]
    res["ref_fp32_dev_keys"] = np.array(keys)
    res["ref_fp32_dev_vals"] = np.array([float(abs(o32[k][1] - o64[k][1]) / (o64[k][1] + 1e-300)) for k in keys])
    res["ref_fp32_loss_dev"] = np.abs(o32["losses"] - o64["losses"]) / np.abs(o64["losses"])
    path = os.path.join(ROOT, "tests", "golden", "style_step32_b2.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "bytes; losses", o64["losses"], "fp32 loss dev", res["ref_fp32_loss_dev"].max())


if __name__ == "__main__":
    main()
