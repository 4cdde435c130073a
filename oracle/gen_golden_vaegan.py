"""Generate tests/golden/vaegan64_b4.npz by running the UNMODIFIED reference VaeGan + the train.py step (TEST INFRASTRUCTURE ONLY).

    python oracle/gen_golden_vaegan.py        # build container only (needs /root/reference)

Config 4 of BASELINE.json, reference models/networks.py:201-281 (VaeGan.forward / VaeGan.loss) and train.py:43-73 (the five
losses and the five accumulating ``backward(retain_graph=True)`` calls).  The reference's modules are executed as they are;
only the two random draws of ``VaeGan.forward`` are replaced by preset tensors (``Tensor.normal_`` -> eps,
``torch.randn`` -> z_p) and ``Tensor.cuda`` is the identity (CPU run), SURVEY.md section 8c.  float64 is the truth; a float32
run of the same thing calibrates the fp32 tolerance (``ref_fp32_dev``).

Stored: forward outputs (x_tilde, disc_class, mus, log_variances, params; a digest of disc_layer), the seven outputs of
``VaeGan.loss``, the five step losses, and every parameter gradient after the five backward calls (full for small tensors,
(sum, l2, max, strided samples) digests for the big ones).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAMBDA_MSE = 1e-6          # train.py:21
FULL_LIMIT = 4096          # gradients up to this many elements are stored in full


def synth_direct_decoder_params(seed=0, z=128):
    """DirectDecoder (networks.py:118-148): eight bias Linears; uniform +-1/sqrt(fan_in) weights, small biases."""
    rs = np.random.RandomState(3000 + seed)
    P = {}
    dims = [("head.0", z, 512), ("head.1", 512, 256), ("head.2", 256, 128), ("head.3", 128, 64),
            ("r_fc.0", 64, 32), ("r_fc.1", 32, 1), ("xy_fc.0", 64, 32), ("xy_fc.1", 32, 2)]
    for name, fi, fo in dims:
        P[f"{name}.weight"] = (rs.uniform(-1, 1, size=(fo, fi)) / np.sqrt(fi)).astype(np.float32)
        P[f"{name}.bias"] = (rs.uniform(-0.1, 0.1, size=fo)).astype(np.float32)
    return P


def synth_vaegan_params(seed=0, img=64, z=128):
    """state_dict of the whole VaeGan(img, z) with the reference's keys."""
    from oracle import vae_numpy as vn
    from oracle.gen_golden_disc import synth_disc_params
    import math
    L = int(math.log2(img // 8))
    P = dict(vn.synth_vae_params(img, z, 1, 1, seed))
    for k, v in synth_disc_params(seed, cin=1, iter_level=L).items():
        P["discriminator." + k] = v
    for k, v in synth_direct_decoder_params(seed, z).items():
        P["param_encoder." + k] = v
    return P


def synth_vaegan_inputs(seed=0, b=4, img=64, z=128):
    rs = np.random.RandomState(4000 + seed)
    x = rs.uniform(0, 1, size=(b, 1, img, img)).astype(np.float32)
    eps = rs.standard_normal(size=(b, z)).astype(np.float32)
    z_p = rs.standard_normal(size=(b, z)).astype(np.float32)
    targets = rs.uniform(0, 1, size=(b, 3)).astype(np.float32)
    return x, eps, z_p, targets


def digest(a, nsamp=256):
    a = np.asarray(a, np.float64).ravel()
    idx = np.linspace(0, a.size - 1, num=min(nsamp, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()], a[idx]])


def run_reference(networks, dtype, seed=0, b=4, img=64, z=128):
    import torch
    import torch.nn.functional as F
    P = synth_vaegan_params(seed, img, z)
    x_np, eps_np, zp_np, t_np = synth_vaegan_inputs(seed, b, img, z)
    net = networks.VaeGan(img, z)
    missing = net.load_state_dict({k: torch.from_numpy(v) for k, v in P.items()}, strict=False)
    assert not missing.unexpected_keys, missing
    assert all("running" in k or "num_batches" in k for k in missing.missing_keys), missing
    net = net.to(dtype).train()
    x = torch.from_numpy(x_np).to(dtype)
    targets = torch.from_numpy(t_np).to(dtype)
    eps, z_p = torch.from_numpy(eps_np).to(dtype), torch.from_numpy(zp_np).to(dtype)
    # the two draws of VaeGan.forward (networks.py:230,241) replaced by the preset tensors; .cuda() is the identity on CPU
    orig_normal, orig_randn, orig_cuda = torch.Tensor.normal_, torch.randn, torch.Tensor.cuda
    torch.Tensor.normal_ = lambda self, *a, **k: self.copy_(eps)
    torch.randn = lambda *a, **k: z_p.clone()
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        x_tilde, disc_class, disc_layer, mus, log_variances, params = net(x)
    finally:
        torch.Tensor.normal_, torch.randn, torch.Tensor.cuda = orig_normal, orig_randn, orig_cuda
    bs = b
    dl_o, dl_p, dl_s = disc_layer[:bs], disc_layer[bs:-bs], disc_layer[-bs:]
    dc_o, dc_p, dc_s = disc_class[:bs], disc_class[bs:-bs], disc_class[-bs:]
    nle, kl, mse, bce_o, bce_p, bce_s, l1 = networks.VaeGan.loss(x, x_tilde, dl_o, dl_p, dl_s, dc_o, dc_p, dc_s, mus, log_variances,
                                                                  targets, params)
    # train.py:62-73
    loss_recon = F.mse_loss(x, x_tilde)
    loss_encoder = torch.sum(kl) + torch.sum(mse)
    loss_discriminator = torch.sum(bce_o) + torch.sum(bce_p) + torch.sum(bce_s)
    loss_decoder = torch.sum(LAMBDA_MSE * mse) - (1.0 - LAMBDA_MSE) * loss_discriminator
    loss_aux = l1
    # well-conditioned single-loss gradients of the discriminator (the accumulated ones below cancel to 1e-6 of their
    # terms: loss_decoder carries -(1 - 1e-6) * loss_discriminator, train.py:66)
    dparams = [(k, p) for k, p in net.named_parameters() if k.startswith("discriminator.")]
    gd = torch.autograd.grad(loss_discriminator, [p for _, p in dparams], retain_graph=True, allow_unused=True)
    gm = torch.autograd.grad(torch.sum(mse), [p for _, p in dparams], retain_graph=True, allow_unused=True)
    net.zero_grad()
    loss_recon.backward(retain_graph=True)
    loss_encoder.backward(retain_graph=True)
    loss_decoder.backward(retain_graph=True)
    loss_discriminator.backward(retain_graph=True)
    loss_aux.backward()
    f = lambda t: t.detach().double().numpy()
    out = {"x_tilde": f(x_tilde), "disc_class": f(disc_class), "disc_layer_digest": digest(f(disc_layer)),
           "disc_layer_shape": np.array(disc_layer.shape), "mus": f(mus), "log_variances": f(log_variances), "params": f(params),
           "nle_digest": digest(f(nle)), "kl": f(kl), "mse": f(mse), "bce_dis_original": f(bce_o), "bce_dis_predicted": f(bce_p),
           "bce_dis_sampled": f(bce_s), "l1_enc_param": f(l1),
           "losses": np.array([float(loss_recon), float(loss_encoder), float(loss_decoder), float(loss_discriminator), float(loss_aux)])}
    grads = {k: f(p.grad) for k, p in net.named_parameters() if p.grad is not None}
    for (k, _), a, b in zip(dparams, gd, gm):
        if a is not None:
            grads["only_loss_discriminator/" + k] = f(a)
        if b is not None:
            grads["only_sum_mse/" + k] = f(b)
    return out, grads


def main(batch=4):
    import torch
    from oracle.gen_golden import import_reference
    networks, _, _ = import_reference()
    out64, g64 = run_reference(networks, torch.float64, b=batch)
    out32, g32 = run_reference(networks, torch.float32, b=batch)
    res = dict(out64)
    dev_keys, dev_vals = [], []
    for k, g in g64.items():
        full = g.size <= (FULL_LIMIT if batch <= 4 else 256)
        res["grad/" + k] = g if full else digest(g)
        res["gradfull/" + k] = np.array([full])
        a, b = g32[k], g
        dev_keys.append(k)
        dev_vals.append(float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300)))
    for k in ("x_tilde", "disc_class", "mus", "log_variances", "params", "kl", "mse", "losses"):
        dev_keys.append(k)
        dev_vals.append(float(np.abs(out32[k] - out64[k]).max() / (np.abs(out64[k]).max() + 1e-300)))
    res["ref_fp32_dev_keys"] = np.array(dev_keys)
    res["ref_fp32_dev_vals"] = np.array(dev_vals)
    res["meta"] = np.array([64, batch, 128, 0])
    path = os.path.join(ROOT, "tests", "golden", f"vaegan64_b{batch}.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "bytes;", len(g64), "gradients; worst fp32 deviation of the reference:",
          max(dev_vals), dev_keys[int(np.argmax(dev_vals))])


if __name__ == "__main__":
    main(4)
    main(16)      # bf16 fixture: BatchNorm1d over 16 / 48 samples instead of 4 / 12
