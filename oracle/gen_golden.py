"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

It imports ``models.networks`` / ``models.blocks`` / ``tools.ops`` from
/root/reference (with the import shims SURVEY.md section 8c lists), loads the
deterministic NumPy-drawn parameters of ``oracle.vae_numpy.synth_vae_params``
into the reference modules, executes forward + loss + backward in float64
(truth) and float32 (tolerance calibration), and stores the results.  The
fixtures are small: full tensors only for small outputs, (sum, l2, strided
samples) digests for the big gradients.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    # shims: SURVEY.md section 8c
    sys.modules.setdefault("turtle", types.ModuleType("turtle"))
    sys.modules["turtle"].shape = None
    sk = types.ModuleType("skimage")
    skm = types.ModuleType("skimage.measure")
    skm.find_contours = lambda *a, **k: []
    sk.measure = skm
    sys.modules.setdefault("skimage", sk)
    sys.modules.setdefault("skimage.measure", skm)
    sys.path.insert(0, REF)
    import models.networks as networks  # noqa
    import models.blocks as blocks  # noqa
    import tools.ops as ops  # noqa
    return networks, blocks, ops


def digest(a, nsamp=64):
    a = np.asarray(a, np.float64).ravel()
    idx = np.linspace(0, a.size - 1, num=min(nsamp, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()], a[idx]])


def run_vae(networks, img, cin, b, z, seed, dtype):
    import math
    import torch
    import torch.nn.functional as F
    from oracle import vae_numpy as vn

    L = int(math.log2(img // 8))
    P = vn.synth_vae_params(img, z, cin, cin, seed)
    x_np, eps_np = vn.synth_batch(b, img, cin, z, seed)
    enc = networks.Encoder(channel_in=cin, z_size=z, iter_level=L)
    dec = networks.Decoder(z_size=z, size=enc.size, channel_out=cin, iter_level=L)
    sd_e = {k[len("encoder."):]: torch.from_numpy(v) for k, v in P.items() if k.startswith("encoder.")}
    sd_d = {k[len("decoder."):]: torch.from_numpy(v) for k, v in P.items() if k.startswith("decoder.")}
    enc.load_state_dict(sd_e, strict=False)
    dec.load_state_dict(sd_d, strict=False)
    enc = enc.to(dtype).train()
    dec = dec.to(dtype).train()
    x = torch.from_numpy(x_np).to(dtype)
    eps = torch.from_numpy(eps_np).to(dtype)

    mu, logvar = enc(x)
    # VaeGan.reparameterize (networks.py:228-231) with the draw replaced by the supplied eps
    orig_normal = torch.Tensor.normal_
    torch.Tensor.normal_ = lambda self, *a, **k: self.copy_(eps)
    try:
        z_t = networks.VaeGan.reparameterize(None, mu, logvar)
    finally:
        torch.Tensor.normal_ = orig_normal
    x_tilde = dec(z_t)
    dummy = torch.zeros(b, 1, dtype=dtype)
    nle, kl, *_ = networks.VaeGan.loss(x, x_tilde, dummy, dummy, dummy, dummy + .5, dummy + .5, dummy + .5,
                                       mu, logvar, torch.zeros(b, 3, dtype=dtype), torch.zeros(b, 3, dtype=dtype))
    loss_recon = F.mse_loss(x, x_tilde)              # train.py:62
    loss = loss_recon + torch.sum(kl)                # train.py:63 (VAE terms)
    enc.zero_grad()
    dec.zero_grad()
    loss.backward()
    out = {
        "mu": mu.detach().numpy(), "logvar": logvar.detach().numpy(), "z": z_t.detach().numpy(),
        "x_tilde": x_tilde.detach().numpy(), "kl": kl.detach().numpy(),
        "nle_sum": np.array(nle.detach().sum().item()),
        "mse": np.array(loss_recon.item()), "loss": np.array(loss.item()),
    }
    for pref, m in (("encoder", enc), ("decoder", dec)):
        for k, p in m.named_parameters():
            out[f"grad/{pref}.{k}"] = digest(p.grad.numpy())
        for k, buf in m.named_buffers():
            if "running" in k:
                out[f"running/{pref}.{k}"] = buf.detach().numpy().astype(np.float64)
    return out


def gen_vae_cases(networks):
    import torch
    cases = {"vae64_c1_b4": (64, 1, 4, 128, 0), "vae64_c3_b4": (64, 3, 4, 128, 1), "vae128_c1_b4": (128, 1, 4, 128, 2)}
    for name, (img, cin, b, z, seed) in cases.items():
        o64 = run_vae(networks, img, cin, b, z, seed, torch.float64)
        o32 = run_vae(networks, img, cin, b, z, seed, torch.float32)
        blob = {"meta": np.array([img, cin, b, z, seed])}
        for k, v in o64.items():
            blob[k] = np.asarray(v, np.float64)
        # fp32-vs-fp64 deviation of the reference itself: the floor for any fp32 tolerance
        dev = {}
        for k in o64:
            a, bb = np.asarray(o64[k], np.float64), np.asarray(o32[k], np.float64)
            dev[k] = float(np.abs(a - bb).max() / (np.abs(a).max() + 1e-30))
        blob["ref_fp32_dev_keys"] = np.array(list(dev.keys()))
        blob["ref_fp32_dev_vals"] = np.array(list(dev.values()))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **blob)
        print(name, "loss", o64["loss"], "max ref fp32 dev", max(dev.values()))


def gen_op_cases(networks, blocks, ops):
    """Per-operator fixtures through the reference's own wrapper classes (full tensors, tiny shapes)."""
    import torch
    import torch.nn.functional as F
    from oracle import vae_numpy as vn
    torch.manual_seed(0)
    blob = {}

    def t(name, shape, lo=-1.0, hi=1.0):
        return torch.from_numpy(vn.synth_tensor(name, shape, 11, lo, hi)).double()

    # EncoderBlock (networks.py:10-30): conv5x5 s2 no-bias + BN(momentum .9) + ReLU, both out modes
    m = networks.EncoderBlock(6, 10).double().train()
    with torch.no_grad():
        m.conv.weight.copy_(t("eb.w", (10, 6, 5, 5), -.2, .2))
        m.bn.weight.copy_(t("eb.g", (10,), .5, 1.5))
        m.bn.bias.copy_(t("eb.b", (10,), -.3, .3))
    x = t("eb.x", (3, 6, 12, 10)).requires_grad_(True)
    y, ypre = m(x, out=True)
    dy = t("eb.dy", tuple(y.shape))
    y.backward(dy)
    blob.update({"eb/x": x.detach(), "eb/w": m.conv.weight.detach(), "eb/g": m.bn.weight.detach(), "eb/b": m.bn.bias.detach(),
                 "eb/y": y.detach(), "eb/ypre": ypre.detach(), "eb/dy": dy, "eb/dx": x.grad,
                 "eb/dw": m.conv.weight.grad, "eb/dg": m.bn.weight.grad, "eb/db": m.bn.bias.grad,
                 "eb/rm": m.bn.running_mean, "eb/rv": m.bn.running_var})

    # DecoderBlock (networks.py:34-46): convT5x5 s2 p2 op1 no-bias + BN + ReLU
    m = networks.DecoderBlock(10, 6).double().train()
    with torch.no_grad():
        m.conv.weight.copy_(t("db.w", (10, 6, 5, 5), -.2, .2))
        m.bn.weight.copy_(t("db.g", (6,), .5, 1.5))
        m.bn.bias.copy_(t("db.b", (6,), -.3, .3))
    x = t("db.x", (3, 10, 5, 7)).requires_grad_(True)
    y = m(x)
    dy = t("db.dy", tuple(y.shape))
    y.backward(dy)
    blob.update({"db/x": x.detach(), "db/w": m.conv.weight.detach(), "db/g": m.bn.weight.detach(), "db/b": m.bn.bias.detach(),
                 "db/y": y.detach(), "db/dy": dy, "db/dx": x.grad, "db/dw": m.conv.weight.grad,
                 "db/dg": m.bn.weight.grad, "db/db": m.bn.bias.grad, "db/rm": m.bn.running_mean, "db/rv": m.bn.running_var})

    # blocks.Conv2d variants (blocks.py:5-34)
    variants = {"c_k3s1_batch_relu": (5, 7, 3, 1, "batch", "relu"), "c_k4s2_inst_lrelu": (4, 6, 4, 2, "instance", "lrelu"),
                "c_k1s1_none_tanh": (6, 3, 1, 1, None, "tanh"), "c_k5s1_none_none": (3, 2, 5, 1, None, None),
                "c_k3s2_batch_lrelu": (4, 8, 3, 2, "batch", "lrelu")}
    for name, (ci, co, k, s, bn, act) in variants.items():
        m = blocks.Conv2d(ci, co, k, stride=s, bn=bn, activate=act).double().train()
        conv = m.conv[0]
        with torch.no_grad():
            conv.weight.copy_(t(name + ".w", tuple(conv.weight.shape), -.3, .3))
            if conv.bias is not None:
                conv.bias.copy_(t(name + ".bias", tuple(conv.bias.shape), -.2, .2))
            if bn == "batch":
                m.conv[1].weight.copy_(t(name + ".g", (co,), .5, 1.5))
                m.conv[1].bias.copy_(t(name + ".b", (co,), -.3, .3))
        x = t(name + ".x", (2, ci, 10, 12)).requires_grad_(True)
        y = m(x)
        dy = t(name + ".dy", tuple(y.shape))
        y.backward(dy)
        blob.update({f"{name}/x": x.detach(), f"{name}/w": conv.weight.detach(), f"{name}/y": y.detach(),
                     f"{name}/dy": dy, f"{name}/dx": x.grad, f"{name}/dw": conv.weight.grad})
        if conv.bias is not None:
            blob[f"{name}/bias"] = conv.bias.detach()
            blob[f"{name}/dbias"] = conv.bias.grad
        if bn == "batch":
            blob.update({f"{name}/g": m.conv[1].weight.detach(), f"{name}/b": m.conv[1].bias.detach(),
                         f"{name}/dg": m.conv[1].weight.grad, f"{name}/db": m.conv[1].bias.grad})

    # ConvTranspose2d k4 s2 p1 bias (network_Style_GAN.py:49)
    ct = torch.nn.ConvTranspose2d(5, 4, 4, 2, 1).double()
    with torch.no_grad():
        ct.weight.copy_(t("ct4.w", (5, 4, 4, 4), -.3, .3))
        ct.bias.copy_(t("ct4.bias", (4,), -.2, .2))
    x = t("ct4.x", (2, 5, 6, 5)).requires_grad_(True)
    y = ct(x)
    dy = t("ct4.dy", tuple(y.shape))
    y.backward(dy)
    blob.update({"ct4/x": x.detach(), "ct4/w": ct.weight.detach(), "ct4/bias": ct.bias.detach(), "ct4/y": y.detach(),
                 "ct4/dy": dy, "ct4/dx": x.grad, "ct4/dw": ct.weight.grad, "ct4/dbias": ct.bias.grad})

    # blocks.Linear (blocks.py:36-50): lrelu slope is 0.2 here (0.02 after convs)
    m = blocks.Linear(9, 7, bias=True, activate="lrelu").double()
    with torch.no_grad():
        m.fc[0].weight.copy_(t("lin.w", (7, 9)))
        m.fc[0].bias.copy_(t("lin.bias", (7,)))
    x = t("lin.x", (5, 9)).requires_grad_(True)
    y = m(x)
    dy = t("lin.dy", (5, 7))
    y.backward(dy)
    blob.update({"lin/x": x.detach(), "lin/w": m.fc[0].weight.detach(), "lin/bias": m.fc[0].bias.detach(), "lin/y": y.detach(),
                 "lin/dy": dy, "lin/dx": x.grad, "lin/dw": m.fc[0].weight.grad, "lin/dbias": m.fc[0].bias.grad})

    # reparameterize + VaeGan.loss KL/nle (networks.py:228-231, 264-270) with the CPU generator's own eps
    mu = t("rp.mu", (6, 16)).requires_grad_(True)
    lv = t("rp.lv", (6, 16)).requires_grad_(True)
    torch.manual_seed(123)
    eps = torch.empty(6, 16, dtype=torch.float64).normal_()
    torch.manual_seed(123)
    z = networks.VaeGan.reparameterize(None, mu, lv)
    d1 = torch.zeros(6, 1, dtype=torch.float64)
    xa, xb = t("rp.x", (6, 1, 4, 4), 0, 1), t("rp.xt", (6, 1, 4, 4), 0, 1)
    nle, kl, *_ = networks.VaeGan.loss(xa, xb, d1, d1, d1, d1 + .5, d1 + .5, d1 + .5, mu, lv,
                                       torch.zeros(6, 3).double(), torch.zeros(6, 3).double())
    dz = t("rp.dz", (6, 16))
    (kl.sum() + (z * dz).sum()).backward()
    blob.update({"rp/mu": mu.detach(), "rp/lv": lv.detach(), "rp/eps": eps, "rp/z": z.detach(), "rp/kl": kl.detach(),
                 "rp/nle": nle.detach(), "rp/x": xa, "rp/xt": xb, "rp/dz": dz, "rp/dmu": mu.grad, "rp/dlv": lv.grad})

    # recon / segmentation losses: F.mse_loss (train.py:62), F.l1_loss (train_Style_GAN.py:220),
    # 0.5*BCEWithLogits + dice(sigmoid) (train_BE.py:58-59, tools/ops.py:12-19)
    a = t("ls.x", (3, 1, 8, 8), 0, 1)
    for nm, fn in (("mse", F.mse_loss), ("l1", F.l1_loss)):
        bt = t("ls.xt", (3, 1, 8, 8), 0, 1).requires_grad_(True)
        l = fn(a, bt)
        l.backward()
        blob.update({f"ls/{nm}": l.detach(), f"ls/{nm}_dxt": bt.grad})
    blob["ls/x"] = a
    blob["ls/xt"] = t("ls.xt", (3, 1, 8, 8), 0, 1)
    logits = t("ls.logits", (3, 1, 8, 8), -3, 3).requires_grad_(True)
    tgt = (t("ls.t", (3, 1, 8, 8), 0, 1) > 0.5).double()
    l = F.binary_cross_entropy_with_logits(logits, tgt) * 0.5 + ops.compute_dice_loss(torch.sigmoid(logits), tgt)
    l.backward()
    blob.update({"ls/logits": logits.detach(), "ls/t": tgt, "ls/bce_dice": l.detach(), "ls/bce_dice_dlogits": logits.grad})

    np.savez_compressed(os.path.join(OUT, "ops.npz"), **{k: np.asarray(v.detach() if hasattr(v, "detach") else v, np.float64)
                                                         for k, v in blob.items()})
    print("ops.npz:", len(blob), "arrays")


def gen_state_dict_keys(networks):
    """state_dict key/shape lists of the stock VaeGan (the drop-in contract of SURVEY.md section 8b)."""
    import json
    out = {}
    for img in (64, 128):
        m = networks.VaeGan(img, 128)
        out[str(img)] = [(k, list(v.shape)) for k, v in m.state_dict().items()]
    json.dump(out, open(os.path.join(OUT, "state_dict_keys.json"), "w"))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    networks, blocks, ops = import_reference()
    gen_state_dict_keys(networks)
    gen_op_cases(networks, blocks, ops)
    gen_vae_cases(networks)
