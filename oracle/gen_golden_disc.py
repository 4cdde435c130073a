"""Generate tests/golden/disc64_b2.npz by running the UNMODIFIED reference Discriminator (TEST INFRASTRUCTURE ONLY).

    python oracle/gen_golden_disc.py        # build container only (needs /root/reference)

models/networks.py:151-195 (VAE-GAN discriminator, config 4 of BASELINE.json): weights drawn by ``synth_disc_params`` below
(shared with tests/test_gpu_parity.py), three input batches, float64.  Stored: the 'REC' features (pre-BatchNorm output of
the last EncoderBlock, flattened in NCHW order), the 'GAN' probabilities, and digests of two parameter gradients of
sum(GAN output * probe).
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_disc_params(seed=0, cin=1, iter_level=3):
    """Deterministic parameters with the reference's state_dict keys (uniform +-1/sqrt(fan_in), BatchNorm gamma in
    [0.5, 1.5], beta in [-0.2, 0.2])."""
    rs = np.random.RandomState(1000 + seed)
    P = {}

    def u(shape, fan):
        return (rs.uniform(-1, 1, size=shape) / np.sqrt(fan)).astype(np.float32)

    P["conv.0.0.weight"] = u((32, cin, 5, 5), cin * 25)
    P["conv.0.0.bias"] = u((32,), 25)
    c = 32
    for i in range(1, iter_level + 1):
        P[f"conv.{i}.conv.weight"] = u((2 * c, c, 5, 5), c * 25)
        P[f"conv.{i}.bn.weight"] = rs.uniform(0.5, 1.5, size=2 * c).astype(np.float32)
        P[f"conv.{i}.bn.bias"] = rs.uniform(-0.2, 0.2, size=2 * c).astype(np.float32)
        c *= 2
    P["fc.0.weight"] = u((512, 64 * c), 64 * c)
    P["fc.1.weight"] = rs.uniform(0.5, 1.5, size=512).astype(np.float32)
    P["fc.1.bias"] = rs.uniform(-0.2, 0.2, size=512).astype(np.float32)
    P["fc.3.weight"] = u((1, 512), 512)
    P["fc.3.bias"] = u((1,), 512)
    return P


def synth_disc_inputs(seed=0, b=2, img=64, cin=1):
    rs = np.random.RandomState(2000 + seed)
    return [rs.uniform(0, 1, size=(b, cin, img, img)).astype(np.float32) for _ in range(3)], rs.uniform(-1, 1, size=(3 * b, 1)).astype(np.float32)


def digest(a, nsamp=256):
    a = np.asarray(a, np.float64).ravel()
    idx = np.linspace(0, a.size - 1, num=min(nsamp, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()], a[idx]])


def main():
    import torch
    from oracle.gen_golden import import_reference
    networks, _, _ = import_reference()
    P = synth_disc_params(0)
    xs, probe = synth_disc_inputs(0)
    d = networks.Discriminator(channel_in=1, recon_level=3, iter_level=3)
    missing = d.load_state_dict({k: torch.from_numpy(v) for k, v in P.items()}, strict=False)
    assert not missing.unexpected_keys, missing
    d = d.double().train()
    xt = [torch.from_numpy(x).double() for x in xs]
    rec = d(*xt, "REC")
    d.zero_grad()
    gan = d(*xt, "GAN")
    (gan * torch.from_numpy(probe).double()).sum().backward()
    out = {"rec_digest": digest(rec.detach().numpy()), "rec_shape": np.array(rec.shape), "gan": gan.detach().numpy(),
           "grad_conv0": digest(d.conv[0][0].weight.grad.numpy()), "grad_conv2": digest(d.conv[2].conv.weight.grad.numpy()),
           "grad_fc3": d.fc[3].weight.grad.numpy()}
    path = os.path.join(ROOT, "tests", "golden", "disc64_b2.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
