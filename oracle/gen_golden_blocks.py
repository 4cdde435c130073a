"""Generate tests/golden/blocks.npz by running the UNMODIFIED reference blocks / Style_GAN modules (TEST INFRASTRUCTURE ONLY).

    python oracle/gen_golden_blocks.py        # build container only (needs /root/reference)

Covers SURVEY.md section 8 rows A10 (Up / bilinear), A11 (dice alone), A12 (StyleEncoder / StyleUp / myConv2d / SCSEBlock /
Generator / Discriminator of models/network_Style_GAN.py) and f3 (AddCoords, Down, AdaptiveAvgPool2d, SelfAttentionBlock,
edge_loss).  Every case: the reference module in float64 with parameters drawn by ``synth_state`` (shared with the GPU tests,
so parameters are not stored), inputs, outputs and gradients (full when small, (sum, l2, max, strided samples) digests
otherwise), plus the reference's own float32-vs-float64 deviation per stored quantity.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FULL_LIMIT = 2048


def synth_state(module, seed):
    """Deterministic parameters for any mirror / reference module: keyed by the state_dict order and shapes only.
    >= 2-D weights: U(-1,1)/sqrt(fan_in); 1-D '.weight' (BatchNorm gamma): U(.5,1.5); biases: U(-.1,.1); 'gamma' (attention): .5."""
    import torch
    rs = np.random.RandomState(7000 + seed)
    sd = {}
    for k, v in module.state_dict().items():
        if "running_" in k or "num_batches" in k:
            continue
        shp = tuple(v.shape)
        if k.endswith("gamma"):
            a = np.full(shp, 0.5)
        elif len(shp) >= 2:
            a = rs.uniform(-1, 1, size=shp) / np.sqrt(np.prod(shp[1:]))
        elif k.endswith(".weight"):
            a = rs.uniform(0.5, 1.5, size=shp)
        else:
            a = rs.uniform(-0.1, 0.1, size=shp)
        sd[k] = torch.from_numpy(a.astype(np.float32))
    return sd


def synth_input(name, shape, lo=-1.0, hi=1.0):
    rs = np.random.RandomState(abs(hash_name(name)) % (2 ** 31))
    return rs.uniform(lo, hi, size=shape).astype(np.float32)


def hash_name(name):
    h = 0
    for ch in name:
        h = (h * 131 + ord(ch)) % 1000003
    return h


def digest(a, nsamp=128):
    a = np.asarray(a, np.float64).ravel()
    idx = np.linspace(0, a.size - 1, num=min(nsamp, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.sqrt((a * a).sum()), np.abs(a).max()], a[idx]])


def store(res, key, a):
    a = np.asarray(a, np.float64)
    full = a.size <= FULL_LIMIT
    res[key] = a if full else digest(a)
    res["full/" + key] = np.array([full])


# case name -> (module builder (given the reference modules), list of (input name, shape, kind), forward closure)
def cases(blocks, style, ops):
    import torch
    import torch.nn.functional as F

    def lab(n):
        return torch.tensor([1.0, 0.0, 1.0][:n])

    return {
        "scse": (lambda: blocks.SCSEBlock(64, 4), [("x", (2, 64, 6, 5))], lambda m, x: m(x)),
        "attn": (lambda: blocks.SelfAttentionBlock(16), [("x", (2, 16, 5, 4))], lambda m, x: m(x)),
        "addcoords": (lambda: blocks.AddCoords(False), [("x", (2, 3, 4, 5))], lambda m, x: m(x)),
        "addcoords_norm": (lambda: blocks.AddCoords(True), [("x", (2, 3, 4, 5))], lambda m, x: m(x)),
        "up": (lambda: blocks.Up(6, 8, if_add_coord=True), [("x", (3, 6, 5, 4))], lambda m, x: m(x)),
        "down": (lambda: blocks.Down(5, 8, 3, True), [("x", (2, 5, 8, 6))], lambda m, x: m(x)),
        "styleup": (lambda: style.StyleUp(64, 64), [("x", (2, 64, 4, 4)), ("skip", (2, 64, 8, 8))], lambda m, x, s: m(x, s)),
        "myconv": (lambda: style.myConv2d(4, 8, 4, 2, bn="instance"), [("x", (3, 4, 8, 8))], lambda m, x: m(x, lab(3).reshape(3, 1, 1, 1).to(x.dtype))),
        "generator": (lambda: style.Generator(32, 16), [("x", (2, 3, 32, 32)), ("style", (2, 16))],
                      lambda m, x, s: m(x, s, lab(2).to(x.dtype))),
        "styleenc": (lambda: style.StyleEncoder(16, 32), [("x", (2, 3, 32, 32))], lambda m, x: torch.cat(m(x), dim=1)),
        "sdisc": (lambda: style.Discriminator(32, 3), [("x", (2, 3, 32, 32)), ("xc", (2, 3, 32, 32))],
                  lambda m, x, xc: torch.cat(m(x, xc, None), dim=1)),
        "avgpool4": (lambda: torch.nn.AdaptiveAvgPool2d((4, 4)), [("x", (2, 8, 9, 10))], lambda m, x: m(x)),
        "bilinear": (lambda: torch.nn.Identity(), [("x", (2, 3, 5, 7))], lambda m, x: F.interpolate(x, scale_factor=2, mode="bilinear")),
        "dice": (lambda: torch.nn.Identity(), [("p", (3, 1, 8, 8))],
                 lambda m, p: ops.compute_dice_loss(p.sigmoid(), (torch.from_numpy(synth_input("dice/t", (3, 1, 8, 8))) > 0).to(p.dtype)).reshape(1)),
        "edge": (lambda: torch.nn.Identity(), [("p", (2, 1, 9, 8))],
                 lambda m, p: ops.edge_loss(p.sigmoid(), (torch.from_numpy(synth_input("edge/t", (2, 1, 9, 8))) > 0).to(p.dtype)).reshape(1)),
    }


def run_case(name, spec, dtype):
    import torch
    build, inputs, fwd = spec
    torch.manual_seed(0)
    m = build()
    sd = synth_state(m, hash_name(name) % 1000)
    m.load_state_dict(sd, strict=False)
    m = m.to(dtype).train()
    xs = [torch.from_numpy(synth_input(f"{name}/{nm}", shp)).to(dtype).requires_grad_(True) for nm, shp in inputs]
    # edge_loss builds its 3x3 kernel with torch.FloatTensor (tools/ops.py:193): give it the run's dtype for the float64 truth
    orig_ft = torch.FloatTensor
    if dtype == torch.float64:
        torch.FloatTensor = torch.DoubleTensor
    try:
        y = fwd(m, *xs)
    finally:
        torch.FloatTensor = orig_ft
    probe = torch.from_numpy(synth_input(f"{name}/probe", tuple(y.shape))).to(dtype)
    (y * probe).sum().backward()
    out = {"y": y.detach().double().numpy()}
    for (nm, _), x in zip(inputs, xs):
        out["d" + nm] = x.grad.double().numpy()
    for k, p in m.named_parameters():
        if p.grad is not None:
            out["g/" + k] = p.grad.double().numpy()
    return out


def main():
    import torch
    from oracle.gen_golden import import_reference
    _, blocks, ops = import_reference()
    import models.network_Style_GAN as style
    res = {}
    dev_keys, dev_vals = [], []
    for name, spec in cases(blocks, style, ops).items():
        o64 = run_case(name, spec, torch.float64)
        o32 = run_case(name, spec, torch.float32)
        for k, v in o64.items():
            store(res, f"{name}/{k}", v)
            dev_keys.append(f"{name}/{k}")
            dev_vals.append(float(np.abs(o32[k] - v).max() / (np.abs(v).max() + 1e-300)))
        print(name, {k: v.shape for k, v in list(o64.items())[:4]}, "worst fp32 dev", max(dev_vals[-len(o64):]))
    res["ref_fp32_dev_keys"] = np.array(dev_keys)
    res["ref_fp32_dev_vals"] = np.array(dev_vals)
    path = os.path.join(ROOT, "tests", "golden", "blocks.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
