"""Diagnostic (not a test): each VAE layer in isolation vs torch fp64 on the GPU, fwd + bwd."""
import sys
import numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, ".")
import vae_play_b200 as vp
import vae_play_b200.functional as VF

def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-300))

def run(kind, cin, cout, hw, B, mode, seed=0, repeat=2):
    vp.set_precision(mode)
    torch.manual_seed(seed)
    dt = VF.act_dtype()
    if kind == "conv":
        layer = VF.TapLayer("conv", cin, cout, k=5, stride=2, pad=2); w = torch.randn(cout, cin, 5, 5, device="cuda") * 0.05
    elif kind == "convT":
        layer = VF.TapLayer("convT", cin, cout, k=5, stride=2, pad=2, out_pad=1); w = torch.randn(cin, cout, 5, 5, device="cuda") * 0.05
    elif kind == "flatten_in":
        layer = VF.TapLayer("flatten_in", cin, cout, spatial=8); w = torch.randn(cout, cin * 64, device="cuda") * 0.02
    else:
        layer = VF.TapLayer("flatten_out", cin, cout, spatial=8); w = torch.randn(cout * 64, cin, device="cuda") * 0.05
    perm = 64 if kind == "flatten_out" else 0
    nfeat = cout * 64 if kind == "flatten_out" else cout
    gamma = (torch.rand(nfeat, device="cuda") + 0.5).requires_grad_(True)
    beta = ((torch.rand(nfeat, device="cuda") - 0.5) * 0.4).requires_grad_(True)
    x32 = torch.randn(B, cin, hw, hw, device="cuda")
    w.requires_grad_(True)
    res = []
    for it in range(repeat):
        x = x32.clone().requires_grad_(True)
        for t in (w, gamma, beta):
            t.grad = None
        bn = (torch.nn.BatchNorm1d if perm else torch.nn.BatchNorm2d)(nfeat, momentum=0.9).cuda().train()
        a, y = VF.fused_layer(VF.to_channels_last(x), w, None, gamma, beta, layer, VF.NormCfg("batch", momentum=0.9, perm_T=perm), "relu", 0.0, True, bn)
        out = VF.from_channels_last(a)
        torch.manual_seed(seed + 1)
        dout = torch.randn(out.shape, device="cuda")
        out.backward(dout)
        res.append((out.detach().clone(), x.grad.clone(), w.grad.clone(), gamma.grad.clone(), beta.grad.clone()))
    # torch fp64 reference
    xd = x32.double().requires_grad_(True); wd = w.detach().double().requires_grad_(True)
    gd = gamma.detach().double().requires_grad_(True); bd = beta.detach().double().requires_grad_(True)
    if kind == "conv": yd = F.conv2d(xd, wd, None, 2, 2)
    elif kind == "convT": yd = F.conv_transpose2d(xd, wd, None, 2, 2, 1)
    elif kind == "flatten_in": yd = F.linear(xd.reshape(B, -1), wd)
    else: yd = F.linear(xd.reshape(B, -1), wd)
    od = F.relu(F.batch_norm(yd, None, None, gd, bd, True, 0.9, 1e-5))
    if kind == "flatten_out": od = od.reshape(B, cout, 8, 8)
    if kind == "flatten_in": od = od.reshape(B, cout, 1, 1)
    od.backward(dout.double())
    names = ["out", "dx", "dw", "dgamma", "dbeta"]
    refs = [od.detach(), xd.grad, wd.grad, gd.grad, bd.grad]
    errs = {n: f"{rel(r0, ref):.1e}" for n, r0, ref in zip(names, res[0], refs)}
    nondet = {n: f"{rel(r1, r0):.1e}" for n, r0, r1 in zip(names, res[0], res[1])}
    print(f"{mode} {kind:11s} cin={cin:4d} cout={cout:4d} hw={hw:2d} B={B:3d} err {errs}  run2-vs-run1 {nondet}", flush=True)

for mode in ("fp32", "bf16"):
    for B in (4, 8, 32):
        run("conv", 1, 64, 64, B, mode)
        run("conv", 64, 128, 32, B, mode)
        run("conv", 128, 256, 16, B, mode)
        run("flatten_in", 256, 1024, 8, B, mode)
        run("flatten_out", 128, 256, 1, B, mode)
        run("convT", 256, 256, 8, B, mode)
        run("convT", 256, 128, 16, B, mode)
        run("convT", 128, 64, 32, B, mode)
