"""Diagnostic: bisect the padded route vs the native in-place route on K = N = 64 stride-1 shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.nn.functional as F
import vae_play_b200 as vp
import vae_play_b200.functional as VF
vp.set_precision("bf16")
g = torch.Generator(device="cuda").manual_seed(1)
def rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max())
def lin(cin, cout, b, bias=True):
    layer = VF.TapLayer("linear", cin, cout)
    w = torch.randn(cout, cin, device="cuda", generator=g) * 0.1
    x = torch.randn(b, 1, 1, cin, device="cuda", generator=g).to(torch.bfloat16)
    bs = torch.randn(cout, device="cuda", generator=g) if bias else None
    y = layer.fwd(x, w, bs, out_dtype=torch.float32)
    want = x.double().reshape(b, cin) @ w.to(torch.bfloat16).double().T + (bs.double() if bias else 0)
    dy = torch.randn(b, 1, 1, cout, device="cuda", generator=g).to(torch.bfloat16)
    dw = layer.wgrad(x, dy, w)
    wdw = dy.double().reshape(b, cout).T @ x.double().reshape(b, cin)
    dx = layer.dgrad(dy, w, tuple(x.shape), out_dtype=torch.float32)
    wdx = dy.double().reshape(b, cout) @ w.to(torch.bfloat16).double()
    print(f"linear {cin}->{cout} b={b} bias={bias}: route cl={layer._cl(torch.bfloat16, w)} fwd {rel(y.reshape(b, cout), want):.2e} dgrad {rel(dx.reshape(b, cin), wdx):.2e} wgrad {rel(dw, wdw):.2e}", flush=True)
def conv(cin, cout, k, s, hw, b, cl=False):
    layer = VF.TapLayer("conv", cin, cout, k=k, stride=s, pad=(k - 1) // 2)
    w = torch.randn(cout, cin, k, k, device="cuda", generator=g) * 0.1
    if cl: w = w.contiguous(memory_format=torch.channels_last)
    x = torch.randn(b, hw, hw, cin, device="cuda", generator=g).to(torch.bfloat16)
    y = layer.fwd(x, w, None, out_dtype=torch.float32)
    xd, wd = x.double().permute(0, 3, 1, 2), w.to(torch.bfloat16).double()
    want = F.conv2d(xd, wd, None, stride=s, padding=(k - 1) // 2)
    dy = torch.randn(y.shape, device="cuda", generator=g).to(torch.bfloat16)
    dyd = dy.double().permute(0, 3, 1, 2)
    dw = layer.wgrad(x, dy, w)
    wdw = torch.nn.grad.conv2d_weight(xd, wd.shape, dyd, stride=s, padding=(k - 1) // 2)
    dx = layer.dgrad(dy, w, tuple(x.shape), out_dtype=torch.float32)
    wdx = torch.nn.grad.conv2d_input(xd.shape, wd, dyd, stride=s, padding=(k - 1) // 2)
    print(f"conv {cin}->{cout} k{k} s{s} hw{hw} b{b} cl={layer._cl(torch.bfloat16, w)}: fwd {rel(y.permute(0, 3, 1, 2), want):.2e} dgrad {rel(dx.permute(0, 3, 1, 2), wdx):.2e} wgrad {rel(dw, wdw):.2e}", flush=True)
for args in ((32, 2, 64), (64, 64, 64), (64, 64, 128), (64, 64, 200), (128, 64, 64), (64, 128, 64), (32, 2, 200), (96, 40, 48)):
    lin(*args)
lin(64, 64, 64, bias=False)
for args in ((64, 64, 3, 1, 24, 3, True), (32, 32, 3, 1, 24, 3), (64, 64, 3, 1, 8, 3, True), (64, 64, 1, 1, 8, 3, True), (64, 64, 5, 1, 8, 2, True), (64, 64, 3, 2, 18, 3, True),
             (64, 128, 3, 1, 24, 3, True), (128, 64, 3, 1, 24, 3, True), (5, 7, 3, 1, 8, 4), (3, 2, 5, 1, 8, 2)):
    conv(*args)
