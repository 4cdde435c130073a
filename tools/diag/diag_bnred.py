import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import ctypes as C
import torch
import vae_play_b200 as vp
import vae_play_b200.functional as VF
from vae_play_b200 import _lib
from vae_play_b200.models.networks import DecoderBlock
vp.set_precision("bf16")
torch.manual_seed(0)
b1, b2 = DecoderBlock(256, 128).cuda().train(), DecoderBlock(128, 64).cuda().train()
x = torch.randn(256, 16, 16, 256, device="cuda").to(torch.bfloat16).requires_grad_(True)
a1 = b1.forward_cl(x)
print("holder on a1:", getattr(a1, "_vp_bn", None) is not None, (getattr(a1, "_vp_bn", {}) or {}).get("ok"))
a2 = b2.forward_cl(a1)
orig = VF.TapLayer.dgrad_bnred
def spy(self, dy, weight, x_shape, prev):
    r = orig(self, dy, weight, x_shape, prev)
    print("dgrad_bnred ->", None if r is None else ("fused", r[2]), "last_error:", _lib.load().vp_last_error())
    return r
VF.TapLayer.dgrad_bnred = spy
n0 = _lib.launch_count()
a2.float().sum().backward()
torch.cuda.synchronize()
print("backward launches", _lib.launch_count() - n0)
g_f = [p.grad.clone() for p in list(b1.parameters()) + list(b2.parameters())] + [x.grad.clone()]
VF.set_fuse_bn_backward(False)
for p in list(b1.parameters()) + list(b2.parameters()): p.grad = None
x.grad = None
a2 = b2.forward_cl(b1.forward_cl(x))
n0 = _lib.launch_count()
a2.float().sum().backward()
torch.cuda.synchronize()
print("backward launches (unfused)", _lib.launch_count() - n0)
g_u = [p.grad.clone() for p in list(b1.parameters()) + list(b2.parameters())] + [x.grad.clone()]
for a, b in zip(g_f, g_u):
    print(tuple(a.shape), float((a.double() - b.double()).norm() / b.double().norm()))
