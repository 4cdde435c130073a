import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import vae_play_b200 as vp
import vae_play_b200.functional as VF
from vae_play_b200 import _lib
from vae_play_b200.models.networks import DecoderBlock
from vae_play_b200.functional import TapLayer, NormCfg
vp.set_precision("bf16")
torch.manual_seed(0)
for (B, H) in ((256, 32), (3, 9), (64, 64)):
    b1 = DecoderBlock(128, 64).cuda().train()
    last = torch.nn.Conv2d(64, 1, 5, 1, 2).cuda()
    lay, nn_ = TapLayer("conv", 64, 1, k=5, stride=1, pad=2), NormCfg(None)
    x = torch.randn(B, H, H, 128, device="cuda").to(torch.bfloat16).requires_grad_(True)
    probe = torch.randn(B, 2 * H, 2 * H, 1, device="cuda")
    def run():
        a1 = b1.forward_cl(x)
        xt, _ = VF.fused_layer(a1, last.weight, last.bias, None, None, lay, nn_, "sigmoid", 0.0, True, None)
        n0 = _lib.launch_count()
        (xt.float() * probe).sum().backward()
        torch.cuda.synchronize()
        ps = list(b1.parameters()) + list(last.parameters())
        g = [p.grad.clone() for p in ps] + [x.grad.clone()]
        for p in ps: p.grad = None
        x.grad = None
        return _lib.launch_count() - n0, g
    VF.set_fuse_bn_backward(True)
    nf, gf = run()
    VF.set_fuse_bn_backward(False)
    nu, gu = run()
    print(B, H, "launches fused/unfused", nf, nu, "simt", _lib.simt_bf16_count())
    for a, b in zip(gf, gu):
        print("  ", tuple(a.shape), float((a.double() - b.double()).norm() / b.double().norm()))
