import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vae_play_b200 as vp
import vae_play_b200.functional as VF
from vae_play_b200 import _lib
from torch.profiler import profile, ProfilerActivity
vp.set_precision("bf16")
def run(name, fn):
    fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    ks = [(e.key[:60], round(e.device_time_total, 1)) for e in prof.key_averages() if e.device_time_total]
    print(name, ks, "last_error:", _lib.load().vp_last_error())
# discriminator conv[1]: EncoderBlock 32 -> 128, k5 s2, on 192 x 128 x 128
lay = VF.TapLayer("conv", 32, 128, k=5, stride=2, pad=2)
w = torch.randn(128, 32, 5, 5, device="cuda") * 0.05
x = torch.randn(192, 128, 128, 32, device="cuda").to(torch.bfloat16)
y = lay.fwd(x, w, None)
dy = torch.randn_like(y)
run("disc.conv1 fwd", lambda: lay.fwd(x, w, None))
run("disc.conv1 dgrad", lambda: lay.dgrad(dy, w, tuple(x.shape)))
run("disc.conv1 wgrad", lambda: lay.wgrad(x, dy, w))
# decoder last block at 128: convT 128 -> 64, 64x64 -> 128x128, batch 64
lay2 = VF.TapLayer("convT", 128, 64, k=5, stride=2, pad=2, out_pad=1)
w2 = (torch.randn(128, 64, 5, 5, device="cuda") * 0.05).contiguous(memory_format=torch.channels_last)
x2 = torch.randn(64, 64, 64, 128, device="cuda").to(torch.bfloat16)
y2 = lay2.fwd(x2, w2, None)
dy2 = torch.randn_like(y2)
run("dec.ct4 fwd", lambda: lay2.fwd(x2, w2, None))
run("dec.ct4 dgrad", lambda: lay2.dgrad(dy2, w2, tuple(x2.shape)))
run("dec.ct4 wgrad", lambda: lay2.wgrad(x2, dy2, w2))
