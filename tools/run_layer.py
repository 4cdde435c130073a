"""Run one contraction of the VAE a few times (ncu target).  usage: run_layer.py [ct3|ct2|ct1|enc2|enc3|ct3_dgrad|ct3_wgrad|out] [iters]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vae_play_b200 as vp
import vae_play_b200.functional as VF

which = sys.argv[1] if len(sys.argv) > 1 else "ct3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
B = 256
vp.set_precision("bf16")
cfg = {"ct3": ("convT", 128, 64, 32), "ct2": ("convT", 256, 128, 16), "ct1": ("convT", 256, 256, 8),
       "enc2": ("conv", 64, 128, 32), "enc3": ("conv", 128, 256, 16), "out": ("conv1", 64, 1, 64), "enc1": ("conv", 1, 64, 64)}
name = which.split("_")[0]
kind, cin, cout, hw = cfg[name]
if kind == "convT":
    layer = VF.TapLayer("convT", cin, cout, k=5, stride=2, pad=2, out_pad=1); w = torch.randn(cin, cout, 5, 5, device="cuda") * 0.05
elif kind == "conv":
    layer = VF.TapLayer("conv", cin, cout, k=5, stride=2, pad=2); w = torch.randn(cout, cin, 5, 5, device="cuda") * 0.05
else:
    layer = VF.TapLayer("conv", cin, cout, k=5, stride=1, pad=2); w = torch.randn(cout, cin, 5, 5, device="cuda") * 0.05
if w.dim() == 4 and cin % 64 == 0 and cout % 64 == 0:
    w = w.contiguous(memory_format=torch.channels_last)      # the layout the models keep their conv weights in
x = torch.randn(B, hw, hw, cin, device="cuda").to(torch.bfloat16)
y = layer.fwd(x, w, None)
dy = torch.randn_like(y)
def once():
    if which.endswith("_dgrad"):
        layer.dgrad(dy, w, tuple(x.shape))
    elif which.endswith("_wgrad"):
        layer.wgrad(x, dy, w)
    else:
        layer.fwd(x, w, None)
if iters > 5:      # warm the clocks up (~0.3 s) before timing; ncu runs use iters <= 5 and skip this
    import time
    t0 = time.time()
    while time.time() - t0 < 0.3:
        once()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    if which.endswith("_dgrad"):
        layer.dgrad(dy, w, tuple(x.shape))
    elif which.endswith("_wgrad"):
        layer.wgrad(x, dy, w)
    else:
        layer.fwd(x, w, None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flops = 2.0 * B * hw * hw * cin * cout * 25 / (4 if kind == "conv" else 1)
print(f"{which}: {ms*1e3:.1f} us per call, {flops/ms/1e9:.1f} TFLOP/s")
