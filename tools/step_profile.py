"""Per-kernel device time of the eager VAE step (torch.profiler / CUPTI).  Diagnostic, not a benchmark."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import vae_play_b200 as vp
import vae_play_b200.functional as VF
from vae_play_b200.models.networks import VaeGan

img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
vp.set_precision("bf16")
torch.manual_seed(0)
model = VaeGan(img, 128).to(dev).train()
params = list(model.encoder.parameters()) + list(model.decoder.parameters())
opt = torch.optim.RMSprop(params, lr=1e-4)
x = torch.rand(B, 1, img, img, device=dev)
off = torch.zeros(1, dtype=torch.int64, device=dev)

def step():
    opt.zero_grad(set_to_none=True)
    xt, mulv, kl = model.vae_forward(x, rng=(0, 0, off))
    VF.philox_advance(off, 4)
    loss = VF.vae_loss(x, xt, kl)
    loss.backward()
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
    if t:
        rows.append((t / N, e.count / N, e.key))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print(f"total device time per step: {tot:.1f} us")
for t, c, k in rows[:40]:
    k = k.replace("void ", "").replace("vp::", "").replace("(anonymous namespace)::", "")
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  n={c:5.1f}  {k[:110]}")
