"""Per-kernel device time of the VAE training step exactly as bench.py runs it (two CUDA graphs per step, fused
optimiser, persistent gradient slots), via torch.profiler / CUPTI.  Diagnostic, not a benchmark.
usage: python tools/step_profile.py [img=64] [batch=256] [--eager]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import vae_play_b200 as vp
import vae_play_b200.functional as VF
from vae_play_b200.models.networks import VaeGan
from vae_play_b200.optim import FusedRMSprop

args = [a for a in sys.argv[1:] if not a.startswith("--")]
eager = "--eager" in sys.argv
img = int(args[0]) if len(args) > 0 else 64
B = int(args[1]) if len(args) > 1 else 256
dev = torch.device("cuda", 0)
vp.set_precision("bf16")
torch.manual_seed(0)
model = VaeGan(img, 128).to(dev).train()
params = list(model.encoder.parameters()) + list(model.decoder.parameters())
opt = FusedRMSprop(params, lr=1e-4, zero_grads=True)
VF.persistent_grads(params)
x = torch.rand(B, 1, img, img, device=dev)
off = torch.zeros(1, dtype=torch.int64, device=dev)

VF.set_async_wgrad("--no-async-wgrad" not in sys.argv)

def fwd_bwd():
    xt, mulv, kl = model.vae_forward(x, rng=(0, 0, off))
    VF.philox_advance(off, 4)
    loss = VF.vae_loss(x, xt, kl)
    loss.backward()
    VF.join_async()
    return loss

def eager_step():
    opt.zero_grad(set_to_none=True)
    fwd_bwd()
    opt.step()

side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        eager_step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
if eager:
    step = eager_step
else:
    VF.invalidate_caches()
    opt.zero_grad(set_to_none=True)
    ga, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(ga):
        fwd_bwd()
    with torch.cuda.graph(gb, pool=ga.pool()):
        opt.step()
    def step():
        ga.replay()
        gb.replay()
for _ in range(3):
    step()
torch.cuda.synchronize()
N = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(N):
    step()
e1.record()
torch.cuda.synchronize()
print(f"wall (CUDA events): {e0.elapsed_time(e1) / N * 1e3:.1f} us per step")
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
print(f"total device time per step: {tot:.1f} us in {sum(r[1] for r in rows):.0f} launches")
for t, c, k in rows[:45]:
    k = k.replace("void ", "").replace("vp::", "").replace("(anonymous namespace)::", "")
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  n={c:5.1f}  {k[:110]}")
