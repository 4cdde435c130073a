"""Device time of every contraction of the VAE step (fwd / dgrad / wgrad per layer), CUDA events, one process.
usage: python tools/bench_layers.py [img=64] [batch=256] [iters=20]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vae_play_b200 as vp
import vae_play_b200.functional as VF

img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
only = sys.argv[4] if len(sys.argv) > 4 else None
vp.set_precision("bf16")
import math
L = int(math.log2(img // 8))
layers = []   # name, TapLayer, weight shape, input (h, cin)
c, h = 1, img
for i in range(L):
    co = 64 if i == 0 else c * 2
    layers.append((f"enc{i+1}", VF.TapLayer("conv", c, co, k=5, stride=2, pad=2), (co, c, 5, 5), h, c))
    c, h = co, h // 2
size = c
layers.append(("enc_fc", VF.TapLayer("linear", 64 * size, 1024), (1024, 64 * size), 1, 64 * size))
layers.append(("dec_fc", VF.TapLayer("linear", 128, 64 * size), (64 * size, 128), 1, 128))
c, h = size, 8
for i in range(L):
    co = c if i == 0 else c // 2
    layers.append((f"ct{i+1}", VF.TapLayer("convT", c, co, k=5, stride=2, pad=2, out_pad=1), (c, co, 5, 5), h, c))
    c, h = co, h * 2
layers.append(("out", VF.TapLayer("conv", c, 1, k=5, stride=1, pad=2), (1, c, 5, 5), h, c))

def timeit(fn):
    """Device time per call: `iters` calls captured into one CUDA graph (no host launch overhead in the timed region)."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn()
        torch.cuda.synchronize()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

t0 = time.time()
x = torch.randn(4096, 4096, device="cuda")
while time.time() - t0 < 0.5:
    x @ x
tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0}
totf = 0.0
print(f"{'layer':8s} {'GFLOP':>8s} | {'fwd us':>8s} {'TF/s':>7s} | {'dgrad us':>8s} {'TF/s':>7s} | {'wgrad us':>8s} {'TF/s':>7s}")
for name, layer, wshape, hin, cin in layers:
    if only and only not in name:
        continue
    w = torch.randn(*wshape, device="cuda") * 0.05
    if w.dim() == 4 and wshape[0] % 64 == 0 and wshape[1] % 64 == 0:
        w = w.contiguous(memory_format=torch.channels_last)
    xin = torch.randn(B, hin, hin, cin, device="cuda").to(torch.bfloat16)
    y = layer.fwd(xin, w, None)
    dy = torch.randn_like(y)
    mac = {"conv": B * y.shape[1] * y.shape[2] * layer.cout * layer.cin * 25, "convT": B * hin * hin * layer.cin * layer.cout * 25,
           "linear": B * wshape[0] * wshape[1]}[layer.kind]
    gf = 2.0 * mac / 1e9
    tf = timeit(lambda: layer.fwd(xin, w, None))
    td = timeit(lambda: layer.dgrad(dy, w, tuple(xin.shape))) if name != "enc1" else 0.0
    tw = timeit(lambda: layer.wgrad(xin, dy, w))
    tot["fwd"] += tf; tot["dgrad"] += td; tot["wgrad"] += tw
    totf += gf * (3 if td else 2)
    r = lambda t: gf / t * 1e3 if t else 0.0
    print(f"{name:8s} {gf:8.2f} | {tf:8.1f} {r(tf):7.0f} | {td:8.1f} {r(td):7.0f} | {tw:8.1f} {r(tw):7.0f}")
s = sum(tot.values())
print(f"total: fwd {tot['fwd']:.0f} us, dgrad {tot['dgrad']:.0f} us, wgrad {tot['wgrad']:.0f} us (incl. weight packing when stale, memset, unpack) = {s:.0f} us; "
      f"{totf:.0f} GFLOP -> {totf / s * 1e3:.0f} TFLOP/s")
