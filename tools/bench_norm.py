"""Achieved HBM GB/s of the norm/activation streaming kernels at the VAE's shapes (diagnostic)."""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vae_play_b200 as vp
from vae_play_b200 import _lib
from vae_play_b200.functional import _ptr, _stream

lib = _lib.load()
def t(fn, iters=20):
    t0 = time.time()
    while time.time() - t0 < 0.2:
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3

for rows, c in [(1 << 20, 64), (1 << 18, 128), (1 << 16, 256), (1 << 14, 256), (1 << 18, 64), (1 << 20, 1), (256, 16384)]:
    x = torch.randn(rows, c, device="cuda").to(torch.bfloat16)
    da = torch.randn(rows, c, device="cuda").to(torch.bfloat16)
    a = torch.empty_like(x)
    sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
    st = torch.rand(4, c, device="cuda") + 0.5
    n = rows * c * 2
    r = {}
    r["stats"] = (t(lambda: _lib.call("vp_norm_stats", _ptr(x), _ptr(sums), 1, 1, rows, c, _stream())), n)
    r["apply"] = (t(lambda: _lib.call("vp_norm_apply_act", _ptr(x), _ptr(st[2]), _ptr(st[3]), _ptr(a), 1, 1, rows, c, 1, 0.0, _stream())), 2 * n)
    r["bwd_reduce"] = (t(lambda: _lib.call("vp_norm_bwd_reduce", _ptr(x), _ptr(da), _ptr(st[0]), _ptr(st[1]), _ptr(st[2]), _ptr(st[3]), _ptr(sums), None, 1, 1, rows, c, 1, 0.0, _stream())), 2 * n)
    r["bwd_apply"] = (t(lambda: _lib.call("vp_norm_bwd_apply", _ptr(x), _ptr(da), _ptr(st[0]), _ptr(st[1]), _ptr(st[2]), _ptr(st[3]), _ptr(sums), _ptr(a), None, None, 1, 1, rows, c, 1, 0.0, _stream())), 3 * n)
    print(f"rows={rows:8d} C={c:5d} " + "  ".join(f"{k}: {us:7.1f} us {b/us/1e3:6.0f} GB/s" for k, (us, b) in r.items()), flush=True)
