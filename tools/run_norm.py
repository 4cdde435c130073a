"""ncu target: the backward-reduce / backward-apply norm passes on the largest VAE activation (1M x 64 bf16)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_play_b200 import _lib
from vae_play_b200.functional import _ptr, _stream
rows, c = 1 << 20, 64
x = torch.randn(rows, c, device="cuda").to(torch.bfloat16)
da = torch.randn(rows, c, device="cuda").to(torch.bfloat16)
a = torch.empty_like(x)
sums = torch.zeros(2 * c, dtype=torch.float64, device="cuda")
st = torch.rand(4, c, device="cuda") + 0.5
for _ in range(3):
    _lib.call("vp_norm_bwd_reduce", _ptr(x), _ptr(da), _ptr(st[0]), _ptr(st[1]), _ptr(st[2]), _ptr(st[3]), _ptr(sums), None, 1, 1, rows, c, 1, 0.0, _stream())
    _lib.call("vp_norm_bwd_apply", _ptr(x), _ptr(da), _ptr(st[0]), _ptr(st[1]), _ptr(st[2]), _ptr(st[3]), _ptr(sums), _ptr(a), None, None, 1, 1, rows, c, 1, 0.0, _stream())
torch.cuda.synchronize()
print("ok")
