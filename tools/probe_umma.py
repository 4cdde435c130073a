"""Probe: can a tcgen05 K-major SW128 A operand start at a non-1024-byte-aligned row of a larger TMA-written tile?
Needs a debug build of the library: VP_DEBUG_PROBES=1 python -m vae_play_b200.build --force"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vae_play_b200 import _lib
from vae_play_b200.functional import _ptr, _stream
x = (torch.arange(256 * 64, device="cuda").reshape(256, 64) % 251).float()
x = (x + torch.arange(256, device="cuda").float()[:, None] * 0.25).to(torch.bfloat16)
ident = torch.eye(64, device="cuda").to(torch.bfloat16)
out = torch.empty(128, 64, device="cuda")
import ctypes as C
_probe = _lib.load().vp_debug_umma_probe      # not in the release ABI / header: bound here
_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
_probe.restype = C.c_int
xf = x.float()
for off in (0, 1, 3, 7, 8, 10, 21):
    for sbo in (8, 10, 18, 12):
        if off + 15 * sbo + 8 > 256:
            continue
        want = torch.stack([xf[off + (m // 8) * sbo + (m % 8)] for m in range(128)])
        res = []
        for bo in sorted({0, off & 7}):
            _lib.check(_probe(_ptr(x), _ptr(ident), _ptr(out), off, sbo, bo, _stream()), "vp_debug_umma_probe")
            torch.cuda.synchronize()
            ok = torch.equal(out, want)
            nbad = int((out != want).any(dim=1).sum())
            res.append(f"base_offset={bo}: {'OK' if ok else f'MISMATCH ({nbad} rows)'}")
        print(f"off={off:3d} sbo_rows={sbo:3d}  " + "   ".join(res), flush=True)
